"""Mirror of the per-MAG score objects of magpurify2/core.py that sit on the two hot paths:
Mag (core.py:35-62), Composition (:224-275), Coverage (:278-336).  Same attributes, same TSV rows.

The classes keep the reference's per-MAG constructor shape; the module drivers do NOT loop over
them for compute -- they score every MAG of a run with one batched GPU call
(score_composition_batch / score_coverage_batch) and hand each object its slice.

tnf_score: the reference derives it from UMAP + HDBSCAN soft-cluster membership
(tools.py:529-575), which is outside this build (not a data-parallel scan, libraries absent,
nondeterministic).  Here it is the north-star score: centred-log-ratio transform of the
pseudocounted canonical 4-mer counts, Euclidean distance to the MAG's length-weighted centroid,
mapped to [0,1] as 1 - max(0, log4(d / weighted median d)).  coverage_cluster_score (UMAP again)
is reported as 1.0.  See DESIGN.md."""
import numpy as np

from . import _scoring, tools
from ._fasta import Fasta

TNF_SCORE_BASE = 4.0


def read_mags(genomes, threads=1, store_sequences=True):
    """[Mag(g) for g in genomes] on a thread pool: the FASTA reader is native code that releases the GIL, so
    the genomes of a run are parsed `threads` at a time (the reference parses them one by one,
    modules/composition.py:48)."""
    genomes = list(genomes)
    threads = max(1, min(int(threads or 1), len(genomes)))
    if threads == 1:
        return [Mag(g, store_sequences=store_sequences) for g in genomes]
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(lambda g: Mag(g, store_sequences=store_sequences), genomes))


class Mag:
    def __init__(self, filepath, store_sequences=True):
        self.genome = filepath.stem
        if tools.is_compressed(filepath) != tools.Compression.noncompressed:
            self.genome = self.genome.rsplit(".", 1)[0]
        with Fasta(filepath) as fa:
            self.contigs = list(fa.ids)
            self.descriptions = list(fa.descriptions)
            self.lengths = np.asarray(fa.lengths, dtype=np.int64).copy()
            # device-friendly form (one byte stream + offsets) -- what the kernels take
            self.offsets = fa.offsets.copy()
            self.bases = fa.bases.copy() if store_sequences else None
        self._sequences = None
        self._store = store_sequences

    @property
    def sequences(self):
        """list[str] as in the reference (built lazily; the drivers use .bases/.offsets)."""
        if not self._store:
            return []
        if self._sequences is None:
            b, o = self.bases, self.offsets
            self._sequences = [b[o[i]:o[i + 1]].tobytes().decode("ascii", "replace") for i in range(len(self))]
        return self._sequences

    def __len__(self):
        return len(self.contigs)

    def __iter__(self):
        if self._store:
            return zip(self.contigs, self.descriptions, self.sequences)
        return iter(self.contigs)

    def __repr__(self):
        return f"{self.genome}: {len(self)} contigs"


def _mag_offsets(mags):
    off = np.zeros(len(mags) + 1, dtype=np.int64)
    np.cumsum([len(m) for m in mags], out=off[1:])
    return off


def score_composition_batch(mags, counts, gc, device=-1):
    """counts uint32[n,136], gc float32[n] for all contigs of `mags` stacked.
    Returns (tnf_scores f64[n], gc_content_scores f64[n])."""
    mag_off = _mag_offsets(mags)
    lengths = np.concatenate([m.lengths for m in mags]) if mags else np.zeros(0, np.int64)
    dist = _scoring.clr_centroid_batch(counts, lengths, mag_off, device=device)
    tnf = _scoring.log_ratio_scores_batch(dist.astype(np.float32), lengths, mag_off, TNF_SCORE_BASE,
                                          min_rows=3, one_sided=True, device=device)
    gcs = _scoring.log_ratio_scores_batch(gc, lengths, mag_off, 2, min_rows=3, device=device)  # core.py:261-263
    return tnf, gcs


def score_coverage_batch(mags, coverages, min_average_coverage, device=-1):
    """coverages float32[n, n_samples] stacked over `mags`.
    Returns (scores f64[n], n_samples int[n_mags], selected bool[n_mags, n_samples])."""
    mag_off = _mag_offsets(mags)
    lengths = np.concatenate([m.lengths for m in mags]) if mags else np.zeros(0, np.int64)
    cov = np.ascontiguousarray(coverages, dtype=np.float32)
    sel, _means = _scoring.select_samples_batch(cov, lengths, mag_off, min_average_coverage, device=device)  # core.py:301-305
    scores = _scoring.log_ratio_scores_batch(cov, lengths, mag_off, 25, col_mask=sel, min_rows=3, device=device)  # :310-312
    return scores, sel.sum(axis=1), sel


class Composition:
    attributes = ["genome", "contig", "tnf_score", "gc_content_score", "contig_length"]

    def __init__(self, mag, composition_dict=None, gc_content=None, n_iterations=None, n_components=None,
                 min_dist=None, n_neighbors=None, set_op_mix_ratio=None, *, scores=None):
        """Reference shape: Composition(mag, composition_dict, gc_content, <UMAP options>)
        (core.py:225-235).  composition_dict[mag.genome] must hold the RAW uint32 counts
        (get_tnf(..., relative=False)) or, as in the reference, the relative profile (the counts are then
        counted again from mag.bases).
        `scores=(tnf, gc)` short-circuits to precomputed batch results."""
        self.attributes = list(Composition.attributes)
        self.genome = mag.genome
        self.contigs = mag.contigs
        self.lengths = mag.lengths
        self.tnf = None if composition_dict is None else composition_dict[mag.genome]
        self.gc_content = gc_content
        if scores is not None:
            self.tnf_scores, self.gc_content_scores = scores
        elif len(self) <= 2:  # core.py:248-250
            self.tnf_scores = np.ones(len(self))
            self.gc_content_scores = np.ones(len(self))
        else:
            counts = np.asarray(self.tnf)
            if counts.dtype != np.uint32:
                # the reference's callers (and tnfs.pickle.gz) hold the RELATIVE profile, which is what
                # tools.get_tnf returns by default; the CLR score needs the raw counts, so they are counted
                # again from the sequences (one kernel launch)
                if getattr(mag, "bases", None) is None:
                    raise TypeError("Composition got a relative TNF profile and a Mag without sequences: pass the raw "
                                    "uint32 counts (get_tnf(..., relative=False)) or read the Mag with "
                                    "store_sequences=True")
                from . import _composition

                counts, _rel, _gc = _composition.tnf_from_concat(mag.bases, mag.offsets, relative=False)
            self.tnf_scores, self.gc_content_scores = score_composition_batch([mag], counts, gc_content)

    def __len__(self):
        return len(self.contigs)

    def __iter__(self):  # core.py:268-275
        return zip(
            np.repeat(self.genome, len(self)),
            self.contigs,
            np.round(self.tnf_scores, 5),
            np.round(self.gc_content_scores, 5),
            self.lengths,
        )


class Coverage:
    attributes = ["genome", "contig", "coverage_score", "coverage_cluster_score", "n_samples"]

    def __init__(self, mag, coverages, min_average_coverage, n_iterations=None, n_components=None, min_dist=None,
                 n_neighbors=None, set_op_mix_ratio=None, *, scores=None):
        """Reference shape: Coverage(mag, coverages, min_average_coverage, <UMAP options>) (core.py:279-289)."""
        self.attributes = list(Coverage.attributes)
        self.genome = mag.genome
        self.contigs = mag.contigs
        self.lengths = mag.lengths
        self.coverages = coverages
        if scores is not None:
            self.scores, self.n_samples, self.selected_samples = scores
        else:
            sc, ns, sel = score_coverage_batch([mag], np.asarray(coverages, dtype=np.float32).reshape(len(mag), -1),
                                               min_average_coverage)
            self.scores, self.n_samples, self.selected_samples = sc, int(ns[0]), sel[0]
        self.cluster_scores = np.ones(len(self))  # UMAP/HDBSCAN score: out of scope (see module docstring)

    def __len__(self):
        return len(self.contigs)

    def __iter__(self):  # core.py:329-336
        return zip(
            np.repeat(self.genome, len(self)),
            self.contigs,
            np.round(self.scores, 5),
            np.round(self.cluster_scores, 5),
            np.repeat(self.n_samples, len(self)),
        )


class ContigClassifier:
    """Mirror of magpurify2/core.py:527-583: contamination probability of every contig from the scores of
    the modules (gradient-boosted trees), fixed or CheckM-driven probability threshold, and the
    genome -> contig -> flagged dictionary the genome writer takes.  The model is evaluated by
    `_xgb.XgbJsonModel` (the xgboost package is not needed)."""

    def __init__(self, genome_contig_matrix, feature_matrix, model_file, fast_mode, probability_threshold,
                 checkm_file=None, threads=1):
        from collections import defaultdict

        from . import tools
        from ._xgb import XgbJsonModel

        self.attributes = ["genome", "contig", "contaminant_probability"]
        self.fast_mode = fast_mode
        self.base_probability_threshold = probability_threshold
        genome_contig_matrix = np.asarray(genome_contig_matrix).reshape(-1, 2)
        self.genomes = genome_contig_matrix[:, 0]
        self.contigs = genome_contig_matrix[:, 1]
        self.probabilities = XgbJsonModel(model_file).predict(np.asarray(feature_matrix).reshape(len(self.genomes), -1))
        self.probability_threshold = probability_threshold
        if checkm_file:
            checkm_scores = tools.get_checkm_contamination(checkm_file)
            if checkm_scores:
                tools.logger.info("Using completeness and contamination estimates to set dynamic thresholds.")
                if len(set(self.genomes)) > len(checkm_scores):
                    tools.logger.warning(
                        "The input CheckM tabular file does not contain all the genomes being filtered. The default "
                        f"probability threshold ({self.base_probability_threshold}) will be used for the genomes not "
                        "found in it.")
                self.probability_threshold = self.get_dynamic_thresholds(checkm_scores)
        self.flagged_contaminants = self.probabilities > self.probability_threshold
        self.mags_contaminants_dict = defaultdict(dict)
        for genome, contig, flagged in zip(self.genomes, self.contigs, self.flagged_contaminants):
            self.mags_contaminants_dict[genome][contig] = bool(flagged)

    def get_dynamic_thresholds(self, checkm_scores):  # core.py:585-596
        thresholds = np.array([
            0.85 + 0.72 * (np.exp(-0.141 * checkm_scores[g]) - 1) if g in checkm_scores
            else self.base_probability_threshold for g in self.genomes])
        return np.clip(thresholds, 0.1, 0.85)

    def __len__(self):
        return len(self.probabilities)

    def __iter__(self):
        return zip(self.genomes, self.contigs, np.round(self.probabilities, 5))
