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


def score_composition_batch(mags, counts, gc):
    """counts uint32[n,136], gc float32[n] for all contigs of `mags` stacked.
    Returns (tnf_scores f64[n], gc_content_scores f64[n])."""
    mag_off = _mag_offsets(mags)
    lengths = np.concatenate([m.lengths for m in mags]) if mags else np.zeros(0, np.int64)
    dist = _scoring.clr_centroid_batch(counts, lengths, mag_off)
    tnf = _scoring.log_ratio_scores_batch(dist.astype(np.float32), lengths, mag_off, TNF_SCORE_BASE,
                                          min_rows=3, one_sided=True)
    gcs = _scoring.log_ratio_scores_batch(gc, lengths, mag_off, 2, min_rows=3)  # core.py:261-263
    return tnf, gcs


def score_coverage_batch(mags, coverages, min_average_coverage):
    """coverages float32[n, n_samples] stacked over `mags`.
    Returns (scores f64[n], n_samples int[n_mags], selected bool[n_mags, n_samples])."""
    mag_off = _mag_offsets(mags)
    lengths = np.concatenate([m.lengths for m in mags]) if mags else np.zeros(0, np.int64)
    cov = np.ascontiguousarray(coverages, dtype=np.float32)
    sel, _means = _scoring.select_samples_batch(cov, lengths, mag_off, min_average_coverage)  # core.py:301-305
    scores = _scoring.log_ratio_scores_batch(cov, lengths, mag_off, 25, col_mask=sel, min_rows=3)  # :310-312
    return scores, sel.sum(axis=1), sel


class Composition:
    attributes = ["genome", "contig", "tnf_score", "gc_content_score", "contig_length"]

    def __init__(self, mag, composition_dict=None, gc_content=None, n_iterations=None, n_components=None,
                 min_dist=None, n_neighbors=None, set_op_mix_ratio=None, *, scores=None):
        """Reference shape: Composition(mag, composition_dict, gc_content, <UMAP options>)
        (core.py:225-235).  composition_dict[mag.genome] must hold the RAW uint32 counts
        (get_tnf(..., relative=False)) or the relative profile together with lengths (the counts are
        then recovered as round(rel * (len - 3)) only when no base was skipped, so pass counts).
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
                raise TypeError("Composition needs the raw uint32 4-mer counts (get_tnf(..., relative=False))")
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
