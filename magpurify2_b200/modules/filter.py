"""`magpurify2 filter`: contamination probability of every contig from the score tables the modules wrote,
and the filtered genomes (mirror of magpurify2/modules/filter.py:32-170).

`--fast_mode` needs only scores/composition_scores.tsv and scores/coverage_scores.tsv -- the two tables this
build writes.  Without it the codon_usage and taxonomy tables of the reference's other modules are read too.
The gradient-boosted model is the reference's JSON file: it is looked up next to an installed `magpurify2`
package (models/fast_model.json, models/full_model.json) or given with --model_file; it is evaluated by
`_xgb.XgbJsonModel`, so xgboost is not needed."""
import importlib.util
import logging
import sys
from pathlib import Path

import numpy as np

from .. import tools
from ..core import ContigClassifier, read_mags

MODULE_NCOL = {"composition": 5, "coverage": 5, "codon_usage": 7, "taxonomy": 6}  # filter.py:93


def find_model_file(fast_mode, model_file=None):
    name = "fast_model.json" if fast_mode else "full_model.json"
    if model_file:
        return Path(model_file)
    spec = importlib.util.find_spec("magpurify2")
    if spec and spec.submodule_search_locations:
        cand = Path(list(spec.submodule_search_locations)[0]).joinpath("models", name)
        if cand.exists():
            return cand
    return None


def load_features(scores_directory, module_list, logger):
    """(genome_contig_matrix, feature_matrix) exactly as filter.py:84-136 builds them."""
    paths = [scores_directory.joinpath(f"{m}_scores.tsv") for m in module_list]
    missing = [str(p) for p in paths if not p.exists()]
    if missing:
        logger.error(f"The following files were not found: '{', '.join(missing)}'.")
        sys.exit(1)
    read_names = lambda p: np.genfromtxt(p, dtype=str, skip_header=1, comments=None, delimiter="\t",  # noqa: E731
                                         usecols=[0, 1]).reshape(-1, 2)
    genome_contig_matrix = read_names(paths[0])
    feature_matrix = []
    for module, path in zip(module_list, paths):
        names = read_names(path)
        if names.shape != genome_contig_matrix.shape or not (names == genome_contig_matrix).all():
            logger.error("It seems that MAGpurify2's modules were executed with different inputs. Please make sure "
                         "to execute all the modules with the exact same inputs in the same order.")
            sys.exit(1)
        feats = np.genfromtxt(path, dtype=float, skip_header=1, comments=None, delimiter="\t",
                              usecols=range(2, MODULE_NCOL[module]))
        feature_matrix.append(feats.reshape(len(genome_contig_matrix), -1))
    feature_matrix = np.column_stack(feature_matrix)
    feature_matrix[:, 2] = np.tanh(np.log10(feature_matrix[:, 2] + 1) * 0.2)  # contig_length
    feature_matrix[:, 5] = np.clip(feature_matrix[:, 5] / 10, 0, 1)            # n_samples
    if "codon_usage" in module_list:  # filter.py:126-135
        feature_matrix[:, 8] = np.tanh(np.log10(feature_matrix[:, 8] + 1))         # n_genes
        feature_matrix[:, 9] = np.tanh(np.log10(feature_matrix[:, 9] + 1) * 0.2)   # total_cds_length
        feature_matrix[:, 12] = feature_matrix[:, 12] / 6                           # genome_rank
        feature_matrix[:, 13] = feature_matrix[:, 13] / 6                           # contig_rank
    return genome_contig_matrix, feature_matrix


def warn_uncalibrated(scores_directory, logger):
    """The reference's trees were fitted on the reference's features.  Two columns of this build's tables are
    not those features (tnf_score is a CLR / centroid-distance score, coverage_cluster_score is the constant
    1.0), so the probabilities -- and with them the contigs that get removed -- are NOT calibrated when the
    tables come from this build.  Say so, loudly, whenever the provenance note shows that."""
    import json

    note = scores_directory.joinpath("provenance.json")
    try:
        known = json.loads(note.read_text()) if note.exists() else {}
    except ValueError:
        known = {}
    ours = sorted(m for m, v in known.items() if isinstance(v, dict) and v.get("producer") == "magpurify2_b200")
    if ours:
        logger.warning(
            f"The {' and '.join(ours)} score table(s) were written by the B200 build: `tnf_score` is a CLR / "
            "centroid-distance score and `coverage_cluster_score` is the constant 1.0, not the UMAP + HDBSCAN "
            "scores the gradient-boosted model was trained on.  The contamination probabilities below are "
            "therefore NOT calibrated and are not comparable with the reference's; treat the filtered genomes as "
            "a screening result.")
    return ours


def main(args):
    logger = logging.getLogger("timestamp")
    args.probability_threshold = tools.validade_input(args.probability_threshold, "probability_threshold", [0.0, 1.0])
    if not args.output_directory.is_dir():
        logger.error(f"The directory '{args.output_directory}' was not found. Please sure you have provided the "
                     "correct path and that MAGpurify2's modules have been executed.")
        sys.exit(1)
    scores_directory = args.output_directory.joinpath("scores")
    if not scores_directory.is_dir():
        logger.error(f"The directory '{scores_directory}' was not found. Please sure you have executed "
                     "MAGpurify2's modules.")
        sys.exit(1)
    if not args.filtered_output_directory.is_dir():
        logger.info(f"'{args.filtered_output_directory}' does not exist. Creating it now.")
        args.filtered_output_directory.mkdir()
    contig_prob_file = scores_directory.joinpath("contamination_probabilities.tsv")
    model_file = find_model_file(args.fast_mode, getattr(args, "model_file", None))
    if model_file is None or not Path(model_file).exists():
        logger.error("The gradient-boosted model was not found: install the `magpurify2` package (its models/ "
                     "directory is used) or pass --model_file <fast_model.json | full_model.json>.")
        sys.exit(1)

    logger.info(f"Reading {len(args.genomes)} genomes.")
    mag_list = read_mags(args.genomes, threads=getattr(args, "threads", 1))
    module_list = ["composition", "coverage"] if args.fast_mode else ["composition", "coverage", "codon_usage", "taxonomy"]
    logger.info(f"Reading contig scores from the {', '.join(module_list)} modules.")
    genome_contig_matrix, feature_matrix = load_features(scores_directory, module_list, logger)

    warn_uncalibrated(scores_directory, logger)
    logger.info("Estimating contamination probabilities and identifying contaminant contigs.")
    classification = ContigClassifier(genome_contig_matrix, feature_matrix, model_file, args.fast_mode,
                                      args.probability_threshold, args.checkm_file, getattr(args, "threads", 1))
    logger.info(f"Writing contig contamination probabilities to: '{contig_prob_file}'.")
    tools.write_module_output([classification], contig_prob_file)

    logger.info(f"Writing filtered genomes to '{args.filtered_output_directory}'.")
    skipped = {m.genome for m in mag_list}.difference(classification.mags_contaminants_dict.keys())
    if skipped:
        logger.warning("The following genomes were not previously analysed by any of MAGpurify2's modules and will "
                       f"be skipped: {', '.join(sorted(skipped))}.")
    for mag in mag_list:
        tools.write_filtered_genome(mag, classification.mags_contaminants_dict, args.filtered_output_directory,
                                    args.suffix)
    return classification
