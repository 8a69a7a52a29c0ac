"""`magpurify2 coverage` driver (mirror of magpurify2/modules/coverage.py:33-137).

Same options, same cache file (coverages.pickle.gz: (signatures, contig_names, matrix)) and the
same TSV.  The per-MAG row gather is a dict lookup instead of the reference's O(N^2) np.where
loop (:113-119), and all MAGs are scored by one batched GPU call."""
import gzip
import logging
import pickle
from functools import reduce

import numpy as np

from .. import tools
from ..core import Coverage, Mag, read_mags, score_coverage_batch  # noqa: F401


def main(args):
    logger = logging.getLogger("timestamp")
    logger.info("Executing MAGpurify2 coverage module.")
    # the reference stores this clamp in args.contig_min_fraction (coverage.py:36-38); min_identity
    # itself is passed on unclamped, which is kept
    args.contig_min_fraction = tools.validade_input(args.min_identity, "min_identity", [0.0, 1.0])
    args.trim_lower = tools.validade_input(args.trim_lower, "trim_lower", [0.0, 1.0])
    args.trim_upper = tools.validade_input(args.trim_upper, "trim_upper", [0.0, 1.0])
    args.min_average_coverage = tools.validade_input(args.min_average_coverage, "min_average_coverage", [0, 999])
    args.n_iterations = tools.validade_input(args.n_iterations, "n_iterations", [1, 999])
    args.n_components = tools.validade_input(args.n_components, "n_components", [1, 999])
    args.n_neighbors = tools.validade_input(args.n_neighbors, "n_neighbors", [1, 999])
    args.set_op_mix_ratio = tools.validade_input(args.set_op_mix_ratio, "set_op_mix_ratio", [0.0, 1.0])
    tools.check_output_directory(args.output_directory)
    scores_directory = args.output_directory.joinpath("scores")
    scores_directory.mkdir(exist_ok=True)
    coverage_score_file = scores_directory.joinpath("coverage_scores.tsv")

    if args.bam_files:
        tools.check_bam_files(args.bam_files)
        input_coverage = args.bam_files
    else:
        input_coverage = [args.coverage_file]

    logger.info(f"Reading {len(args.genomes)} genomes.")
    mag_list = read_mags(args.genomes, threads=getattr(args, "threads", 1), store_sequences=False)
    coverage_data_file = args.output_directory.joinpath("coverages.pickle.gz")
    signatures = [tools.get_file_signature(filepath) for filepath in input_coverage]

    skip_coverage = False
    if coverage_data_file.exists():
        with gzip.open(coverage_data_file) as fin:
            loaded_signatures, contig_names, coverage_matrix = pickle.load(fin)
            if set(signatures).issubset(loaded_signatures):
                logger.info("Skipping contig coverages computation.")
                skip_coverage = True
                if len(loaded_signatures) > len(signatures):
                    selected_input_files = [signature in signatures for signature in loaded_signatures]
                    coverage_matrix = coverage_matrix[:, selected_input_files]
            else:
                coverage_data_file.unlink()
    if not skip_coverage:
        contig_set = reduce(lambda x, y: x.union(y.contigs), mag_list, set())
        if args.bam_files:
            logger.info(f"Computing contig coverages from {len(input_coverage)} BAM files.")
            contig_names, coverage_matrix = tools.get_bam_coverages(
                [str(filepath) for filepath in input_coverage],
                min_identity=args.min_identity,
                trim_lower=args.trim_lower,
                trim_upper=args.trim_upper,
                contig_set=contig_set,
                threads=args.threads,
            )
            contig_names = np.array(contig_names)
        else:
            logger.info(f"Reading contig coverages from '{input_coverage[0]}'.")
            contig_names, coverage_matrix = tools.get_tsv_coverages(input_coverage[0], contig_set=contig_set)
        logger.info(f"Saving contig coverages data to '{coverage_data_file}'.")
        with gzip.open(coverage_data_file, "wb") as fout:
            pickle.dump((signatures, contig_names, coverage_matrix), fout)

    # rows of every MAG by name (first occurrence, as np.where(...)[0][0] picks, coverage.py:114)
    index = {}
    for i, name in enumerate(contig_names):
        index.setdefault(str(name), i)
    rows = np.fromiter((index[c] for m in mag_list for c in m.contigs), dtype=np.int64,
                       count=sum(len(m) for m in mag_list))
    stacked = np.asarray(coverage_matrix)[rows]
    row = np.cumsum([0] + [len(m) for m in mag_list])

    logger.info("Computing contig scores.")
    scores, n_samples, selected = score_coverage_batch(mag_list, stacked, args.min_average_coverage)
    mag_coverage_list = [
        Coverage(m, stacked[row[i]:row[i + 1]], args.min_average_coverage,
                 scores=(scores[row[i]:row[i + 1]], int(n_samples[i]), selected[i]))
        for i, m in enumerate(mag_list)
    ]
    logger.info(f"Writing output to: '{coverage_score_file}'.")
    tools.write_module_output(mag_coverage_list, coverage_score_file)
    return mag_coverage_list
