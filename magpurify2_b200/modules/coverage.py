"""`magpurify2 coverage` driver (entry point of magpurify2/modules/coverage.py:33-137).

Same options, same cache file (coverages.pickle.gz: (signatures, contig_names, matrix), reused when it holds
every input file of the run -- columns picked by signature -- and deleted otherwise) and the same TSV.  The
rows of a MAG are found through a dictionary instead of the reference's O(N^2) np.where scan (:113-119), the
samples are spread over the GPUs given with --devices, and all MAGs are scored by one batched GPU call."""
import logging

import numpy as np

from .. import tools
from ..core import Coverage, Mag, read_mags, score_coverage_batch  # noqa: F401
from ..multi import bam_coverages_on_devices, parse_devices
from ._common import COVERAGE_LIMITS, EMBEDDING_LIMITS, PickleCache, clamp_options, note_provenance, scores_path


def _measure(args, inputs, mags, logger):
    """(contig_names, matrix) of this run's inputs: BAM files on the GPUs, or the tabular file as it is."""
    wanted = set()
    for m in mags:
        wanted.update(m.contigs)
    if args.bam_files:
        logger.info(f"Computing contig coverages from {len(inputs)} BAM files.")
        names, matrix = bam_coverages_on_devices([str(p) for p in inputs], parse_devices(args),
                                                 min_identity=args.min_identity, trim_lower=args.trim_lower,
                                                 trim_upper=args.trim_upper, contig_set=wanted, threads=args.threads)
        return np.array(names), matrix
    logger.info(f"Reading contig coverages from '{inputs[0]}'.")
    return tools.get_tsv_coverages(inputs[0], contig_set=wanted)


def main(args):
    logger = logging.getLogger("timestamp")
    logger.info("Executing MAGpurify2 coverage module.")
    # the reference stores the clamped min_identity in args.contig_min_fraction and passes min_identity on as
    # given (coverage.py:36-38): kept
    args.contig_min_fraction = tools.validade_input(args.min_identity, "min_identity", [0.0, 1.0])
    clamp_options(args, COVERAGE_LIMITS)
    clamp_options(args, EMBEDDING_LIMITS)
    score_file = scores_path(args, "coverage")
    if args.bam_files:
        tools.check_bam_files(args.bam_files)
    inputs = list(args.bam_files) if args.bam_files else [args.coverage_file]

    logger.info(f"Reading {len(args.genomes)} genomes.")
    mags = read_mags(args.genomes, threads=getattr(args, "threads", 1), store_sequences=False)
    signatures = [tools.get_file_signature(p) for p in inputs]
    cache = PickleCache(args.output_directory.joinpath("coverages.pickle.gz"), "contig coverages")
    hit = cache.load(lambda t: isinstance(t, tuple) and len(t) == 3 and set(signatures).issubset(t[0]))
    if hit is not None:
        stored, names, matrix = hit
        if len(stored) > len(signatures):  # the cache knows more files than this run asked for
            matrix = matrix[:, [s in signatures for s in stored]]
    else:
        names, matrix = _measure(args, inputs, mags, logger)
        cache.store((signatures, names, matrix))

    # rows of every MAG by contig name (the first occurrence, as np.where(...)[0][0] picks, coverage.py:114)
    row_of = {}
    for i, name in enumerate(names):
        row_of.setdefault(str(name), i)
    n_rows = sum(len(m) for m in mags)
    stacked = np.asarray(matrix)[np.fromiter((row_of[c] for m in mags for c in m.contigs), dtype=np.int64, count=n_rows)]
    first_row = np.cumsum([0] + [len(m) for m in mags])

    logger.info("Computing contig scores.")
    scores, n_samples, selected = score_coverage_batch(mags, stacked, args.min_average_coverage)
    per_mag = [
        Coverage(m, stacked[first_row[i]:first_row[i + 1]], args.min_average_coverage,
                 scores=(scores[first_row[i]:first_row[i + 1]], int(n_samples[i]), selected[i]))
        for i, m in enumerate(mags)
    ]
    logger.info(f"Writing output to: '{score_file}'.")
    tools.write_module_output(per_mag, score_file)
    note_provenance(args, "coverage")
    return per_mag
