"""`magpurify2 composition` driver (entry point of magpurify2/modules/composition.py:31-89).

Same options, same cache file (tnfs.pickle.gz: dict genome -> float32[n,136] relative TNF, reused when it
holds every genome of the run, deleted otherwise) and the same TSV.  Differences: every MAG of the run goes
through ONE native call (the reference calls the native function once per MAG, :63-66), the MAGs are spread
over the GPUs given with --devices, and the scores come from batched GPU kernels instead of a joblib fan-out."""
import logging

import numpy as np

from .. import tools
from ..core import Composition, Mag, read_mags, score_composition_batch  # noqa: F401
from ..multi import composition_on_devices, parse_devices
from ._common import EMBEDDING_LIMITS, PickleCache, clamp_options, note_provenance, scores_path


def main(args):
    logger = logging.getLogger("timestamp")
    logger.info("Executing MAGpurify2 composition module.")
    clamp_options(args, EMBEDDING_LIMITS)
    score_file = scores_path(args, "composition")
    logger.info(f"Reading {len(args.genomes)} genomes.")
    mags = read_mags(args.genomes, threads=getattr(args, "threads", 1))
    first_row = np.cumsum([0] + [len(m) for m in mags])

    # counts, relative TNF, GC and both scores of every contig; MAGs are split over the devices by base pairs
    logger.info("Computing tetranucleotide frequencies and contig scores.")
    res = composition_on_devices(mags, parse_devices(args), threads=getattr(args, "threads", None))

    cache = PickleCache(args.output_directory.joinpath("tnfs.pickle.gz"), "tetranucleotide frequencies")
    wanted = {m.genome for m in mags}
    tnf_of = cache.load(lambda d: isinstance(d, dict) and wanted.issubset(d.keys()))
    if tnf_of is None:
        tnf_of = {m.genome: res["rel"][first_row[i]:first_row[i + 1]] for i, m in enumerate(mags)}
        cache.store(tnf_of)

    per_mag = [
        Composition(m, tnf_of, res["gc"][first_row[i]:first_row[i + 1]],
                    scores=(res["tnf_score"][first_row[i]:first_row[i + 1]], res["gc_score"][first_row[i]:first_row[i + 1]]))
        for i, m in enumerate(mags)
    ]
    logger.info(f"Writing output to: '{score_file}'.")
    tools.write_module_output(per_mag, score_file)
    note_provenance(args, "composition")
    return per_mag
