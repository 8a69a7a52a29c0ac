"""`magpurify2 composition` driver (mirror of magpurify2/modules/composition.py:31-89).

Same options, same cache file (tnfs.pickle.gz: dict genome -> float32[n,136] relative TNF) and
the same TSV.  Differences: every MAG of the run goes through ONE mp2_tnf call (the reference
calls the native function once per MAG, :63-66) and the scores are computed by batched GPU
kernels instead of a joblib fan-out."""
import gzip
import logging
import pickle

import numpy as np

from .. import _composition, tools
from ..core import Composition, Mag, read_mags, score_composition_batch  # noqa: F401


def main(args):
    logger = logging.getLogger("timestamp")
    logger.info("Executing MAGpurify2 composition module.")
    args.n_iterations = tools.validade_input(args.n_iterations, "n_iterations", [1, 999])
    args.n_components = tools.validade_input(args.n_components, "n_components", [1, 999])
    args.n_neighbors = tools.validade_input(args.n_neighbors, "n_neighbors", [1, 999])
    args.set_op_mix_ratio = tools.validade_input(args.set_op_mix_ratio, "set_op_mix_ratio", [0.0, 1.0])
    tools.check_output_directory(args.output_directory)
    scores_directory = args.output_directory.joinpath("scores")
    scores_directory.mkdir(exist_ok=True)
    composition_score_file = scores_directory.joinpath("composition_scores.tsv")

    logger.info(f"Reading {len(args.genomes)} genomes.")
    mag_list = read_mags(args.genomes, threads=getattr(args, "threads", 1))
    composition_data_file = args.output_directory.joinpath("tnfs.pickle.gz")

    # one byte stream for the whole run
    offsets = np.zeros(sum(len(m) for m in mag_list) + 1, dtype=np.int64)
    np.cumsum(np.concatenate([m.lengths for m in mag_list]), out=offsets[1:])
    bases = np.concatenate([m.bases for m in mag_list]) if mag_list else np.zeros(0, np.uint8)
    logger.info("Computing tetranucleotide frequencies.")
    counts, rel, gc = _composition.tnf_from_concat(bases, offsets, relative=True, want_gc=True, want_counts=True,
                                                   device=getattr(args, "device", -1))
    row = np.cumsum([0] + [len(m) for m in mag_list])
    composition_dict = {m.genome: rel[row[i]:row[i + 1]] for i, m in enumerate(mag_list)}
    logger.info(f"Saving tetranucleotide frequencies data to '{composition_data_file}'.")
    with gzip.open(composition_data_file, "wb") as fout:
        pickle.dump(composition_dict, fout)

    logger.info("Computing contig scores.")
    tnf_scores, gc_scores = score_composition_batch(mag_list, counts, gc)
    mag_composition_list = [
        Composition(m, composition_dict, gc[row[i]:row[i + 1]],
                    scores=(tnf_scores[row[i]:row[i + 1]], gc_scores[row[i]:row[i + 1]]))
        for i, m in enumerate(mag_list)
    ]
    logger.info(f"Writing output to: '{composition_score_file}'.")
    tools.write_module_output(mag_composition_list, composition_score_file)
    return mag_composition_list
