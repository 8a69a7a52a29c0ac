"""`magpurify2 composition` / `magpurify2 coverage` command line (the two sub-commands of
magpurify2/cli.py:244-412 that run on the GPU paths; same flags and defaults, cli.py:30-62) and
`magpurify2 filter` (cli.py:544-600), which consumes their score tables.

    python -m magpurify2_b200.cli composition genomes/*.fna out
    python -m magpurify2_b200.cli coverage genomes/*.fna out --bam_files a.bam b.bam
"""
import argparse
import logging
import multiprocessing
import sys
from pathlib import Path

from . import __version__, modules

default_values = {
    "composition": {"n_iterations": 4, "n_components": 3, "min_dist": 0.1, "n_neighbors": 15,
                    "set_op_mix_ratio": 1.0},
    "coverage": {"min_identity": 0.97, "trim_lower": 0.05, "trim_upper": 0.05, "min_average_coverage": 1.0,
                 "n_iterations": 4, "n_components": 3, "min_dist": 0.15, "n_neighbors": 15,
                 "set_op_mix_ratio": 0.6},
    "filter": {"probability_threshold": 0.5, "suffix": "filtered"},  # cli.py:61
}


def _embedding_options(options, module):
    d = default_values[module]
    options.add_argument("--n_iterations", default=d["n_iterations"], type=int,
                         help="Accepted for compatibility (UMAP/HDBSCAN scoring is not part of this build).")
    options.add_argument("--n_components", default=d["n_components"], type=int, help="Accepted for compatibility.")
    options.add_argument("--min_dist", default=d["min_dist"], type=float, help="Accepted for compatibility.")
    options.add_argument("--n_neighbors", default=d["n_neighbors"], type=int, help="Accepted for compatibility.")
    options.add_argument("--set_op_mix_ratio", default=d["set_op_mix_ratio"], type=float,
                         help="Accepted for compatibility.")


def _other(other):
    other.add_argument("-t", "--threads", default=multiprocessing.cpu_count(), type=int,
                       help="Host threads (BGZF inflate / BAM record filter). All by default.")
    other.add_argument("--device", default=-1, type=int, help="CUDA device (default: current).")
    other.add_argument("--devices", default=None, type=str,
                       help="Several CUDA devices, e.g. '0,1,2,3' or 'all': MAGs (composition) and BAM files "
                            "(coverage) are dealt out to them; overrides --device.")
    other.add_argument("-q", "--quiet", help="Suppress the logger output", action="store_true")


def build_parser():
    parser = argparse.ArgumentParser(prog="magpurify2", description=f"MAGpurify2 (B200 build) v{__version__}")
    sub = parser.add_subparsers(dest="command")
    p = sub.add_parser("composition", help="Identify putative contaminants using tetranucleotide frequences.")
    p.set_defaults(func=modules.composition.main)
    p.add_argument("genomes", nargs="+", type=Path, help="Input genomes in the FASTA format ('.bz2', '.gz', '.xz' ok).")
    p.add_argument("output_directory", type=Path, help="Directory for logs, supporting data and scores.")
    _embedding_options(p.add_argument_group("Data processing options"), "composition")
    _other(p.add_argument_group("Other options"))
    c = sub.add_parser("coverage", help="Identify putative contaminants using coverage profiles.")
    c.set_defaults(func=modules.coverage.main)
    c.add_argument("genomes", nargs="+", type=Path)
    c.add_argument("output_directory", type=Path)
    mex = c.add_argument_group("Coverage input").add_mutually_exclusive_group(required=True)
    mex.add_argument("--bam_files", nargs="+", type=Path, help="Input sorted BAM files.")
    mex.add_argument("--coverage_file", type=Path, help="Tabular file with contig coverage data.")
    o = c.add_argument_group("Data processing options")
    d = default_values["coverage"]
    o.add_argument("--min_identity", default=d["min_identity"], type=float)
    o.add_argument("--trim_lower", default=d["trim_lower"], type=float)
    o.add_argument("--trim_upper", default=d["trim_upper"], type=float)
    o.add_argument("--min_average_coverage", default=d["min_average_coverage"], type=float)
    _embedding_options(o, "coverage")
    _other(c.add_argument_group("Other options"))
    f = sub.add_parser("filter", help="Remove putative contaminants using the scores of the modules.")
    f.set_defaults(func=modules.filter.main)
    f.add_argument("genomes", nargs="+", type=Path)
    f.add_argument("output_directory", type=Path, help="The directory the modules wrote their scores to.")
    f.add_argument("filtered_output_directory", type=Path, help="Directory for the filtered MAGs.")
    o = f.add_argument_group("Data processing options")
    o.add_argument("--checkm_file", type=Path, help="CheckM tabular file: thresholds as a function of MAG quality.")
    o.add_argument("--probability_threshold", default=default_values["filter"]["probability_threshold"], type=float)
    o.add_argument("--fast_mode", action="store_true",
                   help="Use only the outputs of the composition and coverage modules.")
    o.add_argument("--suffix", default=default_values["filter"]["suffix"], type=str)
    o.add_argument("--model_file", type=Path,
                   help="XGBoost JSON model (default: models/ of an installed magpurify2 package).")
    _other(f.add_argument_group("Other options"))
    return parser


def cli(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if not getattr(args, "func", None):
        parser.print_help()
        return 0
    args.output_directory.mkdir(parents=True, exist_ok=True)
    log = logging.getLogger("timestamp")
    log.setLevel(logging.INFO)
    log.handlers.clear()
    fmt = logging.Formatter("[%(asctime)s] %(levelname)s: %(message)s", "%Y-%m-%d %H:%M:%S")
    fh = logging.FileHandler(args.output_directory / f"{args.command}.log")
    fh.setFormatter(fmt)
    log.addHandler(fh)
    if not args.quiet:
        sh = logging.StreamHandler(sys.stderr)
        sh.setFormatter(fmt)
        log.addHandler(sh)
    args.func(args)
    return 0


if __name__ == "__main__":
    sys.exit(cli())
