"""Pieces the module drivers share: option clamping, the scores/ directory, the result caches and the
provenance note that tells `magpurify2 filter` which build wrote a score table.

The reference repeats these steps in each driver (modules/composition.py:34-71, modules/coverage.py:36-109);
the behaviour that is a contract -- which options are clamped to what, the names and the Python objects of
the two cache files, when a cache is reused and when it is deleted -- is kept, the code is this build's own."""
import gzip
import json
import logging
import pickle

from .. import tools

logger = logging.getLogger("timestamp")

# option -> allowed interval (cli.py:30-62 defaults; clamps of modules/composition.py:34-39, coverage.py:39-49)
EMBEDDING_LIMITS = {"n_iterations": (1, 999), "n_components": (1, 999), "n_neighbors": (1, 999),
                    "set_op_mix_ratio": (0.0, 1.0)}
COVERAGE_LIMITS = {"trim_lower": (0.0, 1.0), "trim_upper": (0.0, 1.0), "min_average_coverage": (0, 999)}

PROVENANCE_FILE = "provenance.json"
PROVENANCE = {
    "composition": {"producer": "magpurify2_b200", "tnf_score": "clr-centroid-distance (not the reference's UMAP+HDBSCAN "
                    "soft membership)", "gc_content_score": "reference formula"},
    "coverage": {"producer": "magpurify2_b200", "coverage_score": "reference formula",
                 "coverage_cluster_score": "constant 1.0 (UMAP+HDBSCAN score not computed)"},
}


def clamp_options(args, limits):
    """Brings every option named in `limits` into its interval (with the reference's warning)."""
    for name, (lo, hi) in limits.items():
        if hasattr(args, name):
            setattr(args, name, tools.validade_input(getattr(args, name), name, [lo, hi]))


def scores_path(args, module):
    """<out>/scores/<module>_scores.tsv, creating the directories like the reference does."""
    tools.check_output_directory(args.output_directory)
    directory = args.output_directory.joinpath("scores")
    directory.mkdir(exist_ok=True)
    return directory.joinpath(f"{module}_scores.tsv")


def note_provenance(args, module):
    """Records in <out>/scores/provenance.json how this module's columns were computed; `filter` reads it."""
    path = args.output_directory.joinpath("scores", PROVENANCE_FILE)
    try:
        known = json.loads(path.read_text()) if path.exists() else {}
    except ValueError:
        known = {}
    known[module] = PROVENANCE[module]
    path.write_text(json.dumps(known, indent=1, sort_keys=True) + "\n")


class PickleCache:
    """A gzip-compressed pickle next to the scores (tnfs.pickle.gz, coverages.pickle.gz).

    `load(covers)` returns the stored object when `covers(obj)` says it answers the current request; an object
    that does not is deleted, as the reference does, so that the next `store` starts clean."""

    def __init__(self, path, what):
        self.path, self.what = path, what

    def load(self, covers):
        if not self.path.exists():
            return None
        try:
            with gzip.open(self.path) as fin:
                obj = pickle.load(fin)
            if covers(obj):
                logger.info(f"Skipping {self.what} computation.")
                return obj
        except (OSError, EOFError, pickle.UnpicklingError, ValueError, TypeError, KeyError):
            logger.warning(f"'{self.path}' could not be read and will be replaced.")
        self.path.unlink()
        return None

    def store(self, obj):
        logger.info(f"Saving {self.what} data to '{self.path}'.")
        with gzip.open(self.path, "wb") as fout:
            pickle.dump(obj, fout)
