"""Generates tests/golden/filter_golden.json.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_filter_golden.py

Inputs + outputs of the reference's OWN host-side glue of `magpurify2 filter`, lifted with `ast` and executed
under this container's NumPy (importing the reference package fails: xgboost, hdbscan, umap and Biopython are not
installed).  Nothing from the reference is copied into the repo; only the vectors are.
  * the feature transformations of magpurify2/modules/filter.py:117-135 (the assignments to feature_matrix[:, k]
    inside main(), fast and full mode);
  * ContigClassifier.get_dynamic_thresholds, magpurify2/core.py:585-596;
  * tools.get_checkm_contamination, magpurify2/tools.py:629-670.
The tree evaluation itself cannot be pinned this way (xgboost absent)."""
import ast
import csv
import json
import logging
import os
import tempfile
from pathlib import Path

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift_transforms():
    tree = ast.parse(open(os.path.join(REF, "magpurify2/modules/filter.py")).read())
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")

    def is_fm_assign(n):
        return (isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Subscript)
                and getattr(n.targets[0].value, "id", None) == "feature_matrix")

    always = [n for n in main.body if is_fm_assign(n)]
    full_only = []
    for n in main.body:
        if isinstance(n, ast.If) and "fast_mode" in ast.dump(n.test):
            full_only += [m for m in n.body if is_fm_assign(m)]
    assert len(always) == 2 and len(full_only) == 4, (len(always), len(full_only))

    def run(feature_matrix, fast_mode):
        ns = {"np": np, "feature_matrix": feature_matrix.copy()}
        body = always + ([] if fast_mode else full_only)
        exec(compile(ast.Module(body=body, type_ignores=[]), "filter.py[lifted]", "exec"), ns)
        return ns["feature_matrix"]
    return run


def lift_dynamic_thresholds():
    tree = ast.parse(open(os.path.join(REF, "magpurify2/core.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ContigClassifier")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_dynamic_thresholds")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "core.py[lifted]", "exec"), ns)
    return ns["get_dynamic_thresholds"]


def lift_checkm():
    tree = ast.parse(open(os.path.join(REF, "magpurify2/tools.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_checkm_contamination")
    ns = {"csv": csv, "logger": logging.getLogger("golden")}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "tools.py[lifted]", "exec"), ns)
    return ns["get_checkm_contamination"]


def main():
    rng = np.random.default_rng(20261020)
    transform, dyn, checkm = lift_transforms(), lift_dynamic_thresholds(), lift_checkm()
    out = {"numpy": np.__version__, "transforms": [], "thresholds": [], "checkm": []}
    for fast in (True, False):
        ncol = 6 if fast else 15
        x = rng.random((12, ncol))
        x[:, 2] = rng.integers(0, 3_000_000, 12)        # contig_length
        x[:, 5] = rng.integers(0, 30, 12)               # n_samples
        if not fast:
            x[:, 8] = rng.integers(0, 3000, 12)         # n_genes
            x[:, 9] = rng.integers(0, 2_500_000, 12)    # total_cds_length
            x[:, 12] = rng.integers(0, 7, 12)           # genome_rank
            x[:, 13] = rng.integers(0, 7, 12)           # contig_rank
        out["transforms"].append({"fast_mode": fast, "input": x.tolist(), "output": transform(x, fast).tolist()})

    class Self:
        pass
    for base in (0.5, 0.05, 0.95):
        s = Self()
        s.genomes = np.array(["g1", "g1", "g2", "g3", "g4", "g5"])
        s.base_probability_threshold = base
        scores = {"g1": 0.0, "g2": 1.5, "g3": 12.25, "g4": 100.0}
        out["thresholds"].append({"base": base, "genomes": s.genomes.tolist(), "checkm": scores,
                                  "output": [float(v) for v in dyn(s, scores)]})
    tables = {
        "ok": "Bin Id\tMarker lineage\tCompleteness\tContamination\tStrain heterogeneity\ng1\tk__Bacteria\t98.2\t1.25\t0\ng2\tk__Archaea\t70\t12.5\t50\n",
        "columns_reordered": "Contamination\tx\tBin Id\tCompleteness\n3.5\ty\tbinA\t88\n",
        "missing_column": "Bin Id\tContamination\ng1\t0\n",
    }
    with tempfile.TemporaryDirectory() as td:
        for name, text in tables.items():
            p = Path(td) / f"{name}.tsv"
            p.write_text(text)
            out["checkm"].append({"name": name, "text": text, "output": checkm(p)})
        out["checkm"].append({"name": "absent", "text": None, "output": checkm(Path(td) / "nope.tsv")})
    with open(os.path.join(HERE, "filter_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote filter_golden.json:", {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
