"""Generates tests/golden/*.json.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

* scoring_golden.json : inputs + outputs of the reference's OWN get_weighted_median /
  get_log_ratio_scores, lifted with `ast` from /root/reference/magpurify2/tools.py:283-351 and
  executed under this container's NumPy (importing magpurify2.tools itself fails: hdbscan, umap
  and Biopython are not installed).  Nothing from the reference is copied into the repo; only the
  vectors are.
* lut_sha256.txt      : SHA-256 of the 256 bytes of REV_COMPLEMENTARY_FOURMERS parsed from
  /root/reference/src/composition.rs:24-37.
* kats.json           : the hand-checkable known-answer vectors of SURVEY.md Appendix B
  (4-mer / GC by string windows; trimmed-mean cases under Appendix A; f32 index checks).
"""
import ast
import hashlib
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift_reference_scoring():
    src = open(os.path.join(REF, "magpurify2/tools.py")).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef)
              and n.name in ("get_weighted_median", "get_log_ratio_scores")]
    assert len(wanted) == 2
    ns = {"np": np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "tools.py[lifted]", "exec"), ns)
    return ns["get_weighted_median"], ns["get_log_ratio_scores"]


def main():
    wmed, lrs = lift_reference_scoring()
    rng = np.random.default_rng(20261019)
    cases = []

    def add(data, weights, base, note):
        data = np.asarray(data, dtype=np.float32)
        weights = np.asarray(weights, dtype=np.int64)
        if data.ndim == 1:
            med = [float(wmed(data, weights))]
        else:
            med = [float(wmed(data[:, j], weights)) for j in range(data.shape[1])]
        sc = np.asarray(lrs(data, weights, base), dtype=np.float64)
        cases.append(dict(note=note, base=base, shape=list(data.shape),
                          data=[float(x) for x in data.ravel()],
                          weights=[int(w) for w in weights],
                          medians=med, scores=[float(x) for x in sc]))

    # SURVEY.md Appendix B.3
    add([0.50, 0.52, 0.48, 0.70, 0.51], [10000, 20000, 5000, 3000, 40000], 2, "B3-gc")
    cov = np.array([[10, 20, 0], [11, 22, .5], [50, 19, 0], [9.5, 0, .2], [10.5, 21, .1]], np.float32)
    add(cov[:, :2], [10000, 20000, 5000, 3000, 40000], 25, "B3-cov-selected")
    add([1, 2, 3, 4], [1, 1, 1, 1], 2, "tie-mean-of-straddlers")
    add([1, 2, 3], [1, 10, 1], 2, "dominant-weight")
    add([3, 1, 2, 1], [5, 5, 5, 5], 2, "duplicate-values")
    add([0.4, 0.4, 0.4], [7, 7, 7], 2, "all-equal")
    add([0.0, 0.0, 0.5, 0.6], [100, 100, 100, 100], 25, "zero-median-half")
    add([2.0, 1.0, 3.0], [5, 5, 10], 2, "weight-exactly-half (not dominant: > is strict)")
    add([5.0, 1.0, 3.0, 2.0], [10, 2, 2, 2], 2, "first-of-equal-max-weights dominant")
    # random per-MAG shaped cases: GC-like (1-D, base 2) and coverage-like (2-D, base 25)
    for k in range(40):
        n = int(rng.integers(3, 400))
        lengths = np.maximum(1, rng.lognormal(9.0, 1.0, n)).astype(np.int64)
        if k % 7 == 0:
            lengths[int(rng.integers(0, n))] = lengths.sum() * 3  # dominant contig
        if k % 2 == 0:
            gc = np.clip(rng.normal(0.5, 0.05, n), 0, 1).astype(np.float32)
            if k % 4 == 0:
                gc = np.round(gc, 2)  # many duplicate values
            add(gc, lengths, 2, f"rand-gc-{k}")
        else:
            s = int(rng.integers(1, 6))
            c = rng.lognormal(3.0, 0.4, (n, s)).astype(np.float32)
            c[rng.random((n, s)) < 0.05] = 0.0
            add(c, lengths, 25, f"rand-cov-{k}")
    with open(os.path.join(HERE, "scoring_golden.json"), "w") as f:
        json.dump(dict(numpy=np.__version__, source="magpurify2/tools.py:283-351 (lifted via ast)",
                       cases=cases), f)

    src = open(os.path.join(REF, "src/composition.rs")).read()
    tab = re.search(r"REV_COMPLEMENTARY_FOURMERS: \[usize; 256\] = \[(.*?)\];", src, re.S).group(1)
    lut = np.array([int(x) for x in tab.replace("\n", " ").split(",") if x.strip()], dtype=np.uint8)
    assert lut.size == 256
    with open(os.path.join(HERE, "lut_sha256.txt"), "w") as f:
        f.write(hashlib.sha256(lut.tobytes()).hexdigest() + "\n")

    kats = dict(
        tnf_gc=[  # seq, sum of counts, {canonical 4-mer: count}, GC (f32; null = NaN)
            ["ACGT", 1, {"ACGT": 1}, 0.5],
            ["AAAAA", 2, {"AAAA": 2}, 0.0],
            ["TTTTT", 2, {"AAAA": 2}, 0.0],
            ["ACGTNACGT", 2, {"ACGT": 2}, 0.44444445],
            ["acgtACGT", 5, {"ACGT": 2, "CGTA": 2, "GTAC": 1}, 0.5],
            ["ACG", 0, {}, 0.6666667],
            ["", 0, {}, None],
            ["ACGTRYACGTT", 3, {"AACG": 1, "ACGT": 2}, 0.36363637],
            ["AAAANAAAA", 2, {"AAAA": 2}, 0.0],
            ["GATTACA", 4, {"AATC": 1, "ATTA": 1, "GTAA": 1, "TACA": 1}, 0.2857143],
            ["ACGUACGT", 1, {"ACGT": 1}, 0.5],
            ["NNNN", 0, {}, 0.0],
            ["AAAACCCCGGGGTTTT", 13,
             {"AAAA": 2, "AAAC": 2, "AACC": 2, "ACCC": 2, "CCCC": 2, "CCCG": 2, "CCGG": 1}, 0.5],
        ],
        palindrome_classes=[15, 27, 38, 48, 69, 78, 86, 93, 107, 113, 118, 122, 129, 132, 134, 135],
        trimmed_mean=[  # length, reads (start,len), trim_lower, trim_upper, (N,min,max,total), f32
            [1000, [[0, 1000], [0, 1000]], 0.05, 0.05, [850, 42, 808, 1534], 2.002611],
            [400, [[100, 100]], 0.05, 0.05, [250, 12, 238, 89], 0.3938053],
            [150, [[0, 150]], 0.05, 0.05, [0, 0, 0, 0], 0.0],
            [151, [[0, 151]], 0.05, 0.05, [1, 0, 1, 1], 1.0],
            [1000, [[0, 75], [925, 75]], 0.05, 0.05, [850, 42, 808, 0], 0.0],
            [1000, [[80 * i, 100] for i in range(12)], 0.0, 0.0, [850, 0, 850, 1070], 1.2588235],
            [1000, [[80 * i, 100] for i in range(12)], 0.05, 0.05, [850, 42, 808, 947], 1.2362925],
            [300, [[250, 50], [100, 100], [120, 100]], 0.05, 0.05, [150, 7, 143, 188], 1.382353],
        ],
        f32_index=[[20, 1, 19], [850, 42, 808], [9850, 492, 9358], [10000, 500, 9500],
                   [2999850, 149992, 2849858], [16777217, 838860, 15938355]],
        identity=[[3, 100, True], [4, 150, True], [5, 150, False], [9, 300, True]],
    )
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", len(cases), "scoring cases")


if __name__ == "__main__":
    main()
