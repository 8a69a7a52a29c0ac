"""`magpurify2 filter`: XGBoost-JSON tree walker, feature assembly from the module score
tables, thresholds, filtered genomes.  Host logic only -- no GPU, no xgboost.

The walker is checked against an independent recursive evaluation of the same JSON (hand-made models here;
the reference's own fast/full models when /root/reference is present -- xgboost itself is not installed, so
its outputs cannot be pinned: "parity unpinned" for the probabilities)."""
import json
import math
import os
from pathlib import Path

import numpy as np
import pytest

from magpurify2_b200 import cli, tools
from magpurify2_b200._xgb import XgbJsonModel
from magpurify2_b200.core import ContigClassifier
from magpurify2_b200.modules import filter as filter_module


def make_model(path, trees, num_feature, base_score="5E-1"):
    doc = {"learner": {"attributes": {}, "objective": {"name": "binary:logistic"},
                       "learner_model_param": {"base_score": base_score, "num_class": "0", "num_feature": str(num_feature)},
                       "gradient_booster": {"name": "gbtree", "model": {
                           "gbtree_model_param": {"num_trees": str(len(trees)), "size_leaf_vector": "0"},
                           "tree_info": [0] * len(trees), "trees": trees}}},
           "version": [1, 0, 0]}
    Path(path).write_text(json.dumps(doc))
    return path


def tree(left, right, feat, cond, default_left):
    return {"left_children": left, "right_children": right, "split_indices": feat, "split_conditions": cond,
            "default_left": default_left, "id": 0}


def walk(t, x):
    """Independent evaluation: plain recursion over the JSON arrays, float32 comparisons."""
    def rec(n):
        if t["left_children"][n] == -1:
            return np.float32(t["split_conditions"][n])
        v = np.float32(x[t["split_indices"][n]])
        if math.isnan(v):
            go_left = bool(t["default_left"][n])
        else:
            go_left = bool(v < np.float32(t["split_conditions"][n]))
        return rec(t["left_children"][n] if go_left else t["right_children"][n])
    return rec(0)


def reference_margin(path, X):
    """Margins by the independent walk (float32 sums in tree order, as xgboost accumulates them)."""
    doc = json.load(open(path))
    trees = doc["learner"]["gradient_booster"]["model"]["trees"]
    base = np.float32(float(doc["learner"]["learner_model_param"]["base_score"]))
    out = []
    for x in X:
        m = -np.log(np.float32(1.0) / base - np.float32(1.0))
        for t in trees:
            m = np.float32(m + walk(t, x))
        out.append(m)
    return np.array(out, dtype=np.float32)


TREES = [
    #        node: 0     1     2     3     4
    tree([1, -1, 3, -1, -1], [2, -1, 4, -1, -1], [0, 0, 1, 0, 0], [0.5, -0.4, 0.25, 0.3, 0.9], [True, False, False, False, False]),
    tree([1, -1, -1], [2, -1, -1], [2, 0, 0], [0.1, 0.2, -0.7], [False, False, False]),
]


def test_tree_walker_hand_model(tmp_path):
    m = XgbJsonModel(make_model(tmp_path / "m.json", TREES, 3, base_score="2.5E-1"))
    X = np.array([[0.4, 0.0, 0.0],      # t0: x0 < 0.5 -> leaf -0.4 ; t1: x2 < 0.1 -> 0.2
                  [0.5, 0.2, 0.5],      # t0: right, x1 < 0.25 -> 0.3 ; t1: right -> -0.7
                  [0.9, 0.25, 0.1],     # t0: right, not (x1 < 0.25) -> 0.9 ; t1: not (x2 < 0.1) -> -0.7
                  [np.nan, np.nan, np.nan]],  # t0: default left -> -0.4 ; t1: default right -> -0.7
                 dtype=np.float32)
    base = -math.log(1 / 0.25 - 1)
    want = [base - 0.4 + 0.2, base + 0.3 - 0.7, base + 0.9 - 0.7, base - 0.4 - 0.7]
    assert np.allclose(m.margin(X), want, rtol=0, atol=1e-6)
    assert np.allclose(m.predict(X), [1 / (1 + math.exp(-w)) for w in want], rtol=0, atol=1e-6)
    assert np.array_equal(m.margin(X), reference_margin(tmp_path / "m.json", X))
    with pytest.raises(ValueError):
        m.predict(np.zeros((2, 4)))


@pytest.mark.parametrize("name,nfeat", [("fast_model.json", 6), ("full_model.json", 15)])
def test_tree_walker_on_reference_models(name, nfeat):
    path = Path("/root/reference/magpurify2/models") / name
    if not path.exists():
        pytest.skip("reference models not present on this machine")
    m = XgbJsonModel(path)
    assert m.num_feature == nfeat and len(m.trees) > 0
    rng = np.random.default_rng(7)
    X = rng.random((400, nfeat)).astype(np.float32)
    X[rng.random(X.shape) < 0.05] = np.nan
    X[:20] = np.round(X[:20], 1)  # values that sit exactly on split conditions such as 0.15 / 0.85 / 0.95
    got = m.predict(X)
    assert got.dtype == np.float32 and np.all((got > 0) & (got < 1))
    want = reference_margin(path, X)
    assert np.array_equal(m.margin(X), want)  # same leaves, same float32 sums
    assert np.allclose(got, 1.0 / (1.0 + np.exp(-want.astype(np.float64))), rtol=0, atol=1e-6)


def write_scores(scores_dir, genomes_contigs, comp, cov):
    scores_dir.mkdir(parents=True, exist_ok=True)
    with open(scores_dir / "composition_scores.tsv", "w") as f:
        f.write("genome\tcontig\ttnf_score\tgc_content_score\tcontig_length\n")
        for (g, c), r in zip(genomes_contigs, comp):
            f.write(f"{g}\t{c}\t{r[0]}\t{r[1]}\t{r[2]}\n")
    with open(scores_dir / "coverage_scores.tsv", "w") as f:
        f.write("genome\tcontig\tcoverage_score\tcoverage_cluster_score\tn_samples\n")
        for (g, c), r in zip(genomes_contigs, cov):
            f.write(f"{g}\t{c}\t{r[0]}\t{r[1]}\t{r[2]}\n")


def test_feature_assembly_and_transforms(tmp_path):
    import logging

    gc = [("g1", "a"), ("g1", "b"), ("g2", "c")]
    comp = [(0.9, 0.8, 999), (0.1, 0.2, 99999), (1.0, 1.0, 9)]
    cov = [(0.7, 1.0, 3), (0.2, 1.0, 25), (1.0, 1.0, 0)]
    write_scores(tmp_path / "scores", gc, comp, cov)
    names, X = filter_module.load_features(tmp_path / "scores", ["composition", "coverage"], logging.getLogger("t"))
    assert names.tolist() == [list(x) for x in gc] and X.shape == (3, 6)
    assert np.allclose(X[:, 2], np.tanh(np.log10(np.array([1000.0, 100000.0, 10.0])) * 0.2))
    assert np.allclose(X[:, 5], [0.3, 1.0, 0.0]) and np.allclose(X[:, [0, 1, 3, 4]], [[.9, .8, .7, 1], [.1, .2, .2, 1], [1, 1, 1, 1]])
    # a table written for other inputs is refused (filter.py:96-113)
    write_scores(tmp_path / "other" / "scores", gc, comp, cov)
    (tmp_path / "other" / "scores" / "coverage_scores.tsv").write_text(
        "genome\tcontig\tcoverage_score\tcoverage_cluster_score\tn_samples\ng1\ta\t1\t1\t1\ng1\tX\t1\t1\t1\ng2\tc\t1\t1\t1\n")
    with pytest.raises(SystemExit):
        filter_module.load_features(tmp_path / "other" / "scores", ["composition", "coverage"], logging.getLogger("t"))


def test_classifier_thresholds(tmp_path):
    # one stump on feature 0: x0 < 0.5 -> +2 (p = 0.88), else -2 (p = 0.12)
    model = make_model(tmp_path / "stump.json", [tree([1, -1, -1], [2, -1, -1], [0, 0, 0], [0.5, 2.0, -2.0], [True, False, False])], 6)
    names = np.array([["g1", "a"], ["g1", "b"], ["g2", "c"], ["g3", "d"]])
    X = np.zeros((4, 6)); X[:, 0] = [0.1, 0.9, 0.1, 0.1]
    c = ContigClassifier(names, X, model, True, 0.5)
    assert np.allclose(c.probabilities, [0.880797, 0.119203, 0.880797, 0.880797], atol=1e-6)
    assert c.flagged_contaminants.tolist() == [True, False, True, True]
    assert c.mags_contaminants_dict["g1"] == {"a": True, "b": False} and list(c)[1][2] == np.float32(0.1192)
    # CheckM: contamination 0 -> 0.85, 50 -> clip(0.85 + 0.72 (e^-7.05 - 1)) = 0.13 -> everything above is flagged;
    # g3 is not in the table: base threshold, clipped to [0.1, 0.85] like the others (core.py:594-595)
    (tmp_path / "checkm.tsv").write_text("Bin Id\tMarker lineage\tCompleteness\tContamination\ng1\tx\t90\t0\ng2\tx\t80\t50\n")
    c = ContigClassifier(names, X, model, True, 0.9, tmp_path / "checkm.tsv")
    assert np.allclose(c.probability_threshold, [0.85, 0.85, 0.85 + 0.72 * (math.exp(-7.05) - 1), 0.85])
    assert c.flagged_contaminants.tolist() == [True, False, True, True]
    (tmp_path / "bad.tsv").write_text("Bin\tContamination\ng1\t0\n")
    assert tools.get_checkm_contamination(tmp_path / "bad.tsv") is None
    assert tools.get_checkm_contamination(tmp_path / "missing.tsv") is None
    assert ContigClassifier(names, X, model, True, 0.5, tmp_path / "bad.tsv").probability_threshold == 0.5


def test_filter_cli_end_to_end(tmp_path):
    genomes = tmp_path / "genomes"
    genomes.mkdir()
    seq_b = "ACGT" * 40  # 160 bases: wrapped at 70 columns
    (genomes / "g1.fna").write_text(f">a first\nAAAA\n>b second contig\n{seq_b}\n")
    (genomes / "g2.fna").write_text(">c\nCCCC\n")
    (genomes / "g9.fna").write_text(">z\nGG\n")  # never scored: skipped with a warning
    out = tmp_path / "out"
    write_scores(out / "scores", [("g1", "a"), ("g1", "b"), ("g2", "c")],
                 [(0.1, 1, 4), (0.9, 1, 160), (0.9, 1, 4)], [(1, 1, 1), (1, 1, 1), (1, 1, 1)])
    model = make_model(tmp_path / "stump.json", [tree([1, -1, -1], [2, -1, -1], [0, 0, 0], [0.5, 2.0, -2.0], [True, False, False])], 6)
    rc = cli.cli(["filter", str(genomes / "g1.fna"), str(genomes / "g2.fna"), str(genomes / "g9.fna"), str(out),
                  str(tmp_path / "filtered"), "--fast_mode", "--model_file", str(model), "-q", "--suffix", "clean"])
    assert rc == 0
    rows = (out / "scores" / "contamination_probabilities.tsv").read_text().splitlines()
    assert rows[0] == "genome\tcontig\tcontaminant_probability" and rows[1:] == ["g1\ta\t0.8808", "g1\tb\t0.1192", "g2\tc\t0.1192"]
    assert (tmp_path / "filtered" / "g1.clean.fna").read_text() == ">b second contig\n" + seq_b[:70] + "\n" + seq_b[70:140] + "\n" + seq_b[140:] + "\n"
    assert (tmp_path / "filtered" / "g2.clean.fna").read_text() == ">c\nCCCC\n"
    assert not (tmp_path / "filtered" / "g9.clean.fna").exists()
    # no model anywhere: a clear error instead of a traceback
    if filter_module.find_model_file(True) is None:
        with pytest.raises(SystemExit):
            cli.cli(["filter", str(genomes / "g2.fna"), str(out), str(tmp_path / "f2"), "--fast_mode", "-q"])


def test_glue_matches_reference_golden(tmp_path, golden_dir):
    """Feature transformations, dynamic thresholds and the CheckM reader against vectors produced by the
    reference's own code (tests/golden/make_filter_golden.py lifts it with `ast`)."""
    import logging

    gold = json.load(open(os.path.join(golden_dir, "filter_golden.json")))
    for case in gold["transforms"]:
        x = np.array(case["input"])
        modules = ["composition", "coverage"] if case["fast_mode"] else ["composition", "coverage", "codon_usage", "taxonomy"]
        ncols = {"composition": 3, "coverage": 3, "codon_usage": 5, "taxonomy": 4}
        header = {"composition": "genome\tcontig\ttnf_score\tgc_content_score\tcontig_length",
                  "coverage": "genome\tcontig\tcoverage_score\tcoverage_cluster_score\tn_samples",
                  "codon_usage": "genome\tcontig\ta\tb\tc\td\te", "taxonomy": "genome\tcontig\ta\tb\tc\td"}
        d = tmp_path / ("fast" if case["fast_mode"] else "full") / "scores"
        d.mkdir(parents=True)
        col = 0
        for m in modules:
            with open(d / f"{m}_scores.tsv", "w") as f:
                f.write(header[m] + "\n")
                for i, row in enumerate(x[:, col:col + ncols[m]]):
                    f.write(f"g\tc{i}\t" + "\t".join(repr(float(v)) for v in row) + "\n")
            col += ncols[m]
        _names, got = filter_module.load_features(d, modules, logging.getLogger("t"))
        assert np.array_equal(got, np.array(case["output"])), case["fast_mode"]
    stump = make_model(tmp_path / "stump.json", [tree([1, -1, -1], [2, -1, -1], [0, 0, 0], [0.5, 2.0, -2.0], [True, False, False])], 6)
    for case in gold["thresholds"]:
        names = np.array([[g, f"c{i}"] for i, g in enumerate(case["genomes"])])
        c = ContigClassifier(names, np.zeros((len(names), 6)), stump, True, case["base"])
        assert np.array_equal(c.get_dynamic_thresholds(case["checkm"]), np.array(case["output"])), case["base"]
    for case in gold["checkm"]:
        p = tmp_path / f"{case['name']}.tsv"
        if case["text"] is not None:
            p.write_text(case["text"])
        assert tools.get_checkm_contamination(p) == case["output"], case["name"]
