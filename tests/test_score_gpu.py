"""GPU parity: weighted median / log-ratio scores / sample selection / CLR-centroid through the
C ABI vs (a) the golden vectors produced by the reference's own NumPy functions and (b) the oracle.

Medians are compared bit-exactly (they are input values or an exact f32 mean of two of them);
scores within 1e-6 absolute (north_star tolerance; they are written at 5 decimals)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "scoring_golden.json")))["cases"]


def test_reference_golden_vectors_per_function(golden):
    from magpurify2_b200 import tools

    for case in golden:
        data = np.asarray(case["data"], dtype=np.float32).reshape(case["shape"])
        w = np.asarray(case["weights"], dtype=np.int64)
        sc = tools.get_log_ratio_scores(data, w, case["base"])
        assert np.allclose(sc, case["scores"], rtol=0, atol=TOL, equal_nan=True), case["note"]
        cols = data.reshape(len(data), -1)
        for j, want in enumerate(case["medians"]):
            assert tools.get_weighted_median(cols[:, j], w) == np.float32(want), case["note"]


def test_reference_golden_vectors_batched(golden):
    """All golden cases with the same column count stacked as MAGs of one batch call."""
    from magpurify2_b200 import _scoring

    for base in (2, 25):
        for ncol in range(1, 6):
            cs = [c for c in golden if c["base"] == base and (c["shape"][1] if len(c["shape"]) > 1 else 1) == ncol]
            if not cs:
                continue
            data = np.concatenate([np.asarray(c["data"], np.float32).reshape(-1, ncol) for c in cs])
            w = np.concatenate([np.asarray(c["weights"], np.int64) for c in cs])
            mag_off = np.cumsum([0] + [len(c["weights"]) for c in cs])
            sc, med = _scoring.log_ratio_scores_batch(data, w, mag_off, base, min_rows=1, want_medians=True)
            assert np.allclose(sc, np.concatenate([c["scores"] for c in cs]), rtol=0, atol=TOL, equal_nan=True)
            assert np.array_equal(med, np.array([c["medians"] for c in cs], dtype=np.float32))


def test_small_mag_rule_and_masks(oracle):
    from magpurify2_b200 import _scoring

    rng = np.random.default_rng(31)
    sizes = [1, 2, 3, 50, 0, 2, 700]
    mag_off = np.cumsum([0] + sizes)
    n = int(mag_off[-1])
    data = rng.lognormal(3, 0.5, (n, 4)).astype(np.float32)
    w = rng.integers(1000, 100000, n)
    mask = rng.random((len(sizes), 4)) < 0.6
    mask[3] = False  # a MAG with no selected sample -> ones (core.py:306-308)
    sc = _scoring.log_ratio_scores_batch(data, w, mag_off, 25, col_mask=mask)
    for m, sz in enumerate(sizes):
        rows = slice(mag_off[m], mag_off[m + 1])
        if sz <= 2 or not mask[m].any():
            assert (sc[rows] == 1.0).all()
        else:
            want, _ = oracle.log_ratio_scores(data[rows][:, mask[m]], w[rows], 25)
            assert np.allclose(sc[rows], want, rtol=0, atol=TOL)


def test_big_mag_uses_scratch_sort(oracle):
    """> 4096 rows: the (value, weight) sort leaves shared memory."""
    from magpurify2_b200 import _scoring

    rng = np.random.default_rng(32)
    n = 9000
    data = np.round(rng.normal(0.5, 0.05, n), 3).astype(np.float32)  # many ties
    w = rng.integers(1, 50, n)
    sc, med = _scoring.log_ratio_scores_batch(data, w, [0, n], 2, want_medians=True)
    want, wmed = oracle.log_ratio_scores(data, w, 2)
    assert med[0, 0] == wmed[0] and np.allclose(sc, want, rtol=0, atol=TOL)


def test_select_samples(oracle):
    from magpurify2_b200 import _scoring

    rng = np.random.default_rng(33)
    sizes = [5, 300, 1, 40]
    mag_off = np.cumsum([0] + sizes)
    n = int(mag_off[-1])
    cov = rng.lognormal(0.0, 1.5, (n, 6)).astype(np.float32)
    w = rng.integers(1500, 200000, n)
    sel, means = _scoring.select_samples_batch(cov, w, mag_off, 1.0)
    for m in range(len(sizes)):
        rows = slice(mag_off[m], mag_off[m + 1])
        wsel, wmeans = oracle.select_samples(cov[rows], w[rows], 1.0)
        assert np.array_equal(sel[m], wsel) and np.array_equal(means[m], wmeans)  # same summation order


def test_clr_centroid(oracle):
    from magpurify2_b200 import _scoring

    rng = np.random.default_rng(34)
    sizes = [3, 120, 1]
    mag_off = np.cumsum([0] + sizes)
    n = int(mag_off[-1])
    counts = rng.poisson(40, (n, 136)).astype(np.uint32)
    counts[5] = 0
    w = rng.integers(1500, 200000, n)
    dist, clr = _scoring.clr_centroid_batch(counts, w, mag_off, want_clr=True)
    for m in range(len(sizes)):
        rows = slice(mag_off[m], mag_off[m + 1])
        wclr, wdist = oracle.clr_centroid(counts[rows], w[rows])
        assert np.allclose(clr[rows], wclr, rtol=1e-6, atol=1e-6)
        assert np.allclose(dist[rows], wdist, rtol=1e-9, atol=1e-12)
