"""CPU oracle vs the committed golden vectors (no GPU).

The oracle (oracle/) is the checker for every GPU parity test, so it is pinned here first:
  * LUT   : SHA-256 of REV_COMPLEMENTARY_FOURMERS parsed from the reference
            (src/composition.rs:24-37; tests/golden/lut_sha256.txt);
  * TNF/GC: SURVEY.md Appendix B.1 known answers + cross-check with the brute-force string-window
            oracle (independent method);
  * scoring: outputs of the reference's OWN get_weighted_median / get_log_ratio_scores
            (magpurify2/tools.py:283-351; tests/golden/scoring_golden.json);
  * coverage: Appendix B.2 vectors (PARITY UNPINNED w.r.t. coverm 0.3.2, see oracle header) +
            cross-check with the per-base brute force.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import brute


@pytest.fixture(scope="module")
def kats(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kats.json")))


def test_lut_matches_reference_sha(oracle, golden_dir):
    want = open(os.path.join(golden_dir, "lut_sha256.txt")).read().strip()
    assert hashlib.sha256(oracle.canonical_lut().tobytes()).hexdigest() == want


def test_lut_palindromes(oracle, kats):
    lut = oracle.canonical_lut()
    pal = sorted({int(lut[c]) for c in range(256)
                  if c == int("".join(str(3 - int(d)) for d in reversed(np.base_repr(c, 4).zfill(4))), 4)})
    assert pal == kats["palindrome_classes"]
    assert lut.max() == 135 and len(set(lut.tolist())) == 136


def test_tnf_gc_kats(oracle, kats):
    for seq, total, classes, gc in kats["tnf_gc"]:
        counts = oracle.get_tnf([seq], relative=False)[0]
        assert counts.sum() == total, seq
        want = np.zeros(136, dtype=np.uint32)
        for k, v in classes.items():
            want[brute.CANONICAL.index(k)] = v
        assert (counts == want).all(), seq
        assert (counts == brute.canonical_counts(seq)).all(), seq
        rel = oracle.get_tnf([seq], relative=True)[0]
        if total:
            assert np.array_equal(rel, counts.astype(np.float32) / np.float32(total))
        else:
            assert (rel == 0).all()
        g = oracle.get_gc([seq])[0]
        if gc is None:
            assert np.isnan(g)
        else:
            assert g == np.float32(gc), seq


def test_tnf_random_vs_brute(oracle):
    rng = np.random.default_rng(1)
    alphabet = np.frombuffer(b"ACGTacgtNRYKMSWnU-*", dtype=np.uint8)
    p = np.array([20, 20, 20, 20, 3, 3, 3, 3, 2, .5, .5, .5, .5, .5, .5, 1, .5, .5, .5])
    seqs = []
    for n in [0, 1, 3, 4, 5, 17, 64, 255, 1000, 4097]:
        seqs.append(bytes(rng.choice(alphabet, n, p=p / p.sum())).decode())
    got = oracle.get_tnf(seqs, relative=False, threads=3)
    for s, row in zip(seqs, got):
        assert (row == brute.canonical_counts(s)).all()
    gc = oracle.get_gc(seqs)
    for s, g in zip(seqs, gc):
        b = brute.gc_fraction(s)
        assert (np.isnan(g) and np.isnan(b)) or g == b


def test_revcomp_invariance(oracle):
    rng = np.random.default_rng(2)
    s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 5000)).decode()
    rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    a, b = oracle.get_tnf([s, rc], relative=False)
    assert (a == b).all() and a.sum() == len(s) - 3


def test_non_ascii_counts_as_one_char(oracle):
    # chars().count() (src/composition.rs:131) counts scalar values, not bytes
    assert oracle.get_gc(["GCéAT"])[0] == np.float32(2) / np.float32(5)


def test_scoring_golden(oracle, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "scoring_golden.json")))
    assert len(g["cases"]) >= 40
    for case in g["cases"]:
        data = np.asarray(case["data"], dtype=np.float32).reshape(case["shape"])
        w = np.asarray(case["weights"], dtype=np.int64)
        scores, med = oracle.log_ratio_scores(data, w, case["base"])
        want_med = np.asarray(case["medians"], dtype=np.float32)
        assert np.array_equal(med, want_med), case["note"]
        want = np.asarray(case["scores"])
        assert np.allclose(scores, want, rtol=0, atol=1e-6, equal_nan=True), case["note"]


def test_trimmed_mean_kats(oracle, kats):
    for length, reads, lo, up, (N, mn, mx, total), want in kats["trimmed_mean"]:
        starts = [r[0] for r in reads]
        ends = [min(r[0] + r[1], length) for r in reads]
        st = oracle.cov_contig(starts, ends, length, 75, lo, up)
        assert (st.observed_len, st.total) == (N, total), (length, reads)
        if N and st.covered:
            assert (st.min_index, st.max_index) == (mn, mx)
        assert st.coverage == np.float32(want)
        cov, info = brute.trimmed_mean(list(zip(starts, ends)), length, 75, lo, up)
        assert cov == st.coverage and info["total"] == st.total


def test_f32_index_arithmetic(oracle, kats):
    for N, mn, mx in kats["f32_index"]:
        st = oracle.trimmed_mean_from_hist(np.array([N], dtype=np.uint32), N, 0, 0.05,
                                           np.float32(1.0) - np.float32(0.05))
        assert (st.min_index, st.max_index) == (mn, mx), N


def test_identity_filter(oracle, kats):
    for nm, aligned, passes in kats["identity"]:
        ev = oracle.record_events(0, 0, 10, [(aligned << 4) | 0], nm, 0.97)
        assert bool(ev) == passes
    # flags: secondary / supplementary / unmapped dropped, improper pair kept
    assert oracle.record_events(0x100, 0, 0, [(50 << 4)], 0, 0.97) == []
    assert oracle.record_events(0x800, 0, 0, [(50 << 4)], 0, 0.97) == []
    assert oracle.record_events(0x4, 0, 0, [(50 << 4)], 0, 0.97) == []
    assert oracle.record_events(0x1, 0, 5, [(50 << 4)], 0, 0.97) == [(5, 55)]
    # CIGAR walk: 10M 2D 5M 3I 4M 7N 2=  1X  5S
    cig = [(10 << 4) | 0, (2 << 4) | 2, (5 << 4) | 0, (3 << 4) | 1, (4 << 4) | 0, (7 << 4) | 3,
           (2 << 4) | 7, (1 << 4) | 8, (5 << 4) | 4]
    assert oracle.record_events(0, 0, 100, cig, 0, 0.97) == [
        (100, 110), (112, 117), (117, 121), (128, 130), (130, 131)]
    with pytest.raises(KeyError):
        oracle.record_events(0, 0, 0, [(50 << 4)], None, 0.97)


def test_cov_random_vs_brute(oracle):
    rng = np.random.default_rng(3)
    for length in [100, 150, 151, 152, 300, 1000, 5000]:
        for lo, up in [(0.0, 0.0), (0.05, 0.05), (0.2, 0.1)]:
            n = int(rng.integers(0, 200))
            st = rng.integers(0, length, n)
            en = np.minimum(st + rng.integers(1, 300, n), length)
            s = oracle.cov_contig(st, en, length, 75, lo, up)
            cov, info = brute.trimmed_mean(list(zip(st.tolist(), en.tolist())), length, 75, lo, up)
            assert s.total == info["total"] and s.observed_len == info["N"]
            assert s.coverage == cov or (np.isnan(s.coverage) and np.isnan(cov))


def test_no_trim_is_exact_mean(oracle):
    rng = np.random.default_rng(4)
    length = 2000
    st = rng.integers(0, length - 200, 100)
    en = st + 150
    s = oracle.cov_contig(st, en, length, 75, 0.0, 0.0)
    d = brute.depth_array(zip(st.tolist(), en.tolist()), length)[75:length - 75]
    assert s.total == int(d.sum()) and s.coverage == np.float32(d.sum()) / np.float32(len(d))


def test_select_samples_and_clr(oracle):
    cov = np.array([[10, 20, 0], [11, 22, .5], [50, 19, 0], [9.5, 0, .2], [10.5, 21, .1]], np.float32)
    w = np.array([10000, 20000, 5000, 3000, 40000])
    sel, means = oracle.select_samples(cov, w, 1.0)
    assert sel.tolist() == [True, True, False]
    assert np.allclose(means, np.average(cov, axis=0, weights=w))
    counts = np.random.default_rng(5).integers(0, 50, (7, 136)).astype(np.uint32)
    clr, dist = oracle.clr_centroid(counts, w[:1].repeat(7))
    lg = np.log(counts.astype(np.float64) + 1.0)
    ref = lg - lg.mean(axis=1, keepdims=True)
    assert np.allclose(clr, ref, atol=1e-6)
    assert np.allclose(dist, np.sqrt(((ref - ref.mean(axis=0)) ** 2).sum(axis=1)))
