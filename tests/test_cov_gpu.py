"""GPU parity: coverage path (difference array -> depth -> windowed histogram -> trimmed mean)
through the C ABI vs the CPU oracle (PARITY UNPINNED w.r.t. coverm itself).

Integers (observed length, covered, min/max index, total = the depth sum, max depth) are compared
bit-exactly; the coverage is one IEEE f32 division of those integers, compared bit-exactly too."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INT_FIELDS = ["observed_len", "covered", "total", "max_depth"]


@pytest.fixture(scope="module")
def covmod():
    from magpurify2_b200 import _coverage

    return _coverage


def check_sample(covmod, oracle, starts, ends, ev_off, lengths, excl=75, lo=0.0, up=0.0):
    lengths = np.asarray(lengths, dtype=np.int64)
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    cov, st = covmod.cov_from_events(starts, ends, ev_off, off, excl, lo, up, want_stats=True)
    wcov, wst = oracle.cov_sample(starts, ends, ev_off, lengths, excl, lo, up, threads=8)
    for f in INT_FIELDS:
        bad = np.nonzero(st[f] != wst[f])[0]
        assert bad.size == 0, (f, bad[:5], st[f][bad[:5]], wst[f][bad[:5]], lengths[bad[:5]])
    live = wst["covered"] > 0
    assert np.array_equal(st["min_index"][live], wst["min_index"][live])
    assert np.array_equal(st["max_index"][live], wst["max_index"][live])
    assert np.array_equal(cov.view(np.uint32), wcov.view(np.uint32)) or np.array_equal(cov, wcov, equal_nan=True)
    assert np.array_equal(st["coverage"], cov, equal_nan=True)
    return cov, st


def test_kats(covmod, oracle, golden_dir):
    kats = json.load(open(os.path.join(golden_dir, "kats.json")))
    for length, reads, lo, up, (N, mn, mx, total), want in kats["trimmed_mean"]:
        starts = np.array([r[0] for r in reads], dtype=np.uint32)
        ends = np.array([min(r[0] + r[1], length) for r in reads], dtype=np.uint32)
        cov, st = check_sample(covmod, oracle, starts, ends, [0, len(reads)], [length], 75, lo, up)
        assert cov[0] == np.float32(want)
        assert (int(st["observed_len"][0]), int(st["total"][0])) == (N, total)


def test_all_kats_in_one_sample(covmod, oracle, golden_dir):
    """The same vectors as contigs of ONE concatenated sample (tests the -1 landing on the next contig)."""
    kats = json.load(open(os.path.join(golden_dir, "kats.json")))
    cases = [k for k in kats["trimmed_mean"] if (k[2], k[3]) == (0.05, 0.05)]
    starts, ends, ev_off, lengths, want = [], [], [0], [], []
    for length, reads, lo, up, _ints, w in cases:
        starts += [r[0] for r in reads]
        ends += [min(r[0] + r[1], length) for r in reads]
        ev_off.append(len(starts))
        lengths.append(length)
        want.append(w)
    cov, _ = check_sample(covmod, oracle, starts, ends, ev_off, lengths, 75, 0.05, 0.05)
    assert np.array_equal(cov, np.array(want, dtype=np.float32))


def test_empty_and_ragged(covmod, oracle):
    # no events at all; zero-length / <=150 bp / 151 bp contigs; events only on some contigs
    lengths = [0, 1, 150, 151, 152, 0, 3000, 149, 400, 0]
    check_sample(covmod, oracle, [], [], [0] * 11, lengths, 75, 0.05, 0.05)
    starts = [0, 0, 100, 2850, 10, 380]
    ends = [151, 152, 250, 3000, 149, 400]
    ev_off = [0, 0, 0, 0, 1, 2, 2, 4, 5, 6, 6]
    for lo, up in [(0.0, 0.0), (0.05, 0.05), (0.3, 0.2)]:
        check_sample(covmod, oracle, starts, ends, ev_off, lengths, 75, lo, up)
    check_sample(covmod, oracle, starts, ends, ev_off, lengths, 0, 0.0, 0.0)   # no end exclusion
    check_sample(covmod, oracle, starts, ends, ev_off, lengths, 10, 0.1, 0.1)


def test_synthetic_sample_c2_shape(covmod, oracle):
    """10 MAGs (~3 Mbp, ~300 contigs each) x one sample of ~4 M reads (BASELINE config 2 shape)."""
    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(10, seed=1)
    for sample, variable in [(0, False), (1, True)]:
        st, en, ev_off = synth.make_events(lengths, mag_off, sample, seed=1, variable=variable)
        check_sample(covmod, oracle, st, en, ev_off, lengths, 75, 0.05, 0.05)
    st, en, ev_off = synth.make_events(lengths, mag_off, 2, seed=1)
    check_sample(covmod, oracle, st, en, ev_off, lengths, 75, 0.0, 0.0)


def test_long_contigs_span_tiles(covmod, oracle):
    """Windows spanning many 65536-slot tiles, contig boundaries exactly on / next to tile edges."""
    rng = np.random.default_rng(21)
    lengths = [65536, 65536 - 75, 75, 65536 + 76, 1, 300000, 65535, 65537, 1500000, 131072, 200, 65536 * 3]
    starts, ends, ev_off = [], [], [0]
    for ln in lengths:
        n = int(ln * 8 / 150) if ln > 150 else 0
        s = np.sort(rng.integers(0, max(1, ln - 150), n)) if n else np.zeros(0, np.int64)
        starts.append(s)
        ends.append(np.minimum(s + rng.integers(50, 250, n), ln))
        ev_off.append(ev_off[-1] + n)
    starts = np.concatenate(starts).astype(np.uint32)
    ends = np.concatenate(ends).astype(np.uint32)
    for lo, up in [(0.05, 0.05), (0.0, 0.0)]:
        check_sample(covmod, oracle, starts, ends, ev_off, lengths, 75, lo, up)


def test_deep_coverage_fallback(covmod, oracle):
    """Depth >= 1023 (the shared-memory histogram's range) and > 8192 (several fallback ranges)."""
    rng = np.random.default_rng(22)
    lengths = [5000, 2000, 100000, 800, 40000]
    depth = [1500, 30, 20000, 1023, 1022]
    starts, ends, ev_off = [], [], [0]
    for ln, dp in zip(lengths, depth):
        n = int(ln * dp / 150)
        s = np.sort(rng.integers(0, ln - 150, n))
        starts.append(s)
        ends.append(s + 150)
        ev_off.append(ev_off[-1] + n)
    starts = np.concatenate(starts).astype(np.uint32)
    ends = np.concatenate(ends).astype(np.uint32)
    _cov, st = check_sample(covmod, oracle, starts, ends, ev_off, lengths, 75, 0.05, 0.05)
    assert st["max_depth"][2] > 8192 * 2 and st["max_depth"][0] >= 1023
    # a pile-up: every read on the same position
    n = 3000
    check_sample(covmod, oracle, np.full(n, 500, np.uint32), np.full(n, 650, np.uint32), [0, n, n], [2000, 900],
                 75, 0.05, 0.05)


def test_no_trim_equals_mean_depth(covmod):
    rng = np.random.default_rng(23)
    ln = 50000
    s = np.sort(rng.integers(0, ln - 150, 3000)).astype(np.uint32)
    cov, st = covmod.cov_from_events(s, s + 150, [0, len(s)], [0, ln], 75, 0.0, 0.0, want_stats=True)
    d = np.zeros(ln + 1, dtype=np.int64)
    np.add.at(d, s, 1)
    np.add.at(d, s + 150, -1)
    depth = np.cumsum(d)[75:ln - 75]
    assert int(st["total"][0]) == int(depth.sum())
    assert cov[0] == np.float32(depth.sum()) / np.float32(len(depth))


def test_device_entry_strided_columns(covmod, oracle):
    """mp2_cov_events_device writing column s of a float32[n_contigs, n_samples] matrix in place."""
    import torch

    from magpurify2_b200 import _lib, synth

    lengths, mag_off, _gc = synth.mag_layout(3, seed=5, mean_bp=5e5, n_contigs=80)
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    n, S = len(lengths), 3
    d_off = torch.from_numpy(off).cuda()
    mat = torch.full((n, S), -1.0, dtype=torch.float32, device="cuda")
    lib = _lib.load()
    want = np.zeros((n, S), dtype=np.float32)
    for s in range(S):
        st, en, ev_off = synth.make_events(lengths, mag_off, s, seed=5)
        want[:, s], _ = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=4)
        d_st = torch.from_numpy(st.view(np.int32)).cuda()
        d_en = torch.from_numpy(en.view(np.int32)).cuda()
        d_eo = torch.from_numpy(ev_off).cuda()
        _lib.check(lib.mp2_cov_events_device(d_st.data_ptr(), d_en.data_ptr(), d_eo.data_ptr(), d_off.data_ptr(),
                                             n, len(st), int(off[-1]), 150, 0, 75, 0.05, 0.05,
                                             mat.data_ptr() + 4 * s, S, None,
                                             torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(mat.cpu().numpy(), want)


def test_rejects_bad_trim(covmod):
    with pytest.raises(ValueError):
        covmod.cov_from_events([], [], [0, 0], [0, 1000], 75, 0.6, 0.5)


def test_unsorted_events_take_the_general_path(covmod, oracle):
    """Blocks in arbitrary order inside a contig (allowed by the ABI): the device check notices and
    the difference array goes through HBM instead of shared memory; same integers either way."""
    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(3, seed=6, mean_bp=6e5, n_contigs=70)
    st, en, ev_off = synth.make_events(lengths, mag_off, 0, seed=6, variable=True)
    rng = np.random.default_rng(6)
    st2, en2 = st.copy(), en.copy()
    for c in range(len(lengths)):
        p = rng.permutation(ev_off[c + 1] - ev_off[c]) + ev_off[c]
        st2[ev_off[c]:ev_off[c + 1]] = st[p]
        en2[ev_off[c]:ev_off[c + 1]] = en[p]
    a, sa = check_sample(covmod, oracle, st, en, ev_off, lengths, 75, 0.05, 0.05)
    b, sb = check_sample(covmod, oracle, st2, en2, ev_off, lengths, 75, 0.05, 0.05)
    assert np.array_equal(a, b) and np.array_equal(sa["total"], sb["total"])


def test_vouched_max_span_matches_checked(covmod):
    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(4, seed=8, mean_bp=8e5, n_contigs=50)
    off = np.concatenate([[0], np.cumsum(lengths)])
    st, en, ev_off = synth.make_events(lengths, mag_off, 1, seed=8, variable=True)
    span = int((en - st).max())
    a = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, max_span=0)
    b = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, max_span=span)
    c = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, max_span=span + 5000)
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_wrong_vouched_span_is_caught(covmod, oracle):
    """A caller that vouches for max_span=100 while blocks are 150 long must still get exact results
    (the kernel notices and the general path redoes the sample)."""
    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(2, seed=9, mean_bp=4e5, n_contigs=30)
    off = np.concatenate([[0], np.cumsum(lengths)])
    st, en, ev_off = synth.make_events(lengths, mag_off, 0, seed=9)
    want, _ = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=4)
    got = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, max_span=100)
    assert np.array_equal(got, want)
    # long blocks (> 512, the streaming kernel's overflow region): general path, same integers
    rng = np.random.default_rng(9)
    ln = 300000
    s = np.sort(rng.integers(0, ln - 5000, 2000)).astype(np.uint32)
    e = (s + rng.integers(600, 5000, 2000)).astype(np.uint32)
    want, _ = oracle.cov_sample(s, e, [0, 2000], [ln], 75, 0.05, 0.05)
    assert np.array_equal(covmod.cov_from_events(s, e, [0, 2000], [0, ln], 75, 0.05, 0.05), want)
    assert np.array_equal(covmod.cov_from_events(s, e, [0, 2000], [0, ln], 75, 0.05, 0.05, max_span=5000), want)


GENERAL = 1  # mp2.h MP2_COV_GENERAL_PATH


def _disorder(st, ev_off, lengths):
    off = np.concatenate([[0], np.cumsum(lengths)])
    cid = np.repeat(np.arange(len(lengths)), np.diff(ev_off))
    g = off[cid] + st.astype(np.int64)
    return int((np.maximum.accumulate(g) - g).max()) if len(g) else 0


@pytest.mark.parametrize("mode", ["measured", "vouched"])
def test_indel_variant_takes_the_fast_path(covmod, oracle, mode):
    """SURVEY 8(d) variant: read length N(150,15), 5 % of the reads with one I / D (two blocks, the second one
    out of order).  Bit-exact, and computed by the shared-memory path (flags == 0 everywhere)."""
    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(5, seed=12, mean_bp=9e5, n_contigs=60)
    off = np.concatenate([[0], np.cumsum(lengths)])
    st, en, ev_off = synth.make_events(lengths, mag_off, 2, seed=12, variable=True, indel_frac=0.05)
    dis = _disorder(st, ev_off, lengths)
    assert dis > 0 and (np.diff(st.astype(np.int64)) < 0).sum() > len(lengths)  # really unsorted
    kw = {} if mode == "measured" else dict(max_span=int((en - st).max()), max_disorder=dis)
    cov, stats = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, want_stats=True, **kw)
    wcov, wst = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    for f in INT_FIELDS:
        assert np.array_equal(stats[f], wst[f]), f
    assert np.array_equal(cov.view(np.uint32), wcov.view(np.uint32))
    assert not (stats["flags"] & GENERAL).any()


def test_large_disorder_goes_to_the_general_path(covmod, oracle):
    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(2, seed=13, mean_bp=4e5, n_contigs=30)
    off = np.concatenate([[0], np.cumsum(lengths)])
    st, en, ev_off = synth.make_events(lengths, mag_off, 0, seed=13)
    rng = np.random.default_rng(13)
    for c in range(len(lengths)):  # shuffle inside every contig: disorder = contig length
        p = rng.permutation(ev_off[c + 1] - ev_off[c]) + ev_off[c]
        st[ev_off[c]:ev_off[c + 1]], en[ev_off[c]:ev_off[c + 1]] = st[p], en[p]
    cov, stats = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, want_stats=True)
    wcov, wst = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    assert np.array_equal(stats["total"], wst["total"]) and np.array_equal(cov.view(np.uint32), wcov.view(np.uint32))
    has_window = lengths > 150
    assert (stats["flags"][has_window] & GENERAL).all()
    # a caller vouching for more disorder than the kernel takes: general path as well, same numbers
    cov2 = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, max_span=150, max_disorder=100000)
    assert np.array_equal(cov2.view(np.uint32), wcov.view(np.uint32))


@pytest.mark.parametrize("shift", [1, 2, 3])
def test_event_arrays_at_any_4_byte_alignment(covmod, oracle, shift):
    """The event ring is filled by 16-byte bulk copies; the arrays may start at any 4-byte boundary."""
    import torch

    from magpurify2_b200 import _lib, synth

    lengths, mag_off, _gc = synth.mag_layout(3, seed=14, mean_bp=6e5, n_contigs=50)
    off = np.concatenate([[0], np.cumsum(lengths)])
    st, en, ev_off = synth.make_events(lengths, mag_off, 1, seed=14, variable=True, indel_frac=0.03)
    want, _ = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    n, R = len(lengths), len(st)
    pad = np.zeros(shift, np.int32)
    d_st = torch.from_numpy(np.concatenate([pad, st.view(np.int32)])).cuda()[shift:]
    d_en = torch.from_numpy(np.concatenate([pad, pad, en.view(np.int32)])).cuda()[2 * shift:]
    d_eo, d_off = torch.from_numpy(ev_off).cuda(), torch.from_numpy(off).cuda()
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    lib = _lib.load()
    _lib.check(lib.mp2_cov_events_device(d_st.data_ptr(), d_en.data_ptr(), d_eo.data_ptr(), d_off.data_ptr(), n, R,
                                         int(off[-1]), 0, 0, 75, 0.05, 0.05, out.data_ptr(), 1, None,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_sparse_and_empty_contigs(covmod, oracle):
    """Runs of contigs without any event, events only at the very end, single events: the contig cursor of the
    event stream has to gallop over empty contigs."""
    rng = np.random.default_rng(15)
    lengths = rng.integers(1, 4000, 3000).astype(np.int64)
    lengths[::7] = rng.integers(60000, 200000, len(lengths[::7]))
    n_reads = np.zeros(len(lengths), dtype=np.int64)
    pick = rng.random(len(lengths)) < 0.03
    n_reads[pick] = rng.integers(1, 400, int(pick.sum()))
    n_reads[lengths < 160] = 0
    ev_off = np.concatenate([[0], np.cumsum(n_reads)])
    cid = np.repeat(np.arange(len(lengths)), n_reads)
    st = np.floor(rng.random(len(cid)) * (lengths[cid] - 150)).astype(np.int64)
    order = np.lexsort((st, cid))
    st = st[order].astype(np.uint32)
    en = (st + 150).astype(np.uint32)
    check_sample(covmod, oracle, st, en, ev_off, lengths, 75, 0.05, 0.05)


def _random_sample(rng, n_contigs, len_lo, len_hi, depth, read_len, jitter, indel_frac, gap_hi):
    """Reads sorted by position; `indel_frac` of them split into two blocks with a reference gap of 0..gap_hi."""
    lengths = rng.integers(len_lo, len_hi, n_contigs).astype(np.int64)
    lengths[rng.random(n_contigs) < 0.1] = rng.integers(0, 160, int((rng.random(n_contigs) < 0.1).sum()) or 1)[0]
    starts, ends, ev_off = [], [], [0]
    for ln in lengths.tolist():
        n = int(rng.poisson(depth * max(ln - read_len, 0) / read_len)) if ln > read_len + 5 else 0
        pos = np.sort(rng.integers(0, max(1, ln - read_len), n))
        rl = np.clip(read_len + rng.integers(-jitter, jitter + 1, n), 20, None)
        two = rng.random(n) < indel_frac
        a = (rl * rng.uniform(0.2, 0.8, n)).astype(np.int64)
        gap = rng.integers(0, gap_hi + 1, n)
        for p, r, t, aa, g in zip(pos.tolist(), rl.tolist(), two.tolist(), a.tolist(), gap.tolist()):
            if t:
                starts += [p, p + aa + g]
                ends += [p + aa, min(p + g + r, ln)]
            else:
                starts.append(p)
                ends.append(min(p + r, ln))
        ev_off.append(len(starts))
    st, en = np.array(starts, dtype=np.uint32), np.array(ends, dtype=np.uint32)
    keep = en > st  # the generator may produce an empty second block at a contig end: keep it, every path drops it
    return st, en, np.array(ev_off, dtype=np.int64), lengths, keep


@pytest.mark.parametrize("case", range(12))
def test_fuzz_shapes_depths_and_disorder(covmod, oracle, case):
    """Random shapes around every limit of the streaming kernel: contigs spanning several 65 536-slot runs, tiny and
    empty contigs, depths beyond the 511-bin shared histogram (deep pool) with out-of-order blocks, block span +
    disorder just inside and just outside the 512-slot overflow region (the latter must come back from the general
    path), vouched and measured."""
    rng = np.random.default_rng(1000 + case)
    shapes = [
        dict(n_contigs=40, len_lo=200, len_hi=30000, depth=20, read_len=150, jitter=20, indel_frac=0.1, gap_hi=5),
        dict(n_contigs=6, len_lo=100000, len_hi=400000, depth=8, read_len=150, jitter=30, indel_frac=0.2, gap_hi=60),
        dict(n_contigs=30, len_lo=1000, len_hi=9000, depth=700, read_len=150, jitter=10, indel_frac=0.05, gap_hi=3),   # deep pool
        dict(n_contigs=200, len_lo=1, len_hi=900, depth=5, read_len=100, jitter=40, indel_frac=0.3, gap_hi=10),        # tiny contigs
        dict(n_contigs=20, len_lo=5000, len_hi=60000, depth=15, read_len=250, jitter=0, indel_frac=0.3, gap_hi=5),     # 250 + ~200: inside
        dict(n_contigs=20, len_lo=5000, len_hi=60000, depth=15, read_len=250, jitter=0, indel_frac=0.3, gap_hi=300),   # beyond 512: general
        dict(n_contigs=10, len_lo=70000, len_hi=140000, depth=40, read_len=100, jitter=50, indel_frac=0.5, gap_hi=200),
        dict(n_contigs=3, len_lo=300000, len_hi=700000, depth=3, read_len=60, jitter=30, indel_frac=0.0, gap_hi=0),
        dict(n_contigs=50, len_lo=151, len_hi=400, depth=30, read_len=60, jitter=20, indel_frac=0.2, gap_hi=20),       # windows of a few slots
        dict(n_contigs=6, len_lo=2000, len_hi=8000, depth=5000, read_len=150, jitter=0, indel_frac=0.0, gap_hi=0),     # beyond the deep pool's range
        dict(n_contigs=25, len_lo=3000, len_hi=50000, depth=60, read_len=200, jitter=60, indel_frac=0.4, gap_hi=40),
        dict(n_contigs=8, len_lo=65000, len_hi=66500, depth=25, read_len=150, jitter=15, indel_frac=0.1, gap_hi=8),    # contig ends near run boundaries
    ]
    st, en, ev_off, lengths, _keep = _random_sample(rng, **shapes[case])
    off = np.concatenate([[0], np.cumsum(lengths)])
    wcov, wst = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    span = int((en.astype(np.int64) - st.astype(np.int64)).clip(0).max()) if len(st) else 0
    dis = _disorder(st, ev_off, lengths)
    for kw in ({}, dict(max_span=max(span, 1), max_disorder=dis)):
        cov, stats = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, want_stats=True, **kw)
        for f in INT_FIELDS:
            bad = np.nonzero(stats[f] != wst[f])[0]
            assert bad.size == 0, (case, kw, f, bad[:5], stats[f][bad[:5]], wst[f][bad[:5]])
        assert np.array_equal(cov.view(np.uint32), wcov.view(np.uint32)), (case, kw)
        general = bool((stats["flags"] & GENERAL).any())
        if case == 5:
            assert general, "span + disorder beyond the overflow region must take the general path"
        elif case != 9:  # 5000x is beyond the deep pool's range: general path
            assert not general, (case, kw, span, dis)


def test_depth_that_wanders_across_the_histogram_range(covmod, oracle):
    """One long contig (five 65 536-slot runs) whose depth ramps 60x -> 1500x -> 30x, next to shallow and uniformly
    deep (800x) ones: the shared histogram's range follows the depth (several re-bases per run, bins migrating to
    the pool, partial histograms of different bases merged across runs).  Bit-exact and not the fallback."""
    rng = np.random.default_rng(77)
    lengths = np.array([5000, 330000, 12000, 9000, 140], dtype=np.int64)
    starts, ev_off = [], [0]
    for c, ln in enumerate(lengths.tolist()):
        if ln < 200:
            ev_off.append(len(starts))
            continue
        pos = np.arange(0, ln - 150)
        if c == 1:
            x = pos / (ln - 150)
            depth = np.where(x < 0.5, 60 + (1500 - 60) * (x / 0.5), 1500 - (1500 - 30) * ((x - 0.5) / 0.5))
        else:
            depth = np.full(len(pos), [20, 0, 800, 3][c], dtype=float)
        n_at = rng.poisson(depth / 150.0)
        starts += np.repeat(pos, n_at).tolist()
        ev_off.append(len(starts))
    st = np.array(starts, dtype=np.uint32)
    en = (st + 150).astype(np.uint32)
    ev_off = np.array(ev_off, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lengths)])
    wcov, wst = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    assert wst["max_depth"][1] > 1200 and wst["max_depth"][2] > 700
    for kw in ({}, dict(max_span=150, max_disorder=0)):
        cov, stats = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, want_stats=True, **kw)
        for f in INT_FIELDS:
            assert np.array_equal(stats[f], wst[f]), (f, stats[f], wst[f])
        assert np.array_equal(cov.view(np.uint32), wcov.view(np.uint32))
        assert not (stats["flags"] & GENERAL).any()


@pytest.mark.parametrize("n_shallow", [0, 200000])
def test_more_deep_contigs_than_the_smallest_pool(covmod, oracle, n_shallow):
    """2 500 contigs at 700x all need a row of the deep pool.  Alone they overflow its 2 048 rows and the sample comes
    back from the general path; inside a call with 200 000 more (empty) contigs the pool has grown (one row per 64
    contigs) and the shared-memory path finishes the sample.  Same numbers either way."""
    rng = np.random.default_rng(31)
    n_deep = 2500
    lengths = np.concatenate([np.full(n_deep, 400, dtype=np.int64), rng.integers(160, 300, n_shallow).astype(np.int64)])
    per = int(700 * 250 / 150)
    st = np.sort(rng.integers(0, 250, (n_deep, per)), axis=1).astype(np.uint32).ravel()
    en = (st + 150).astype(np.uint32)
    ev_off = np.concatenate([np.arange(n_deep + 1, dtype=np.int64) * per, np.full(n_shallow, n_deep * per, dtype=np.int64)])
    off = np.concatenate([[0], np.cumsum(lengths)])
    wcov, wst = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    assert (wst["max_depth"][:n_deep] > 600).all()
    cov, stats = covmod.cov_from_events(st, en, ev_off, off, 75, 0.05, 0.05, want_stats=True, max_span=150, max_disorder=0)
    for f in INT_FIELDS:
        assert np.array_equal(stats[f], wst[f]), f
    assert np.array_equal(cov.view(np.uint32), wcov.view(np.uint32))
    assert bool((stats["flags"] & GENERAL).any()) == (n_shallow == 0)


@pytest.mark.parametrize("env", [{"MP2_COV_RUN_ORDER": "0"}, {"MP2_COV_PATH": "general"}, {"MP2_COV_STREAM": "1"}])
def test_development_switches_keep_parity(env, oracle, tmp_path):
    """The A/B switches of the coverage path (runs drawn in array order, general path only, round 1's streaming kernel)
    are read once per process: run each in a child on a sample with a 1 000x organism and compare with the oracle."""
    import subprocess
    import sys

    from magpurify2_b200 import synth

    lengths, mag_off, _gc = synth.mag_layout(4, seed=21, mean_bp=7e5, n_contigs=70)
    st, en, ev_off = synth.make_events(lengths, mag_off, 1, seed=21)
    deep = np.sort(np.random.default_rng(21).integers(0, lengths[3] - 150, int(1000 * lengths[3] / 150))).astype(np.uint32)
    st = np.concatenate([st[:ev_off[3]], deep, st[ev_off[4]:]])
    en = np.concatenate([en[:ev_off[3]], deep + 150, en[ev_off[4]:]]).astype(np.uint32)
    ev_off = ev_off.copy()
    ev_off[4:] += len(deep) - (ev_off[4] - ev_off[3])
    off = np.concatenate([[0], np.cumsum(lengths)])
    want, _ = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=8)
    for name, a in (("st", st), ("en", en), ("ev_off", ev_off), ("off", off)):
        np.save(tmp_path / f"{name}.npy", a)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from magpurify2_b200 import _coverage as c; "
            "d = %r; st, en, eo, off = (np.load(d + '/' + n + '.npy') for n in ('st', 'en', 'ev_off', 'off')); "
            "np.save(d + '/cov.npy', c.cov_from_events(st, en, eo, off, 75, 0.05, 0.05))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path)))
    subprocess.run([sys.executable, "-c", code], check=True, env={**os.environ, **env}, timeout=300)
    assert np.array_equal(np.load(tmp_path / "cov.npy").view(np.uint32), want.view(np.uint32)), env
