"""GPU parity: composition path (get_tnf / get_gc) through the C ABI vs the CPU oracle.

Counts are compared bit-exactly; relative frequencies and GC are single IEEE f32 divisions of
exact integers, so they are compared bit-exactly as well (tolerance in north_star: 1e-6 rel)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same_f32(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def comp():
    from magpurify2_b200 import _composition

    return _composition


def test_kats(comp, oracle, golden_dir):
    kats = json.load(open(os.path.join(golden_dir, "kats.json")))
    seqs = [k[0] for k in kats["tnf_gc"]]
    counts = comp.get_tnf(seqs, relative=False)
    assert counts.dtype == np.uint32 and counts.shape == (len(seqs), 136)
    assert np.array_equal(counts, oracle.get_tnf(seqs, relative=False))
    for row, k in zip(counts, kats["tnf_gc"]):
        assert row.sum() == k[1]
    rel = comp.get_tnf(seqs)
    assert rel.dtype == np.float32
    assert _same_f32(rel, oracle.get_tnf(seqs, relative=True))
    gc = comp.get_gc(seqs)
    assert gc.dtype == np.float32 and _same_f32(gc, oracle.get_gc(seqs))
    assert np.isnan(gc[seqs.index("")])


def test_signature_matches_reference(comp):
    # positional and keyword use (PyO3 #[pyfunction(relative=true, threads=1)], composition.rs:94)
    a = comp.get_tnf(["ACGTTGCA"], True, 4)
    b = comp.get_tnf(seq_list=["ACGTTGCA"], relative=True, threads=1)
    assert np.array_equal(a, b)
    assert comp.get_tnf([]).shape == (0, 136)
    assert comp.get_gc([]).shape == (0,)
    with pytest.raises(TypeError):
        comp.get_tnf("ACGT")
    with pytest.raises(TypeError):
        comp.get_tnf([1, 2])


def test_non_ascii_strings(comp, oracle):
    # a non-ASCII character is one `char` to the reference (chars().count(), src/composition.rs:131) and its UTF-8 bytes
    # only reset the k-mer window (composition.rs:48-61): GC denominators count characters, k-mers never span it
    seqs = ["GCéAT", "ACGTé", "éACGTACGT", "ACGT\u4e2dACGT\U0001F9ECTTTT", "é", "ACGTACGTé" * 40]
    counts = comp.get_tnf(seqs, relative=False)
    assert np.array_equal(counts, oracle.get_tnf(seqs, relative=False))
    assert counts[3].sum() == 1 + 1 + 1 and counts[0].sum() == 0
    assert _same_f32(comp.get_tnf(seqs), oracle.get_tnf(seqs, relative=True))
    gc = comp.get_gc(seqs)
    assert _same_f32(gc, oracle.get_gc(seqs))
    assert gc[0] == np.float32(2) / np.float32(5) and gc[4] == 0.0


def test_ragged_random(comp, oracle):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTacgtNRYKMSWnU-*", dtype=np.uint8)
    p = np.array([20, 20, 20, 20, 3, 3, 3, 3, 2, .5, .5, .5, .5, .5, .5, 1, .5, .5, .5])
    lens = [0, 1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 1023, 1024, 1025, 4096, 65535,
            65536, 65537, 70000, 0, 0, 3, 200000, 7, 131072, 1] + rng.integers(0, 3000, 300).tolist()
    seqs = [bytes(rng.choice(alphabet, n, p=p / p.sum())).decode() for n in lens]
    assert np.array_equal(comp.get_tnf(seqs, relative=False), oracle.get_tnf(seqs, relative=False, threads=4))
    assert _same_f32(comp.get_tnf(seqs), oracle.get_tnf(seqs, threads=4))
    assert _same_f32(comp.get_gc(seqs), oracle.get_gc(seqs))


def test_all_boundaries_mod_16(comp, oracle):
    """Every contig start alignment (mod 16) x N positions around the 16-byte groups."""
    rng = np.random.default_rng(12)
    seqs = []
    for pad in range(0, 40):
        seqs.append("ACGT"[pad % 4] * pad)
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 2100).copy()
        s[rng.integers(0, 2100, 3)] = ord("N")
        seqs.append(bytes(s).decode())
    assert np.array_equal(comp.get_tnf(seqs, relative=False), oracle.get_tnf(seqs, relative=False))
    assert _same_f32(comp.get_gc(seqs), oracle.get_gc(seqs))


def test_revcomp_invariance_and_sum(comp):
    rng = np.random.default_rng(13)
    s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 300001)).decode()
    rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    a, b = comp.get_tnf([s, rc], relative=False)
    assert np.array_equal(a, b) and a.sum() == len(s) - 3


def test_synthetic_mags_c1(comp, oracle):
    """BASELINE config 1 shape: 10 MAGs x ~3 Mbp x ~300 contigs, full parity."""
    from magpurify2_b200 import synth

    d = synth.make_mags(10, seed=1)
    counts, rel, gc = comp.tnf_from_concat(d["bases"], d["offsets"], relative=True, want_gc=True,
                                           want_counts=True)
    want = oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False, threads=8)
    assert np.array_equal(counts, want)
    assert _same_f32(rel, oracle.get_tnf_concat(d["bases"], d["offsets"], relative=True, threads=8))
    assert _same_f32(gc, oracle.get_gc_concat(d["bases"], d["offsets"]))


def test_long_contig_c5(comp, oracle):
    """BASELINE config 5 shape: few very large contigs (work per CTA independent of contig size)."""
    from magpurify2_b200 import synth

    d = synth.make_mags(2, seed=2, long_contigs=True)
    counts, _rel, gc = comp.tnf_from_concat(d["bases"], d["offsets"], relative=False, want_gc=True)
    assert np.array_equal(counts, oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False, threads=8))
    assert _same_f32(gc, oracle.get_gc_concat(d["bases"], d["offsets"]))


def test_device_pointer_entry(comp, oracle):
    """mp2_tnf_device on torch buffers (tensors are buffers only), unaligned base pointer too."""
    import torch

    from magpurify2_b200 import _lib, synth

    d = synth.make_mags(2, seed=3, mean_bp=4e5, n_contigs=60)
    n = len(d["offsets"]) - 1
    want = oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False, threads=4)
    want_gc = oracle.get_gc_concat(d["bases"], d["offsets"])
    lib = _lib.load()
    for shift in (0, 1, 5, 15):
        buf = torch.zeros(len(d["bases"]) + 64, dtype=torch.uint8, device="cuda")
        view = buf[shift:shift + len(d["bases"])]
        view.copy_(torch.from_numpy(d["bases"]))
        off = torch.from_numpy(d["offsets"]).cuda()
        counts = torch.full((n, 136), 7, dtype=torch.int32, device="cuda")
        rel = torch.empty((n, 136), dtype=torch.float32, device="cuda")
        gc = torch.empty(n, dtype=torch.float32, device="cuda")
        _lib.check(lib.mp2_tnf_device(view.data_ptr(), off.data_ptr(), n, len(d["bases"]),
                                      counts.data_ptr(), rel.data_ptr(), gc.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy().view(np.uint32), want), shift
        assert _same_f32(gc.cpu().numpy(), want_gc)


def test_normalize_device_equals_relative_output(comp, oracle):
    """mp2_tnf_normalize_device on the counts of mp2_tnf_device == the relative output of the same call == the oracle
    (empty and < 4 bp sequences: all-zero rows, NaN -> 0)."""
    import torch

    from magpurify2_b200 import _lib, synth

    d = synth.make_mags(2, seed=9, mean_bp=3e5, n_contigs=70)
    n = len(d["offsets"]) - 1
    want = oracle.get_tnf_concat(d["bases"], d["offsets"], relative=True, threads=4)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    d_b, d_o = torch.from_numpy(d["bases"]).cuda(), torch.from_numpy(d["offsets"]).cuda()
    counts = torch.empty((n, 136), dtype=torch.int32, device="cuda")
    rel_a = torch.empty((n, 136), dtype=torch.float32, device="cuda")
    rel_b = torch.full((n, 136), 5.0, dtype=torch.float32, device="cuda")
    _lib.check(lib.mp2_tnf_device(d_b.data_ptr(), d_o.data_ptr(), n, len(d["bases"]), counts.data_ptr(), rel_a.data_ptr(), None, st))
    _lib.check(lib.mp2_tnf_normalize_device(counts.data_ptr(), n, rel_b.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(rel_a, rel_b)
    assert _same_f32(rel_b.cpu().numpy(), want)
    assert lib.mp2_tnf_normalize_device(None, n, rel_b.data_ptr(), st) != 0


def test_device_entry_defines_every_row(comp, oracle):
    """mp2_tnf_device does not clear the count rows wholesale: contigs that lie inside one piece are STORED by their
    one flush, only rows of contigs cut by a piece boundary and of empty contigs are cleared first.  Output buffers
    full of garbage, contigs of every kind (empty ones in runs, at the very start and end, shorter than a k-mer, longer
    than several pieces, ending exactly on piece boundaries), twice into the same buffers."""
    import torch

    from magpurify2_b200 import _lib

    rng = np.random.default_rng(41)
    alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    lens = ([0, 0, 3, 16384, 0, 16384 - 3, 1, 0, 0, 70000, 2, 16384 * 3, 5, 0, 40000, 0] +
            rng.integers(0, 6000, 400).tolist() + [0, 200000, 0, 0, 7, 32768, 0])
    bases = rng.choice(alphabet, int(sum(lens)), p=[.24, .24, .24, .24, .01, .01, .005, .005, .01])
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = len(lens)
    want = oracle.get_tnf_concat(bases, offsets, relative=False, threads=4)
    want_gc = oracle.get_gc_concat(bases, offsets)
    lib = _lib.load()
    d_b, d_o = torch.from_numpy(bases).cuda(), torch.from_numpy(offsets).cuda()
    counts = torch.full((n, 136), -559038737, dtype=torch.int32, device="cuda")
    gc = torch.full((n,), 123.0, dtype=torch.float32, device="cuda")
    for _ in range(2):
        _lib.check(lib.mp2_tnf_device(d_b.data_ptr(), d_o.data_ptr(), n, len(bases), counts.data_ptr(), None,
                                      gc.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy().view(np.uint32), want)
        assert _same_f32(gc.cpu().numpy(), want_gc)


@pytest.mark.parametrize("env", [{"MP2_TNF_MODE": "4"}, {"MP2_TNF_PIECE_SHIFT": "15"}, {"MP2_TNF_SNAP": "0"},
                                 {"MP2_TNF_PIECE_SHIFT": "15", "MP2_TNF_SNAP": "0"}])
def test_development_switches_keep_parity(env, oracle, tmp_path):
    """The A/B switches of the counting kernel (one increment per base, 32 KiB pieces on a small input,
    nominal piece boundaries) are read once per process: run each in a child and compare with the oracle."""
    import subprocess
    import sys

    from magpurify2_b200 import synth

    d = synth.make_mags(3, seed=5, mean_bp=6e5, n_contigs=90)
    want = oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False, threads=4)
    np.save(tmp_path / "bases.npy", d["bases"])
    np.save(tmp_path / "offsets.npy", d["offsets"])
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from magpurify2_b200 import _composition as c; "
            "b = np.load(%r); o = np.load(%r); "
            "k, _r, g = c.tnf_from_concat(b, o, relative=False, want_gc=True); np.save(%r, k)"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "bases.npy"),
               str(tmp_path / "offsets.npy"), str(tmp_path / "counts.npy")))
    subprocess.run([sys.executable, "-c", code], check=True, env={**os.environ, **env}, timeout=300)
    assert np.array_equal(np.load(tmp_path / "counts.npy"), want), env
