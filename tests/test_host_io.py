"""Host-side ingest (no GPU): libmp2's BAM reader against the oracle's per-record restatement of
the reference filter + CIGAR walk, and the FASTA reader against the
reference's record rules (tools.py:210-240, core.py:35-50)."""
import bz2
import gzip
import lzma
import os

import numpy as np
import pytest

from magpurify2_b200 import _bam, _fasta, _lib, bamio, tools


def _random_records(rng, lengths, n):
    recs = []
    tids = np.sort(rng.integers(0, len(lengths), n))
    last_pos = {}
    for tid in tids.tolist():
        ln = lengths[tid]
        if ln < 260:
            continue
        pos = int(rng.integers(last_pos.get(tid, 0), max(last_pos.get(tid, 0) + 1, ln - 255)))
        last_pos[tid] = pos
        kind = rng.integers(0, 6)
        if kind == 0:
            cig = [("M", 150)]
        elif kind == 1:
            cig = [("S", 10), ("M", 70), ("I", 2), ("M", 68)]
        elif kind == 2:
            cig = [("M", 60), ("D", 3), ("M", 90)]
        elif kind == 3:
            cig = [("=", 50), ("X", 1), ("=", 99)]
        elif kind == 4:
            cig = [("H", 5), ("M", 40), ("N", 20), ("M", 110), ("P", 1)]
        else:
            cig = [("M", int(rng.integers(30, 250)))]
        flag = int(rng.choice([0, 0, 0, 0, 0x1, 0x2, 0x10, 0x100, 0x800, 0x400, 0x200]))
        nm = int(rng.binomial(150, 0.012))
        recs.append(dict(tid=int(tid), pos=pos, cigar=cig, flag=flag, nm=nm,
                         nm_type=str(rng.choice(["C", "c", "S", "s", "I", "i"])),
                         extra_aux=b"ASC\x05XSZabc\x00" if rng.random() < 0.3 else b""))
    recs.append(dict(tid=-1, pos=-1, cigar=[], flag=0x4, nm=None, seq_len=20))
    return recs


@pytest.mark.parametrize("min_identity", [0.97, 0.0, 0.99])
def test_bam_reader_matches_record_oracle(tmp_path, oracle, min_identity):
    rng = np.random.default_rng(41)
    names = [f"contig_{i}" for i in range(40)]
    lengths = rng.integers(100, 5000, 40).tolist()
    recs = _random_records(rng, lengths, 3000)
    path = tmp_path / "x.bam"
    bamio.write_bam(str(path), names, lengths, recs)
    want = [[] for _ in names]
    kept = 0
    for r in recs:
        cig = [(ln << 4) | bamio.CIGAR_OPS.index(op) for op, ln in r["cigar"]]
        ev = oracle.record_events(r.get("flag", 0), r["tid"], r["pos"], cig, r["nm"], min_identity)
        if ev:
            kept += 1
            want[r["tid"]] += ev
    with _bam.BamEvents(path, min_identity, threads=3) as b:
        assert b.names == names and b.lengths.tolist() == lengths
        assert b.n_records == len(recs) and b.n_kept == kept
        assert "SO:coordinate" in b.header_text
        for t in range(len(names)):
            got = list(zip(b.starts[b.ev_off[t]:b.ev_off[t + 1]].tolist(), b.ends[b.ev_off[t]:b.ev_off[t + 1]].tolist()))
            assert got == want[t], t
        assert b.max_span >= 150


def test_bam_reader_many_blocks_and_threads(tmp_path):
    """Records straddling BGZF blocks; result independent of the inflate thread count."""
    rng = np.random.default_rng(42)
    n = 60000
    lengths = [200000, 300000]
    tid = np.sort(rng.integers(0, 2, n))
    pos = np.concatenate([np.sort(rng.integers(0, lengths[t] - 150, int((tid == t).sum()))) for t in (0, 1)])
    nm = rng.binomial(150, 0.01, n)
    path = tmp_path / "big.bam"
    bamio.write_bam_fixed(str(path), ["a", "b"], lengths, tid, pos, nm)
    keep = (np.float32(1.0) - nm.astype(np.float32) / np.float32(150)) >= np.float32(0.97)
    with _bam.BamEvents(path, 0.97, threads=1) as b1, _bam.BamEvents(path, 0.97, threads=4) as b4:
        assert b1.n_records == n and b1.n_kept == int(keep.sum())
        assert np.array_equal(b1.starts, pos[keep]) and np.array_equal(b1.ends, pos[keep] + 150)
        assert np.array_equal(b1.starts, b4.starts) and np.array_equal(b1.ev_off, b4.ev_off)
        assert b1.ev_off.tolist() == [0, int((keep & (tid == 0)).sum()), int(keep.sum())]


def test_bam_reader_small_batches(tmp_path, monkeypatch):
    """Inflate rounds far smaller than the file: records (and the header) straddle the rounds, the leftover
    bytes of one round are the head of the next.  Same events as one big round."""
    rng = np.random.default_rng(43)
    n = 20000
    names = [f"contig_{i}" for i in range(40)]
    lengths = [50000] * 40
    tid = np.sort(rng.integers(0, 40, n))
    pos = np.concatenate([np.sort(rng.integers(0, 50000 - 150, int((tid == t).sum()))) for t in range(40)])
    nm = rng.binomial(150, 0.01, n)
    path = tmp_path / "rounds.bam"
    bamio.write_bam_fixed(str(path), names, lengths, tid, pos, nm)
    with _bam.BamEvents(path, 0.97, threads=3) as ref:
        want = (ref.starts.copy(), ref.ends.copy(), ref.ev_off.copy(), ref.n_records, ref.n_kept, ref.max_span)
    for kb in ("1", "65", "300"):  # 1 KiB rounds are smaller than a BGZF block: one block per round
        monkeypatch.setenv("MP2_BAM_BATCH_KB", kb)
        with _bam.BamEvents(path, 0.97, threads=3) as b:
            assert (b.n_records, b.n_kept, b.max_span) == want[3:], kb
            assert np.array_equal(b.starts, want[0]) and np.array_equal(b.ends, want[1]), kb
            assert np.array_equal(b.ev_off, want[2]) and b.names == names, kb


def test_bam_reader_errors(tmp_path):
    p = tmp_path / "nm.bam"
    bamio.write_bam(str(p), ["c"], [1000], [dict(tid=0, pos=1, cigar=[("M", 50)], nm=None)])
    with pytest.raises(_lib.Mp2Error) as ei:  # the reference panics on a missing NM tag
        _bam.BamEvents(p, 0.97)
    assert ei.value.code == _lib.EFORMAT and "NM" in str(ei.value)
    with _bam.BamEvents(p, 0.0) as b:  # no identity filter -> NM is not needed
        assert b.n_kept == 1
    q = tmp_path / "unsorted.bam"
    bamio.write_bam(str(q), ["c", "d"], [1000, 1000], [dict(tid=1, pos=1, cigar=[("M", 50)]),
                                                       dict(tid=0, pos=1, cigar=[("M", 50)])])
    with pytest.raises(_lib.Mp2Error):
        _bam.BamEvents(q, 0.97)
    with pytest.raises(OSError):
        _bam.BamEvents(tmp_path / "missing.bam")
    junk = tmp_path / "junk.bam"
    junk.write_bytes(b"not a bam file at all" * 10)
    with pytest.raises(OSError):
        _bam.BamEvents(junk)


def test_sorted_header_check(tmp_path):
    from pathlib import Path

    a, b = tmp_path / "s.bam", tmp_path / "u.bam"
    bamio.write_bam(str(a), ["c"], [1000], [])
    bamio.write_bam(str(b), ["c"], [1000], [], sort_order="unsorted")
    assert tools.is_sorted_bam(Path(a)) and not tools.is_sorted_bam(Path(b))


FA = ">c~1 first contig\nACGT\nacgtN\r\n>c2\n\n>c3\tx y\nAC GT\nTT\n"


@pytest.mark.parametrize("opener,ext", [(open, ".fa"), (gzip.open, ".fa.gz"), (bz2.open, ".fa.bz2"), (lzma.open, ".fa.xz")])
def test_fasta_reader(tmp_path, opener, ext):
    from pathlib import Path

    from magpurify2_b200.core import Mag

    p = tmp_path / ("genome.v1" + ext)
    with opener(p, "wt") as f:
        f.write(FA)
    recs = list(tools.read_fasta(p))
    assert recs == [("c_1", "c~1 first contig", "ACGTacgtN"), ("c2", "c2", ""), ("c3", "c3\tx y", "ACGTTT")]
    mag = Mag(Path(p))
    assert mag.genome == "genome.v1"  # double stem for compressed files (core.py:37-39)
    assert mag.contigs == ["c_1", "c2", "c3"] and mag.lengths.tolist() == [9, 0, 6]
    assert mag.sequences == ["ACGTacgtN", "", "ACGTTT"] and len(mag) == 3
    assert Mag(Path(p), store_sequences=False).sequences == []


def test_fasta_reader_odd_lines(tmp_path):
    """Lines that need the byte-wise path (inner blanks, CRLF), an inner tab (kept: Biopython's parser under
    tools.read_fasta only strips trailing white space and removes ' ' and CR), text before the first header, no
    final newline, '>' inside a sequence line, and a file bigger than the reader's first window."""
    rng = np.random.default_rng(3)
    big = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), 3_000_000)).decode()
    wrapped = "\n".join(big[i:i + 61] for i in range(0, len(big), 61))
    text = "junk before\n>a d1\r\nAC\tGT \r\n A>C \n>b\n" + wrapped + "\n>c\nTTTT"
    p = tmp_path / "odd.fa"
    p.write_text(text)
    recs = list(tools.read_fasta(p))
    assert [r[0] for r in recs] == ["a", "b", "c"] and recs[0][1] == "a d1"
    assert recs[0][2] == "AC\tGTA>C" and recs[1][2] == big and recs[2][2] == "TTTT"


def test_read_mags_thread_pool(tmp_path):
    from magpurify2_b200.core import read_mags

    rng = np.random.default_rng(4)
    paths = []
    for m in range(7):
        p = tmp_path / f"g{m}.fna"
        with open(p, "w") as f:
            for c in range(int(rng.integers(1, 6))):
                f.write(f">g{m}_c{c}\n" + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(0, 5000)))).decode() + "\n")
        paths.append(p)
    one, many = read_mags(paths, threads=1), read_mags(paths, threads=4)
    assert [m.genome for m in many] == [f"g{m}" for m in range(7)]  # input order is kept
    for a, b in zip(one, many):
        assert a.contigs == b.contigs and np.array_equal(a.bases, b.bases) and np.array_equal(a.offsets, b.offsets)
    assert read_mags([], threads=8) == []


def test_validade_input_and_signature(tmp_path):
    assert tools.validade_input(5, "x", [1, 3]) == 3 and tools.validade_input(0.5, "x", [0.0, 1.0]) == 0.5
    p = tmp_path / "f.bin"
    p.write_bytes(os.urandom(200000))
    import hashlib

    assert tools.get_file_signature(p) == hashlib.sha256(p.read_bytes()[-128000:]).hexdigest()
    small = tmp_path / "s.bin"
    small.write_bytes(b"x" * 10)
    with pytest.raises(OSError):  # as the reference: seek(-128000) fails on small files
        tools.get_file_signature(small)


def _pack_reference(b):
    L = len(b)
    nq = (L + 63) // 64
    hb = np.frombuffer(bytes(b).upper(), dtype=np.uint8)
    lut = np.full(256, 255, np.uint8)
    lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
    c = lut[hb]
    v = c != 255
    pad = nq * 64 - L
    c2 = np.concatenate([np.where(v, c, 0), np.zeros(pad, np.uint8)]).reshape(-1, 16).astype(np.uint32)
    codes = (c2 << (2 * (15 - np.arange(16, dtype=np.uint32)))).sum(axis=1).astype(np.uint32)
    v2 = np.concatenate([v, np.zeros(pad, bool)]).reshape(-1, 32).astype(np.uint64)
    valid = (v2 << (31 - np.arange(32, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
    return codes, valid


def test_host_packer_matches_the_layout_definition():
    """mp2_pack_host (AVX-512 / scalar, threaded) on the virtual concatenation of several buffers: 2 bits per base,
    first base most significant, validity bit per base, padded to quads of 64 -- against a NumPy restatement."""
    from magpurify2_b200 import _composition

    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNRYnX-*\x00\xff", dtype=np.uint8)
    for trial in range(40):
        k = 8 if trial % 3 == 0 else len(alphabet)
        parts = [alphabet[rng.integers(0, k, int(rng.integers(0, 900)))].copy() for _ in range(int(rng.integers(1, 9)))]
        if trial == 7:
            parts = [np.zeros(0, np.uint8)]
        codes, valid = _composition.pack_parts(parts, threads=int(rng.integers(1, 5)))
        want_c, want_v = _pack_reference(np.concatenate(parts))
        assert np.array_equal(codes, want_c) and np.array_equal(valid, want_v), trial
    big = alphabet[rng.integers(0, 9, 3_000_001)].copy()  # many tasks, a tail that is not a whole quad
    codes, valid = _composition.pack_parts([big[:1_000_003], big[1_000_003:]], threads=4)
    want_c, want_v = _pack_reference(big)
    assert np.array_equal(codes, want_c) and np.array_equal(valid, want_v)


def test_get_bam_coverages_propagates_decode_errors_without_hanging(tmp_path):
    """The decoder thread runs ahead of the device threads: a file that cannot be decoded (no NM tag: the reference
    panics there) or whose @SQ lines differ must surface as the call's exception, not as a stuck pipeline.  Both
    fail on the host, before any GPU work."""
    from magpurify2_b200 import _coverage

    names, lengths = ["a", "b"], [5000, 6000]
    bad = tmp_path / "no_nm.bam"
    bamio.write_bam(str(bad), names, lengths, [dict(tid=0, pos=10, cigar=[("M", 100)], nm=None)])
    with pytest.raises(_lib.Mp2Error):
        _coverage.get_bam_coverages([str(bad)], threads=2)
    ok = tmp_path / "ok.bam"
    bamio.write_bam(str(ok), names, lengths, [dict(tid=0, pos=10, cigar=[("M", 100)], nm=0)])
    with pytest.raises(_lib.Mp2Error):  # the good file first: its device thread fails on the CPU box, or runs on a GPU box
        _coverage.get_bam_coverages([str(ok), str(bad), str(ok)], threads=2, devices=[0, 0])
    with pytest.raises((FileNotFoundError, OSError)):
        _coverage.get_bam_coverages([str(tmp_path / "missing.bam")])


def test_bam_reader_reports_block_span_and_disorder(tmp_path, oracle):
    """max_block / max_disorder: what get_bam_coverages vouches for; the CG:B,I tag of a read with a placeholder CIGAR."""
    import struct

    names, lengths = ["c0"], [100000]
    real = [("M", 40), ("D", 7), ("M", 50), ("N", 200), ("M", 60)]
    packed = [(ln << 4) | bamio.CIGAR_OPS.index(op) for op, ln in real]
    cg = b"CGBI" + struct.pack("<I", len(packed)) + b"".join(struct.pack("<I", c) for c in packed)
    recs = [
        dict(tid=0, pos=100, cigar=[("M", 150)], nm=0),
        dict(tid=0, pos=120, cigar=[("M", 30), ("D", 2), ("M", 120)], nm=2),       # second block starts at 152
        dict(tid=0, pos=130, cigar=[("S", 150), ("N", 357)], nm=0, seq_len=150, extra_aux=cg),  # placeholder + CG tag
        dict(tid=0, pos=140, cigar=[("M", 150)], nm=0),
    ]
    path = tmp_path / "cg.bam"
    bamio.write_bam(str(path), names, lengths, recs)
    with _bam.BamEvents(path, 0.0, threads=2) as b:
        got = list(zip(b.starts.tolist(), b.ends.tolist()))
        assert got == [(100, 250), (120, 150), (152, 272), (130, 170), (177, 227), (427, 487), (140, 290)]
        assert b.max_block == 150
        assert b.max_disorder == 427 - 130          # the last block of the spliced read starts 297 after its position
        assert b.max_span == 487 - 130
