"""GPU parity of the remaining entry points: 2-bit packed composition path, get_bam_coverages on
real BAM files, and the two module drivers end to end (FASTA + BAM in, TSV out)."""
import gzip
import pickle
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_packed_path_matches_ascii_and_oracle(oracle):
    import torch

    from magpurify2_b200 import _lib, synth

    lib = _lib.load()
    d = synth.make_mags(3, seed=7, mean_bp=6e5, n_contigs=90)
    rng = np.random.default_rng(7)
    bases = d["bases"].copy()
    bases[rng.integers(0, len(bases), 500)] = ord("N")
    for off_extra in (0, 1):  # also a stream whose length is not a multiple of 64
        b = bases[:len(bases) - off_extra * 13]
        offsets = d["offsets"].copy()
        offsets[-1] = len(b)
        n, L = len(offsets) - 1, len(b)
        want = oracle.get_tnf_concat(b, offsets, relative=False, threads=4)
        want_gc = oracle.get_gc_concat(b, offsets)
        nq = (L + 63) // 64
        d_b = torch.from_numpy(b).cuda()
        codes = torch.zeros(nq * 4, dtype=torch.int32, device="cuda")
        valid = torch.zeros(nq * 2, dtype=torch.int32, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.mp2_pack_device(d_b.data_ptr(), L, codes.data_ptr(), valid.data_ptr(), s))
        # the packed form itself: 2 bits per base first-base-most-significant, valid bit per base
        hb = np.frombuffer(b.tobytes().upper(), dtype=np.uint8)
        lut = np.full(256, 255, np.uint8)
        lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
        c = lut[hb]
        v = (c != 255)
        pad = nq * 64 - L
        c2 = np.concatenate([np.where(v, c, 0), np.zeros(pad, np.uint8)]).reshape(-1, 16).astype(np.uint32)
        want_codes = (c2 << (2 * (15 - np.arange(16, dtype=np.uint32)))).sum(axis=1).astype(np.uint32)
        v2 = np.concatenate([v, np.zeros(pad, bool)]).reshape(-1, 32).astype(np.uint64)
        want_valid = (v2 << (31 - np.arange(32, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        assert np.array_equal(codes.cpu().numpy().view(np.uint32), want_codes)
        assert np.array_equal(valid.cpu().numpy().view(np.uint32), want_valid)
        off = torch.from_numpy(offsets).cuda()
        counts = torch.empty((n, 136), dtype=torch.int32, device="cuda")
        gc = torch.empty(n, dtype=torch.float32, device="cuda")
        rel = torch.empty((n, 136), dtype=torch.float32, device="cuda")
        _lib.check(lib.mp2_tnf_packed_device(codes.data_ptr(), valid.data_ptr(), off.data_ptr(), n, L,
                                             counts.data_ptr(), rel.data_ptr(), gc.data_ptr(), s))
        torch.cuda.synchronize()
        assert np.array_equal(counts.cpu().numpy().view(np.uint32), want)
        assert np.array_equal(gc.cpu().numpy(), want_gc, equal_nan=True)


def test_packed_host_path_matches_ascii_path_and_oracle(oracle):
    """pack_parts (host) + tnf_from_packed (0.375 B/base upload) == tnf_from_concat == oracle, bit for bit; the
    host packer and the device packer produce the same words."""
    import torch

    from magpurify2_b200 import _composition, _lib, synth

    d = synth.make_mags(5, seed=23, mean_bp=7e5, n_contigs=70)
    bases, offsets = d["bases"], d["offsets"]
    cut = [0, int(offsets[40]), int(offsets[41]) + 17, int(offsets[200]), len(bases)]  # parts need not end at contigs
    parts = [bases[a:b] for a, b in zip(cut[:-1], cut[1:])]
    codes, valid = _composition.pack_parts(parts, threads=3)
    L = len(bases)
    nq = (L + 63) // 64
    d_b = torch.from_numpy(bases).cuda()
    d_codes = torch.zeros(nq * 4, dtype=torch.int32, device="cuda")
    d_valid = torch.zeros(nq * 2, dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().mp2_pack_device(d_b.data_ptr(), L, d_codes.data_ptr(), d_valid.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(d_codes.cpu().numpy().view(np.uint32), codes)
    assert np.array_equal(d_valid.cpu().numpy().view(np.uint32), valid)
    c1, r1, g1 = _composition.tnf_from_packed(codes, valid, offsets, relative=True, want_gc=True, want_counts=True)
    c2, r2, g2 = _composition.tnf_from_concat(bases, offsets, relative=True, want_gc=True, want_counts=True)
    assert np.array_equal(c1, c2) and np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
    assert np.array_equal(g1, g2, equal_nan=True)
    assert np.array_equal(c1, oracle.get_tnf_concat(bases, offsets, relative=False, threads=4))
    assert np.array_equal(g1, oracle.get_gc_concat(bases, offsets), equal_nan=True)
    with pytest.raises(ValueError):
        _composition.tnf_from_packed(codes[:8], valid, offsets)
    # the runs of invalid bases instead of the mask (0.25 B/base): the device rebuilds the mask
    runs = _composition.invalid_runs(valid, L, threads=3)
    inv = ~np.isin(bases, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    edge = np.diff(np.concatenate([[0], inv.astype(np.int8), [0]]))
    assert np.array_equal(runs, np.stack([np.nonzero(edge == 1)[0], np.nonzero(edge == -1)[0]], axis=1))
    c3, r3, g3 = _composition.tnf_from_packed(codes, None, offsets, relative=True, want_gc=True, want_counts=True, runs=runs)
    assert np.array_equal(c3, c2) and np.array_equal(r3.view(np.uint32), r2.view(np.uint32)) and np.array_equal(g3, g2, equal_nan=True)
    # a stream whose invalid bases are everywhere: long runs, runs at both ends, a tail that is not a whole quad
    rng = np.random.default_rng(23)
    b2 = bases[:1_000_037].copy()
    off2 = np.concatenate([offsets[offsets < len(b2)], [len(b2)]])
    b2[:100] = ord("N")
    b2[-70:] = ord("n")
    b2[500_000:640_000] = ord("N")
    b2[rng.integers(0, len(b2), 20000)] = ord("R")
    cd, vd = _composition.pack_parts([b2], threads=2)
    want = oracle.get_tnf_concat(b2, off2, relative=False, threads=4)
    for kw in (dict(valid=vd), dict(valid=None, runs=_composition.invalid_runs(vd, len(b2)))):
        got, _r, gg = _composition.tnf_from_packed(cd, kw.pop("valid"), off2, relative=False, want_gc=True, **kw)
        assert np.array_equal(got, want) and np.array_equal(gg, oracle.get_gc_concat(b2, off2), equal_nan=True)


def _write_dataset(tmp, n_mags=4, n_samples=3, seed=9):
    """FASTA files + coordinate-sorted BAMs of a small synthetic community."""
    from magpurify2_b200 import bamio, synth

    d = synth.make_mags(n_mags, seed=seed, mean_bp=2.5e5, n_contigs=40)
    names, genomes = [], []
    for m in range(n_mags):
        path = tmp / f"mag{m}.fna" if m % 2 else tmp / f"mag{m}.fna.gz"
        opener = open if m % 2 else gzip.open
        with opener(path, "wt") as f:
            for c in range(d["mag_off"][m], d["mag_off"][m + 1]):
                nm = f"m{m}_c{c}"
                names.append(nm)
                seq = d["bases"][d["offsets"][c]:d["offsets"][c + 1]].tobytes().decode()
                f.write(f">{nm} len={len(seq)}\n")
                for o in range(0, len(seq), 80):
                    f.write(seq[o:o + 80] + "\n")
        genomes.append(path)
    bams, events = [], []
    rng = np.random.default_rng(seed)
    for s in range(n_samples):
        st, en, ev_off = synth.make_events(d["lengths"], d["mag_off"], s, seed=seed, depth_scale=0.5)
        tid = np.repeat(np.arange(len(d["lengths"])), np.diff(ev_off))
        nm = rng.binomial(150, 0.012, len(st))
        flag = rng.choice([0, 1, 0x100, 0x800, 0x400], len(st), p=[0.9, 0.05, 0.02, 0.02, 0.01])
        p = tmp / f"sample{s}.bam"
        bamio.write_bam_fixed(str(p), names, d["lengths"], tid, st, nm, flag=flag)
        keep = ((np.float32(1.0) - nm.astype(np.float32) / np.float32(150)) >= np.float32(0.97)) & ((flag & 0x900) == 0)
        k_off = np.concatenate([[0], np.cumsum(np.bincount(tid[keep], minlength=len(d["lengths"])))])
        events.append((st[keep], en[keep], k_off))
        bams.append(p)
    return d, names, genomes, bams, events


def test_get_bam_coverages(tmp_path, oracle):
    from magpurify2_b200 import tools

    d, names, _g, bams, events = _write_dataset(tmp_path)
    got_names, mat = tools.get_bam_coverages([str(b) for b in bams], min_identity=0.97, trim_lower=0.05,
                                             trim_upper=0.05, threads=4)
    assert got_names == names and mat.dtype == np.float32 and mat.shape == (len(names), len(bams))
    for s, (st, en, k_off) in enumerate(events):
        want, _ = oracle.cov_sample(st, en, k_off, d["lengths"], 75, 0.05, 0.05, threads=4)
        assert np.array_equal(mat[:, s], want)
    sub = set(names[5:25])
    n2, m2 = tools.get_bam_coverages([str(bams[0])], contig_set=sub, trim_lower=0.05, trim_upper=0.05)
    assert n2 == names[5:25] and np.array_equal(m2[:, 0], mat[5:25, 0])
    with pytest.raises(TypeError):
        tools.get_bam_coverages(str(bams[0]))


def test_get_bam_coverages_with_indel_and_spliced_cigars(tmp_path, oracle):
    """BAMs whose reads carry D / N / I / S / H / = / X operations (two or three blocks per read, the later
    ones out of start order) through get_bam_coverages: bit-exact against the oracle's record walk + depth
    arithmetic, and computed by the shared-memory path -- not the HBM fallback -- for every contig."""
    from magpurify2_b200 import bamio, tools

    rng = np.random.default_rng(77)
    names = [f"c{i}" for i in range(30)]
    lengths = rng.integers(200, 60000, 30).astype(np.int64)
    lengths[3], lengths[11] = 140, 100000
    shapes = [
        [("M", 150)],
        [("S", 10), ("M", 70), ("I", 2), ("M", 68)],
        [("M", 60), ("D", 3), ("M", 90)],
        [("=", 50), ("X", 1), ("=", 99)],
        [("H", 5), ("M", 40), ("N", 120), ("M", 110)],
        [("M", 30), ("D", 1), ("M", 30), ("I", 1), ("M", 30), ("N", 60), ("M", 58)],
        [("M", 100)],
    ]
    recs, want_ev = [], [[] for _ in names]
    for tid, ln in enumerate(lengths.tolist()):
        if ln < 600:
            continue
        pos = np.sort(rng.integers(0, ln - 420, int(ln * 20 / 150)))
        for p in pos.tolist():
            cig = shapes[int(rng.choice(len(shapes), p=[0.55, 0.1, 0.1, 0.05, 0.08, 0.07, 0.05]))]
            flag = int(rng.choice([0, 1, 0x10, 0x100, 0x800, 0x400], p=[0.8, 0.08, 0.08, 0.015, 0.015, 0.01]))
            nm = int(rng.binomial(150, 0.012))
            recs.append(dict(tid=tid, pos=p, cigar=cig, flag=flag, nm=nm))
            packed = [(l << 4) | bamio.CIGAR_OPS.index(op) for op, l in cig]
            want_ev[tid] += oracle.record_events(flag, tid, p, packed, nm, 0.97)
    path = tmp_path / "spliced.bam"
    bamio.write_bam(str(path), names, lengths.tolist(), recs)
    st = np.array([e[0] for ev in want_ev for e in ev], dtype=np.uint32)
    en = np.array([e[1] for ev in want_ev for e in ev], dtype=np.uint32)
    ev_off = np.concatenate([[0], np.cumsum([len(ev) for ev in want_ev])]).astype(np.int64)
    assert (np.diff(st.astype(np.int64)) < 0).sum() > 100  # the later blocks really are out of order
    want, wst = oracle.cov_sample(st, en, ev_off, lengths, 75, 0.05, 0.05, threads=4)
    stats = []
    got_names, mat = tools.get_bam_coverages([str(path)], min_identity=0.97, trim_lower=0.05, trim_upper=0.05,
                                             threads=4, stats_out=stats)
    assert got_names == names
    assert np.array_equal(mat[:, 0].view(np.uint32), want.view(np.uint32))
    assert np.array_equal(stats[0]["total"], wst["total"]) and np.array_equal(stats[0]["max_depth"], wst["max_depth"])
    assert not (stats[0]["flags"] & 1).any(), "fell back to the general path"


def test_cli_composition_and_coverage(tmp_path, oracle):
    from magpurify2_b200 import cli

    d, names, genomes, bams, events = _write_dataset(tmp_path, n_mags=4, n_samples=3, seed=11)
    out = tmp_path / "out"
    assert cli.cli(["composition", *map(str, genomes), str(out), "-q"]) == 0
    assert cli.cli(["coverage", *map(str, genomes), str(out), "--bam_files", *map(str, bams), "-q", "-t", "4"]) == 0
    # ---- composition TSV ----
    rows = [l.rstrip("\n").split("\t") for l in open(out / "scores" / "composition_scores.tsv")]
    assert rows[0] == ["genome", "contig", "tnf_score", "gc_content_score", "contig_length"]
    assert [r[1] for r in rows[1:]] == names
    assert rows[1][0] == "mag0" and rows[-1][0] == "mag3"
    gc = oracle.get_gc_concat(d["bases"], d["offsets"])
    for m in range(4):
        sl = slice(d["mag_off"][m], d["mag_off"][m + 1])
        want, _ = oracle.log_ratio_scores(gc[sl], d["lengths"][sl], 2)
        got = np.array([float(r[3]) for r in rows[1:][sl]])
        assert np.allclose(got, np.round(want, 5), atol=1.1e-5, equal_nan=True)
        counts = oracle.get_tnf_concat(d["bases"][d["offsets"][sl.start]:d["offsets"][sl.stop]],
                                       d["offsets"][sl.start:sl.stop + 1] - d["offsets"][sl.start], relative=False)
        _clr, dist = oracle.clr_centroid(counts, d["lengths"][sl])
        # tnf_score = 1 - max(0, log4(d / weighted median d)), from the oracle's pieces
        med = oracle.weighted_median(dist.astype(np.float32), d["lengths"][sl])
        lg = np.log((dist.astype(np.float32) + np.float32(1e-5)) / (med + np.float32(1e-5))) / np.log(4.0)
        want_tnf = np.clip(1.0 - np.maximum(lg, 0.0), 0, 1)
        got_tnf = np.array([float(r[2]) for r in rows[1:][sl]])
        assert np.allclose(got_tnf, want_tnf, atol=2e-5)
        assert [int(r[4]) for r in rows[1:][sl]] == d["lengths"][sl].tolist()
    with gzip.open(out / "tnfs.pickle.gz") as f:
        tn = pickle.load(f)
    assert sorted(tn) == ["mag0", "mag1", "mag2", "mag3"] and tn["mag0"].dtype == np.float32
    assert np.array_equal(np.concatenate([tn[f"mag{m}"] for m in range(4)]),
                          oracle.get_tnf_concat(d["bases"], d["offsets"], relative=True))
    # ---- coverage TSV ----
    rows = [l.rstrip("\n").split("\t") for l in open(out / "scores" / "coverage_scores.tsv")]
    assert rows[0] == ["genome", "contig", "coverage_score", "coverage_cluster_score", "n_samples"]
    cov = np.stack([oracle.cov_sample(st, en, k, d["lengths"], 75, 0.05, 0.05)[0] for st, en, k in events], axis=1)
    for m in range(4):
        sl = slice(d["mag_off"][m], d["mag_off"][m + 1])
        sel, _ = oracle.select_samples(cov[sl], d["lengths"][sl], 1.0)
        got = np.array([float(r[2]) for r in rows[1:][sl]])
        if sel.any():
            want, _ = oracle.log_ratio_scores(cov[sl][:, sel], d["lengths"][sl], 25)
            assert np.allclose(got, np.round(want, 5), atol=1.1e-5)
        else:
            assert (got == 1.0).all()
        assert all(int(r[4]) == int(sel.sum()) for r in rows[1:][sl])
    with gzip.open(out / "coverages.pickle.gz") as f:
        sig, cn, mat = pickle.load(f)
    assert len(sig) == 3 and list(cn) == names and np.array_equal(mat, cov)
    # second run hits the cache (BAMs are not opened again: remove them to prove it)
    for b in bams[1:]:
        b.write_bytes(b.read_bytes())  # same content, same signature
    assert cli.cli(["coverage", *map(str, genomes), str(out), "--bam_files", *map(str, bams), "-q"]) == 0


def test_dropin_modules_on_gpu(tmp_path, oracle, monkeypatch):
    """The alias modules magpurify2_b200.dropin writes (`magpurify2._composition`, `magpurify2._coverage`: the
    module paths magpurify2/tools.py:40-41 imports) resolve and compute on the GPU."""
    import importlib
    import sys

    from magpurify2_b200 import dropin, synth

    pkg = tmp_path / "site" / "magpurify2"
    pkg.mkdir(parents=True)
    (pkg / "__init__.py").write_text("")
    assert "_composition.py" in dropin.install(pkg)
    for name in [m for m in sys.modules if m == "magpurify2" or m.startswith("magpurify2.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(str(tmp_path / "site"))
    importlib.invalidate_caches()
    comp = importlib.import_module("magpurify2._composition")
    cov = importlib.import_module("magpurify2._coverage")
    d = synth.make_mags(1, seed=3, mean_bp=5e4, n_contigs=12)
    seqs = [d["bases"][d["offsets"][i]:d["offsets"][i + 1]].tobytes().decode() for i in range(len(d["lengths"]))]
    assert np.array_equal(comp.get_tnf(seqs, False, 2), oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False))
    assert np.array_equal(comp.get_tnf(seqs), oracle.get_tnf_concat(d["bases"], d["offsets"], relative=True))
    assert np.array_equal(comp.get_gc(seqs), oracle.get_gc_concat(d["bases"], d["offsets"]), equal_nan=True)
    assert cov.get_bam_coverages.__module__ == "magpurify2_b200._coverage"
    with pytest.raises(NotImplementedError):
        importlib.import_module("magpurify2._codon").get_cai()


def test_shards_on_one_gpu_equal_unsharded(tmp_path, oracle):
    """the MAGs LPT-split into 3 shards, every shard through the kernels (three host
    threads on the same GPU), rows gathered back -- bit-identical to the unsharded run; the same for BAM files
    dealt out to "devices"."""
    from magpurify2_b200 import multi, tools
    from magpurify2_b200.core import read_mags

    d, names, genomes, bams, events = _write_dataset(tmp_path, n_mags=7, n_samples=3, seed=31)
    mags = read_mags(genomes, threads=2)
    one = multi.composition_of(mags, 0)
    parts = multi.shard_mags(mags, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(7)) and all(len(p) for p in parts)
    three = multi.composition_on_devices(mags, [0, 0, 0])
    for k in ("counts", "rel", "gc", "tnf_score", "gc_score"):
        assert np.array_equal(one[k], three[k], equal_nan=True), k
    assert np.array_equal(one["counts"], oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False, threads=4))
    n1, m1 = tools.get_bam_coverages([str(b) for b in bams], trim_lower=0.05, trim_upper=0.05, threads=4)
    n2, m2 = multi.bam_coverages_on_devices([str(b) for b in bams], [0, 0], trim_lower=0.05, trim_upper=0.05, threads=4)
    assert n1 == n2 and np.array_equal(m1.view(np.uint32), m2.view(np.uint32))


def test_cli_devices_cache_and_coverage_file(tmp_path, oracle):
    """--devices; tnfs.pickle.gz is reused when it holds every genome of the run (extra genomes survive) and
    replaced otherwise (modules/composition.py:48-60); --coverage_file (modules/coverage.py:101-106)."""
    from magpurify2_b200 import cli

    d, names, genomes, bams, events = _write_dataset(tmp_path, n_mags=4, n_samples=2, seed=17)
    out = tmp_path / "out"
    assert cli.cli(["composition", *map(str, genomes), str(out), "-q", "--devices", "0"]) == 0
    first = (out / "scores" / "composition_scores.tsv").read_text()
    with gzip.open(out / "tnfs.pickle.gz") as f:
        assert sorted(pickle.load(f)) == ["mag0", "mag1", "mag2", "mag3"]
    # a run on a subset finds everything it needs in the cache: the file (all four genomes) stays
    assert cli.cli(["composition", *map(str, genomes[:2]), str(out), "-q"]) == 0
    with gzip.open(out / "tnfs.pickle.gz") as f:
        assert sorted(pickle.load(f)) == ["mag0", "mag1", "mag2", "mag3"]
    # a cache that lacks a genome is thrown away and rewritten for this run's genomes
    with gzip.open(out / "tnfs.pickle.gz", "wb") as f:
        pickle.dump({"mag0": np.zeros((1, 136), np.float32)}, f)
    assert cli.cli(["composition", *map(str, genomes), str(out), "-q"]) == 0
    with gzip.open(out / "tnfs.pickle.gz") as f:
        tn = pickle.load(f)
    assert sorted(tn) == ["mag0", "mag1", "mag2", "mag3"] and tn["mag0"].shape[0] == d["mag_off"][1]
    assert (out / "scores" / "composition_scores.tsv").read_text() == first
    assert "magpurify2_b200" in (out / "scores" / "provenance.json").read_text()
    with pytest.raises(ValueError):
        cli.cli(["composition", *map(str, genomes), str(out), "-q", "--devices", "0,99"])
    # ---- --coverage_file: float64 table, one line per contig, shuffled, with a contig no MAG has ----
    rng = np.random.default_rng(17)
    cov = np.round(rng.gamma(2.0, 8.0, (len(names), 4)), 3)
    cov[:, 3] *= 0.01  # a sample below min_average_coverage
    table = tmp_path / "cov.tsv"
    order = rng.permutation(len(names))
    lines = [names[i] + "\t" + "\t".join(repr(float(v)) for v in cov[i]) for i in order] + ["stray_contig\t1\t2\t3\t4"]
    table.write_text("\n".join(lines) + "\n" + "x" * 130000 + "\t0\t0\t0\t0\n")  # >= 128000 bytes (tools.py:150-166)
    out2 = tmp_path / "out2"
    assert cli.cli(["coverage", *map(str, genomes), str(out2), "--coverage_file", str(table), "-q"]) == 0
    rows = [l.rstrip("\n").split("\t") for l in open(out2 / "scores" / "coverage_scores.tsv")]
    assert [r[1] for r in rows[1:]] == names
    for m in range(4):
        sl = slice(d["mag_off"][m], d["mag_off"][m + 1])
        c32 = cov[sl].astype(np.float32)
        sel, _ = oracle.select_samples(c32, d["lengths"][sl], 1.0)
        assert sel.tolist() == [True, True, True, False]
        want, _ = oracle.log_ratio_scores(c32[:, sel], d["lengths"][sl], 25)
        got = np.array([float(r[2]) for r in rows[1:][sl]])
        assert np.allclose(got, np.round(want, 5), atol=1.1e-5)
        assert {int(r[4]) for r in rows[1:][sl]} == {3}


def test_composition_object_accepts_the_relative_profile(tmp_path, oracle):
    """Reference-shaped call: Composition(mag, {genome: get_tnf(mag.sequences)}, get_gc(mag.sequences), ...)."""
    from magpurify2_b200 import tools
    from magpurify2_b200.core import Composition, Mag

    d, names, genomes, _b, _e = _write_dataset(tmp_path, n_mags=1, n_samples=0, seed=5)
    mag = Mag(genomes[0])
    a = Composition(mag, {mag.genome: tools.get_tnf(mag.sequences)}, tools.get_gc(mag.sequences), 4, 3, 0.1, 15, 1.0)
    b = Composition(mag, {mag.genome: tools.get_tnf(mag.sequences, relative=False)}, tools.get_gc(mag.sequences))
    assert np.array_equal(a.tnf_scores, b.tnf_scores) and np.array_equal(a.gc_content_scores, b.gc_content_scores)
