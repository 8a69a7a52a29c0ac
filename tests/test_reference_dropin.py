"""The reference's OWN drivers, unchanged, on top of the drop-in modules ("magpurify2 composition
and magpurify2 coverage run unchanged").

A throw-away `magpurify2` package is assembled from symlinks to the reference's Python sources
(/root/reference/magpurify2/*.py -- read at test time, never copied into this repository) plus the alias
modules written by magpurify2_b200.dropin.install(); the third-party packages the reference imports and this
container lacks (Bio, umap, hdbscan, taxopy, xgboost, biolib) are stubbed in sys.modules.  Then
magpurify2.modules.composition.main / coverage.main run on a small FASTA + BAM dataset.

This container has no GPU and the GPU box has no /root/reference, so the single native call behind each shim
(_composition._run -> mp2_tnf, _coverage.cov_from_events -> mp2_cov_events) is answered by the CPU oracle here:
what is under test is the boundary -- module paths, signatures, positional/keyword use, dtypes, row order, the
cache files and the TSV schema as the reference's code consumes them.  tests/test_pipeline_gpu.py checks the
same shims against the oracle with the real kernels (test_dropin_modules_on_gpu drives the alias modules)."""
import gzip
import importlib
import importlib.metadata
import io
import os
import pickle
import struct
import sys
import types
import zlib
from argparse import Namespace
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/magpurify2")
pytestmark = pytest.mark.skipif(not REF.is_dir(), reason="the reference tree is only present in the build container")


def _stub_third_party(monkeypatch):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    class _Rec:
        def __init__(self, header, seq):
            self.id = header.split()[0] if header.split() else ""
            self.description = header
            self.seq = seq

    def parse(handle, fmt):  # Bio.SeqIO.parse(handle, "fasta")
        assert fmt == "fasta"
        header, chunks = None, []
        for line in handle:
            line = line.rstrip("\n\r")
            if line.startswith(">"):
                if header is not None:
                    yield _Rec(header, "".join(chunks))
                header, chunks = line[1:], []
            elif header is not None:
                chunks.append(line.strip())
        if header is not None:
            yield _Rec(header, "".join(chunks))

    class BgzfReader:  # Bio.bgzf.BgzfReader(path, "rb"): enough for tools.is_sorted_bam's readline()
        def __init__(self, path, mode="rb"):
            raw = open(path, "rb").read(1 << 16)
            xlen = struct.unpack_from("<H", raw, 10)[0]
            self._buf = io.BytesIO(zlib.decompressobj(-15).decompress(raw[12 + xlen:]))

        def readline(self):
            return self._buf.readline()

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    seqio = mod("Bio.SeqIO", parse=parse)
    bgzf = mod("Bio.bgzf", BgzfReader=BgzfReader)
    mod("Bio", SeqIO=seqio, bgzf=bgzf)
    for name in ("hdbscan", "umap", "taxopy", "xgboost", "imp"):
        mod(name)
    mod("biolib")
    mod("biolib.logger", logger_setup=lambda *a, **k: None)
    mod("biolib.external")
    mod("biolib.external.prodigal", Prodigal=object)
    real_version = importlib.metadata.version
    monkeypatch.setattr(importlib.metadata, "version", lambda n: "0+reference" if n == "magpurify2" else real_version(n))


@pytest.fixture
def reference_package(tmp_path, monkeypatch, oracle):
    """The reference's sources + the installed drop-in, importable as `magpurify2`; native calls -> oracle."""
    from magpurify2_b200 import _composition, _coverage, dropin

    site = tmp_path / "site"
    pkg = site / "magpurify2"
    pkg.mkdir(parents=True)
    for entry in REF.iterdir():
        if entry.name != "__pycache__":
            os.symlink(entry, pkg / entry.name)
    written = dropin.install(pkg)
    assert sorted(written) == ["_codon.py", "_composition.py", "_coverage.py", "_mmseqs2.py"]
    _stub_third_party(monkeypatch)
    for name in [m for m in sys.modules if m == "magpurify2" or m.startswith("magpurify2.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(str(site))

    def fake_run(bases, offsets, counts, rel, gc, device=-1):
        if counts is not None:
            counts[...] = oracle.get_tnf_concat(bases, offsets, relative=False, threads=2)
        if rel is not None:
            rel[...] = oracle.get_tnf_concat(bases, offsets, relative=True, threads=2)
        if gc is not None:
            gc[...] = oracle.get_gc_concat(bases, offsets)

    def fake_cov(starts, ends, ev_off, contig_off, contig_end_exclusion=75, trim_lower=0.0, trim_upper=0.0,
                 want_stats=False, device=-1, max_span=0, max_disorder=0):
        cov, stats = oracle.cov_sample(np.asarray(starts), np.asarray(ends), np.asarray(ev_off),
                                       np.diff(np.asarray(contig_off)), contig_end_exclusion, trim_lower, trim_upper)
        return (cov, stats) if want_stats else cov

    monkeypatch.setattr(_composition, "_run", fake_run)
    monkeypatch.setattr(_coverage, "cov_from_events", fake_cov)
    importlib.invalidate_caches()
    pkg_mod = importlib.import_module("magpurify2")
    tools = importlib.import_module("magpurify2.tools")
    assert Path(tools.__file__).resolve() == (REF / "tools.py").resolve()  # the reference's file, not ours
    # UMAP + HDBSCAN scoring is outside this build (and its libraries are absent): neutral value
    monkeypatch.setattr(tools, "get_cluster_score_from_embedding", lambda data, lengths, **kw: np.ones(len(lengths)))
    return pkg_mod


def _dataset(tmp, oracle, seed=21, n_mags=3, n_samples=3):
    from magpurify2_b200 import bamio, synth

    d = synth.make_mags(n_mags, seed=seed, mean_bp=1.2e5, n_contigs=25)
    names, genomes = [], []
    for m in range(n_mags):
        path = tmp / (f"mag{m}.fna" if m % 2 else f"mag{m}.fna.gz")
        with (open if m % 2 else gzip.open)(path, "wt") as f:
            for c in range(d["mag_off"][m], d["mag_off"][m + 1]):
                nm = f"m{m}~c{c}" if c % 5 == 0 else f"m{m}_c{c}"  # '~' becomes '_' (tools.py:239)
                names.append(nm.replace("~", "_"))
                seq = d["bases"][d["offsets"][c]:d["offsets"][c + 1]].tobytes().decode()
                f.write(f">{nm} len={len(seq)}\n")
                f.writelines(seq[o:o + 70] + "\n" for o in range(0, len(seq), 70))
        genomes.append(path)
    rng = np.random.default_rng(seed)
    bams, cols = [], []
    for s in range(n_samples):
        st, en, ev_off = synth.make_events(d["lengths"], d["mag_off"], s, seed=seed, depth_scale=1.5)
        tid = np.repeat(np.arange(len(d["lengths"])), np.diff(ev_off))
        nmv = rng.binomial(150, 0.012, len(st))
        p = tmp / f"sample{s}.bam"
        bamio.write_bam_fixed(str(p), names, d["lengths"], tid, st, nmv)
        keep = (np.float32(1.0) - nmv.astype(np.float32) / np.float32(150)) >= np.float32(0.97)
        k_off = np.concatenate([[0], np.cumsum(np.bincount(tid[keep], minlength=len(d["lengths"])))])
        cols.append(oracle.cov_sample(st[keep], en[keep], k_off, d["lengths"], 75, 0.05, 0.05)[0])
    return d, names, genomes, bams + [tmp / f"sample{s}.bam" for s in range(n_samples)], np.stack(cols, axis=1)


def _args(genomes, out, **kw):
    base = dict(genomes=genomes, output_directory=out, threads=1, n_iterations=4, n_components=3, min_dist=0.1,
                n_neighbors=15, set_op_mix_ratio=1.0)
    base.update(kw)
    return Namespace(**base)


def test_reference_drivers_run_unchanged_on_the_dropin(tmp_path, oracle, reference_package):
    mp = reference_package
    d, names, genomes, bams, cov_want = _dataset(tmp_path, oracle)
    out = tmp_path / "out"
    out.mkdir()
    # ---- magpurify2 composition (reference modules/composition.py:31-89) ----
    mp.modules.composition.main(_args(genomes, out))
    rows = [l.rstrip("\n").split("\t") for l in open(out / "scores" / "composition_scores.tsv")]
    assert rows[0] == ["genome", "contig", "tnf_score", "gc_content_score", "contig_length"]
    assert [r[1] for r in rows[1:]] == names and [int(r[4]) for r in rows[1:]] == d["lengths"].tolist()
    gc = oracle.get_gc_concat(d["bases"], d["offsets"])
    for m in range(len(genomes)):
        sl = slice(d["mag_off"][m], d["mag_off"][m + 1])
        want, _ = oracle.log_ratio_scores(gc[sl], d["lengths"][sl], 2)
        got = np.array([float(r[3]) for r in rows[1:][sl]])
        assert np.allclose(got, np.round(want, 5), atol=1.1e-5)
        assert {r[0] for r in rows[1:][sl]} == {f"mag{m}"}
    with gzip.open(out / "tnfs.pickle.gz") as f:
        tn = pickle.load(f)
    assert sorted(tn) == [f"mag{m}" for m in range(len(genomes))]
    got_tnf = np.concatenate([tn[f"mag{m}"] for m in range(len(genomes))])
    assert got_tnf.dtype == np.float32
    assert np.array_equal(got_tnf, oracle.get_tnf_concat(d["bases"], d["offsets"], relative=True))
    # ---- magpurify2 coverage (reference modules/coverage.py:33-137) ----
    mp.modules.coverage.main(_args(genomes, out, bam_files=bams, coverage_file=None, min_identity=0.97, trim_lower=0.05,
                                   trim_upper=0.05, min_average_coverage=1.0, min_dist=0.15, set_op_mix_ratio=0.6))
    rows = [l.rstrip("\n").split("\t") for l in open(out / "scores" / "coverage_scores.tsv")]
    assert rows[0] == ["genome", "contig", "coverage_score", "coverage_cluster_score", "n_samples"]
    assert [r[1] for r in rows[1:]] == names
    for m in range(len(genomes)):
        sl = slice(d["mag_off"][m], d["mag_off"][m + 1])
        sel, _ = oracle.select_samples(cov_want[sl], d["lengths"][sl], 1.0)
        want = np.ones(sl.stop - sl.start)
        if sel.any() and sl.stop - sl.start > 2:
            want, _ = oracle.log_ratio_scores(cov_want[sl][:, sel], d["lengths"][sl], 25)
        got = np.array([float(r[2]) for r in rows[1:][sl]])
        assert np.allclose(got, np.round(want, 5), atol=1.1e-5)
        assert {int(r[4]) for r in rows[1:][sl]} == {int(sel.sum())}
    with gzip.open(out / "coverages.pickle.gz") as f:
        sig, cnames, mat = pickle.load(f)
    assert len(sig) == len(bams) and list(cnames) == names
    assert mat.dtype == np.float32 and np.array_equal(mat, cov_want)


def test_dropin_refuses_to_shadow_a_compiled_extension(tmp_path):
    from magpurify2_b200 import dropin

    pkg = tmp_path / "magpurify2"
    pkg.mkdir()
    (pkg / "_composition.cpython-312-x86_64-linux-gnu.so").write_bytes(b"")
    with pytest.raises(FileExistsError):
        dropin.install(pkg)
    (pkg / "_composition.cpython-312-x86_64-linux-gnu.so").unlink()
    (pkg / "_codon.cpython-312-x86_64-linux-gnu.so").write_bytes(b"")  # a built _codon is kept
    assert sorted(dropin.install(pkg)) == ["_composition.py", "_coverage.py", "_mmseqs2.py"]
    assert "magpurify2_b200._coverage" in (pkg / "_coverage.py").read_text()
