"""Host logic of the module drivers (no GPU): option clamping, the result caches, the tabular coverage reader,
the TSV writer, the provenance note `filter` warns from, and the MAG sharding used by --devices."""
import gzip
import logging
import pickle
from argparse import Namespace

import numpy as np
import pytest

from magpurify2_b200 import multi, sharding, tools
from magpurify2_b200.modules import _common
from magpurify2_b200.modules import filter as filter_module


def test_clamp_options_and_validade_input(caplog):
    args = Namespace(n_iterations=0, n_components=5, n_neighbors=5000, set_op_mix_ratio=1.5, other=7)
    with caplog.at_level(logging.WARNING, logger="timestamp"):
        _common.clamp_options(args, _common.EMBEDDING_LIMITS)
    assert (args.n_iterations, args.n_components, args.n_neighbors, args.set_op_mix_ratio, args.other) == (1, 5, 999, 1.0, 7)
    assert sum("must be between" in r.message for r in caplog.records) == 3
    assert tools.validade_input(0.5, "x", [1.0, 0.0]) == 0.5 and tools.validade_input(-3, "x", [0, 9]) == 0


def test_pickle_cache_reuse_and_replacement(tmp_path):
    cache = _common.PickleCache(tmp_path / "tnfs.pickle.gz", "tetranucleotide frequencies")
    assert cache.load(lambda d: True) is None
    cache.store({"a": 1, "b": 2})
    assert cache.load(lambda d: {"a"}.issubset(d)) == {"a": 1, "b": 2}       # covers the request: reused as it is
    assert cache.load(lambda d: {"a", "c"}.issubset(d)) is None               # does not: deleted
    assert not (tmp_path / "tnfs.pickle.gz").exists()
    (tmp_path / "tnfs.pickle.gz").write_bytes(b"not a gzip file")
    assert cache.load(lambda d: True) is None and not (tmp_path / "tnfs.pickle.gz").exists()
    cache.store(("sig", np.arange(3)))
    with gzip.open(tmp_path / "tnfs.pickle.gz") as f:
        assert pickle.load(f)[0] == "sig"   # a plain gzip pickle: the reference can read it


def test_get_tsv_coverages_matches_genfromtxt(tmp_path):
    p = tmp_path / "cov.tsv"
    p.write_text("c1\t1.5\t2\nc#2\t0\t1e3\nc3\t7\tnan\n")
    names, mat = tools.get_tsv_coverages(p)
    want = np.genfromtxt(p, dtype=float, comments=None, delimiter="\t", usecols=[1, 2])
    assert names.tolist() == ["c1", "c#2", "c3"] and mat.dtype == np.float64
    assert np.array_equal(mat, want, equal_nan=True)
    n2, m2 = tools.get_tsv_coverages(p, contig_set={"c3", "c1", "zz"})
    assert n2.tolist() == ["c1", "c3"] and np.array_equal(m2, want[[0, 2]], equal_nan=True)
    one = tmp_path / "one.tsv"
    one.write_text("only\t4.25\n")
    n3, m3 = tools.get_tsv_coverages(one)
    assert n3.tolist() == ["only"] and m3.shape == (1, 1) and m3[0, 0] == 4.25
    bad = tmp_path / "bad.tsv"
    bad.write_text("just_a_name\n")
    with pytest.raises(SystemExit):
        tools.get_tsv_coverages(bad)


def test_write_module_output_format(tmp_path):
    class Scores:
        attributes = ["genome", "contig", "score"]

        def __init__(self, rows):
            self.rows = rows

        def __iter__(self):
            return iter(self.rows)

    out = tmp_path / "x.tsv"
    tools.write_module_output([Scores([("g", "c1", np.float64(0.5))]), Scores([("h", "c2", 1)])], out)
    assert out.read_text() == "genome\tcontig\tscore\ng\tc1\t0.5\nh\tc2\t1\n"


def test_provenance_note_and_filter_warning(tmp_path, caplog):
    args = Namespace(output_directory=tmp_path)
    (tmp_path / "scores").mkdir()
    log = logging.getLogger("timestamp")
    assert filter_module.warn_uncalibrated(tmp_path / "scores", log) == []   # tables of unknown origin: silent
    _common.note_provenance(args, "composition")
    _common.note_provenance(args, "coverage")
    with caplog.at_level(logging.WARNING, logger="timestamp"):
        assert filter_module.warn_uncalibrated(tmp_path / "scores", log) == ["composition", "coverage"]
    assert any("NOT calibrated" in r.message for r in caplog.records)


def test_shard_and_gather_round_trip():
    class M:
        def __init__(self, n, bp):
            self.offsets = np.array([0, bp])
            self.n = n

        def __len__(self):
            return self.n

    rng = np.random.default_rng(2)
    mags = [M(int(rng.integers(1, 9)), int(rng.integers(1000, 90000))) for _ in range(11)]
    parts = multi.shard_mags(mags, 4)
    assert sorted(np.concatenate(parts).tolist()) == list(range(11))
    loads = [sum(int(mags[i].offsets[-1]) for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(int(m.offsets[-1]) for m in mags)   # LPT bound
    first = np.concatenate([[0], np.cumsum([len(m) for m in mags])])
    truth = np.arange(first[-1] * 3, dtype=np.float32).reshape(-1, 3)
    shards = [dict(x=np.concatenate([truth[first[i]:first[i + 1]] for i in p.tolist()])) if len(p) else None for p in parts]
    assert np.array_equal(multi.gather_shards(mags, parts, shards)["x"], truth)
    assert [p.tolist() for p in sharding.lpt_partition([5, 1, 1, 1, 1, 1], 2)] == [[0], [1, 2, 3, 4, 5]]
