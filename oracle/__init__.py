"""ctypes front-end of the CPU oracle (oracle/mp2_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (magpurify2_b200) never imports it.

Parity status: composition and scoring are pinned (reference LUT + reference NumPy functions,
see tests/golden/); the coverage arithmetic lives in the un-vendored crate coverm 0.3.2 and is
"parity unpinned" (restated from SURVEY.md Appendix A).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmp2_oracle.so")
NCANON = 136


def build(force=False):
    src = os.path.join(_HERE, "mp2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libmp2_oracle.so"])
    return _SO


class CovStats(C.Structure):
    _fields_ = [
        ("observed_len", C.c_uint64),
        ("covered", C.c_uint64),
        ("min_index", C.c_uint64),
        ("max_index", C.c_uint64),
        ("total", C.c_uint64),
        ("max_depth", C.c_uint64),
        ("coverage", C.c_float),
    ]


COV_STATS_DTYPE = np.dtype(
    [
        ("observed_len", "<u8"),
        ("covered", "<u8"),
        ("min_index", "<u8"),
        ("max_index", "<u8"),
        ("total", "<u8"),
        ("max_depth", "<u8"),
        ("coverage", "<f4"),
        ("_pad", "<u4"),
    ]
)
assert COV_STATS_DTYPE.itemsize == C.sizeof(CovStats)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.mp2o_rowsum_f32.restype = C.c_float
        _lib.mp2o_weighted_median.restype = C.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def concat(seqs):
    """list[str|bytes] -> (uint8[L], int64[n+1])."""
    bs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return bases, offsets


def canonical_lut():
    out = np.zeros(256, dtype=np.uint8)
    lib().mp2o_canonical_lut(_p(out, C.c_uint8))
    return out


def get_tnf_concat(bases, offsets, relative=True, threads=1):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    counts = np.zeros((n, NCANON), dtype=np.uint32)
    rel = np.zeros((n, NCANON), dtype=np.float32)
    lib().mp2o_get_tnf(
        _p(bases, C.c_uint8), _p(offsets, C.c_int64), C.c_int64(n), C.c_int(threads),
        _p(counts, C.c_uint32), _p(rel, C.c_float) if relative else None,
    )
    return rel if relative else counts


def get_tnf(seq_list, relative=True, threads=1):
    """Same call shape as magpurify2._composition.get_tnf (src/composition.rs:94-95)."""
    bases, offsets = concat(seq_list)
    return get_tnf_concat(bases, offsets, relative, threads)


def get_gc_concat(bases, offsets, passes5=False):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    gc = np.zeros(n, dtype=np.float32)
    lib().mp2o_get_gc(_p(bases, C.c_uint8), _p(offsets, C.c_int64), C.c_int64(n),
                      C.c_int(1 if passes5 else 0), _p(gc, C.c_float))
    return gc


def get_gc(seq_list):
    """Same call shape as magpurify2._composition.get_gc (src/composition.rs:148-149)."""
    bases, offsets = concat(seq_list)
    return get_gc_concat(bases, offsets)


def trimmed_mean_from_hist(counts, N, covered, trim_lo, trim_hi):
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    st = CovStats()
    lib().mp2o_trimmed_mean_from_hist(
        _p(counts, C.c_uint32), C.c_uint64(len(counts)), C.c_uint64(N), C.c_uint64(covered),
        C.c_float(trim_lo), C.c_float(trim_hi), C.byref(st),
    )
    return st


def cov_contig(starts, ends, length, excl=75, trim_lo=0.0, trim_upper=0.0):
    starts = np.ascontiguousarray(starts, dtype=np.uint32)
    ends = np.ascontiguousarray(ends, dtype=np.uint32)
    st = CovStats()
    hi = np.float32(1.0) - np.float32(trim_upper)
    rc = lib().mp2o_cov_contig(
        _p(starts, C.c_uint32), _p(ends, C.c_uint32), C.c_int64(len(starts)),
        C.c_uint64(length), C.c_uint32(excl), C.c_float(trim_lo), C.c_float(hi), C.byref(st),
    )
    if rc:
        raise MemoryError("oracle allocation failed")
    return st


def cov_sample(starts, ends, ev_off, contig_len, excl=75, trim_lo=0.0, trim_upper=0.0, threads=1):
    """One sample: events grouped by contig. Returns (cov f32[n], stats recarray[n])."""
    starts = np.ascontiguousarray(starts, dtype=np.uint32)
    ends = np.ascontiguousarray(ends, dtype=np.uint32)
    ev_off = np.ascontiguousarray(ev_off, dtype=np.int64)
    contig_len = np.ascontiguousarray(contig_len, dtype=np.int64)
    n = len(contig_len)
    cov = np.zeros(n, dtype=np.float32)
    stats = np.zeros(n, dtype=COV_STATS_DTYPE)
    rc = lib().mp2o_cov_sample(
        _p(starts, C.c_uint32), _p(ends, C.c_uint32), _p(ev_off, C.c_int64),
        _p(contig_len, C.c_int64), C.c_int64(n), C.c_uint32(excl), C.c_float(trim_lo),
        C.c_float(trim_upper), C.c_int(threads), _p(cov, C.c_float),
        stats.ctypes.data_as(C.c_void_p),
    )
    if rc:
        raise MemoryError("oracle allocation failed")
    return cov, stats


def record_events(flag, tid, pos, cigar, nm, min_identity, max_out=64):
    cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
    s = np.zeros(max_out, dtype=np.uint32)
    e = np.zeros(max_out, dtype=np.uint32)
    n = lib().mp2o_record_events(
        C.c_uint16(flag), C.c_int32(tid), C.c_int32(pos), _p(cigar, C.c_uint32),
        C.c_int(len(cigar)), C.c_int64(-1 if nm is None else nm), C.c_float(min_identity),
        _p(s, C.c_uint32), _p(e, C.c_uint32), C.c_int(max_out),
    )
    if n < 0:
        raise KeyError("NM tag missing")
    return list(zip(s[:n].tolist(), e[:n].tolist()))


def weighted_median(data, weights):
    data = np.ascontiguousarray(data, dtype=np.float32)
    weights = np.ascontiguousarray(weights, dtype=np.int64)
    return np.float32(lib().mp2o_weighted_median(_p(data, C.c_float), _p(weights, C.c_int64),
                                                 C.c_int64(len(data))))


def log_ratio_scores(data, weights, base):
    data = np.ascontiguousarray(data, dtype=np.float32)
    weights = np.ascontiguousarray(weights, dtype=np.int64)
    d2 = data.reshape(len(data), -1)
    n, s = d2.shape
    scores = np.zeros(n, dtype=np.float64)
    med = np.zeros(s, dtype=np.float32)
    lib().mp2o_log_ratio_scores(_p(d2, C.c_float), _p(weights, C.c_int64), C.c_int64(n),
                                C.c_int64(s), C.c_double(base), _p(scores, C.c_double),
                                _p(med, C.c_float))
    return scores, med


def select_samples(cov, lengths, min_average_coverage):
    cov = np.ascontiguousarray(cov, dtype=np.float32)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    n, s = cov.shape
    sel = np.zeros(s, dtype=np.uint8)
    means = np.zeros(s, dtype=np.float64)
    lib().mp2o_select_samples(_p(cov, C.c_float), _p(lengths, C.c_int64), C.c_int64(n),
                              C.c_int64(s), C.c_double(min_average_coverage),
                              _p(sel, C.c_uint8), _p(means, C.c_double))
    return sel.astype(bool), means


def clr_centroid(counts, lengths):
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    n = counts.shape[0]
    clr = np.zeros((n, NCANON), dtype=np.float32)
    dist = np.zeros(n, dtype=np.float64)
    lib().mp2o_clr_centroid(_p(counts, C.c_uint32), _p(lengths, C.c_int64), C.c_int64(n),
                            _p(clr, C.c_float), _p(dist, C.c_double))
    return clr, dist


def max_threads():
    return int(lib().mp2o_max_threads())
