"""Drop-in for the reference's PyO3 module `magpurify2._coverage`
(/root/reference/src/coverage.rs:208-212): same function name, arguments and results, with the
depth arithmetic done by the sm_100a kernels in libmp2.so and BAM decoding by libmp2's own
BGZF/BAM reader (htslib is not used).

    get_bam_coverages(bam_list, contig_end_exclusion=75, min_identity=0.97, trim_lower=0.0,
                      trim_upper=0.0, contig_set=None, threads=1)
        -> (list[str] names, float32[n_contigs, n_bams])                    (coverage.rs:81-98)
"""
import ctypes as C

import numpy as np

from . import _lib

COV_STATS_DTYPE = _lib.COV_STATS_DTYPE


def cov_from_events(starts, ends, ev_off, contig_off, contig_end_exclusion=75, trim_lower=0.0,
                    trim_upper=0.0, want_stats=False, device=-1, max_span=0, max_disorder=0):
    """One sample on host arrays: alignment blocks grouped by contig -> trimmed-mean coverage.

    starts/ends: uint32[n_events] (0-based start, exclusive end, relative to the contig);
    ev_off: int64[n_contigs+1]; contig_off: int64[n_contigs+1] cumulative contig lengths.
    max_span > 0 vouches for the order of the blocks: end - start <= max_span, and no start more than
    max_disorder below the largest start before it in its contig (0 = sorted; a BAM with D / N
    operations is not: BamEvents reports both numbers).  max_span = 0 lets the library measure both
    on the device.
    Returns float32[n_contigs] (and the per-contig integer statistics when want_stats)."""
    starts = np.ascontiguousarray(starts, dtype=np.uint32)
    ends = np.ascontiguousarray(ends, dtype=np.uint32)
    ev_off = np.ascontiguousarray(ev_off, dtype=np.int64)
    contig_off = np.ascontiguousarray(contig_off, dtype=np.int64)
    n = len(contig_off) - 1
    if len(ev_off) != n + 1:
        raise ValueError("ev_off and contig_off must have the same length")
    cov = np.empty(n, dtype=np.float32)
    stats = np.zeros(n, dtype=COV_STATS_DTYPE) if want_stats else None
    lib = _lib.load()
    _lib.check(lib.mp2_cov_events(_lib.ptr(starts), _lib.ptr(ends), _lib.ptr(ev_off), _lib.ptr(contig_off),
                                  n, len(starts), int(max_span), int(max_disorder), int(contig_end_exclusion), float(trim_lower),
                                  float(trim_upper), device, _lib.ptr(cov), 1, _lib.ptr(stats)))
    return (cov, stats) if want_stats else cov


def get_bam_coverages(bam_list, contig_end_exclusion=75, min_identity=0.97, trim_lower=0.0, trim_upper=0.0,
                      contig_set=None, threads=1, *, device=-1, stats_out=None):
    """Trimmed-mean coverage of every contig in every BAM (src/coverage.rs:81-206).

    Returns (names, float32[n_contigs, n_bams]); columns follow bam_list.  Rows follow the @SQ
    order of the first BAM (the reference returns them in Rust HashMap order, i.e. arbitrary:
    callers index by name, magpurify2/modules/coverage.py:113-119)."""
    from . import _bam

    if isinstance(bam_list, (str, bytes)):
        raise TypeError("bam_list must be a list of paths")
    bam_list = [str(b) for b in bam_list]
    if not bam_list:
        raise ValueError("bam_list is empty")
    names = None
    cols = []
    for path in bam_list:
        with _bam.BamEvents(path, min_identity=min_identity, threads=threads) as b:
            if names is None:
                names, lengths = b.names, b.lengths
                contig_off = np.zeros(len(lengths) + 1, dtype=np.int64)
                np.cumsum(lengths, out=contig_off[1:])
            elif b.names != names:
                raise ValueError(f"{path}: @SQ lines differ from {bam_list[0]} (the reference assumes identical headers)")
            # the reader emits blocks in BAM order and measures, while it decodes, the longest block and how
            # far a block of a read with D / N operations can start behind the next reads' first blocks
            res = cov_from_events(b.starts, b.ends, b.ev_off, contig_off, contig_end_exclusion,
                                  trim_lower, trim_upper, device=device, max_span=b.max_block,
                                  max_disorder=b.max_disorder, want_stats=stats_out is not None)
            if stats_out is not None:  # per-contig integers of every BAM (tests: which path computed them)
                stats_out.append(res[1])
                res = res[0]
            cols.append(res)
    if not names:
        raise ValueError("no reference sequences in the BAM header")  # the reference panics (coverage.rs:200)
    mat = np.stack(cols, axis=1).astype(np.float32, copy=False)
    if contig_set is not None:
        keep = np.fromiter((nm in contig_set for nm in names), dtype=bool, count=len(names))
        names = [nm for nm, k in zip(names, keep) if k]
        mat = mat[keep]
        if not names:
            raise ValueError("contig_set matches no reference sequence")
    return names, np.ascontiguousarray(mat)
