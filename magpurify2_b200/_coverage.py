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
                      contig_set=None, threads=1, *, device=-1, devices=None, stats_out=None):
    """Trimmed-mean coverage of every contig in every BAM (src/coverage.rs:81-206).

    Returns (names, float32[n_contigs, n_bams]); columns follow bam_list.  Rows follow the @SQ
    order of the first BAM (the reference returns them in Rust HashMap order, i.e. arbitrary:
    callers index by name, magpurify2/modules/coverage.py:113-119).

    The BAMs are decoded on the host (`threads` inflate / filter workers, one file at a time) while the
    GPUs work on the files decoded before: file s+1 is being decoded while file s is uploaded and
    reduced (the reference reads the BAMs strictly one after the other, src/coverage.rs:125-142).
    `devices` (keyword only, not in the reference) spreads the files over several GPUs: samples are
    independent columns of the result, so there is no exchange between devices."""
    import threading
    from concurrent.futures import ThreadPoolExecutor

    from . import _bam

    if isinstance(bam_list, (str, bytes)):
        raise TypeError("bam_list must be a list of paths")
    bam_list = [str(b) for b in bam_list]
    if not bam_list:
        raise ValueError("bam_list is empty")
    devices = [device] if not devices else list(devices)
    n_bams = len(bam_list)
    header = {}                      # names / lengths / contig_off of the first file, set by whoever decodes it
    header_lock = threading.Lock()
    in_flight = threading.Semaphore(len(devices) + 1)  # decoded samples waiting for, or on, a GPU
    cols = [None] * n_bams
    stats = [None] * n_bams

    def decode(s):
        in_flight.acquire()
        try:
            b = _bam.BamEvents(bam_list[s], min_identity=min_identity, threads=threads)
        except BaseException:
            in_flight.release()
            raise
        with header_lock:
            if not header:
                off = np.zeros(len(b.lengths) + 1, dtype=np.int64)
                np.cumsum(b.lengths, out=off[1:])
                header.update(names=b.names, contig_off=off, first=bam_list[s])
            elif b.names != header["names"]:
                b.close()
                in_flight.release()
                raise ValueError(f"{bam_list[s]}: @SQ lines differ from {header['first']} (the reference assumes "
                                 "identical headers)")
        return b

    def reduce_on(dev, s, fut):
        b = fut.result()
        try:
            # the reader measured, while decoding, the longest block and how far a block of a read with D / N
            # operations can start behind the first blocks of the following reads
            res = cov_from_events(b.starts, b.ends, b.ev_off, header["contig_off"], contig_end_exclusion,
                                  trim_lower, trim_upper, device=dev, max_span=b.max_block,
                                  max_disorder=b.max_disorder, want_stats=stats_out is not None)
            if stats_out is not None:
                cols[s], stats[s] = res
            else:
                cols[s] = res
        finally:
            b.close()
            in_flight.release()

    with ThreadPoolExecutor(max_workers=1) as decoder:
        lanes = [ThreadPoolExecutor(max_workers=1) for _ in devices]
        try:
            decoded = [decoder.submit(decode, s) for s in range(n_bams)]
            done = [lanes[s % len(devices)].submit(reduce_on, devices[s % len(devices)], s, decoded[s])
                    for s in range(n_bams)]
            for f in done:
                f.result()
        finally:
            for ln in lanes:
                ln.shutdown(wait=True, cancel_futures=True)
    names = header.get("names")
    if not names:
        raise ValueError("no reference sequences in the BAM header")  # the reference panics (coverage.rs:200)
    if stats_out is not None:  # per-contig integers of every BAM (tests: which path computed them)
        stats_out.extend(stats)
    mat = np.stack(cols, axis=1).astype(np.float32, copy=False)
    if contig_set is not None:
        keep = np.fromiter((nm in contig_set for nm in names), dtype=bool, count=len(names))
        names = [nm for nm, k in zip(names, keep) if k]
        mat = mat[keep]
        if not names:
            raise ValueError("contig_set matches no reference sequence")
    return names, np.ascontiguousarray(mat)
