"""Drop-in for the reference's PyO3 module `magpurify2._composition`
(/root/reference/src/composition.rs:163-168): same function names, arguments and results, with the
compute done by the sm_100a kernels in libmp2.so.

    get_tnf(seq_list, relative=True, threads=1) -> float32[n,136] | uint32[n,136]   (:94-124)
    get_gc(seq_list)                            -> float32[n]                        (:148-161)
"""
import numpy as np

from . import _lib

NCANON = _lib.NCANON


def _concat(seq_list):
    """list[str] -> (uint8[L], int64[n+1]).

    The reference borrows the UTF-8 buffers of the Python strings (Vec<&str>); the device needs
    one contiguous byte stream, so this is the single host copy of the call.  A non-ASCII
    character becomes ONE '?' byte: like its UTF-8 bytes it only resets the k-mer window
    (src/composition.rs:54), and it keeps the GC denominator equal to chars().count() (:131)."""
    if isinstance(seq_list, (str, bytes)):
        raise TypeError("seq_list must be a list of str")
    enc = []
    for i, s in enumerate(seq_list):
        if isinstance(s, str):
            enc.append(s.encode("ascii") if s.isascii() else s.encode("ascii", "replace"))
        elif isinstance(s, (bytes, bytearray, memoryview)):
            enc.append(bytes(s))
        else:
            raise TypeError(f"seq_list[{i}] must be str, got {type(s).__name__}")
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum([len(b) for b in enc], out=offsets[1:])
    bases = np.frombuffer(b"".join(enc), dtype=np.uint8)
    return bases, offsets


def _out(shape, dtype, small_too=False):
    """Result buffer: page-locked when it is large (the D2H copy of a pageable 160 MB array costs
    more than the kernels), plain NumPy otherwise.  torch is used as the pinned allocator only."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes >= (8 << 20) or (small_too and nbytes > 0):
        try:
            import torch

            if torch.cuda.is_available():
                tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.uint32): torch.int32}[np.dtype(dtype)]
                return torch.empty(shape, dtype=tdt, pin_memory=True).numpy().view(dtype)
        except Exception:
            pass
    return np.empty(shape, dtype=dtype)


def pinned_empty(shape, dtype):
    """Page-locked result buffer to pass as out_* (reuse it across calls of a streaming driver)."""
    return _out(shape, dtype)


def _check_out(a, shape, dtype, name):
    if not (isinstance(a, np.ndarray) and a.shape == tuple(shape) and a.dtype == np.dtype(dtype) and a.flags.c_contiguous):
        raise ValueError(f"{name} must be a C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
    return a


def tnf_from_concat(bases, offsets, relative=True, want_gc=False, want_counts=False, device=-1,
                    out_counts=None, out_rel=None, out_gc=None):
    """Batch entry point on one concatenated byte stream (what the scalable drivers call).
    out_*: optional preallocated result arrays (NumPy `out=` style; page-locked ones from
    pinned_empty() avoid the staging copy of the device->host transfer)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    counts = rel = gc = None
    if want_counts or not relative or out_counts is not None:
        counts = _check_out(out_counts, (n, NCANON), np.uint32, "out_counts") if out_counts is not None \
            else _out((n, NCANON), np.uint32)
    if relative or out_rel is not None:
        rel = _check_out(out_rel, (n, NCANON), np.float32, "out_rel") if out_rel is not None \
            else _out((n, NCANON), np.float32)
    if want_gc or out_gc is not None:
        gc = _check_out(out_gc, (n,), np.float32, "out_gc") if out_gc is not None else np.empty(n, dtype=np.float32)
    _run(bases, offsets, counts, rel, gc, device)
    return counts, rel, gc


def _run(bases, offsets, counts, rel, gc, device=-1):
    """The one native call of this module: mp2_tnf on host arrays (any of counts / rel / gc may be None)."""
    _lib.check(_lib.load().mp2_tnf(_lib.ptr(bases), _lib.ptr(offsets), len(offsets) - 1, device,
                                   _lib.ptr(counts), _lib.ptr(rel), _lib.ptr(gc)))


def pack_parts(parts, threads=None):
    """2-bit packs the virtual concatenation of `parts` (uint8 arrays: the genomes of a run as the FASTA reader
    left them) on the host: (codes uint32[4q], valid uint32[2q]) for q quads of 64 bases, in page-locked
    memory when there is a GPU to copy them to.  0.375 bytes per base instead of 1."""
    import os

    parts = [np.ascontiguousarray(p, dtype=np.uint8) for p in parts]
    lens = np.array([len(p) for p in parts], dtype=np.int64)
    total = int(lens.sum())
    quads = (total + 63) // 64
    codes = _out((4 * quads,), np.uint32) if quads else np.zeros(0, np.uint32)
    valid = _out((2 * quads,), np.uint32) if quads else np.zeros(0, np.uint32)
    ptrs = (_lib.C.c_void_p * max(1, len(parts)))(*[p.ctypes.data for p in parts])
    rc = _lib.load().mp2_pack_host(ptrs, _lib.ptr(lens), len(parts), _lib.ptr(codes), _lib.ptr(valid),
                                   int(threads or os.cpu_count() or 1))
    if rc != _lib.OK:
        raise ValueError("mp2_pack_host: bad argument")
    return codes, valid


def invalid_runs(valid, n_bases, threads=None):
    """int64[k, 2] (begin, end) runs of invalid bases of a validity mask -- what tnf_from_packed uploads instead of
    the mask when it is the smaller of the two (it nearly always is: a genome is ACGT almost everywhere)."""
    import os

    valid = np.ascontiguousarray(valid, dtype=np.uint32)
    lib = _lib.load()
    th = int(threads or os.cpu_count() or 1)
    cap = 4096
    while True:
        runs = np.empty((cap, 2), dtype=np.int64)
        k = lib.mp2_invalid_runs(_lib.ptr(valid), int(n_bases), _lib.ptr(runs), cap, th)
        if k >= 0:
            return runs[:k]
        if k == _lib.EINVAL:
            raise ValueError("mp2_invalid_runs: bad argument")
        cap = -k


def tnf_from_packed(codes, valid, offsets, relative=True, want_gc=False, want_counts=False, device=-1,
                    out_counts=None, out_rel=None, out_gc=None, runs=None):
    """tnf_from_concat on the 2-bit packed stream of pack_parts (offsets in bases).  `runs` (invalid_runs(valid))
    is uploaded instead of the mask when given and smaller; pass valid=None to send only the runs."""
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    quads = (int(offsets[-1]) + 63) // 64 if n >= 0 and len(offsets) else 0
    if valid is None and runs is None:
        raise ValueError("give the validity mask, the runs of invalid bases, or both")
    if valid is not None:
        valid = np.ascontiguousarray(valid, dtype=np.uint32)
        if len(valid) < 2 * quads:
            raise ValueError("valid is shorter than offsets[-1] bases")
    if runs is not None:
        runs = np.ascontiguousarray(runs, dtype=np.int64).reshape(-1, 2)
        if valid is not None and runs.nbytes >= 8 * quads:  # the mask is the smaller upload
            runs = None
        else:
            valid = None
    if len(codes) < 4 * quads:
        raise ValueError("codes is shorter than offsets[-1] bases")
    counts = rel = gc = None
    if want_counts or not relative or out_counts is not None:
        counts = _check_out(out_counts, (n, NCANON), np.uint32, "out_counts") if out_counts is not None \
            else _out((n, NCANON), np.uint32)
    if relative or out_rel is not None:
        rel = _check_out(out_rel, (n, NCANON), np.float32, "out_rel") if out_rel is not None \
            else _out((n, NCANON), np.float32)
    if want_gc or out_gc is not None:
        gc = _check_out(out_gc, (n,), np.float32, "out_gc") if out_gc is not None else _out((n,), np.float32, small_too=n > 4096)
    _lib.check(_lib.load().mp2_tnf_packed(_lib.ptr(codes), _lib.ptr(valid), _lib.ptr(runs), 0 if runs is None else len(runs),
                                          _lib.ptr(offsets), n, device, _lib.ptr(counts), _lib.ptr(rel), _lib.ptr(gc)))
    return counts, rel, gc


def get_tnf(seq_list, relative=True, threads=1):
    """Canonical 4-mer profile of every sequence (rows in input order).

    `threads` is accepted for signature compatibility (rayon workers in the reference,
    src/composition.rs:96-99); the GPU path has no use for it."""
    bases, offsets = _concat(seq_list)
    counts, rel, _gc = tnf_from_concat(bases, offsets, relative=bool(relative))
    return rel if relative else counts


def get_gc(seq_list):
    """GC fraction of every sequence: (#G+#g+#C+#c) / number of characters; '' -> NaN."""
    bases, offsets = _concat(seq_list)
    n = len(offsets) - 1
    gc = np.empty(n, dtype=np.float32)
    _run(bases, offsets, None, None, gc)
    return gc
