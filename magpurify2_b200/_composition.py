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


def tnf_from_concat(bases, offsets, relative=True, want_gc=False, want_counts=False, device=-1):
    """Batch entry point on one concatenated byte stream (what the scalable drivers call)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    lib = _lib.load()
    counts = np.empty((n, NCANON), dtype=np.uint32) if (want_counts or not relative) else None
    rel = np.empty((n, NCANON), dtype=np.float32) if relative else None
    gc = np.empty(n, dtype=np.float32) if want_gc else None
    _lib.check(lib.mp2_tnf(_lib.ptr(bases), _lib.ptr(offsets), n, device,
                           _lib.ptr(counts), _lib.ptr(rel), _lib.ptr(gc)))
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
    _lib.check(_lib.load().mp2_tnf(_lib.ptr(bases), _lib.ptr(offsets), n, -1, None, None,
                                   _lib.ptr(gc)))
    return gc
