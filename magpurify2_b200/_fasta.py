"""ctypes view of libmp2's FASTA reader (mp2_fasta_*): one contiguous byte stream + offsets,
the layout the composition kernels take (no per-record Python str round trip)."""
import bz2
import ctypes as C
import lzma

import numpy as np

from . import _lib


class Fasta:
    """ids / descriptions (lists of str), bases uint8[L] and offsets int64[n+1] (NumPy views that
    stay valid until close())."""

    def __init__(self, path):
        lib = _lib.load()
        h = C.c_void_p()
        path = str(path)
        with open(path, "rb") as fin:
            sig = fin.read(8)
        if sig[:3] == b"BZh" or sig[:7] == b"\xfd7zXZ\x00\x00":  # tools.py:140-145
            raw = (bz2 if sig[:3] == b"BZh" else lzma).open(path, "rb").read()
            buf = np.frombuffer(raw, dtype=np.uint8)
            _lib.check(lib.mp2_fasta_open_mem(_lib.ptr(buf), len(buf), C.byref(h)))
        else:
            _lib.check(lib.mp2_fasta_open(path.encode(), C.byref(h)))
        self._h, self._lib = h, lib
        n = lib.mp2_fasta_n(h)
        self.ids = [lib.mp2_fasta_id(h, i).decode("utf-8", "replace") for i in range(n)]
        self.descriptions = [lib.mp2_fasta_description(h, i).decode("utf-8", "replace") for i in range(n)]
        self.offsets = np.ctypeslib.as_array(C.cast(lib.mp2_fasta_offsets(h), C.POINTER(C.c_int64)), (n + 1,)).copy()
        L = int(self.offsets[-1])
        self.bases = (np.ctypeslib.as_array(C.cast(lib.mp2_fasta_bases(h), C.POINTER(C.c_uint8)), (L,))
                      if L else np.zeros(0, np.uint8))
        self.lengths = np.diff(self.offsets)

    def sequence(self, i):
        return self.bases[self.offsets[i]:self.offsets[i + 1]].tobytes().decode("ascii", "replace")

    def close(self):
        if self._h:
            self.bases = None
            self._lib.mp2_fasta_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
