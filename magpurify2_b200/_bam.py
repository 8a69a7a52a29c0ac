"""ctypes view of libmp2's BAM reader (mp2_bam_*): the filtered alignment blocks of one
coordinate-sorted BAM, grouped by reference sequence, as zero-copy NumPy arrays over the
library's (pinned) buffers."""
import ctypes as C

import numpy as np

from . import _lib


class BamEvents:
    def __init__(self, path, min_identity=0.97, threads=1):
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.mp2_bam_open(str(path).encode(), float(min_identity), int(threads), C.byref(h)))
        self._h, self._lib = h, lib
        n = lib.mp2_bam_n_contigs(h)
        self.names = [lib.mp2_bam_contig_name(h, i).decode() for i in range(n)]
        self.lengths = np.ctypeslib.as_array(C.cast(lib.mp2_bam_contig_len(h), C.POINTER(C.c_int64)), (n,)).copy() \
            if n else np.zeros(0, np.int64)
        self.ev_off = np.ctypeslib.as_array(C.cast(lib.mp2_bam_ev_off(h), C.POINTER(C.c_int64)), (n + 1,)).copy()
        ne = lib.mp2_bam_n_events(h)
        self.n_events = ne
        self.n_records = lib.mp2_bam_n_records(h)
        self.n_kept = lib.mp2_bam_n_kept(h)
        self.max_span = lib.mp2_bam_max_span(h)          # longest reference footprint of a kept read
        self.max_block = lib.mp2_bam_max_block(h)        # longest M/=/X block
        self.max_disorder = lib.mp2_bam_max_disorder(h)  # largest (block start - read position)
        if ne:
            self.starts = np.ctypeslib.as_array(C.cast(lib.mp2_bam_starts(h), C.POINTER(C.c_uint32)), (ne,))
            self.ends = np.ctypeslib.as_array(C.cast(lib.mp2_bam_ends(h), C.POINTER(C.c_uint32)), (ne,))
        else:
            self.starts = np.zeros(0, np.uint32)
            self.ends = np.zeros(0, np.uint32)
        ln = C.c_int64()
        p = lib.mp2_bam_header_text(h, C.byref(ln))
        self.header_text = C.string_at(p, ln.value).decode("utf-8", "replace") if p and ln.value else ""

    def close(self):
        if self._h:
            self.starts = self.ends = None  # views into memory that is about to be freed
            self._lib.mp2_bam_close(self._h)
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
