"""ctypes binding of libmp2.so (C ABI in include/mp2.h).

This is the thin shim that takes the place of the reference's PyO3 glue
(/root/reference/src/lib.rs:19-22, setup.py:33-38).  It fails loudly when the library is missing:
the product has no CPU path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmp2.so")
NCANON = 136

OK, EINVAL, ENODEV, ECUDA, ENOMEM, EOVERFLOW, EIO, EFORMAT = 0, -1, -2, -3, -4, -5, -6, -7


class Mp2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmp2 error {code}: {msg}")
        self.code = code


class CovStats(C.Structure):
    _fields_ = [
        ("observed_len", C.c_uint64),
        ("covered", C.c_uint64),
        ("min_index", C.c_uint64),
        ("max_index", C.c_uint64),
        ("total", C.c_uint64),
        ("max_depth", C.c_uint64),
        ("coverage", C.c_float),
        ("flags", C.c_uint32),
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
        ("flags", "<u4"),
    ]
)
assert COV_STATS_DTYPE.itemsize == C.sizeof(CovStats) == 56

_vp, _i64, _i32, _u32, _f32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol include/mp2.h declares
SIGNATURES = {
    "mp2_last_error": (C.c_char_p, []),
    "mp2_version": (C.c_char_p, []),
    "mp2_device_count": (_i32, []),
    "mp2_canonical_lut": (_i32, [_vp]),
    "mp2_profile_enable": (_i32, [_i32]),
    "mp2_profile_reset": (_i32, []),
    "mp2_profile_read": (_i32, [C.c_char_p, C.POINTER(_f64), C.POINTER(_i64)]),
    "mp2_launch_count": (_i64, []),
    "mp2_tnf": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mp2_tnf_device": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mp2_tnf_normalize_device": (_i32, [_vp, _i64, _vp, _vp]),
    "mp2_tnf_packed_device": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mp2_pack_device": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "mp2_packed_words": (_i64, [_i64, C.POINTER(_i64)]),
    "mp2_pack_host": (_i32, [_vp, _vp, _i64, _vp, _vp, _i32]),
    "mp2_invalid_runs": (_i64, [_vp, _i64, _vp, _i64, _i32]),
    "mp2_tnf_packed": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mp2_cov_events_device": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _u32, _u32, _u32, _f32, _f32, _vp, _i64, _vp, _vp]),
    "mp2_cov_order_device": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(_u32), C.POINTER(_u32), _vp]),
    "mp2_cov_events": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _u32, _u32, _u32, _f32, _f32, _i32, _vp, _i64, _vp]),
    "mp2_logratio_device": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _f64, _f32, _f64, _f64, _i64, _i32, _vp, _vp, _vp]),
    "mp2_logratio": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _f64, _f32, _f64, _f64, _i64, _i32, _i32, _vp, _vp]),
    "mp2_select_samples_device": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _f64, _vp, _vp, _vp]),
    "mp2_clr_centroid_device": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "mp2_bam_open": (_i32, [C.c_char_p, _f32, _i32, C.POINTER(_vp)]),
    "mp2_bam_n_contigs": (_i64, [_vp]),
    "mp2_bam_n_events": (_i64, [_vp]),
    "mp2_bam_n_records": (_i64, [_vp]),
    "mp2_bam_n_kept": (_i64, [_vp]),
    "mp2_bam_max_span": (_u32, [_vp]),
    "mp2_bam_max_block": (_u32, [_vp]),
    "mp2_bam_max_disorder": (_u32, [_vp]),
    "mp2_bam_contig_name": (C.c_char_p, [_vp, _i64]),
    "mp2_bam_contig_len": (_vp, [_vp]),
    "mp2_bam_ev_off": (_vp, [_vp]),
    "mp2_bam_starts": (_vp, [_vp]),
    "mp2_bam_ends": (_vp, [_vp]),
    "mp2_bam_header_text": (_vp, [_vp, C.POINTER(_i64)]),
    "mp2_bam_close": (None, [_vp]),
    "mp2_fasta_open": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "mp2_fasta_open_mem": (_i32, [_vp, _i64, C.POINTER(_vp)]),
    "mp2_fasta_n": (_i64, [_vp]),
    "mp2_fasta_bases": (_vp, [_vp]),
    "mp2_fasta_offsets": (_vp, [_vp]),
    "mp2_fasta_id": (C.c_char_p, [_vp, _i64]),
    "mp2_fasta_description": (C.c_char_p, [_vp, _i64]),
    "mp2_fasta_close": (None, [_vp]),
}

_lib = None


def load():
    """Loads libmp2.so or raises -- never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C magpurify2_b200/csrc`. magpurify2_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().mp2_last_error().decode("utf-8", "replace")
        if rc == ENOMEM:
            raise MemoryError(msg)
        if rc == EINVAL:
            raise ValueError(msg)
        if rc in (EIO,):
            raise OSError(msg)
        raise Mp2Error(rc, msg)


def ptr(a):
    """numpy array / torch tensor / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch.Tensor used purely as a device (or pinned host) buffer


def device_count():
    n = load().mp2_device_count()
    return max(n, 0)


def current_stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
