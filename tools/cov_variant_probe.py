#!/usr/bin/env python
"""Where the time of one coverage call goes when the order of the blocks is measured inside the call (max_span = 0):
per-kernel times (mp2_profile_*) for a headline sample and a SURVEY 8(d) variant sample.  usage: python tools/cov_variant_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpurify2_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
clen, mag_off, gc = synth.mag_layout(1000, seed=0)
offsets = np.zeros(len(clen) + 1, dtype=np.int64)
np.cumsum(clen, out=offsets[1:])
n, L = len(clen), int(offsets[-1])
d_off = torch.from_numpy(offsets).to(dev)
cov = torch.empty(n, dtype=torch.float32, device=dev)
stats = torch.zeros(n * 56, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
NAMES = ["cov_init_kernel", "cov_map_kernel", "cov_order_kernel", "cov_order_scan_kernel", "cov_bounds_kernel", "cov_run_order_kernel",
         "cov_stream2_kernel", "cov_zero_kernel", "cov_scatter_kernel", "cov_tile_scan_kernel", "cov_depth_kernel", "cov_deep_kernel"]
for variant in (False, True):
    st, en, eo, R = synth.make_events_device(clen, mag_off, 0, dev, seed=0, variant=variant)
    sp, ds = C.c_uint32(0), C.c_uint32(0)
    _lib.check(lib.mp2_cov_order_device(st.data_ptr(), en.data_ptr(), eo.data_ptr(), d_off.data_ptr(), n, R, C.byref(sp), C.byref(ds), stream))
    for span, dis in ((sp.value, ds.value), (0, 0)):
        for rep in range(3):
            if rep == 2:
                lib.mp2_profile_reset(); lib.mp2_profile_enable(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.mp2_cov_events_device(st.data_ptr(), en.data_ptr(), eo.data_ptr(), d_off.data_ptr(), n, R, L, span, dis,
                                                 75, 0.05, 0.05, cov.data_ptr(), 1, stats.data_ptr(), stream))
            b.record()
            torch.cuda.synchronize()
        lib.mp2_profile_enable(0)
        stt = stats.cpu().numpy().view(_lib.COV_STATS_DTYPE)
        parts = []
        for nm in NAMES:
            ms, k = C.c_double(0), C.c_int64(0)
            lib.mp2_profile_read(nm.encode(), C.byref(ms), C.byref(k))
            if k.value:
                parts.append(f"{nm} {ms.value:.3f}")
        print(f"variant {variant} R {R / 1e6:.0f} M span {sp.value} disorder {ds.value} passed ({span},{dis}): call {a.elapsed_time(b):.3f} ms  general {bool((stt['flags'] & 1).any())}  " + "  ".join(parts))
    del st, en, eo
