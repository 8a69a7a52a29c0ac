#!/usr/bin/env python
"""Per-sample kernel time of the coverage fast path on the C3 samples of bench.py (device events around single
calls), next to each sample's number of blocks and depth statistics.  usage: python tools/cov_per_sample.py [n_samples]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magpurify2_b200 import _lib, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 10
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
for s in range(S):
    st, en, eo, R = synth.make_events_device(clen, mag_off, s, dev, seed=0)
    sp, ds = C.c_uint32(0), C.c_uint32(0)
    _lib.check(lib.mp2_cov_order_device(st.data_ptr(), en.data_ptr(), eo.data_ptr(), d_off.data_ptr(), n, R, C.byref(sp), C.byref(ds), stream))
    times = []
    for rep in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.mp2_cov_events_device(st.data_ptr(), en.data_ptr(), eo.data_ptr(), d_off.data_ptr(), n, R, L, sp.value, ds.value,
                                             75, 0.05, 0.05, cov.data_ptr(), 1, stats.data_ptr(), stream))
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    stt = stats.cpu().numpy().view(_lib.COV_STATS_DTYPE)
    deep = stt["max_depth"] >= 511
    print(f"sample {s}: R {R / 1e6:.0f} M  call ms {min(times[1:]):.3f}  contigs with depth >= 511: {int(deep.sum())} "
          f"({int(clen[deep].sum()) / 1e6:.1f} Mbp)  max depth {int(stt['max_depth'].max())}  general {bool((stt['flags'] & 1).any())}")
    del st, en, eo
