#!/usr/bin/env python
"""Isolated times of the scoring calls at C3 shapes (one stream, device events, best of 10):
mp2_clr_centroid_device, mp2_logratio_device (weighted median + log-ratio) on 1 column and on 10 columns."""
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
n, n_mags = len(clen), len(mag_off) - 1
rng = np.random.default_rng(0)
counts = torch.from_numpy(rng.poisson(np.maximum(clen[:, None] / 136.0, 0.1), (n, 136)).astype(np.int32)).to(dev)
d_len = torch.from_numpy(clen.astype(np.int64)).to(dev)
d_mo = torch.from_numpy(mag_off.astype(np.int64)).to(dev)
dist = torch.empty(n, dtype=torch.float64, device=dev)
x1 = torch.from_numpy(rng.random(n).astype(np.float32)).to(dev)
x10 = torch.from_numpy((rng.random((n, 10)) * 30).astype(np.float32)).to(dev)
sc = torch.empty(n, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def best(f, reps=10):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[2:]) * 1e3


def prof(f, names):
    lib.mp2_profile_reset(); lib.mp2_profile_enable(1); f(); torch.cuda.synchronize(); lib.mp2_profile_enable(0)
    out = []
    for nm in names:
        ms, k = C.c_double(0), C.c_int64(0)
        lib.mp2_profile_read(nm.encode(), C.byref(ms), C.byref(k))
        out.append(f"{nm} {ms.value * 1e3:.1f} us")
    return "  ".join(out)


clr = lambda: _lib.check(lib.mp2_clr_centroid_device(counts.data_ptr(), d_len.data_ptr(), d_mo.data_ptr(), n_mags, n, None, dist.data_ptr(), stream))
lr1 = lambda: _lib.check(lib.mp2_logratio_device(x1.data_ptr(), d_len.data_ptr(), d_mo.data_ptr(), n_mags, n, 1, None, 2.0, 1e-5, 0.0, 1.0, 3, 0, sc.data_ptr(), None, stream))
lr10 = lambda: _lib.check(lib.mp2_logratio_device(x10.data_ptr(), d_len.data_ptr(), d_mo.data_ptr(), n_mags, n, 10, None, 25.0, 1e-5, 0.0, 1.0, 3, 0, sc.data_ptr(), None, stream))
print(f"clr_centroid call {best(clr):.1f} us   {prof(clr, ['clr_mag_kernel'])}")
print(f"logratio 1 column call {best(lr1):.1f} us   {prof(lr1, ['wmedian_kernel', 'logratio_kernel'])}")
print(f"logratio 10 columns call {best(lr10):.1f} us   {prof(lr10, ['wmedian_kernel', 'logratio_kernel'])}")
