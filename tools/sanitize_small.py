"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from magpurify2_b200 import _composition, _coverage, _scoring, synth

d = synth.make_mags(2, seed=3, mean_bp=1.5e5, n_contigs=25)
counts, rel, gc = _composition.tnf_from_concat(d["bases"], d["offsets"], relative=True, want_gc=True, want_counts=True)
assert np.array_equal(counts, oracle.get_tnf_concat(d["bases"], d["offsets"], relative=False))
for variable in (False, True):
    st, en, eo = synth.make_events(d["lengths"], d["mag_off"], 0, seed=3, variable=variable, depth_scale=0.5)
    cov = _coverage.cov_from_events(st, en, eo, d["offsets"], 75, 0.05, 0.05)
    want, _ = oracle.cov_sample(st, en, eo, d["lengths"], 75, 0.05, 0.05)
    assert np.array_equal(cov, want)
# deep contig (deep pool) + unsorted (general path)
n = 20000
s = np.sort(np.random.default_rng(1).integers(0, 4000, n)).astype(np.uint32)
cov = _coverage.cov_from_events(s, s + 150, [0, n], [0, 5000], 75, 0.05, 0.05)
want, _ = oracle.cov_sample(s, s + 150, [0, n], [5000], 75, 0.05, 0.05)
assert np.array_equal(cov, want)
p = np.random.default_rng(2).permutation(n)
cov = _coverage.cov_from_events(s[p], (s + 150)[p], [0, n], [0, 5000], 75, 0.05, 0.05)
assert np.array_equal(cov, want)
sc = _scoring.log_ratio_scores_batch(gc, d["lengths"], d["mag_off"], 2)
dist = _scoring.clr_centroid_batch(counts, d["lengths"], d["mag_off"])
print("sanitize_small ok", float(np.nansum(sc)), float(dist.sum()))
