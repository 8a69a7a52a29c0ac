"""Brute-force, obviously-correct Python oracles, independent of the rolling-code method.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Small inputs only: pure-Python loops.

  * canonical 4-mers by string windows + min(kmer, revcomp)   (behaviour of
    src/composition.rs:42-74 expressed without the rolling code / LUT);
  * per-base depth arrays + literal sorted-rank trimmed mean  (SURVEY.md Appendix A.3/A.4).
"""
import itertools
import math

import numpy as np

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
CANONICAL = sorted(
    {min("".join(k), "".join(_COMP[b] for b in reversed(k))) for k in itertools.product("ACGT", repeat=4)}
)
assert len(CANONICAL) == 136
_CLASS = {}
for _i, _k in enumerate(CANONICAL):
    _CLASS[_k] = _i
    _CLASS["".join(_COMP[b] for b in reversed(_k))] = _i


def canonical_counts(seq):
    """uint32[136] by looking at every window of 4 characters."""
    if isinstance(seq, bytes):
        seq = seq.decode("latin-1")
    out = np.zeros(136, dtype=np.uint32)
    up = seq.upper() if seq.isascii() else "".join(c.upper() if c.isascii() else "?" for c in seq)
    for i in range(len(up) - 3):
        w = up[i : i + 4]
        cls = _CLASS.get(w)
        if cls is not None:
            out[cls] += 1
    return out


def gc_fraction(seq):
    if isinstance(seq, bytes):
        seq = seq.decode("utf-8")
    gc = sum(1 for c in seq if c in "GgCc")
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float32(gc) / np.float32(len(seq))


def depth_array(events, length):
    """events: iterable of (start, end) with 0-based start, exclusive end."""
    d = np.zeros(length, dtype=np.int64)
    for s, e in events:
        if s >= length:
            continue
        d[s : min(e, length)] += 1
    return d


def trimmed_mean(events, length, excl=75, trim_lo=0.0, trim_upper=0.0):
    """Appendix A.3/A.4 on a brute-force per-base depth array (slice assignment per read, no
    difference array, no prefix sum): np.bincount of the window depths, then the literal
    histogram walk.  Returns (coverage f32, dict of integers)."""
    if 2 * excl >= length:
        return np.float32(0.0), dict(N=0, covered=0, min_index=0, max_index=0, total=0)
    d = depth_array(events, length)[excl : length - excl]
    n = len(d)
    covered = int((d > 0).sum())
    lo = np.float32(trim_lo)
    hi = np.float32(1.0) - np.float32(trim_upper)
    min_index = int(math.floor(float(lo * np.float32(n))))
    max_index = int(math.ceil(float(hi * np.float32(n))))
    info = dict(N=n, covered=covered, min_index=min_index, max_index=max_index, total=0)
    if covered == 0:
        return np.float32(0.0), info
    counts = np.bincount(d)
    acc = 0
    total = 0
    started = False
    for i, c in enumerate(counts.tolist()):
        acc += c
        if not started:
            if acc >= min_index:
                total = (max_index - min_index + 1) * i if acc > max_index else (acc - min_index + 1) * i
                started = True
        else:
            if acc > max_index:
                excess = acc - c
                total += (max_index - excess + 1 if max_index >= excess else 0) * i
                break
            total += c * i
    info["total"] = total
    with np.errstate(invalid="ignore", divide="ignore"):
        cov = np.float32(total) / np.float32(max_index - min_index)
    return cov, info
