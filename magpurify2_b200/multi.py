"""The GPUs of one box as a product feature: `--devices 0,1,2,3` on the command line.

MAGs (composition) and samples (coverage) are independent, so the work is simply dealt out -- whole MAGs by
longest-processing-time-first on their base pairs (sharding.lpt_partition), BAM files round-robin -- to one
host thread per device; the native calls release the GIL and select their device themselves, so the threads
run concurrently.  There is no exchange between devices; the "final gather" is the concatenation of the
per-device rows back into input order.  (The reference fans the same units out over joblib processes on the
CPU: modules/composition.py:74-86, modules/coverage.py:113-135.)"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _composition, _coverage, _lib, sharding
from .core import score_composition_batch


def parse_devices(args):
    """[device, ...] from --devices ("all", "0,2,3"), else the single --device (default: current)."""
    spec = getattr(args, "devices", None)
    if spec in (None, "", []):
        return [getattr(args, "device", -1)]
    if isinstance(spec, (list, tuple)):
        devs = [int(d) for d in spec]
    elif str(spec).strip().lower() == "all":
        devs = list(range(_lib.device_count()))
    else:
        devs = [int(tok) for tok in str(spec).replace(" ", "").split(",") if tok != ""]
    n = _lib.device_count()
    bad = [d for d in devs if d < 0 or d >= n]
    if bad or not devs:
        raise ValueError(f"--devices {spec!r}: this box has {n} CUDA device(s)")
    if len(set(devs)) != len(devs):
        raise ValueError(f"--devices {spec!r}: a device is listed twice")
    return devs


def shard_mags(mags, n_parts):
    """Indices of the MAGs of every part: LPT on base pairs, parts in ascending MAG order."""
    return sharding.lpt_partition([int(m.offsets[-1]) for m in mags], n_parts)


def _bound(fn, device, *a, **k):
    """Runs fn on a worker thread that first moves to the CPUs of the device's NUMA node (numa.bind_to_gpu_node)."""
    from . import numa

    if device is not None and device >= 0:
        numa.bind_to_gpu_node(device)
    return fn(*a, **k)


def composition_of(mags, device=-1, threads=None, packed=True):
    """counts, relative TNF, GC and the two scores of every contig of `mags` (stacked in order) on ONE device.
    The genomes are 2-bit packed on the host straight from the buffers the FASTA reader filled (no ASCII
    concatenation of the run) and uploaded at 0.375 bytes per base; packed=False uploads ASCII (A/B switch)."""
    n = sum(len(m) for m in mags)
    offsets = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.concatenate([m.lengths for m in mags]), out=offsets[1:])
    if packed:
        codes, valid = _composition.pack_parts([m.bases for m in mags], threads=threads)
        runs = _composition.invalid_runs(valid, int(offsets[-1]), threads=threads)
        counts, rel, gc = _composition.tnf_from_packed(codes, valid, offsets, relative=True, want_gc=True,
                                                       want_counts=True, device=device, runs=runs)
    else:
        bases = np.concatenate([m.bases for m in mags]) if mags else np.zeros(0, np.uint8)
        counts, rel, gc = _composition.tnf_from_concat(bases, offsets, relative=True, want_gc=True, want_counts=True,
                                                       device=device)
    tnf_score, gc_score = score_composition_batch(mags, counts, gc, device=device)
    return dict(counts=counts, rel=rel, gc=gc, tnf_score=tnf_score, gc_score=gc_score)


def composition_on_devices(mags, devices, threads=None):
    """The same, with the MAGs dealt out to `devices`; rows come back in the order of `mags`."""
    import os

    devices = list(devices)
    threads = int(threads or os.cpu_count() or 1)
    if len(devices) <= 1 or len(mags) <= 1:
        return composition_of(mags, devices[0] if devices else -1, threads)
    parts = shard_mags(mags, len(devices))
    per = max(1, threads // len(devices))  # host threads of every shard's packer
    with ThreadPoolExecutor(max_workers=len(devices)) as pool:
        futs = [pool.submit(_bound, composition_of, dev, [mags[i] for i in part.tolist()], dev, per) if len(part) else None
                for part, dev in zip(parts, devices)]
        shards = [f.result() if f is not None else None for f in futs]
    return gather_shards(mags, parts, shards)


def gather_shards(mags, parts, shards):
    """Final gather on the host: every shard's rows go back to the rows of its MAGs."""
    sizes = np.array([len(m) for m in mags], dtype=np.int64)
    first = np.concatenate([[0], np.cumsum(sizes)])
    out = None
    for part, res in zip(parts, shards):
        if res is None:
            continue
        if out is None:
            out = {k: np.empty((int(first[-1]),) + v.shape[1:], dtype=v.dtype) for k, v in res.items()}
        src = 0
        for m in part.tolist():
            k = int(sizes[m])
            for name, v in res.items():
                out[name][first[m]:first[m] + k] = v[src:src + k]
            src += k
    return out


def bam_coverages_on_devices(bam_list, devices, **kw):
    """get_bam_coverages with the files dealt out to `devices` (see _coverage.get_bam_coverages)."""
    return _coverage.get_bam_coverages(bam_list, devices=list(devices), **kw)
