"""Batch scoring entry points on libmp2 (all MAGs of a run in one call).

These are what the scalable drivers use; magpurify2_b200.tools keeps the reference's per-MAG
function names (get_weighted_median, get_log_ratio_scores) on top of them."""
import numpy as np

from . import _lib

NCANON = _lib.NCANON


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def log_ratio_scores_batch(data, weights, mag_off, base, col_mask=None, pseudo=1e-5, clip=(0.0, 1.0),
                           min_rows=3, one_sided=False, device=-1, want_medians=False):
    """data: float32[n_rows] or [n_rows, n_cols] for all MAGs stacked; weights int64[n_rows];
    mag_off int64[n_mags+1].  Returns float64[n_rows] (and float32[n_mags, n_cols] medians)."""
    data = _c(data, np.float32)
    d2 = data.reshape(len(data), -1)
    n_rows, n_cols = d2.shape
    weights = _c(weights, np.int64)
    mag_off = _c(mag_off, np.int64)
    n_mags = len(mag_off) - 1
    if len(weights) != n_rows:
        raise ValueError("weights and data disagree on the number of rows")
    mask = None if col_mask is None else _c(col_mask, np.uint8).reshape(n_mags, n_cols)
    scores = np.empty(n_rows, dtype=np.float64)
    med = np.empty((n_mags, n_cols), dtype=np.float32)
    if n_cols == 0:
        scores[:] = 1.0
        return (scores, med) if want_medians else scores
    _lib.check(_lib.load().mp2_logratio(_lib.ptr(d2), _lib.ptr(weights), _lib.ptr(mag_off), n_mags, n_rows,
                                        n_cols, _lib.ptr(mask), float(base), float(pseudo), float(clip[0]),
                                        float(clip[1]), int(min_rows), int(bool(one_sided)), device,
                                        _lib.ptr(scores), _lib.ptr(med)))
    return (scores, med) if want_medians else scores


def _torch_dev(device):
    import torch

    if not torch.cuda.is_available():
        raise _lib.Mp2Error(_lib.ENODEV, "no usable CUDA device; magpurify2_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device() if device < 0 else device)


def select_samples_batch(cov, weights, mag_off, min_average_coverage, device=-1):
    """core.py:301-305 for all MAGs: (bool[n_mags, n_cols] selected, float64 means)."""
    import torch

    dev = _torch_dev(device)
    cov = _c(cov, np.float32)
    n_rows, n_cols = cov.shape
    mag_off = _c(mag_off, np.int64)
    n_mags = len(mag_off) - 1
    with torch.cuda.device(dev):
        d_cov = torch.from_numpy(cov).to(dev)
        d_w = torch.from_numpy(_c(weights, np.int64)).to(dev)
        d_mo = torch.from_numpy(mag_off).to(dev)
        d_sel = torch.empty((n_mags, n_cols), dtype=torch.uint8, device=dev)
        d_mean = torch.empty((n_mags, n_cols), dtype=torch.float64, device=dev)
        _lib.check(_lib.load().mp2_select_samples_device(
            d_cov.data_ptr(), d_w.data_ptr(), d_mo.data_ptr(), n_mags, n_rows, n_cols,
            float(min_average_coverage), d_sel.data_ptr(), d_mean.data_ptr(), _lib.current_stream_ptr()))
        return d_sel.cpu().numpy().astype(bool), d_mean.cpu().numpy()


def clr_centroid_batch(counts, weights, mag_off, device=-1, want_clr=False):
    """North-star composition score input: CLR rows + Euclidean distance to the MAG's
    length-weighted centroid.  counts uint32[n_rows,136] -> float64[n_rows] (and float32 CLR)."""
    import torch

    dev = _torch_dev(device)
    counts = _c(counts, np.uint32)
    n_rows = counts.shape[0]
    mag_off = _c(mag_off, np.int64)
    n_mags = len(mag_off) - 1
    with torch.cuda.device(dev):
        d_c = torch.from_numpy(counts.view(np.int32)).to(dev)
        d_w = torch.from_numpy(_c(weights, np.int64)).to(dev)
        d_mo = torch.from_numpy(mag_off).to(dev)
        d_dist = torch.empty(n_rows, dtype=torch.float64, device=dev)
        d_clr = torch.empty((n_rows, NCANON), dtype=torch.float32, device=dev) if want_clr else None
        _lib.check(_lib.load().mp2_clr_centroid_device(
            d_c.data_ptr(), d_w.data_ptr(), d_mo.data_ptr(), n_mags, n_rows,
            d_clr.data_ptr() if want_clr else None, d_dist.data_ptr(), _lib.current_stream_ptr()))
        dist = d_dist.cpu().numpy()
        return (dist, d_clr.cpu().numpy()) if want_clr else dist
