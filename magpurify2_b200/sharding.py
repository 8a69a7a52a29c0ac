"""MAG sharding across the GPUs of one box (one process per GPU, torch.distributed).

MAGs are independent, so the only exchange on these paths is the FINAL GATHER of the per-contig
result rows (north_star): whole MAGs are assigned to ranks with longest-processing-time-first on
their base pairs, every rank scores its MAGs, and the rows are gathered back in input order.
Works with the NCCL backend (CUDA tensors, NVLink/NVSwitch) and with gloo (CPU tensors; used by
the world_size-2 tests)."""
import numpy as np


def lpt_partition(weights, n_parts):
    """Greedy LPT: items sorted by weight descending go to the currently lightest part.
    Returns a list of n_parts int64 index arrays (each ascending)."""
    weights = np.asarray(weights, dtype=np.float64)
    order = np.argsort(-weights, kind="stable")
    load = np.zeros(n_parts)
    parts = [[] for _ in range(n_parts)]
    for i in order.tolist():
        p = int(np.argmin(load))
        parts[p].append(i)
        load[p] += weights[i]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def mags_of_rank(mag_bp, rank, world):
    return lpt_partition(mag_bp, world)[rank]


def gather_rows(local_rows, local_mags, mag_sizes, group=None):
    """Final gather.  local_rows: torch tensor [n_local_rows, ...] holding the rows of this rank's
    MAGs (`local_mags`, ascending MAG indices, rows in that order); mag_sizes[m] = rows of MAG m.
    Returns the tensor of ALL rows in MAG order on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mag_sizes = np.asarray(mag_sizes, dtype=np.int64)
    parts = lpt_partition_from_assignment(local_mags, len(mag_sizes), group)
    counts = [int(mag_sizes[p].sum()) for p in parts]
    mx = max(counts + [1])
    tail = tuple(local_rows.shape[1:])
    pad = torch.zeros((mx,) + tail, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = torch.empty((world * mx,) + tail, dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    row_off = np.zeros(len(mag_sizes) + 1, dtype=np.int64)
    np.cumsum(mag_sizes, out=row_off[1:])
    full = torch.empty((int(row_off[-1]),) + tail, dtype=local_rows.dtype, device=local_rows.device)
    for r, p in enumerate(parts):
        src = r * mx
        for m in p.tolist():
            k = int(mag_sizes[m])
            full[row_off[m]:row_off[m] + k] = out[src:src + k]
            src += k
    return full


def lpt_partition_from_assignment(local_mags, n_mags, group=None):
    """Every rank learns every rank's MAG list (tiny all_gather of index vectors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.full((n_mags,), -1, dtype=torch.int64, device=dev)
    mine[:len(local_mags)] = torch.as_tensor(np.asarray(local_mags, dtype=np.int64), device=dev)
    allm = torch.empty((world * n_mags,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allm, mine, group=group)
    allm = allm.cpu().numpy().reshape(world, n_mags)
    return [row[row >= 0] for row in allm]
