"""Multi-GPU plumbing of the path: RVEs are independent, so ranks take the reference's rank-strided ids
(`np.arange(run, ns, num_runs)`, simulation_snapshots.py:76; `[run::num_runs]`, tangents_prediction.py:44)
and the only exchange is ONE gather of the small per-RVE outputs at the end.
Works on any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
import numpy as np


def shard_ids(ns, rank, world):
    return np.arange(rank, ns, world).astype(np.int64)


def gather_rows(ids, arrays, dst=0, device=None):
    """Gather per-RVE rows from every rank on `dst` and order them by id.

    ids: (n_local,) int64 global ids of the local rows; arrays: dict name -> (n_local, ...) float64.
    Returns (ids_sorted, dict of concatenated arrays sorted by id) on dst, (None, None) elsewhere.
    Without an initialised process group it is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    names = sorted(arrays)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        order = np.argsort(ids, kind="stable")
        return np.asarray(ids)[order], {k: np.asarray(arrays[k])[order] for k in names}
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n_local = len(ids)
    widths = [int(np.prod(np.shape(arrays[k])[1:], dtype=np.int64)) for k in names]
    flat = np.concatenate([np.asarray(ids, dtype=np.float64).reshape(n_local, 1)] +
                          [np.asarray(arrays[k], dtype=np.float64).reshape(n_local, w) for k, w in zip(names, widths)], axis=1)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = n_local
    dist.all_reduce(counts)
    nmax = int(counts.max().item())
    buf = torch.zeros((nmax, flat.shape[1]), dtype=torch.float64, device=dev)
    buf[:n_local] = torch.from_numpy(flat).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None, None
    rows = np.concatenate([out[r][:int(counts[r].item())].cpu().numpy() for r in range(world)], axis=0)
    order = np.argsort(rows[:, 0], kind="stable")
    rows = rows[order]
    res, o = {}, 1
    for k, w in zip(names, widths):
        res[k] = rows[:, o:o + w].reshape((rows.shape[0],) + tuple(np.shape(arrays[k])[1:]))
        o += w
    return rows[:, 0].astype(np.int64), res
