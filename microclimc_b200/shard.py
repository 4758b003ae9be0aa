"""Static column partition for multi-GPU runs: columns are independent, so a batch is split into
contiguous ranges, one per GPU / rank, with NO data-path collective; the only exchange is the final host gather.

Two launch styles use the same partition:
 * in-process: `Context(m, sm, n_gpus=G)` — the C library splits [0, ncol) over G devices (one host thread + stream
   per device, mc_lib.cu `mc_set_columns`);
 * one process per GPU under torchrun (bench.py): every rank owns `column_range(ncol, world, rank)` and rank 0
   collects per-column results with `gather_columns` (torch.distributed gather over gloo or NCCL).
"""
import numpy as np


def column_range(ncol, world, rank):
    """[c0, c1) of rank `rank`: ceil(ncol/world) columns per rank, the tail rank takes what is left (mc_set_columns)."""
    per = (ncol + world - 1) // world
    c0 = min(rank * per, ncol)
    return c0, min(c0 + per, ncol)


def slice_columns(arr, c0, c1):
    """Columns [c0, c1) of a [rows][ncol] (or [ncol]) host array, contiguous."""
    a = np.asarray(arr)
    return np.ascontiguousarray(a[..., c0:c1])


def gather_columns(local, ncol, group=None, dst=0):
    """Final host gather: every rank passes its [rows][c1-c0] block; rank `dst` gets [rows][ncol], others None."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    rows = local.shape[:-1]
    per = (ncol + world - 1) // world
    pad = np.zeros(rows + (per,), dtype=np.float64)
    pad[..., :local.shape[-1]] = local
    t = torch.from_numpy(pad)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.empty(rows + (ncol,), dtype=np.float64)
    for r in range(world):
        c0, c1 = column_range(ncol, world, r)
        out[..., c0:c1] = bufs[r].cpu().numpy()[..., :c1 - c0]
    return out
