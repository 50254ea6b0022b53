"""Multi-GPU host logic: packets shard trivially -- one process per GPU, contiguous index
blocks so that concatenation in rank order reproduces the reference's row order bit for bit, the flow
replicated on every rank, no collective in the step loop; a gather only at packet-output frames."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Rows [lo, hi) of rank `rank`: lo = rank n // world."""
    return rank * n // world, (rank + 1) * n // world


def initial_wavepackets_host(L, k0, sqrtN, first=0, count=None):
    """Host (NumPy) twin of the device generator, raytracing/RaytracingDriver.jl:27-47, for a contiguous
    block of packets -- used for small runs and to check the sharded device generator."""
    N = sqrtN * sqrtN
    count = N - first if count is None else count
    p0 = np.arange(first, first + count, dtype=np.int64)
    offset = L / sqrtN / 2
    xk = np.empty((count, 4), order="F")
    xk[:, 0] = (p0 % sqrtN + 1).astype(np.float64) * L / sqrtN - L / 2 - offset
    xk[:, 1] = (p0 // sqrtN + 1).astype(np.float64) * L / sqrtN - L / 2 - offset
    phase = 2 * np.pi * (p0 + 1).astype(np.float64) / N
    xk[:, 2] = k0 * np.cos(phase)
    xk[:, 3] = k0 * np.sin(phase)
    sign = np.where(p0 % 2 == 0, -1.0, 1.0)
    return xk, sign


def gather_rows(local: np.ndarray, dist=None, dst=0):
    """Concatenate per-rank row blocks in rank order on `dst` (ncclAllGather / gloo all_gather_object at
    output frames only).  Returns the full array on `dst`, None elsewhere.  `dist` = torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [None] * world
    dist.all_gather_object(counts, int(local.shape[0]))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    mx = max(counts)
    buf = torch.zeros((mx, width), dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local).reshape(local.shape[0], width)).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    if rank != dst:
        return None
    parts = [o[:c].cpu().numpy().reshape((c,) + local.shape[1:]) for o, c in zip(outs, counts)]
    return np.concatenate(parts, axis=0)


def max_over_ranks(x: float, dist=None):
    """Device timings are reported as the max over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
