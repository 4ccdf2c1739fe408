"""Multi-GPU: queries are the independent unit (the reference already treats them so — OpenMP
loop with a private VisitedTable / distance computer per thread, faiss/IndexHNSW.cpp:361-399).
One process per GPU, graph + passage store + encoder replicated, the query batch split into
contiguous slices by rank, NO per-hop communication; one all_gather of (labels, distances) at the
end of the call (NCCL over NVLink on GPU tensors, gloo on CPU tensors in the tests).
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of [0, n) owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_search(local_search: Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]], query: np.ndarray, k: int,
                   device: torch.device | str | None = None, group=None) -> tuple[np.ndarray, np.ndarray]:
    """Runs `local_search(query[lo:hi]) -> (D [m,k] f32, I [m,k] i64)` on this rank's slice and
    returns the full (D, I) on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_search(query)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nq = query.shape[0]
    lo, hi = shard_bounds(nq, world, rank)
    D_loc, I_loc = local_search(query[lo:hi]) if hi > lo else (np.zeros((0, k), np.float32), np.zeros((0, k), np.int64))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    cap = (nq + world - 1) // world  # equal-sized buffers for all_gather
    # ONE collective per call: labels and distances travel together as 12-byte (int64, float32) records
    rec = np.dtype([("i", "<i8"), ("d", "<f4")])
    mine = np.zeros((cap, k), rec)
    mine["i"] = -1
    mine["i"][: hi - lo] = I_loc
    mine["d"][: hi - lo] = D_loc
    buf = torch.from_numpy(mine.view(np.uint8).reshape(cap, k * rec.itemsize)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    for r in range(world):
        a, b = shard_bounds(nq, world, r)
        got = out[r].cpu().numpy().reshape(-1).view(rec).reshape(cap, k)
        D[a:b] = got["d"][: b - a]
        I[a:b] = got["i"][: b - a]
    return D, I
