"""N>1 path on CPU: world_size-2 gloo, query batch split by rank, one all_gather of results."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from leann_b200.parallel import shard_bounds, sharded_search


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, nq, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = np.arange(nq * 4, dtype=np.float32).reshape(nq, 4)
    calls = []

    def local(qs):  # stands in for B200HnswSearcher on this rank's GPU
        calls.append(len(qs))
        base = qs[:, :1]
        return (base + np.arange(k, dtype=np.float32)[None]), (base.astype(np.int64) * 10 + np.arange(k)[None])

    D, I = sharded_search(local, q, k, device="cpu")
    lo, hi = shard_bounds(nq, world, rank)
    ret[rank] = (D, I, calls, (lo, hi))
    dist.destroy_process_group()


def test_world2_gloo_sharded_search_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    nq, k, world = 11, 3, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, nq, k, ret), nprocs=world, join=True)
    q = np.arange(nq * 4, dtype=np.float32).reshape(nq, 4)
    expD = q[:, :1] + np.arange(k, dtype=np.float32)[None]
    expI = q[:, :1].astype(np.int64) * 10 + np.arange(k)[None]
    for r in range(world):
        D, I, calls, (lo, hi) = ret[r]
        assert np.array_equal(D, expD) and np.array_equal(I, expI)
        assert calls == [hi - lo]  # each rank searched only its own slice
