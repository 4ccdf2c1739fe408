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


def _plugin_worker(rank, world, port, tmp, ret):
    """Both plugin searchers with shard_queries=True under a 2-rank gloo group; the device handle is a fake that scores
    from the query values so that every rank can check the gathered result."""
    import json
    from pathlib import Path

    from leann_b200 import backend, capi, diskann_backend

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    class Fake:
        def __init__(self, *a, **kw):
            self.info = type("I", (), dict(d=4, ntotal=100, has_vectors=1))()
            self.dinfo = type("D", (), dict(dim=4))()
            self.last_stats = capi.SearchStats()

        def configure(self, *a): pass
        def set_option(self, *a): pass
        def close(self): pass

        def search(self, q, k, params):
            seen.append(len(q))
            base = q[:, :1]
            return (base + np.arange(k, dtype=np.float32)[None]), (base.astype(np.int64) + np.arange(k)[None])

    capi.Index = Fake
    capi.DiskannIndex = Fake
    d = Path(tmp)
    meta = {"backend_name": "x", "embedding_model": "synthetic/tiny-bert", "dimensions": 4, "backend_kwargs": {"distance_metric": "l2"},
            "is_compact": True, "is_pruned": False}
    out = {}
    for name, cls, marker in (("h", backend.B200HnswSearcher, "h.index"), ("d", diskann_backend.B200DiskannSearcher, "d_pq_compressed.bin")):
        if rank == 0:
            (d / f"{name}.leann.meta.json").write_text(json.dumps(meta))
            (d / marker).write_bytes(b"x")
        dist.barrier()
        s = cls(str(d / f"{name}.leann"), shard_queries=True)
        assert s.device == rank  # one process per GPU: LOCAL_RANK picks the device
        q = np.arange(9 * 4, dtype=np.float32).reshape(9, 4)
        seen.clear()
        res = s.search(q, 3, recompute_embeddings=False, **({"skip_search_reorder": True} if name == "d" else {}))
        out[name] = (res["labels"], res["distances"], list(seen))
    ret[rank] = out
    dist.destroy_process_group()


def test_world2_plugin_searchers_shard_queries(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_plugin_worker, args=(2, port, str(tmp_path), ret), nprocs=2, join=True)
    q = np.arange(9 * 4, dtype=np.float32).reshape(9, 4)
    expD = q[:, :1] + np.arange(3, dtype=np.float32)[None]
    expL = [[str(int(v)) for v in row] for row in (q[:, :1].astype(np.int64) + np.arange(3)[None])]
    for r in range(2):
        lo, hi = shard_bounds(9, 2, r)
        for name in ("h", "d"):
            labels, D, seen = ret[r][name]
            assert labels == expL and np.array_equal(D, expD)   # every rank holds the full, ordered result
            assert seen == [hi - lo]                            # ... having searched only its own slice
