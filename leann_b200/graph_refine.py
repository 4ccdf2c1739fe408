"""Tooling (not the hot path): search-based refinement of a level-0 HNSW adjacency.

The batch builder of graph_build.py needs k-nearest-neighbour candidate lists.  Above a few million points those come from a
partition-restricted search whose recall is the limit of the final graph (recall@10 0.80 at 10 M).  A graph
that is merely *navigable* is, however, itself a much better kNN oracle than the partitions it was built from: searching it with
every point as the query (ef ~ 128) returns lists of far higher recall, and rebuilding level 0 from those lists gives a better
graph — the same idea as the reference's incremental construction (every insertion is a search over the graph so far,
faiss/impl/HNSW.cpp:600-894), applied as whole-graph sweeps.  `search_fn` is any batch searcher over the current graph: the CUDA
stored-vector search (capi.Index.search(..., recompute=False), 6e5 queries/s, i.e. tens of seconds per sweep at 10 M) in
production, the CPU oracle in the unit test.

Status: validated on CPU at small scale (tests/test_graph_refine.py); bench.py exposes the sweeps as --sweeps (default 0: the
insertion-as-search builder of graph_build.py reaches the target recall without them).  `prune_degrees` (below) is the
reference's per-node degree cap for the pruned-degree configuration.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from .csr import METRIC_INNER_PRODUCT, CSRGraph, csr_from_padded
from .graph_build import _add_reverse_and_cap, _heuristic_prune


def upper_levels(g: CSRGraph, M: int) -> dict[int, tuple[np.ndarray, np.ndarray]]:
    """{level >= 1: (member ids int64 [n_l], neighbours int32 [n_l, M] padded with -1)} — the shape csr_from_padded takes."""
    out = {}
    base = g.node_offsets[:-1].astype(np.int64)
    for l in range(1, int(g.max_level) + 1):
        ids = np.nonzero(g.levels > l)[0].astype(np.int64)
        if ids.size == 0:
            continue
        st = g.level_ptr[base[ids] + l].astype(np.int64)
        en = g.level_ptr[base[ids] + l + 1].astype(np.int64)
        deg = en - st
        nb = np.full((ids.size, max(M, int(deg.max()) if deg.size else 0)), -1, np.int32)
        col = np.arange(nb.shape[1])[None, :]
        m = col < deg[:, None]
        nb[m] = g.neighbors[(st[:, None] + col)[m]]
        out[l] = (ids, nb)
    return out


@torch.no_grad()
def lists_from_search(search_fn: Callable[[np.ndarray], np.ndarray], x: np.ndarray, k: int, block: int = 65536) -> np.ndarray:
    """ids [n, k] (−1 padded): for every point the k nearest *other* points the searcher finds.  search_fn(q [b, d]) -> ids
    [b, >= k + 1] in ascending-distance order (the point itself is usually first and is dropped wherever it appears)."""
    n = x.shape[0]
    out = np.full((n, k), -1, np.int64)
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        ids = np.asarray(search_fn(x[b0:b1]), np.int64)
        me = np.arange(b0, b1)[:, None]
        keep = (ids != me) & (ids >= 0)
        rank = np.cumsum(keep, 1) - 1
        sel = keep & (rank < k)
        rows = np.nonzero(sel)[0]
        out[b0 + rows, rank[sel]] = ids[sel]
    return out


@torch.no_grad()
def rebuild_level0(x, g: CSRGraph, cand: np.ndarray, M: int = 32, alpha: float = 1.0, union_factor: int = 2,
                   device: str | None = None) -> CSRGraph:
    """New CSRGraph whose level-0 adjacency is built from `cand` [n, k] (−1 padded) with the batch builder's own steps
    (neighbour-selection heuristic, reverse edges, re-prune to 2M); levels, upper levels and the entry point are kept."""
    metric_ip = g.metric_type == METRIC_INNER_PRODUCT
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, np.float32))
    xt = xt.to(dev, torch.float32)
    ci = torch.from_numpy(np.ascontiguousarray(cand)).to(dev)
    n, k = ci.shape
    cd = torch.empty((n, k), dtype=torch.float32, device=dev)
    xs = xt.half() if xt.is_cuda else xt
    sq = (xt * xt).sum(1)
    for b0 in range(0, n, 1 << 16):  # exact distances of the candidates, row blocks
        b1 = min(n, b0 + (1 << 16))
        cc = ci[b0:b1].clamp(min=0)
        ip = torch.bmm(xs[cc], xs[b0:b1].unsqueeze(2)).squeeze(2).float()
        d = -ip if metric_ip else (sq[b0:b1, None] + sq[cc] - 2 * ip)
        cd[b0:b1] = torch.where(ci[b0:b1] >= 0, d, torch.full_like(d, float("inf")))
    o = torch.argsort(cd, dim=1, stable=True)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    cap = 2 * M
    fwd = _heuristic_prune(xt, ci, cd, cap, metric_ip, fill=False, alpha=alpha)
    ui, ud = _add_reverse_and_cap(xt, fwd, union_factor * cap, metric_ip)
    both = _heuristic_prune(xt, ui, ud, cap, metric_ip, fill=True, alpha=alpha)
    level0 = both.cpu().numpy().astype(np.int32)
    return csr_from_padded(g.d, g.metric_type, g.levels, level0, upper_levels(g, M), g.entry_point, M=M)


def level0_padded(g: CSRGraph, width: int) -> np.ndarray:
    """Current level-0 adjacency as int64 [n, width] padded with -1."""
    base = g.node_offsets[:-1].astype(np.int64)
    st = g.level_ptr[base].astype(np.int64)
    deg = g.level_ptr[base + 1].astype(np.int64) - st
    out = np.full((g.ntotal, max(width, int(deg.max()) if deg.size else 0)), -1, np.int64)
    col = np.arange(out.shape[1])[None, :]
    m = col < deg[:, None]
    out[m] = g.neighbors[(st[:, None] + col)[m]]
    return out


def refine_graph_by_search(x, g: CSRGraph, make_search_fn: Callable[[CSRGraph], Callable[[np.ndarray], np.ndarray]], M: int = 32,
                           k: int = 48, rounds: int = 1, device: str | None = None) -> CSRGraph:
    """`rounds` sweeps of: search the current graph with every point -> candidate lists -> rebuild level 0.
    make_search_fn(graph) returns the batch searcher over that graph (it owns ef and the result width k + 1).  The current
    adjacency stays in the candidate set: it carries the long links (nearest members of the sparser levels) that pure
    nearest-neighbour lists lack and that keep the graph navigable."""
    xn = x.cpu().numpy() if isinstance(x, torch.Tensor) else np.ascontiguousarray(x, np.float32)
    for _ in range(rounds):
        found = lists_from_search(make_search_fn(g), xn, k)
        cand = np.concatenate([found, level0_padded(g, 2 * M)], axis=1)
        cand.sort(axis=1)                      # group equal ids, then blank the repeats
        cand[:, 1:][cand[:, 1:] == cand[:, :-1]] = -1
        g = rebuild_level0(x, g, cand, M=M, device=device)
    return g


def gpu_searcher(vectors_device_ptr: int, workdir, ef: int = 128, k: int = 48, device: int = 0):
    """make_search_fn for refine_graph_by_search backed by the CUDA stored-vector search: writes the current graph in the
    reference's CSR format, opens it through the C ABI, attaches the (device-resident) vectors and searches with
    recompute_embeddings=False.  Untested at the time of writing (round 1 ran out of GPU time); the pieces it calls
    (write_compact_index, capi.Index, set_vectors_device, search) are the ones bench.py's stored-vector extra uses."""
    from pathlib import Path

    from . import capi, csr

    def make(g: CSRGraph):
        f = Path(workdir) / "refine_sweep.index"
        csr.write_compact_index(str(f), g)
        idx = capi.Index(str(f), device)
        idx.set_vectors_device(vectors_device_ptr)
        params = capi.make_params(ef, 1, 0, True, recompute=False)

        def search(q: np.ndarray) -> np.ndarray:
            return idx.search(np.ascontiguousarray(q, np.float32), k + 1, params)[1]

        search.close = idx.close  # type: ignore[attr-defined]
        return search

    return make


@torch.no_grad()
def prune_degrees(x, g: CSRGraph, M: int = 32, hub_fraction: float = 0.02, low: tuple[int, int] = (6, 7), seed: int = 789,
                  by: str = "in", device: str | None = None) -> tuple[CSRGraph, np.ndarray]:
    """High-degree-preserving pruning of the level-0 adjacency — the per-node degree cap `ems` of the reference's builder
    (faiss/IndexHNSW.cpp:130-225, hnsw.ems; consumed at faiss/impl/HNSW.cpp:762-763 where the new node's candidate list is
    shrunk to ems[pt_id] instead of M0).  The reference ships that branch switched off (`bool prune = false`), with the
    policy written out beside it: nodes whose degree is in the top 2 % keep M0 = 2M links, every other node keeps 6 or 7
    (`6 + rng.rand_int(2)`), chosen from its links by the usual neighbour-selection heuristic (shrink_neighbor_list).  This
    is that policy applied to a finished graph, as BASELINE.json's pruned-degree configuration (C5) needs:

      hubs  = the `hub_fraction` of nodes with the largest degree (`by` = "in": number of level-0 lists a node appears in —
              what makes a node a hub for the traversal; "out": length of its own list, which is what a degree file written
              from a stored graph holds)
      ems_i = 2M for hubs, low[0] + (0 or 1) for the others
      list_i <- neighbour-selection heuristic over list_i (ascending distance), at most ems_i kept, no fill

    Upper levels, levels and the entry point are untouched (the reference sets ems = infinity above level 0).  Returns the
    new graph and ems [n]."""
    if by not in ("in", "out"):
        raise ValueError("by must be 'in' or 'out'")
    metric_ip = g.metric_type == METRIC_INNER_PRODUCT
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, np.float32))
    xt = xt.to(dev, torch.float32)
    n = g.ntotal
    l0 = level0_padded(g, 2 * M)
    deg_out = (l0 >= 0).sum(1)
    degree = np.bincount(l0[l0 >= 0], minlength=n) if by == "in" else deg_out
    n_hub = max(1, int(n * hub_fraction))
    threshold = np.sort(degree)[::-1][n_hub - 1]
    hub = degree >= threshold
    rng = np.random.default_rng(seed)
    ems = np.where(hub, 2 * M, low[0] + rng.integers(0, low[1] - low[0] + 1, n)).astype(np.int32)
    # candidates of every node = its own list in ascending distance
    ci = torch.from_numpy(l0).to(dev)
    cd = torch.full(ci.shape, float("inf"), dtype=torch.float32, device=dev)
    sq = (xt * xt).sum(1)
    for b0 in range(0, n, 1 << 16):
        b1 = min(n, b0 + (1 << 16))
        cc = ci[b0:b1].clamp(min=0)
        ip = torch.bmm(xt[cc], xt[b0:b1].unsqueeze(2)).squeeze(2)
        d = -ip if metric_ip else (sq[b0:b1, None] + sq[cc] - 2 * ip)
        cd[b0:b1] = torch.where(ci[b0:b1] >= 0, d, cd[b0:b1])
    o = torch.argsort(cd, dim=1, stable=True)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    kept = _heuristic_prune(xt, ci, cd, low[1], metric_ip, fill=False).cpu().numpy()   # [n, low[1]], selection order
    col = np.arange(kept.shape[1])[None, :]
    kept[col >= ems[:, None]] = -1                                                       # the nodes drawn with the smaller cap
    new0 = np.full_like(l0, -1)
    new0[:, : kept.shape[1]] = kept
    new0[hub] = l0[hub]
    g2 = csr_from_padded(g.d, g.metric_type, g.levels, new0.astype(np.int32), upper_levels(g, M), g.entry_point, M=M)
    return g2, ems
