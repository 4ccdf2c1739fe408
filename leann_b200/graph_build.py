"""Tooling (not the hot path): builds an HNSW-shaped graph in the reference's CSR format
from a matrix of embeddings, with torch ops (CUDA when available, CPU for small tests).

SURVEY.md §8(f) row 1 — the reference builds with faiss' incremental insertion
(faiss/IndexHNSW.cpp:59-280, impl/HNSW.cpp:426-894), hours at 10 M points on CPU.  The search
path only needs *a* navigable HNSW-format graph that both the reference CPU search and the CUDA
search traverse identically, so this builder takes the batch route:
  * levels drawn like HNSW::random_level (impl/HNSW.cpp:182-193, level_mult = 1/ln M);
  * per level: candidates = exact k-nearest members (blocked GEMM + top-k) plus the nearest members
    of the next two levels up (sparser samples -> longer links, standing in for HNSW's early
    insertions), pruned with the HNSW neighbour-selection heuristic (keep a candidate only if it is closer to
    the node than to every already kept neighbour — shrink_neighbor_list, impl/HNSW.cpp:426-480),
    then reverse edges added and each list capped to M (2M on level 0) nearest;
  * entry point = first node of the top level.
Output is a leann_b200.csr.CSRGraph, written with csr.write_compact_index().
"""
from __future__ import annotations

import numpy as np
import torch

from .csr import METRIC_INNER_PRODUCT, METRIC_L2, CSRGraph, csr_from_padded


def draw_levels(n: int, M: int, seed: int = 12345) -> np.ndarray:
    rng = np.random.default_rng(seed)
    mult = 1.0 / np.log(M)
    lv = np.floor(-np.log(1.0 - rng.random(n)) * mult).astype(np.int32) + 1
    return np.minimum(lv, 8)


@torch.no_grad()
def _knn(xq: torch.Tensor, xb: torch.Tensor, k: int, metric_ip: bool, self_pos: torch.Tensor | None = None,
         block: int = 4096) -> tuple[torch.Tensor, torch.Tensor]:
    """Exact kNN of every row of xq among the rows of xb.  self_pos[i] = row of xb that IS query i
    (excluded) or -1.  Returns (idx [n,k] into xb, 'distance' [n,k]) with distance = -ip or squared
    L2 (smaller = closer), ascending."""
    n = xq.shape[0]
    k = max(1, min(k, xb.shape[0] - 1))
    xqs, xbs = (xq.half(), xb.half()) if xq.is_cuda else (xq, xb)
    sqq, sqb = (xq * xq).sum(1), (xb * xb).sum(1)
    idx = torch.empty((n, k), dtype=torch.int64, device=xq.device)
    dist = torch.empty((n, k), dtype=torch.float32, device=xq.device)
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        ip = (xqs[b0:b1] @ xbs.T).float()
        d = -ip if metric_ip else (sqq[b0:b1, None] + sqb[None, :] - 2 * ip)
        if self_pos is not None:
            sp = self_pos[b0:b1]
            m = sp >= 0
            d[torch.nonzero(m)[:, 0], sp[m]] = float("inf")
        dv, di = torch.topk(d, k, dim=1, largest=False)
        idx[b0:b1] = di
        dist[b0:b1] = dv
    return idx, dist


@torch.no_grad()
def _knn_ivf(x: torch.Tensor, k: int, metric_ip: bool, n_clusters: int | None = None, n_probe: int = 12,
             seed: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """Approximate kNN of every row among all rows for large n: k-means partition, then exact search of each
    cluster's members inside the `n_probe` clusters nearest to its centroid (~n_probe / n_clusters of the
    data).  Same return convention as _knn (self excluded)."""
    n, d = x.shape
    dev = x.device
    C = n_clusters or int(round((n / 10000) ** 0.5 * 32))  # ~1000 clusters at 10 M
    C = max(1, min(max(8, min(C, n // 64)), n))  # tiny member sets (upper levels when IVF is forced): never more cells than points
    g = torch.Generator(device=dev).manual_seed(seed)
    samp = x[torch.randint(0, n, (min(n, 256 * C),), device=dev, generator=g)]
    cent = samp[torch.randperm(samp.shape[0], device=dev, generator=g)[:C]].clone()
    for _ in range(8):  # Lloyd iterations on the sample
        sc = samp @ cent.T if metric_ip else -(torch.cdist(samp, cent) ** 2)
        a = sc.argmax(1)
        cent.zero_().index_add_(0, a, samp)
        cnt = torch.bincount(a, minlength=C).clamp(min=1).unsqueeze(1)
        cent = cent / cnt
        if metric_ip:
            cent = torch.nn.functional.normalize(cent, dim=1)
    assign = torch.empty(n, dtype=torch.int64, device=dev)
    for b0 in range(0, n, 1 << 18):
        xb = x[b0:b0 + (1 << 18)]
        sc = xb @ cent.T if metric_ip else -(torch.cdist(xb, cent) ** 2)
        assign[b0:b0 + (1 << 18)] = sc.argmax(1)
    order = torch.argsort(assign, stable=True)
    counts = torch.bincount(assign, minlength=C)
    starts = torch.cumsum(counts, 0) - counts
    csc = cent @ cent.T if metric_ip else -(torch.cdist(cent, cent) ** 2)
    probes = torch.topk(csc, min(n_probe, C), dim=1).indices  # includes the cluster itself
    xs = x.half()
    sq = (x * x).sum(1)
    k = min(k, n - 1)
    idx = torch.empty((n, k), dtype=torch.int64, device=dev)
    dist = torch.full((n, k), float("inf"), dtype=torch.float32, device=dev)
    counts_h, starts_h = counts.tolist(), starts.tolist()
    for c in range(C):
        if counts_h[c] == 0:
            continue
        qi = order[starts_h[c]: starts_h[c] + counts_h[c]]
        bi = torch.cat([order[starts_h[p]: starts_h[p] + counts_h[p]] for p in probes[c].tolist()])
        xb = xs[bi]
        kk = min(k, bi.numel() - 1)
        for q0 in range(0, qi.numel(), 8192):
            qq = qi[q0:q0 + 8192]
            ip = (xs[qq] @ xb.T).float()
            dm = -ip if metric_ip else (sq[qq][:, None] + sq[bi][None, :] - 2 * ip)
            dm[qq[:, None] == bi[None, :]] = float("inf")  # exclude self
            dv, di = torch.topk(dm, kk, dim=1, largest=False)
            idx[qq, :kk] = bi[di]
            dist[qq, :kk] = dv
            if kk < k:
                idx[qq, kk:] = -1
    return idx, dist


@torch.no_grad()
def _refine_knn(x: torch.Tensor, idx: torch.Tensor, dist: torch.Tensor, metric_ip: bool, rounds: int = 6, fan: int = 24,
                block: int = 4096) -> tuple[torch.Tensor, torch.Tensor]:
    """NN-descent style repair of approximate kNN lists: a node's candidates are the lists of its `fan` nearest
    current neighbours ("a neighbour of a neighbour is likely a neighbour"), scored exactly and merged with the current
    list.  One round reads fan*k rows of x per node — seconds at 10 M points — and lifts the recall of the
    partition-restricted lists of _knn_ivf, whose misses are the true neighbours that fell into unprobed cells
    (1 M points, 12 probes: list recall 0.41 -> 0.49 -> 0.58 -> 0.63 over three rounds at fan 16, search recall@10
    0.870 -> 0.911 against 0.936 with exact lists; profiles/r01e_ivf_refine_recall_1m.log)."""
    n, k = idx.shape
    xs = x.half()
    sq = (x * x).sum(1)
    for _ in range(rounds):
        new_idx = torch.empty_like(idx)
        new_dist = torch.empty_like(dist)
        for b0 in range(0, n, block):
            b1 = min(n, b0 + block)
            cur = idx[b0:b1]                                   # [B, k]
            nb = cur[:, :fan].clamp(min=0)
            cand = idx[nb.reshape(-1)].reshape(b1 - b0, -1)    # [B, fan*k]
            cand = torch.cat([cur, cand], 1)
            me = torch.arange(b0, b1, device=x.device)[:, None]
            cand = torch.where(cand == me, torch.full_like(cand, -1), cand)
            cand, _ = torch.sort(cand, dim=1)
            dup = torch.zeros_like(cand, dtype=torch.bool)
            dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
            bad = dup | (cand < 0)
            cc = cand.clamp(min=0)
            ip = torch.bmm(xs[cc], xs[b0:b1].unsqueeze(2)).squeeze(2).float()   # [B, C]
            d = -ip if metric_ip else (sq[b0:b1, None] + sq[cc] - 2 * ip)
            d = torch.where(bad, torch.full_like(d, float("inf")), d)
            dv, di = torch.topk(d, k, dim=1, largest=False)
            sel = torch.gather(cand, 1, di)
            new_idx[b0:b1] = torch.where(torch.isinf(dv), torch.full_like(sel, -1), sel)
            new_dist[b0:b1] = dv
        idx, dist = new_idx, new_dist
    return idx, dist


@torch.no_grad()
def _heuristic_prune(x: torch.Tensor, cand: torch.Tensor, cdist: torch.Tensor, keep: int, metric_ip: bool,
                     fill: bool, alpha: float = 1.0, block: int = 8192) -> torch.Tensor:
    """HNSW neighbour selection over candidates sorted by distance (padding = -1 / +inf at the end);
    returns idx [n, keep] padded with -1.  fill=True tops the list up with the nearest pruned candidates."""
    n, K = cand.shape
    out = torch.full((n, keep), -1, dtype=torch.int64, device=x.device)
    xs = x.half() if x.is_cuda else x
    sq = (x * x).sum(1)
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        c = cand[b0:b1]
        valid = c >= 0
        c = c.clamp(min=0)
        cv = xs[c]  # [b, K, d]
        ip = torch.bmm(cv, cv.transpose(1, 2)).float()
        pd = -ip if metric_ip else (sq[c][:, :, None] + sq[c][:, None, :] - 2 * ip)  # dist(c_i, c_j)
        dn = cdist[b0:b1]  # dist(node, c_j)
        if alpha != 1.0:
            # Vamana-style slack: a kept neighbour dominates c_j only if it is alpha times closer.  Needs a
            # non-negative distance: for inner product use 2 - 2 ip (exact for unit vectors).
            if metric_ip:
                pd, dn = 2 + 2 * pd, 2 + 2 * dn
            pd = pd * alpha
        kept = torch.zeros((b1 - b0, K), dtype=torch.bool, device=x.device)
        nk = torch.zeros(b1 - b0, dtype=torch.int64, device=x.device)
        for j in range(K):
            # c_j is dominated if some kept c_i is closer to c_j than the node is
            dom = ((pd[:, :, j] < dn[:, j:j + 1]) & kept).any(1)
            ok = (~dom) & (nk < keep) & valid[:, j]
            kept[:, j] = ok
            nk += ok.long()
        # kept first (in distance order), then — if fill — the nearest pruned candidates
        rank = torch.arange(K, device=x.device)[None, :].expand_as(kept)
        prio = torch.where(kept, rank, rank + K)
        prio = torch.where(valid & (kept | fill), prio, torch.full_like(prio, 4 * K))
        order = torch.argsort(prio, dim=1)[:, :keep]
        sel = torch.gather(c, 1, order)
        selp = torch.gather(prio, 1, order)
        sel = torch.where(selp < 4 * K, sel, torch.full_like(sel, -1))
        out[b0:b1, : sel.shape[1]] = sel
    return out


@torch.no_grad()
def _add_reverse_and_cap(x: torch.Tensor, nbr: torch.Tensor, cap: int, metric_ip: bool):
    """Symmetrise: union of forward and reverse edges, keep the `cap` nearest per node, sorted by
    distance.  Returns (idx [n, cap] padded -1, dist [n, cap] padded +inf)."""
    n, K = nbr.shape
    dev = x.device
    src = torch.arange(n, device=dev)[:, None].expand(n, K).reshape(-1)
    dst = nbr.reshape(-1)
    m = dst >= 0
    src, dst = src[m], dst[m]
    a = torch.cat([src, dst])
    b = torch.cat([dst, src])
    del src, dst, m
    key = a * n + b
    del a, b
    key = torch.unique(key)
    a, b = key // n, key % n
    del key
    xs = x.half() if x.is_cuda else x
    d = torch.empty(a.numel(), dtype=torch.float32, device=dev)
    step = 1 << 22  # edge chunks: the gathered [chunk, dim] operands stay small
    for e0 in range(0, a.numel(), step):
        xa, xb = xs[a[e0:e0 + step]].float(), xs[b[e0:e0 + step]].float()
        d[e0:e0 + step] = -(xa * xb).sum(1) if metric_ip else ((xa - xb) ** 2).sum(1)
    # sort by (a, d): stable two-pass
    o = torch.argsort(d, stable=True)
    a, b, d = a[o], b[o], d[o]
    del o
    o = torch.argsort(a, stable=True)
    a, b, d = a[o], b[o], d[o]
    del o
    counts = torch.bincount(a, minlength=n)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(a.numel(), device=dev) - starts[a]
    keepm = rank < cap
    out = torch.full((n, cap), -1, dtype=torch.int64, device=dev)
    out[a[keepm], rank[keepm]] = b[keepm]
    outd = torch.full((n, cap), float("inf"), dtype=torch.float32, device=dev)
    outd[a[keepm], rank[keepm]] = d[keepm].float()
    return out, outd


@torch.no_grad()
def _build_level_exact(x, levels, members, l, cap, metric_ip, knn_factor=1.5, n_scales=2, alpha=1.0, union_factor=2,
                       ivf_threshold=2_500_000, ivf_refine_rounds=6) -> np.ndarray:
    """One level of the batch builder: links among `members` (global ids, all with levels > l).  Returns global
    neighbour ids int32 [len(members), cap], -1 padded."""
    dev = x.device
    if len(members) <= 1:
        return np.full((len(members), cap), -1, np.int32)
    mt = torch.from_numpy(members).to(dev)
    xm = x[mt]
    nm = len(members)
    # multi-scale candidates: nearest members, plus nearest among the (sparser) members of
    # the next levels up.  The sparse samples play the role of HNSW's early insertions and
    # supply the long links a pure kNN graph lacks.
    if nm > ivf_threshold:  # brute force is O(n^2): partition-restricted search beyond a few million points
        ci, cd = _knn_ivf(xm, int(cap * knn_factor), metric_ip)
        ci, cd = _refine_knn(xm, ci, cd, metric_ip, rounds=ivf_refine_rounds)
    else:
        ci, cd = _knn(xm, xm, int(cap * knn_factor), metric_ip, torch.arange(nm, device=dev))
    cis, cds = [ci], [cd]
    for s_up in range(1, n_scales + 1):
        sub = np.nonzero(levels[members] > l + s_up)[0]
        if len(sub) < 2:
            break
        st = torch.from_numpy(sub).to(dev)
        pos = torch.full((nm,), -1, dtype=torch.int64, device=dev)
        pos[st] = torch.arange(len(sub), device=dev)
        di, dv = _knn(xm, xm[st], max(2, cap // 2), metric_ip, pos)
        cis.append(st[di])
        cds.append(dv)
    ci, cd = torch.cat(cis, 1), torch.cat(cds, 1)
    o = torch.argsort(cd, dim=1, stable=True)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    o = torch.argsort(ci, dim=1, stable=True)  # group equal ids (distance order kept inside a group)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    dup = torch.zeros_like(ci, dtype=torch.bool)
    dup[:, 1:] = ci[:, 1:] == ci[:, :-1]
    ci = torch.where(dup, torch.full_like(ci, -1), ci)
    cd = torch.where(dup, torch.full_like(cd, float("inf")), cd)
    o = torch.argsort(cd, dim=1, stable=True)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    fwd = _heuristic_prune(xm, ci, cd, cap, metric_ip, fill=False, alpha=alpha)
    ui, ud = _add_reverse_and_cap(xm, fwd, union_factor * cap, metric_ip)
    both = _heuristic_prune(xm, ui, ud, cap, metric_ip, fill=True, alpha=alpha)
    gl = torch.where(both >= 0, mt[both.clamp(min=0)], torch.full_like(both, -1))
    return gl.cpu().numpy().astype(np.int32)


@torch.no_grad()
def build_hnsw_graph(emb, M: int = 32, metric: str = "mips", seed: int = 12345, device: str | None = None,
                     knn_factor: float = 1.5, n_scales: int = 2, alpha: float = 1.0, union_factor: int = 2,
                     ivf_threshold: int = 2_500_000, ivf_refine_rounds: int = 6, verbose: bool = False) -> CSRGraph:
    metric_ip = metric.lower() in ("mips", "cosine", "ip")
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    x = emb if isinstance(emb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(emb, np.float32))
    x = x.to(dev, torch.float32)
    n, d = x.shape
    levels = draw_levels(n, M, seed)
    max_level = int(levels.max()) - 1
    level0 = None
    upper = {}
    for l in range(max_level, -1, -1):
        members = np.nonzero(levels > l)[0]
        cap = 2 * M if l == 0 else M
        nb = _build_level_exact(x, levels, members, l, cap, metric_ip, knn_factor, n_scales, alpha, union_factor,
                                ivf_threshold, ivf_refine_rounds)
        if l == 0:
            level0 = nb
        else:
            upper[l] = (members.astype(np.int64), nb)
        if verbose:
            print(f"  level {l}: {len(members)} nodes, mean degree {(nb >= 0).sum(1).mean():.1f}")
    entry = int(np.nonzero(levels == levels.max())[0][0])
    return csr_from_padded(d, METRIC_INNER_PRODUCT if metric_ip else METRIC_L2, levels, level0, upper, entry, M=M)


# ------------------------------------------------------------------------------------------------
# Incremental (insertion-as-search) builder: the reference's construction, batch-parallel on the GPU.
# Device stages: csrc/graph_build.cu (lb2_build_insert_search, lb2_build_select); this file owns the
# batch schedule and the link bookkeeping (torch tensors as device buffers).

@torch.no_grad()
def _select_rows(xs: torch.Tensor, me: torch.Tensor, cand: torch.Tensor, cd: torch.Tensor, keep: int, metric_ip: bool,
                 sq: torch.Tensor | None, block: int = 4096, fill: int = 0):
    """Neighbour-selection heuristic (HNSW::shrink_neighbor_list, faiss/impl/HNSW.cpp:426-468) over candidate rows
    sorted by distance (cand int32 [b, K] -1 padded at the end, cd fp32 [b, K]).  Pairwise candidate distances come from
    one batched GEMM per row block; the sequential keep/drop scan is lb2_build_select (one warp per row).
    Returns (ids int32 [b, keep] -1 padded, dist fp32 [b, keep])."""
    from . import capi

    b, K = cand.shape
    out_i = torch.empty((b, keep), dtype=torch.int32, device=cand.device)
    out_d = torch.empty((b, keep), dtype=torch.float32, device=cand.device)
    for b0 in range(0, b, block):
        b1 = min(b, b0 + block)
        c = cand[b0:b1].clamp(min=0).long()
        cv = xs[c]                                             # [r, K, d] fp16
        pd = torch.bmm(cv, cv.transpose(1, 2))                 # [r, K, K] fp16 inner products
        if metric_ip:
            pd = -pd
        else:
            s = sq[c]
            pd = (s[:, :, None] + s[:, None, :]).half() - 2 * pd
        pd = pd.contiguous()
        ci = cand[b0:b1].contiguous()
        di = cd[b0:b1].contiguous()
        capi.build_select(pd.data_ptr(), False, di.data_ptr(), ci.data_ptr(), b1 - b0, K, keep,
                          out_i[b0:b1].data_ptr(), out_d[b0:b1].data_ptr(), fill=fill)
        del cv, pd
    return out_i, out_d


@torch.no_grad()
def _merge_incoming(xs, sq, adj, adjd, dst, src, dd, cap: int, metric_ip: bool, r_in: int = 32):
    """Reverse links of a batch (add_link, faiss/impl/HNSW.cpp:510-552, applied per target instead of per edge): the
    incoming edges (dst <- src at distance dd) are appended to dst's list while there is room; a list that would
    overflow is re-selected with the heuristic over (current list + its r_in nearest incoming)."""
    if dst.numel() == 0:
        return
    dev = adj.device
    o = torch.argsort(dd, stable=True)
    dst, src, dd = dst[o], src[o], dd[o]
    o = torch.argsort(dst, stable=True)
    dst, src, dd = dst[o], src[o], dd[o]
    uniq, inv, counts = torch.unique_consecutive(dst, return_inverse=True, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(dst.numel(), device=dev) - starts[inv]
    m = rank < r_in
    u = uniq.numel()
    inc_i = torch.full((u, r_in), -1, dtype=torch.int32, device=dev)
    inc_d = torch.full((u, r_in), float("inf"), dtype=torch.float32, device=dev)
    inc_i[inv[m], rank[m]] = src[m].int()
    inc_d[inv[m], rank[m]] = dd[m]
    ex_i, ex_d = adj[uniq], adjd[uniq]
    mi = torch.cat([ex_i, inc_i], 1)
    md = torch.cat([ex_d, inc_d], 1)
    md = torch.where(mi >= 0, md, torch.full_like(md, float("inf")))
    o = torch.argsort(md, dim=1, stable=True)
    mi, md = torch.gather(mi, 1, o), torch.gather(md, 1, o)
    total = (mi >= 0).sum(1)
    over = total > cap
    new_i, new_d = mi[:, :cap].clone(), md[:, :cap].clone()
    if bool(over.any()):
        rows = torch.nonzero(over)[:, 0]
        si, sd = _select_rows(xs, uniq[rows], mi[rows], md[rows], cap, metric_ip, sq)
        new_i[rows], new_d[rows] = si, sd
    adj[uniq] = new_i
    adjd[uniq] = new_d


def upper_level_arrays(levels: np.ndarray, upper: dict, M: int, dev):
    """The upper levels in lb2_build_insert_search's layout: up_row int32 [n] = first row of a node's level-1 list in
    up_adj int32 [rows, M] (level l = row up_row + l - 1), -1 for nodes that live on level 0 only."""
    n_up = np.maximum(np.asarray(levels).astype(np.int64) - 1, 0)
    up_row_h = np.where(n_up > 0, np.cumsum(n_up) - n_up, -1).astype(np.int32)
    up_adj = torch.full((max(1, int(n_up.sum())), M), -1, dtype=torch.int32, device=dev)
    for l, (ids, nb) in upper.items():
        rows = torch.from_numpy(up_row_h[ids].astype(np.int64) + l - 1).to(dev)
        up_adj[rows] = torch.from_numpy(np.ascontiguousarray(nb[:, :M], np.int32)).to(dev)
    return torch.from_numpy(up_row_h).to(dev), up_adj


@torch.no_grad()
def build_hnsw_graph_incremental(emb, M: int = 32, metric: str = "mips", seed: int = 12345, device: str | None = None,
                                 ef_construction: int = 200, growth: float = 0.25, min_seed: int = 20000,
                                 max_batch: int = 1 << 20, sweeps: int = 0, fill: int = 0, cover: int = 0,
                                 verbose: bool = False) -> CSRGraph:
    """HNSW construction the way the reference does it — every point is inserted by searching the graph built so far
    (hnsw_add_vertices, faiss/IndexHNSW.cpp:59-280: upper levels first, level-0-only points last) — run batch-parallel
    on the GPU: the points of a batch search concurrently (lb2_build_insert_search, one warp per point), select their
    links with the heuristic (lb2_build_select) and their reverse links are merged per target.  A batch never exceeds
    `growth` x the points already inserted, so a new point misses at most that share of its potential neighbours at
    insertion time; later insertions link back to it, and `sweeps` optional passes re-search every point on the finished
    graph to repair what the batches missed.  `fill` > 0 tops forward lists up to that many links with the nearest
    rejected candidates (the reference's keep_max_size_level0 idea with a settable floor).  `cover` > 0 runs that many
    rounds of level-1 coverage promotion after the level-0 graph is complete (see below).  The upper levels (3 % of the points) and the level-0 seed among them are
    built exactly (brute-force lists) by the batch builder above.  CUDA only."""
    from . import capi

    metric_ip = metric.lower() in ("mips", "cosine", "ip")
    dev = torch.device(device or "cuda")
    if dev.type != "cuda":
        raise RuntimeError("build_hnsw_graph_incremental needs a CUDA device (csrc/graph_build.cu); use build_hnsw_graph on CPU")
    x = emb if isinstance(emb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(emb, np.float32))
    x = x.to(dev, torch.float32)
    n, d = x.shape
    if d % 8:
        raise ValueError("dimension must be a multiple of 8")
    cap0 = 2 * M
    levels = draw_levels(n, M, seed)
    max_level = int(levels.max()) - 1
    upper = {}
    for l in range(max_level, 0, -1):
        members = np.nonzero(levels > l)[0]
        upper[l] = (members.astype(np.int64), _build_level_exact(x, levels, members, l, M, metric_ip))
    entry = int(np.nonzero(levels == levels.max())[0][0])
    # seed of level 0: every node that also lives above, topped up with the first level-0-only points
    is_seed = levels > 1
    short = min(n, min_seed) - int(is_seed.sum())
    if short > 0:
        is_seed[np.nonzero(~is_seed)[0][:short]] = True
    seed_ids = np.nonzero(is_seed)[0]
    rest = np.nonzero(~is_seed)[0]
    seed_nb = _build_level_exact(x, levels, seed_ids, 0, cap0, metric_ip)
    xs = x.half().contiguous()
    sq = None if metric_ip else (x * x).sum(1)
    adj = torch.full((n, cap0), -1, dtype=torch.int32, device=dev)
    adjd = torch.full((n, cap0), float("inf"), dtype=torch.float32, device=dev)
    st = torch.from_numpy(seed_ids).to(dev)
    snb = torch.from_numpy(seed_nb).to(dev)
    adj[st] = snb
    for b0 in range(0, len(seed_ids), 1 << 16):  # distances of the seed links
        rows = st[b0:b0 + (1 << 16)]
        c = snb[b0:b0 + (1 << 16)].clamp(min=0).long()
        ip = torch.bmm(xs[c], xs[rows].unsqueeze(2)).squeeze(2).float()
        dd = -ip if metric_ip else (sq[rows][:, None] + sq[c] - 2 * ip)
        adjd[rows] = torch.where(snb[b0:b0 + (1 << 16)] >= 0, dd, torch.full_like(dd, float("inf")))
    up_row, up_adj = upper_level_arrays(levels, upper, M, dev)
    ws_bytes = capi.build_workspace_bytes(ef_construction, cap0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def search(points: torch.Tensor):
        b = points.numel()
        oi = torch.empty((b, ef_construction), dtype=torch.int32, device=dev)
        od = torch.empty((b, ef_construction), dtype=torch.float32, device=dev)
        capi.build_insert_search(xs.data_ptr(), n, d, metric_ip, adj.data_ptr(), cap0, up_row.data_ptr(), up_adj.data_ptr(), M,
                                 entry, max_level, points.data_ptr(), b, ef_construction, oi.data_ptr(), od.data_ptr(),
                                 ws.data_ptr(), ws_bytes)
        return oi, od

    n_in = len(seed_ids)
    pos = 0
    rest_t = torch.from_numpy(rest.astype(np.int32)).to(dev)
    while pos < len(rest):
        b = int(min(max(1024, growth * n_in), max_batch, len(rest) - pos))
        pts = rest_t[pos:pos + b].contiguous()
        ci, cd = search(pts)
        fi, fd = _select_rows(xs, pts, ci, cd, cap0, metric_ip, sq, fill=fill)
        p64 = pts.long()
        adj[p64] = fi
        adjd[p64] = fd
        v = fi >= 0
        src = p64[:, None].expand_as(fi)[v]
        _merge_incoming(xs, sq, adj, adjd, fi[v].long(), src, fd[v], cap0, metric_ip)
        pos += b
        n_in += b
        if verbose:
            print(f"  inserted {n_in}/{n} (batch {b}), mean degree {(adj[p64] >= 0).sum(1).float().mean().item():.1f}", flush=True)
    for _ in range(sweeps):  # repair pass: re-search every point on the finished graph, merge into its list
        for b0 in range(0, n, max_batch):
            pts = torch.arange(b0, min(n, b0 + max_batch), dtype=torch.int32, device=dev)
            ci, cd = search(pts)
            p64 = pts.long()
            mi = torch.cat([adj[p64], ci], 1)
            md = torch.cat([adjd[p64], cd], 1)
            o = torch.argsort(mi, dim=1, stable=True)          # blank repeated ids
            mi, md = torch.gather(mi, 1, o), torch.gather(md, 1, o)
            dup = torch.zeros_like(mi, dtype=torch.bool)
            dup[:, 1:] = mi[:, 1:] == mi[:, :-1]
            md = torch.where(dup | (mi < 0), torch.full_like(md, float("inf")), md)
            mi = torch.where(dup, torch.full_like(mi, -1), mi)
            o = torch.argsort(md, dim=1, stable=True)
            mi, md = torch.gather(mi, 1, o), torch.gather(md, 1, o)
            old = adj[p64].clone()
            fi, fd = _select_rows(xs, pts, mi.contiguous(), md.contiguous(), cap0, metric_ip, sq, fill=fill)
            adj[p64] = fi
            adjd[p64] = fd
            # reverse links for edges that are new
            isnew = (fi[:, :, None] != old[:, None, :]).all(2) & (fi >= 0)
            src = p64[:, None].expand_as(fi)[isnew]
            _merge_incoming(xs, sq, adj, adjd, fi[isnew].long(), src, fd[isnew], cap0, metric_ip)
    if cover > 0 and max_level >= 1:
        # Coverage of level 1.  Random level draws leave ~1/e of all neighbourhoods of ~M points without any member above
        # level 0, and a query whose neighbourhood has no such member can only be reached through level-0 links.  Promote a
        # spread-out subset of the uncovered points (no level >= 1 point in their own level-0 list; local minima by id, so
        # promoted points are not adjacent) to level 1 and rebuild that level.  The format and the search are unchanged:
        # levels are a free choice of the builder (the reference draws them at random, HNSW::random_level).
        lv = torch.from_numpy(levels.astype(np.int32)).to(dev)
        ids = torch.arange(n, device=dev)
        for _ in range(cover):
            up = lv > 1
            nb_up = torch.zeros(n, dtype=torch.bool, device=dev)
            for b0 in range(0, n, 1 << 20):
                a = adj[b0:b0 + (1 << 20)].long()
                nb_up[b0:b0 + (1 << 20)] = (up[a.clamp(min=0)] & (a >= 0)).any(1)
            unc = ~(up | nb_up)
            if not bool(unc.any()):
                break
            promote = torch.zeros(n, dtype=torch.bool, device=dev)
            for b0 in range(0, n, 1 << 20):
                a = adj[b0:b0 + (1 << 20)].long()
                rows = ids[b0:b0 + (1 << 20)]
                cand = torch.where((a >= 0) & unc[a.clamp(min=0)], a, torch.full_like(a, n))  # uncovered neighbours
                promote[b0:b0 + (1 << 20)] = unc[rows] & (rows < cand.min(1).values)
            lv = torch.where(promote, torch.full_like(lv, 2), lv)
            if verbose:
                print(f"  coverage round: {int(unc.sum())} uncovered points, {int(promote.sum())} promoted to level 1", flush=True)
        levels = lv.cpu().numpy().astype(np.int32)
        members = np.nonzero(levels > 1)[0]
        upper[1] = (members.astype(np.int64), _build_level_exact(x, levels, members, 1, M, metric_ip))
    level0 = adj.cpu().numpy()
    return csr_from_padded(d, METRIC_INNER_PRODUCT if metric_ip else METRIC_L2, levels, level0, upper, entry, M=M)
