"""Tooling (not the hot path): a Vamana-shaped flat graph + PQ codes in the reference's DiskANN file formats.

The reference builds with `diskannpy.build_disk_float_index` (diskann_backend.py:262-274; Vamana incremental
insertion with alpha-pruning, R = graph_degree = 32, L = complexity = 64), which is not buildable here.  The
search path only needs *a* navigable degree-bounded graph with a medoid entry point that the CPU oracle and the
CUDA searcher traverse identically, so this takes the same batch route as graph_build.py: multi-scale
k-nearest-neighbour candidates (exact nearest neighbours plus nearest members of two sparser random samples,
which supply the long links), the alpha-slack neighbour-selection rule, reverse edges, re-prune to R.
"""
from __future__ import annotations

import numpy as np
import torch

from . import diskann_format as dfmt
from .graph_build import _add_reverse_and_cap, _heuristic_prune, _knn, _knn_ivf, _refine_knn


@torch.no_grad()
def build_vamana_graph(coords, R: int = 32, alpha: float = 1.2, seed: int = 12345, device: str | None = None,
                       knn_factor: float = 1.5, ivf_threshold: int = 2_500_000) -> dfmt.VamanaGraph:
    """coords: [n, dim] float32 in the L2 space the index searches in (diskann_format.prepare_base)."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    x = coords if isinstance(coords, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(coords, np.float32))
    x = x.to(dev, torch.float32)
    n = x.shape[0]
    if n <= 1:
        return dfmt.VamanaGraph(np.full((n, R), -1, np.int32), 0)
    if n > ivf_threshold:
        ci, cd = _knn_ivf(x, int(R * knn_factor), False)
        ci, cd = _refine_knn(x, ci, cd, False)
    else:
        ci, cd = _knn(x, x, int(R * knn_factor), False, torch.arange(n, device=dev))
    cis, cds = [ci], [cd]
    rng = np.random.default_rng(seed)
    for frac in (32, 1024):
        m = n // frac
        if m < 2:
            break
        sub = torch.from_numpy(np.sort(rng.choice(n, m, replace=False))).to(dev)
        pos = torch.full((n,), -1, dtype=torch.int64, device=dev)
        pos[sub] = torch.arange(m, device=dev)
        di, dv = _knn(x, x[sub], max(2, R // 2), False, pos)
        cis.append(sub[di])
        cds.append(dv)
    ci, cd = torch.cat(cis, 1), torch.cat(cds, 1)
    o = torch.argsort(ci, dim=1, stable=True)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    dup = torch.zeros_like(ci, dtype=torch.bool)
    dup[:, 1:] = ci[:, 1:] == ci[:, :-1]
    ci = torch.where(dup, torch.full_like(ci, -1), ci)
    cd = torch.where(dup, torch.full_like(cd, float("inf")), cd)
    o = torch.argsort(cd, dim=1, stable=True)
    ci, cd = torch.gather(ci, 1, o), torch.gather(cd, 1, o)
    fwd = _heuristic_prune(x, ci, cd, R, False, fill=False, alpha=alpha)
    ui, ud = _add_reverse_and_cap(x, fwd, 2 * R, False)
    both = _heuristic_prune(x, ui, ud, R, False, fill=True, alpha=alpha)
    nb = both.cpu().numpy().astype(np.int32)
    # valid ids first
    order = np.argsort(nb < 0, axis=1, kind="stable")
    nb = np.take_along_axis(nb, order, axis=1)
    mean = x.mean(0, keepdim=True)
    medoid = int(((x - mean) ** 2).sum(1).argmin().item())
    return dfmt.VamanaGraph(nb, medoid)


def build_diskann_index(dirpath, name: str, emb: np.ndarray, metric: str = "mips", R: int = 32, n_chunks: int | None = None,
                        partition: bool = True, keep_disk_index: bool = True, seed: int = 0, device: str | None = None,
                        alpha: float = 1.2):
    """Embeddings -> the file set the reference's DiskannSearcher opens.  Returns (prefix, graph, coords, pq, codes,
    max_base_norm)."""
    coords, max_norm = dfmt.prepare_base(emb, metric)
    g = build_vamana_graph(coords, R=R, alpha=alpha, seed=seed + 12345, device=device)
    nch = n_chunks or dfmt.default_num_chunks(coords.shape[0], coords.shape[1])
    pq = dfmt.train_pq(coords, nch, zero_mean=(metric.lower() != "mips"), seed=seed, device=device)
    codes = dfmt.encode_pq(coords, pq, device=device)
    prefix = dfmt.write_diskann_index(dirpath, name, emb, g, coords, max_norm, pq, codes, metric, partition=partition,
                                      keep_disk_index=keep_disk_index)
    return prefix, g, coords, pq, codes, max_norm
