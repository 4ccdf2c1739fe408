"""Small host-side utilities shared by the benchmark, smoke() and the tests: stub graphs, index
directory layout, recall."""
from __future__ import annotations

import json
import tempfile
from pathlib import Path

import numpy as np

from . import csr

def stub_graph(n: int, d: int, metric_type: int = 0) -> csr.CSRGraph:
    """N isolated nodes: enough to open an index and drive the recompute stage alone."""
    levels = np.ones(n, np.int32)
    node_offsets = (np.arange(n + 1, dtype=np.uint64) * 2)
    level_ptr = np.zeros(2 * n, np.uint64)
    return csr.CSRGraph(d=d, ntotal=n, metric_type=metric_type, levels=levels, level_ptr=level_ptr,
                        node_offsets=node_offsets, neighbors=np.zeros(0, np.int32), entry_point=0 if n else -1,
                        max_level=0 if n else -1)


def open_encoder_only(preset, blob, corpus, device=0):
    """capi.Index over a stub graph with passages + encoder attached."""
    from leann_b200 import capi

    f = Path(tempfile.mkdtemp()) / "stub.index"
    csr.write_compact_index(str(f), stub_graph(corpus.n, preset.hidden))
    idx = capi.Index(str(f), device)
    idx.set_passages(corpus.tokens, corpus.offsets)
    idx.set_encoder(preset.config(), blob)
    return idx


def write_leann_index(dirpath: Path, name: str, graph: csr.CSRGraph, preset=None, corpus=None, metric="mips",
                      weight_seed=0, is_pruned=True):
    """Lays out the files LeannSearcher expects: <name>.index, <name>.leann.meta.json (+ token sidecars)."""
    dirpath.mkdir(parents=True, exist_ok=True)
    index_path = dirpath / f"{name}.leann"
    csr.write_compact_index(str(dirpath / f"{name}.index"), graph)
    meta = {"version": "1.0", "backend_name": "hnsw_b200", "embedding_model": preset.name if preset else "none",
            "dimensions": graph.d, "backend_kwargs": {"distance_metric": metric, "M": 16},
            "embedding_mode": "sentence-transformers", "is_compact": True, "is_pruned": is_pruned,
            "b200_synthetic_weights": True, "b200_weight_seed": weight_seed}
    (dirpath / f"{name}.leann.meta.json").write_text(json.dumps(meta))
    if corpus is not None:
        np.save(dirpath / f"{name}.leann.tokens.npy", corpus.tokens)
        np.save(dirpath / f"{name}.leann.tokoffsets.npy", corpus.offsets)
    return index_path


def write_diskann_leann_index(dirpath: Path, name: str, emb: np.ndarray, preset=None, corpus=None, metric="mips", R=32,
                              n_chunks=None, partition=True, weight_seed=0, device=None):
    """Lays out what LeannSearcher + the DiskANN searcher expect: <name>_pq_pivots.bin, _pq_compressed.bin, _disk.index
    (+ partition pair), <name>.leann.meta.json (+ token sidecars).  Returns (index_path, build artefacts)."""
    from .vamana_build import build_diskann_index

    dirpath.mkdir(parents=True, exist_ok=True)
    index_path = dirpath / f"{name}.leann"
    art = build_diskann_index(dirpath, name, emb, metric=metric, R=R, n_chunks=n_chunks, partition=partition, device=device)
    meta = {"version": "1.0", "backend_name": "diskann_b200", "embedding_model": preset.name if preset else "none",
            "dimensions": int(emb.shape[1]), "backend_kwargs": {"distance_metric": metric, "graph_degree": R},
            "embedding_mode": "sentence-transformers", "b200_synthetic_weights": True, "b200_weight_seed": weight_seed}
    (dirpath / f"{name}.leann.meta.json").write_text(json.dumps(meta))
    if corpus is not None:
        np.save(dirpath / f"{name}.leann.tokens.npy", corpus.tokens)
        np.save(dirpath / f"{name}.leann.tokoffsets.npy", corpus.offsets)
    return index_path, art


def recall_at_k(I, gt):
    return float(np.mean([len(set(a.tolist()) & set(b.tolist())) / len(b) for a, b in zip(I, gt)]))
