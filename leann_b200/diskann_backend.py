"""LEANN backend plugin: the DiskANN/Vamana search path on one B200.

Mirrors ``leann_backend_diskann.diskann_backend`` (packages/leann-backend-diskann/leann_backend_diskann/
diskann_backend.py): ``DiskannBackend`` factory :113-121, ``DiskannSearcher.__init__`` :302-357,
``_ensure_index_loaded`` :359-381 and ``search`` :383-472 — same keyword arguments, same return dict, same
exceptions — with ``self._index.batch_search(...)`` (pybind -> PQFlashIndex::cached_beam_search -> protobuf/ZMQ
embedding fetch) replaced by one C-ABI call, ``lb2_diskann_search``.  Opens the files the stock searcher opens
(``<stem>_pq_pivots.bin``, ``_pq_compressed.bin``, ``_disk.index`` … or the partition pair) plus, for
``recompute_embeddings=True``, the pre-tokenised passage sidecars and encoder weights described in backend.py.
"""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Any, Literal, Optional

import numpy as np

from . import capi
from .backend import _B200SearcherBase
from .interface import (
    LeannBackendBuilderInterface,
    LeannBackendFactoryInterface,
    LeannBackendSearcherInterface,
    register_backend,
)

logger = logging.getLogger(__name__)
BACKEND_NAME = "diskann_b200"


class B200DiskannBuilder(LeannBackendBuilderInterface):
    """``DiskannBuilder`` counterpart (diskann_backend.py:124-299): same on-disk artefacts, built with the torch
    tooling of leann_b200.vamana_build instead of diskannpy.build_disk_float_index (tooling, not the hot path)."""

    def __init__(self, **kwargs):
        self.build_params = kwargs

    def build(self, data: np.ndarray, ids: list[str], index_path: str, **kwargs):
        from .vamana_build import build_diskann_index

        path = Path(index_path)
        path.parent.mkdir(parents=True, exist_ok=True)
        if data.dtype != np.float32:
            logger.warning(f"Converting data to float32, shape: {data.shape}")
            data = data.astype(np.float32)
        kw = {**self.build_params, **kwargs}
        if "backend_kwargs" in kw:
            kw.update(kw.pop("backend_kwargs"))
        metric = kw.get("distance_metric", "mips").lower()
        if metric not in capi.DISKANN_METRICS:
            raise ValueError(f"Unsupported distance_metric '{kw.get('distance_metric', 'unknown')}'.")
        if ids is not None and list(ids) != [str(i) for i in range(len(ids))]:
            raise ValueError("labels are the decimal strings of the internal ids (leann-core api.py:376)")
        is_recompute = bool(kw.get("is_recompute", False))
        build_diskann_index(path.parent, path.stem, data, metric=metric, R=int(kw.get("graph_degree", 32)),
                            n_chunks=kw.get("pq_chunks"), partition=is_recompute,
                            keep_disk_index=not is_recompute)  # _safe_cleanup_after_partition, diskann_backend.py:128-190


class B200DiskannSearcher(_B200SearcherBase):
    def __init__(self, index_path: str, **kwargs):
        self._init_base(index_path, kwargs)
        # --- DiskannSearcher.__init__ (diskann_backend.py:310-357)
        self.distance_metric = kwargs.get("distance_metric",
                                          self.meta.get("backend_kwargs", {}).get("distance_metric", "mips")).lower()
        if self.distance_metric not in capi.DISKANN_METRICS:
            raise ValueError(f"Unsupported distance_metric '{self.distance_metric}'.")
        self.num_threads = kwargs.get("num_threads", 8)  # accepted for signature parity; the GPU has no use for it
        index_name = self.index_path.stem
        self._index_prefix = str(self.index_dir / index_name)
        graph_f, part_f = Path(self._index_prefix + "_disk_graph.index"), Path(self._index_prefix + "_partition.bin")
        self._partition_prefix = self._index_prefix if graph_f.exists() and part_f.exists() else ""
        if not Path(self._index_prefix + "_pq_compressed.bin").exists():
            raise FileNotFoundError(f"DiskANN index files not found at prefix {self._index_prefix}")
        self._current_zmq_port = None
        self._tuning = {k: kwargs[k] for k in ("passages_per_pass",) if kwargs.get(k)}

    def _open_index(self) -> None:
        try:
            self._index = capi.DiskannIndex(self._index_prefix, self.distance_metric, self._partition_prefix, self.device)
        except capi.Lb2Error as e:
            raise RuntimeError(str(e)) from e
        if self._index.dinfo.dim != int(self.dimensions):
            raise ValueError(f"index dimension {self._index.dinfo.dim} != meta dimensions {self.dimensions}")
        if self._tuning:
            self._index.configure(0, int(self._tuning["passages_per_pass"]))

    def _ensure_index_loaded(self, zmq_port: int):
        """diskann_backend.py:359-381 loads the index lazily and reloads it when the port changes; there is no port here,
        so the index is loaded once (and again after cleanup() released it)."""
        self._ensure_open()
        self._current_zmq_port = zmq_port

    def search(self, query: np.ndarray, top_k: int, complexity: int = 64, beam_width: int = 1, prune_ratio: float = 0.0,
               recompute_embeddings: bool = False,
               pruning_strategy: Literal["global", "local", "proportional"] = "global", zmq_port: Optional[int] = None,
               batch_recompute: bool = False, dedup_node_dis: bool = False, **kwargs) -> dict[str, Any]:
        if recompute_embeddings:
            if zmq_port is None:
                raise ValueError("zmq_port must be provided if recompute_embeddings is True")
            self._ensure_index_loaded(zmq_port)
            self._attach_recompute_stage()
        elif self._index is None:
            self._ensure_index_loaded(6666)
        if pruning_strategy == "proportional":
            raise NotImplementedError("DiskANN backend does not support 'proportional' pruning strategy. "
                                      "Use 'global' or 'local' instead.")
        if query.dtype != np.float32:
            query = query.astype(np.float32)
        params = capi.make_diskann_params(
            complexity, beam_width, recompute_embeddings=recompute_embeddings,
            skip_search_reorder=kwargs.get("skip_search_reorder", False), recompute_neighbors=False,
            dedup_node_dis=dedup_node_dis, prune_ratio=prune_ratio, batch_recompute=batch_recompute,
            global_pruning=(pruning_strategy != "local"))
        try:
            distances, labels = self._run_search(query, top_k, params)
        except capi.Lb2Error as e:
            if "not implemented" in str(e):
                raise NotImplementedError(str(e)) from e
            raise RuntimeError(str(e)) from e
        string_labels = [[str(int_label) for int_label in batch_labels] for batch_labels in labels]
        return {"labels": string_labels, "distances": distances}


@register_backend(BACKEND_NAME)
class B200DiskannBackend(LeannBackendFactoryInterface):
    @staticmethod
    def builder(**kwargs) -> LeannBackendBuilderInterface:
        return B200DiskannBuilder(**kwargs)

    @staticmethod
    def searcher(index_path: str, **kwargs) -> LeannBackendSearcherInterface:
        return B200DiskannSearcher(index_path, **kwargs)


if os.environ.get("LEANN_B200_OVERRIDE_DISKANN") == "1":  # same-process override of the stock backend
    register_backend("diskann")(B200DiskannBackend)
