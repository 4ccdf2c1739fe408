"""LEANN backend plugin: the HNSW selective-recompute search path on one B200.

Mirrors ``leann_backend_hnsw.hnsw_backend`` (packages/leann-backend-hnsw/leann_backend_hnsw/
hnsw_backend.py): ``HNSWBackend`` factory :37-46, ``HNSWSearcher.__init__`` :120-151 and
``HNSWSearcher.search`` :153-253 — same keyword arguments, same return dict, same exceptions —
but ``self._index.search(...)`` (SWIG -> faiss fork -> ZMQ embedding server) is
``capi.Index.search`` (one C-ABI call into libleann_b200.so; traversal, recompute and scoring
all on the GPU).  Reads the same artefacts: ``<stem>.index`` (compact CSR) and
``<name>.meta.json``; the recompute stage additionally needs the passages pre-tokenised
(``<name>.tokens.npy`` + ``<name>.tokoffsets.npy`` sidecars, produced once by
``tokenize_passages``) and the encoder weights (HF checkpoint if present, or a
``b200_weights`` blob / synthetic preset named in the meta file).
"""
from __future__ import annotations

import json
import logging
import os
from pathlib import Path
from typing import Any, Literal, Optional

import numpy as np

from . import capi, synth
from .csr import METRIC_INNER_PRODUCT, METRIC_L2, write_compact_index
from .interface import (
    LeannBackendBuilderInterface,
    LeannBackendFactoryInterface,
    LeannBackendSearcherInterface,
    register_backend,
)

logger = logging.getLogger(__name__)
BACKEND_NAME = "hnsw_b200"


def normalize_l2(data: np.ndarray) -> np.ndarray:
    norms = np.linalg.norm(data, axis=1, keepdims=True)
    norms[norms == 0] = 1
    return data / norms


def _metric_enum(name: str) -> int:
    table = {"mips": METRIC_INNER_PRODUCT, "cosine": METRIC_INNER_PRODUCT, "l2": METRIC_L2}
    if name.lower() not in table:
        raise ValueError(f"Unsupported distance_metric '{name}'.")
    return table[name.lower()]


class B200HnswBuilder(LeannBackendBuilderInterface):
    """``HNSWBuilder`` counterpart (hnsw_backend.py:49-117).  Graph construction is tooling
    (leann_b200.graph_build, torch ops), not part of the accelerated path."""

    def __init__(self, **kwargs):
        self.build_params = kwargs.copy()
        self.is_compact = self.build_params.setdefault("is_compact", True)
        self.is_recompute = self.build_params.setdefault("is_recompute", True)
        self.M = self.build_params.setdefault("M", 32)
        self.efConstruction = self.build_params.setdefault("efConstruction", 200)
        self.distance_metric = self.build_params.setdefault("distance_metric", "mips")
        self.dimensions = self.build_params.get("dimensions")
        if not self.is_compact:
            raise ValueError("the B200 backend only writes the compact CSR format (is_compact=True)")

    def build(self, data: np.ndarray, ids: list[str], index_path: str, **kwargs):
        from .graph_build import build_hnsw_graph

        path = Path(index_path)
        path.parent.mkdir(parents=True, exist_ok=True)
        if data.dtype != np.float32:
            data = data.astype(np.float32)
        _metric_enum(self.distance_metric)
        if self.distance_metric.lower() == "cosine":
            data = normalize_l2(data)
        g = build_hnsw_graph(data, M=self.M, metric=self.distance_metric)
        g.ef_construction = self.efConstruction
        if not self.is_recompute:
            g.vectors = np.ascontiguousarray(data, np.float32)
        write_compact_index(str(path.parent / f"{path.stem}.index"), g)


class _InProcessRecomputeStage:
    """What LeannSearcher.cleanup() / BaseSearcher.__del__ look for (leann-core api.py:798-806, searcher_base.py:194-197):
    an ``embedding_server_manager`` with ``stop_server()``.  The reference stops its embedding-server process there; here
    the "server" is the GPU-resident recompute stage, so stopping it releases the device handle (graph, passages, encoder).
    The next search re-opens it, like the reference re-spawns its server."""

    def __init__(self, owner):
        self._owner = owner

    def stop_server(self):
        self._owner._release_device()


class _B200SearcherBase(LeannBackendSearcherInterface):
    """What ``BaseSearcher`` gives the reference's searchers (leann-core/src/leann/searcher_base.py:18-160), with the
    embedding *server* replaced by the in-process GPU recompute stage."""

    _index: capi.Index

    def _init_base(self, index_path: str, kwargs: dict) -> None:
        # --- BaseSearcher.__init__ contract (searcher_base.py:18-56)
        self.index_path = Path(index_path)
        self.index_dir = self.index_path.parent
        self.meta = kwargs.get("meta") or self._load_meta()
        if not self.meta:
            raise ValueError("Searcher requires metadata from .meta.json.")
        self.dimensions = self.meta.get("dimensions")
        if not self.dimensions:
            raise ValueError("Dimensions not found in Leann metadata.")
        self.embedding_model = self.meta.get("embedding_model")
        self.embedding_mode = self.meta.get("embedding_mode", "sentence-transformers")
        self.device = int(kwargs.get("device", os.environ.get("LOCAL_RANK", 0)))
        self._tokenizer = None
        self._recompute_ready = False
        self.preset: synth.ModelPreset | None = None
        self._index = None
        self.embedding_server_manager = _InProcessRecomputeStage(self)
        # one process per GPU (torchrun): split each call's query batch over the ranks and all_gather the results
        # (leann_b200/parallel.py).  Off by default: a lone process answers its whole batch like the reference.
        self._shard_queries = bool(kwargs.get("shard_queries", False))

    def _run_search(self, query: np.ndarray, top_k: int, params):
        run = lambda qs: self._index.search(np.ascontiguousarray(qs), int(top_k), params)  # noqa: E731
        if self._shard_queries:
            from .parallel import sharded_search

            return sharded_search(run, query, int(top_k))
        return run(query)

    def _open_index(self) -> None:  # subclasses: create self._index
        raise NotImplementedError

    def _ensure_open(self) -> None:
        if self._index is None:
            self._open_index()

    def _release_device(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None
        self._recompute_ready = False

    # ------------------------------------------------------------------ helpers
    def _load_meta(self) -> dict[str, Any]:
        meta_path = self.index_dir / f"{self.index_path.name}.meta.json"
        if not meta_path.exists():
            raise FileNotFoundError(f"Leann metadata file not found at {meta_path}")
        with open(meta_path, encoding="utf-8") as f:
            return json.load(f)

    def _sidecar(self, suffix: str) -> Path:
        return self.index_dir / f"{self.index_path.name}.{suffix}"

    def _load_encoder_weights(self) -> tuple[synth.ModelPreset, np.ndarray]:
        name = self.embedding_model or ""
        blob_file = self.meta.get("b200_weights")
        if name in synth.PRESETS and (blob_file or name.startswith("synthetic/") or self.meta.get("b200_synthetic_weights")):
            preset = synth.PRESETS[name]
            if blob_file:
                return preset, np.load(self.index_dir / blob_file)
            seed = int(self.meta.get("b200_weight_seed", 0))
            return preset, synth.pack_weights(preset, synth.synthetic_weights(preset, seed))
        # real checkpoint through transformers (what sentence-transformers wraps)
        try:
            from transformers import AutoModel  # type: ignore

            hf = AutoModel.from_pretrained(name)
        except Exception as e:  # no network / no cache
            raise RuntimeError(f"cannot load encoder weights for '{name}': {e}. Provide 'b200_weights' in the "
                               "meta file or use a synthetic preset.") from e
        return weights_from_hf(hf, name)

    def _attach_recompute_stage(self):
        self._ensure_open()
        if self._recompute_ready:
            return
        tok_f, off_f = self._sidecar("tokens.npy"), self._sidecar("tokoffsets.npy")
        if not (tok_f.exists() and off_f.exists()):
            raise RuntimeError(f"pre-tokenised passage store not found ({tok_f.name}); run "
                               "leann_b200.backend.tokenize_passages() once for this index")
        self._index.set_passages(np.load(tok_f, mmap_mode="r"), np.load(off_f))
        self.preset, blob = self._load_encoder_weights()
        self._index.set_encoder(self.preset.config(), blob)
        self._recompute_ready = True

    # ------------------------------------------------------------------ plugin API
    def _ensure_server_running(self, passages_source_file: str, port: Optional[int], **kwargs) -> int:
        """No server process exists: the recompute stage lives in this process on the GPU.
        Keeps the reference's contract (searcher_base.py:58-84): raise if recompute is impossible,
        return the port that the caller hands back as ``zmq_port``."""
        if not self.embedding_model:
            raise ValueError("Cannot use recompute mode without 'embedding_model' in meta.json.")
        try:
            self._attach_recompute_stage()
        except Exception as e:
            raise RuntimeError(f"Failed to start embedding server on port {port}: {e}") from e
        return int(port) if port is not None else 5557

    def compute_query_embedding(self, query, use_server_if_available: bool = True,
                                zmq_port: Optional[int] = None) -> np.ndarray:
        """(1, D) float32 embedding of the query with the SAME GPU encoder as the passages.
        ``query`` is a string (needs the model's WordPiece tokenizer on disk) or an array of token ids."""
        self._attach_recompute_stage()  # (re)opens the handle if cleanup() released it
        if isinstance(query, str):
            if self._tokenizer is None:
                try:
                    from transformers import AutoTokenizer  # type: ignore

                    self._tokenizer = AutoTokenizer.from_pretrained(self.embedding_model, use_fast=True)
                except Exception as e:
                    raise RuntimeError(f"tokenizer for '{self.embedding_model}' unavailable: {e}") from e
            ids = self._tokenizer(query, truncation=True, max_length=self.preset.max_pos)["input_ids"]
        else:
            ids = query
        ids = np.asarray(ids, np.int64).reshape(-1)
        if ids.size == 0 or ids.min() < 0 or ids.max() >= self.preset.vocab_size or ids.max() > 65535:
            raise ValueError(f"query token ids must lie in [0, {min(self.preset.vocab_size, 65536)}) and be non-empty")
        toks = ids.astype(np.uint16)
        emb = self._index.encode_tokens(toks, np.array([0, toks.size], np.uint64))
        return emb.reshape(1, -1)

    @property
    def last_stats(self) -> dict[str, Any]:
        return self._index.last_stats.as_dict() if self._index is not None else {}

    def cleanup(self):
        self._release_device()


class B200HnswSearcher(_B200SearcherBase):
    def __init__(self, index_path: str, **kwargs):
        self._init_base(index_path, kwargs)
        # --- HNSWSearcher.__init__ (hnsw_backend.py:128-151)
        self.distance_metric = self.meta.get("backend_kwargs", {}).get("distance_metric", "mips").lower()
        _metric_enum(self.distance_metric)
        self.is_compact = self.meta.get("is_compact", True)
        self.is_pruned = self.meta.get("is_pruned", True)
        index_file = self.index_dir / f"{self.index_path.stem}.index"
        if not index_file.exists():
            raise FileNotFoundError(f"HNSW index file not found at {index_file}")
        if not self.is_compact:
            raise RuntimeError("the B200 backend reads compact (CSR) HNSW indexes only")
        self._index_file = str(index_file)
        self._tuning = {k: kwargs[k] for k in ("slots", "passages_per_pass", "dedup_scope") if kwargs.get(k) is not None}
        self._open_index()

    def _open_index(self) -> None:
        self._index = capi.Index(self._index_file, self.device)
        if self._index.info.d != int(self.dimensions):
            raise ValueError(f"index dimension {self._index.info.d} != meta dimensions {self.dimensions}")
        t = self._tuning
        if t.get("slots") or t.get("passages_per_pass"):
            self._index.configure(int(t.get("slots", 0)), int(t.get("passages_per_pass", 0)))
        if t.get("dedup_scope") is not None:  # "hop" (default) | "call"
            self._index.set_option("dedup_scope", {"hop": 0, "call": 1}[str(t["dedup_scope"])])
        # PQ-guided pruning tables (HNSW::load_pq_pruning_data, HNSW_search.cpp:253-297): DiskANN-format sidecars
        self._has_pq = self._sidecar("pq_pivots.bin").exists() and self._sidecar("pq_compressed.bin").exists()
        if self._has_pq:
            self._index.load_pq_pruning(str(self._sidecar("pq_pivots.bin")), str(self._sidecar("pq_compressed.bin")))

    def search(self, query: np.ndarray, top_k: int, zmq_port: Optional[int] = None, complexity: int = 64,
               beam_width: int = 1, prune_ratio: float = 0.0, recompute_embeddings: bool = True,
               pruning_strategy: Literal["global", "local", "proportional"] = "global", batch_size: int = 0,
               **kwargs) -> dict[str, Any]:
        self._ensure_open()
        if not recompute_embeddings and self.is_pruned and not self._index.info.has_vectors:
            raise RuntimeError("Recompute is required for pruned/compact HNSW index. "
                               "Re-run search with --recompute, or rebuild with --no-recompute and --no-compact.")
        if recompute_embeddings:
            if zmq_port is None:
                raise ValueError("zmq_port must be provided if recompute_embeddings is True")
            self._attach_recompute_stage()
        if query.dtype != np.float32:
            query = query.astype(np.float32)
        if self.distance_metric == "cosine":
            query = normalize_l2(query)
        embedding_model = (self.meta.get("embedding_model") or "").lower()
        check_rel = not (self.distance_metric == "cosine"
                         and any(m in embedding_model for m in ["text-embedding", "openai"]))  # hnsw_backend.py:213-220
        local_prune, send_ratio = False, 0.0
        if pruning_strategy == "local":
            local_prune = True
        elif pruning_strategy == "proportional":
            send_ratio = 1.0
        if not getattr(self, "_has_pq", False):
            # the reference ignores prune_ratio AND the strategy flags when no PQ files are loaded
            # (perform_pq_pruning needs hnsw.pq_data_loader, HNSW_search.cpp:442-445): the same call
            # succeeds on the stock backend, so it must not raise here
            prune_ratio, local_prune, send_ratio = 0.0, False, 0.0
        params = capi.make_params(complexity, beam_width, batch_size, check_rel, prune_ratio, local_prune, send_ratio,
                                  recompute_embeddings)
        try:
            distances, labels = self._run_search(query, top_k, params)
        except capi.Lb2Error as e:
            if "not implemented" in str(e):
                raise NotImplementedError(str(e)) from e
            raise RuntimeError(str(e)) from e
        string_labels = [[str(int_label) for int_label in batch_labels] for batch_labels in labels]
        return {"labels": string_labels, "distances": distances}


@register_backend(BACKEND_NAME)
class B200HnswBackend(LeannBackendFactoryInterface):
    @staticmethod
    def builder(**kwargs) -> LeannBackendBuilderInterface:
        return B200HnswBuilder(**kwargs)

    @staticmethod
    def searcher(index_path: str, **kwargs) -> LeannBackendSearcherInterface:
        return B200HnswSearcher(index_path, **kwargs)


if os.environ.get("LEANN_B200_OVERRIDE_HNSW") == "1":  # same-process override of the stock backend
    register_backend("hnsw")(B200HnswBackend)


# ---------------------------------------------------------------------------------------------
def weights_from_hf(hf_model, name: str) -> tuple[synth.ModelPreset, np.ndarray]:
    """Pack a transformers BertModel's state_dict into the lb2_set_encoder blob order."""
    cfg = hf_model.config
    sd = {k: v.detach().float().cpu().numpy() for k, v in hf_model.state_dict().items()}
    known = synth.PRESETS.get(name)
    preset = synth.ModelPreset(name, cfg.vocab_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                               cfg.intermediate_size, known.max_pos if known else min(512, cfg.max_position_embeddings),
                               cfg.type_vocab_size, cfg.layer_norm_eps, known.pooling if known else 0, 1)
    g = lambda k: sd[k] if k in sd else sd["bert." + k]
    w = {"word_emb": g("embeddings.word_embeddings.weight"),
         "pos_emb": g("embeddings.position_embeddings.weight")[: preset.max_pos],
         "type_emb": g("embeddings.token_type_embeddings.weight"),
         "emb_ln_g": g("embeddings.LayerNorm.weight"), "emb_ln_b": g("embeddings.LayerNorm.bias")}
    for l in range(preset.layers):
        p = f"encoder.layer.{l}."
        w[f"l{l}.w_qkv"] = np.concatenate([g(p + f"attention.self.{n}.weight") for n in ("query", "key", "value")])
        w[f"l{l}.b_qkv"] = np.concatenate([g(p + f"attention.self.{n}.bias") for n in ("query", "key", "value")])
        w[f"l{l}.w_o"] = g(p + "attention.output.dense.weight"); w[f"l{l}.b_o"] = g(p + "attention.output.dense.bias")
        w[f"l{l}.ln1_g"] = g(p + "attention.output.LayerNorm.weight"); w[f"l{l}.ln1_b"] = g(p + "attention.output.LayerNorm.bias")
        w[f"l{l}.w_1"] = g(p + "intermediate.dense.weight"); w[f"l{l}.b_1"] = g(p + "intermediate.dense.bias")
        w[f"l{l}.w_2"] = g(p + "output.dense.weight"); w[f"l{l}.b_2"] = g(p + "output.dense.bias")
        w[f"l{l}.ln2_g"] = g(p + "output.LayerNorm.weight"); w[f"l{l}.ln2_b"] = g(p + "output.LayerNorm.bias")
    return preset, synth.pack_weights(preset, w)


def tokenize_passages(index_path: str, texts, tokenizer, max_len: int) -> None:
    """One-off: WordPiece-tokenise the passages (same tokenizer call the reference makes per hop,
    embedding_compute.py:299-305) into the uint16 sidecars the recompute stage reads."""
    toks, offs = [], [0]
    for t in texts:
        ids = np.asarray(tokenizer(t, truncation=True, max_length=max_len)["input_ids"], np.int64)
        if ids.size == 0 or ids.min() < 0 or ids.max() > 65535:
            raise ValueError("token ids must fit the uint16 passage store (vocabularies up to 65 536 entries) and passages must not be empty")
        toks.append(ids.astype(np.uint16))
        offs.append(offs[-1] + len(ids))
    p = Path(index_path)
    np.save(p.parent / f"{p.name}.tokens.npy", np.concatenate(toks) if toks else np.zeros(0, np.uint16))
    np.save(p.parent / f"{p.name}.tokoffsets.npy", np.asarray(offs, np.uint64))
