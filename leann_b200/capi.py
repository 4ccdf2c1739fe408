"""ctypes binding of libleann_b200.so (include/leann_b200.h).

The shared library is the product; this file only marshals numpy / torch buffers into
plain pointers.  There is no fallback: if the library is missing or no sm_100 device
is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "libleann_b200.so"
_lib = None

EXPORTED_SYMBOLS = [
    "lb2_last_error", "lb2_version", "lb2_open", "lb2_close", "lb2_info", "lb2_set_vectors", "lb2_set_vectors_device",
    "lb2_set_passages", "lb2_load_pq_pruning",
    "lb2_encoder_weight_count", "lb2_set_encoder", "lb2_default_params", "lb2_search", "lb2_search_device",
    "lb2_last_query_stats", "lb2_encode_ids", "lb2_encode_tokens", "lb2_encode_range_device", "lb2_configure",
    "lb2_set_option",
    "lb2_diskann_open", "lb2_diskann_info", "lb2_diskann_default_params", "lb2_diskann_search",
    "lb2_diskann_search_device", "lb2_diskann_last_expansions",
    "lb2_build_insert_search", "lb2_build_workspace_bytes", "lb2_build_select",
    "lb2_test_gemm_f16", "lb2_test_gemm_grouped_f16", "lb2_test_layernorm_f16", "lb2_test_gemm_res_ln_f16", "lb2_test_attention_f16",
]


class SearchParams(C.Structure):
    _fields_ = [("efSearch", C.c_int32), ("beam_size", C.c_int32), ("batch_size", C.c_int32),
                ("check_relative_distance", C.c_int32), ("pq_pruning_ratio", C.c_float),
                ("local_prune", C.c_int32), ("send_neigh_times_ratio", C.c_float), ("recompute", C.c_int32)]


class SearchStats(C.Structure):
    _fields_ = [("ndis", C.c_int64), ("nhops", C.c_int64), ("n_recomputed", C.c_int64), ("n_requested", C.c_int64),
                ("n_tokens", C.c_int64), ("n_steps", C.c_int64), ("n_kernel_launches", C.c_int64),
                ("gpu_ms", C.c_double), ("encoder_ms", C.c_double), ("gemm_ms", C.c_double),
                ("gemm_flops", C.c_double), ("attention_ms", C.c_double), ("norm_ms", C.c_double), ("n_encoder_passes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class IndexInfo(C.Structure):
    _fields_ = [("ntotal", C.c_int64), ("d", C.c_int32), ("metric_type", C.c_int32), ("entry_point", C.c_int32),
                ("max_level", C.c_int32), ("n_edges", C.c_int64), ("max_degree_level0", C.c_int32),
                ("max_degree_upper", C.c_int32), ("has_vectors", C.c_int32), ("has_passages", C.c_int32),
                ("has_encoder", C.c_int32), ("device", C.c_int32)]


class EncoderConfig(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("ffn", C.c_int32), ("max_pos", C.c_int32), ("type_vocab", C.c_int32), ("ln_eps", C.c_float),
                ("pooling", C.c_int32), ("normalize", C.c_int32)]


class DiskannParams(C.Structure):
    _fields_ = [("complexity", C.c_int32), ("beam_width", C.c_int32), ("deferred_fetch", C.c_int32),
                ("skip_search_reorder", C.c_int32), ("recompute_neighbors", C.c_int32), ("dedup_node_dis", C.c_int32),
                ("batch_recompute", C.c_int32), ("global_pruning", C.c_int32), ("prune_ratio", C.c_float),
                ("io_limit", C.c_uint32)]


class DiskannInfo(C.Structure):
    _fields_ = [("npts", C.c_int64), ("dim", C.c_int32), ("data_dim", C.c_int32), ("n_chunks", C.c_int32),
                ("max_degree", C.c_int32), ("metric", C.c_int32), ("n_medoids", C.c_int32), ("partitioned", C.c_int32),
                ("has_coords", C.c_int32), ("max_base_norm", C.c_float), ("pad", C.c_int32), ("n_edges", C.c_int64)]


DISKANN_METRICS = {"l2": 0, "mips": 1, "cosine": 2}


def library_path() -> Path:
    return _LIB_PATH


def load():
    """dlopen the library (building nothing: see leann_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(f"{_LIB_PATH} is missing: run `python -m leann_b200.build` (needs nvcc). "
                           "leann_b200 has no CPU path.")
    lib = C.CDLL(str(_LIB_PATH))
    lib.lb2_last_error.restype = C.c_char_p
    lib.lb2_open.restype = C.c_void_p
    lib.lb2_open.argtypes = [C.c_char_p, C.c_int]
    lib.lb2_close.argtypes = [C.c_void_p]
    lib.lb2_close.restype = None
    lib.lb2_info.argtypes = [C.c_void_p, C.POINTER(IndexInfo)]
    lib.lb2_set_vectors.argtypes = [C.c_void_p, C.c_void_p]
    lib.lb2_set_vectors_device.argtypes = [C.c_void_p, C.c_void_p]
    lib.lb2_set_passages.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lb2_load_pq_pruning.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    lib.lb2_encoder_weight_count.restype = C.c_size_t
    lib.lb2_encoder_weight_count.argtypes = [C.POINTER(EncoderConfig)]
    lib.lb2_set_encoder.argtypes = [C.c_void_p, C.POINTER(EncoderConfig), C.c_void_p, C.c_size_t]
    lib.lb2_default_params.argtypes = [C.POINTER(SearchParams)]
    lib.lb2_default_params.restype = None
    for name in ("lb2_search", "lb2_search_device"):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                       C.POINTER(SearchParams), C.POINTER(SearchStats)]
    lib.lb2_last_query_stats.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.lb2_encode_ids.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.lb2_encode_tokens.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lb2_encode_range_device.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.lb2_configure.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.lb2_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.lb2_diskann_open.restype = C.c_void_p
    lib.lb2_diskann_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    lib.lb2_diskann_info.argtypes = [C.c_void_p, C.POINTER(DiskannInfo)]
    lib.lb2_diskann_default_params.argtypes = [C.POINTER(DiskannParams)]
    lib.lb2_diskann_default_params.restype = None
    for name in ("lb2_diskann_search", "lb2_diskann_search_device"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.POINTER(DiskannParams), C.POINTER(SearchStats)]
    lib.lb2_diskann_last_expansions.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.lb2_test_gemm_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int]
    lib.lb2_test_gemm_res_ln_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                             C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.lb2_test_layernorm_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float]
    lib.lb2_test_attention_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lb2_test_gemm_grouped_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lb2_build_insert_search.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.lb2_build_workspace_bytes.restype = C.c_size_t
    lib.lb2_build_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.lb2_build_select.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


class Lb2Error(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().lb2_last_error().decode("utf-8", "replace")
        raise Lb2Error(f"{what} failed ({rc}): {msg}")


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def make_params(complexity=64, beam_width=1, batch_size=0, check_relative_distance=True, prune_ratio=0.0,
                local_prune=False, send_neigh_times_ratio=0.0, recompute=True) -> SearchParams:
    return SearchParams(int(complexity), int(beam_width), int(batch_size), int(bool(check_relative_distance)),
                        float(prune_ratio), int(bool(local_prune)), float(send_neigh_times_ratio),
                        int(bool(recompute)))


class Index:
    """Owning wrapper of an lb2_index handle."""

    def __init__(self, index_path: str, device: int = 0):
        lib = load()
        h = lib.lb2_open(str(index_path).encode(), int(device))
        if not h:
            raise Lb2Error(f"lb2_open('{index_path}') failed: {lib.lb2_last_error().decode('utf-8', 'replace')}")
        self._h = C.c_void_p(h)
        self._lib = lib
        self.info = IndexInfo()
        _check(lib.lb2_info(self._h, C.byref(self.info)), "lb2_info")
        self.last_stats = SearchStats()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lb2_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh_info(self):
        _check(self._lib.lb2_info(self._h, C.byref(self.info)), "lb2_info")
        return self.info

    def set_vectors(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.float32)
        if x.shape != (self.info.ntotal, self.info.d):
            raise ValueError(f"vectors must be [{self.info.ntotal}, {self.info.d}], got {x.shape}")
        _check(self._lib.lb2_set_vectors(self._h, _np_ptr(x)), "lb2_set_vectors")

    def set_vectors_device(self, d_ptr: int):
        _check(self._lib.lb2_set_vectors_device(self._h, C.c_void_p(d_ptr)), "lb2_set_vectors_device")
        self.refresh_info()

    def set_passages(self, tokens: np.ndarray, offsets: np.ndarray):
        tokens = np.ascontiguousarray(tokens, np.uint16)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        if offsets.shape != (self.info.ntotal + 1,):
            raise ValueError("offsets must have ntotal + 1 entries")
        if int(offsets[-1]) != tokens.size:
            raise ValueError("offsets[-1] must equal the token count")
        _check(self._lib.lb2_set_passages(self._h, _np_ptr(tokens), _np_ptr(offsets)), "lb2_set_passages")

    def load_pq_pruning(self, pq_pivots_path: str, pq_compressed_path: str):
        _check(self._lib.lb2_load_pq_pruning(self._h, str(pq_pivots_path).encode(), str(pq_compressed_path).encode()),
               "lb2_load_pq_pruning")

    def set_encoder(self, cfg: EncoderConfig, weights: np.ndarray):
        weights = np.ascontiguousarray(weights, np.float32)
        _check(self._lib.lb2_set_encoder(self._h, C.byref(cfg), _np_ptr(weights), weights.size), "lb2_set_encoder")

    def configure(self, slots: int = 0, passages_per_pass: int = 0):
        _check(self._lib.lb2_configure(self._h, int(slots), int(passages_per_pass)), "lb2_configure")

    def set_option(self, key: str, value: int):
        _check(self._lib.lb2_set_option(self._h, key.encode(), int(value)), f"lb2_set_option({key})")

    def search(self, q: np.ndarray, k: int, params: SearchParams | None = None):
        q = np.ascontiguousarray(q, np.float32)
        if q.ndim != 2 or q.shape[1] != self.info.d:
            raise ValueError(f"query must be [B, {self.info.d}], got {q.shape}")
        nq = q.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        p = params or make_params()
        _check(self._lib.lb2_search(self._h, nq, _np_ptr(q), int(k), _np_ptr(D), _np_ptr(I), C.byref(p),
                                    C.byref(self.last_stats)), "lb2_search")
        return D, I

    def search_device(self, d_q: int, nq: int, k: int, d_D: int, d_I: int, params: SearchParams | None = None):
        """Raw device pointers (e.g. torch.Tensor.data_ptr())."""
        p = params or make_params()
        _check(self._lib.lb2_search_device(self._h, int(nq), C.c_void_p(d_q), int(k), C.c_void_p(d_D),
                                           C.c_void_p(d_I), C.byref(p), C.byref(self.last_stats)), "lb2_search_device")

    def last_query_stats(self, nq: int):
        ndis = np.zeros(nq, np.int64)
        nhops = np.zeros(nq, np.int64)
        _check(self._lib.lb2_last_query_stats(self._h, nq, _np_ptr(ndis), _np_ptr(nhops)), "lb2_last_query_stats")
        return ndis, nhops

    def encode_ids(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.int64)
        out = np.empty((ids.size, self.info.d), np.float32)
        _check(self._lib.lb2_encode_ids(self._h, ids.size, _np_ptr(ids), _np_ptr(out)), "lb2_encode_ids")
        return out

    def encode_tokens(self, tokens: np.ndarray, offsets: np.ndarray, dim: int | None = None) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, np.uint16)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        out = np.empty((n, dim or self.info.d), np.float32)
        _check(self._lib.lb2_encode_tokens(self._h, n, _np_ptr(tokens), _np_ptr(offsets), _np_ptr(out)),
               "lb2_encode_tokens")
        return out

    def encode_range_device(self, first: int, n: int, d_out: int):
        _check(self._lib.lb2_encode_range_device(self._h, int(first), int(n), C.c_void_p(d_out)),
               "lb2_encode_range_device")


def make_diskann_params(complexity=64, beam_width=1, recompute_embeddings=True, skip_search_reorder=False,
                        recompute_neighbors=False, dedup_node_dis=False, prune_ratio=0.0, batch_recompute=False,
                        global_pruning=True, io_limit=0) -> DiskannParams:
    """Argument-for-argument the tail of StaticDiskIndex::batch_search (static_disk_index.cpp:88-93)."""
    return DiskannParams(int(complexity), int(beam_width), int(bool(recompute_embeddings)), int(bool(skip_search_reorder)),
                         int(bool(recompute_neighbors)), int(bool(dedup_node_dis)), int(bool(batch_recompute)),
                         int(bool(global_pruning)), float(prune_ratio), int(io_limit))


class DiskannIndex(Index):
    """Owning wrapper of a handle opened with lb2_diskann_open; passages / encoder / encode_* as on Index."""

    def __init__(self, index_prefix: str, metric: str = "mips", partition_prefix: str | None = None, device: int = 0):
        lib = load()
        if metric.lower() not in DISKANN_METRICS:
            raise ValueError(f"Unsupported distance_metric '{metric}'.")
        h = lib.lb2_diskann_open(str(index_prefix).encode(), (partition_prefix or "").encode(),
                                 DISKANN_METRICS[metric.lower()], int(device))
        if not h:
            raise Lb2Error(f"lb2_diskann_open('{index_prefix}') failed: {lib.lb2_last_error().decode('utf-8', 'replace')}")
        self._h = C.c_void_p(h)
        self._lib = lib
        self.info = IndexInfo()
        _check(lib.lb2_info(self._h, C.byref(self.info)), "lb2_info")
        self.dinfo = DiskannInfo()
        _check(lib.lb2_diskann_info(self._h, C.byref(self.dinfo)), "lb2_diskann_info")
        self.last_stats = SearchStats()

    def search(self, q: np.ndarray, k: int, params: DiskannParams | None = None):
        q = np.ascontiguousarray(q, np.float32)
        if q.ndim != 2 or q.shape[1] != self.dinfo.dim:
            raise ValueError(f"query must be [B, {self.dinfo.dim}], got {q.shape}")
        nq = q.shape[0]
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        p = params or make_diskann_params()
        _check(self._lib.lb2_diskann_search(self._h, nq, _np_ptr(q), int(k), _np_ptr(D), _np_ptr(I), C.byref(p),
                                            C.byref(self.last_stats)), "lb2_diskann_search")
        return D, I

    def search_device(self, d_q: int, nq: int, k: int, d_D: int, d_I: int, params: DiskannParams | None = None):
        p = params or make_diskann_params()
        _check(self._lib.lb2_diskann_search_device(self._h, int(nq), C.c_void_p(d_q), int(k), C.c_void_p(d_D),
                                                   C.c_void_p(d_I), C.byref(p), C.byref(self.last_stats)),
               "lb2_diskann_search_device")

    def last_expansions(self, nq: int, cap: int):
        ids = np.empty((nq, cap), np.uint32)
        n_full = np.empty(nq, np.int32)
        _check(self._lib.lb2_diskann_last_expansions(self._h, nq, cap, _np_ptr(ids), _np_ptr(n_full)),
               "lb2_diskann_last_expansions")
        return ids, n_full


# ---- graph construction stages (device pointers; tooling, see leann_b200/graph_build.py)
def build_workspace_bytes(ef: int, cap0: int) -> int:
    return int(load().lb2_build_workspace_bytes(int(ef), int(cap0)))


def build_insert_search(d_x_f16: int, n: int, d: int, metric_ip: bool, d_adj0: int, cap0: int, d_up_row: int, d_up_adj: int,
                        capU: int, entry: int, max_level: int, d_points: int, npts: int, ef: int, d_out_ids: int,
                        d_out_dist: int, d_workspace: int, workspace_bytes: int) -> None:
    _check(load().lb2_build_insert_search(C.c_void_p(d_x_f16), int(n), int(d), int(bool(metric_ip)), C.c_void_p(d_adj0),
                                          int(cap0), C.c_void_p(d_up_row), C.c_void_p(d_up_adj), int(capU), int(entry),
                                          int(max_level), C.c_void_p(d_points), int(npts), int(ef), C.c_void_p(d_out_ids),
                                          C.c_void_p(d_out_dist), C.c_void_p(d_workspace), int(workspace_bytes)),
           "lb2_build_insert_search")


def build_select(d_pd: int, pd_is_f32: bool, d_dn: int, d_cand: int, b: int, K: int, keep: int, d_out_ids: int,
                 d_out_dist: int, fill: int = 0) -> None:
    _check(load().lb2_build_select(C.c_void_p(d_pd), int(bool(pd_is_f32)), C.c_void_p(d_dn), C.c_void_p(d_cand), int(b), int(K),
                                   int(keep), int(fill), C.c_void_p(d_out_ids), C.c_void_p(d_out_dist)), "lb2_build_select")
