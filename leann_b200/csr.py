"""Host-side reader/writer for the reference's compact-CSR HNSW ``.index`` file.

Layout (all little endian), as written by the reference's
``leann_backend_hnsw/convert_to_csr.py:182-237`` (``write_compact_format``) and
read by its faiss fork ``faiss/impl/index_read.cpp:523-813`` (``read_HNSW``)
and ``:1402-1490`` (``IHNf`` header):

    u32  fourcc 'IHNf'      i32 d        i64 ntotal     i64 dummy  i64 dummy
    u8   is_trained         i32 metric_type  [f32 metric_arg if metric_type > 1]
    vec<f64> assign_probas  vec<i32> cum_nneighbor_per_level   vec<i32> levels
    u8   storage_is_compact (=1)
    vec<u64> compact_level_ptr          vec<u64> compact_node_offsets (ntotal+1)
    i32  entry_point  i32 max_level  i32 efConstruction  i32 efSearch  i32 dummy
    u32  storage fourcc ('null' when the embeddings are pruned)
    vec<i32> compact_neighbors_data
    [storage index blob when storage fourcc != 'null']

``vec<T>`` is ``u64 count`` followed by ``count`` items.  This module is the
Python twin of the C++ loader in ``csrc/index_io.cpp`` (which is what the
search path uses); it exists for tooling: writing synthetic / GPU-built graphs
in the reference format and checking the C++ loader in tests.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

FOURCC_IHNF = int.from_bytes(b"IHNf", "little")
FOURCC_NULL = int.from_bytes(b"null", "little")
FOURCC_IXFI = int.from_bytes(b"IxFI", "little")  # IndexFlatIP
FOURCC_IXF2 = int.from_bytes(b"IxF2", "little")  # IndexFlatL2
METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


@dataclass
class CSRGraph:
    d: int
    ntotal: int
    metric_type: int
    levels: np.ndarray  # int32 [ntotal], level count per node (base level = 1)
    level_ptr: np.ndarray  # uint64 [sum(levels)+ntotal]
    node_offsets: np.ndarray  # uint64 [ntotal+1]
    neighbors: np.ndarray  # int32 [nnz]
    entry_point: int
    max_level: int
    ef_construction: int = 200
    ef_search: int = 16
    assign_probas: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))
    cum_nneighbor_per_level: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    metric_arg: float = 0.0
    storage_fourcc: int = FOURCC_NULL
    vectors: np.ndarray | None = None  # fp32 [ntotal, d] when a flat storage blob follows

    def neighbors_of(self, node: int, level: int) -> np.ndarray:
        ps, pe = int(self.node_offsets[node]), int(self.node_offsets[node + 1])
        nlev = max(pe - ps - 1, 0)
        if level < 0 or level >= nlev:
            return self.neighbors[:0]
        return self.neighbors[int(self.level_ptr[ps + level]) : int(self.level_ptr[ps + level + 1])]


def default_level_tables(M: int, M0: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """assign_probas / cum_nneighbor_per_level as HNSW::set_default_probas fills
    them (faiss/impl/HNSW.cpp:195-212): 2*M links on level 0, M above."""
    M0 = 2 * M if M0 is None else M0
    level_mult = 1.0 / np.log(M)
    probas, cum = [], [0]
    nn = 0
    level = 0
    while True:
        p = np.exp(-level / level_mult) * (1 - np.exp(-1 / level_mult))
        if p < 1e-9:
            break
        probas.append(p)
        nn += M0 if level == 0 else M
        cum.append(nn)
        level += 1
    return np.asarray(probas, np.float64), np.asarray(cum, np.int32)


def csr_from_level_lists(d, metric_type, levels, adj_by_level, entry_point, M=32, **kw) -> CSRGraph:
    """Build a CSRGraph from ``levels`` (int32 [N], >=1) and
    ``adj_by_level[l][node] -> iterable of neighbour ids`` (dict or list)."""
    levels = np.asarray(levels, np.int32)
    n = len(levels)
    node_offsets = np.zeros(n + 1, np.uint64)
    node_offsets[1:] = np.cumsum(levels.astype(np.int64) + 1)
    level_ptr = np.zeros(int(node_offsets[-1]), np.uint64)
    chunks = []
    pos = 0
    for i in range(n):
        base = int(node_offsets[i])
        for l in range(int(levels[i])):
            level_ptr[base + l] = pos
            nb = np.asarray(adj_by_level[l][i], np.int32)
            chunks.append(nb)
            pos += len(nb)
        level_ptr[base + int(levels[i])] = pos
    neighbors = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
    probas, cum = default_level_tables(M)
    return CSRGraph(d=d, ntotal=n, metric_type=metric_type, levels=levels, level_ptr=level_ptr,
                    node_offsets=node_offsets, neighbors=neighbors.astype(np.int32),
                    entry_point=int(entry_point), max_level=int(levels.max()) - 1 if n else -1,
                    assign_probas=probas, cum_nneighbor_per_level=cum, **kw)


def csr_from_padded(d, metric_type, levels, level0, upper, entry_point, M=32, **kw) -> CSRGraph:
    """Vectorised constructor for big graphs.

    level0: int32 [N, M0] padded with -1 ; upper: dict level(>=1) -> (node_ids int64 [n_l],
    nbrs int32 [n_l, M] padded with -1).  Rows are compacted (the -1 padding is dropped)."""
    levels = np.asarray(levels, np.int32)
    n = len(levels)
    node_offsets = np.zeros(n + 1, np.uint64)
    node_offsets[1:] = np.cumsum(levels.astype(np.int64) + 1)
    nptr = int(node_offsets[-1])
    # degree of (node, level) laid out in level_ptr order
    deg = np.zeros(nptr, np.int64)  # deg[p] = degree of the (node, level) that starts at p ; last slot per node = 0
    base = node_offsets[:-1].astype(np.int64)
    deg0 = (level0 >= 0).sum(1)
    deg[base] = deg0
    for l, (ids, nb) in upper.items():
        deg[base[ids] + l] = (nb >= 0).sum(1)
    level_ptr = np.zeros(nptr, np.uint64)
    # exclusive scan over the (node, level) slots; the trailing slot of each node has deg 0
    level_ptr[1:] = np.cumsum(deg[:-1])
    # But the last slot of node i must equal the end of its last level == start of node i+1's first level.
    # With deg[last slot] = 0 the exclusive scan already gives that.
    nnz = int(deg.sum())
    neighbors = np.empty(nnz, np.int32)
    # scatter level 0
    start0 = level_ptr[base].astype(np.int64)
    mask0 = level0 >= 0
    col_rank = np.cumsum(mask0, 1) - 1
    dst = (start0[:, None] + col_rank)[mask0]
    neighbors[dst] = level0[mask0]
    for l, (ids, nb) in upper.items():
        st = level_ptr[base[ids] + l].astype(np.int64)
        m = nb >= 0
        cr = np.cumsum(m, 1) - 1
        neighbors[(st[:, None] + cr)[m]] = nb[m]
    probas, cum = default_level_tables(M)
    return CSRGraph(d=d, ntotal=n, metric_type=metric_type, levels=levels, level_ptr=level_ptr,
                    node_offsets=node_offsets, neighbors=neighbors, entry_point=int(entry_point),
                    max_level=int(levels.max()) - 1 if n else -1, assign_probas=probas,
                    cum_nneighbor_per_level=cum, **kw)


def _wvec(f, arr: np.ndarray, dtype) -> None:
    arr = np.ascontiguousarray(arr, dtype=dtype)
    f.write(struct.pack("<Q", arr.size))
    f.write(arr.tobytes())


def _rvec(buf: memoryview, pos: int, dtype) -> tuple[np.ndarray, int]:
    (count,) = struct.unpack_from("<Q", buf, pos)
    pos += 8
    nbytes = count * np.dtype(dtype).itemsize
    arr = np.frombuffer(buf, dtype=dtype, count=count, offset=pos)
    return arr, pos + nbytes


def write_compact_index(path: str, g: CSRGraph) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<I", FOURCC_IHNF))
        f.write(struct.pack("<i", g.d))
        f.write(struct.pack("<q", g.ntotal))
        f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<?", True))
        f.write(struct.pack("<i", g.metric_type))
        if g.metric_type > 1:
            f.write(struct.pack("<f", g.metric_arg))
        _wvec(f, g.assign_probas, np.float64)
        _wvec(f, g.cum_nneighbor_per_level, np.int32)
        _wvec(f, g.levels, np.int32)
        f.write(struct.pack("<?", True))
        _wvec(f, g.level_ptr, np.uint64)
        _wvec(f, g.node_offsets, np.uint64)
        f.write(struct.pack("<5i", g.entry_point, g.max_level, g.ef_construction, g.ef_search, 1))
        has_vec = g.vectors is not None
        fourcc = (FOURCC_IXFI if g.metric_type == METRIC_INNER_PRODUCT else FOURCC_IXF2) if has_vec else FOURCC_NULL
        f.write(struct.pack("<I", fourcc))
        _wvec(f, g.neighbors, np.int32)
        if has_vec:
            # IndexFlat blob (faiss/impl/index_write.cpp:81-92 write_index_header, :419-426):
            # the common header again, then WRITEXBVECTOR(codes): u64 count of 4-byte
            # words (= floats, impl/io_macros.h:73-79) followed by the raw bytes
            f.write(struct.pack("<i", g.d))
            f.write(struct.pack("<q", g.ntotal))
            f.write(struct.pack("<q", 1 << 20))
            f.write(struct.pack("<q", 1 << 20))
            f.write(struct.pack("<?", True))
            f.write(struct.pack("<i", g.metric_type))
            v = np.ascontiguousarray(g.vectors, np.float32)
            f.write(struct.pack("<Q", v.size))
            f.write(v.tobytes())


def read_compact_index(path: str) -> CSRGraph:
    buf = memoryview(np.fromfile(path, dtype=np.uint8))
    pos = 0
    (fourcc,) = struct.unpack_from("<I", buf, pos); pos += 4
    if fourcc != FOURCC_IHNF:
        raise ValueError(f"{path}: not an IndexHNSWFlat file (fourcc {fourcc:08x})")
    d, = struct.unpack_from("<i", buf, pos); pos += 4
    ntotal, _, _ = struct.unpack_from("<3q", buf, pos); pos += 24
    pos += 1  # is_trained
    metric_type, = struct.unpack_from("<i", buf, pos); pos += 4
    metric_arg = 0.0
    if metric_type > 1:
        metric_arg, = struct.unpack_from("<f", buf, pos); pos += 4
    probas, pos = _rvec(buf, pos, np.float64)
    cum, pos = _rvec(buf, pos, np.int32)
    levels, pos = _rvec(buf, pos, np.int32)
    flag = buf[pos]; pos += 1
    if flag != 1:
        raise ValueError(f"{path}: expected the compact-CSR flag, got {flag}")
    level_ptr, pos = _rvec(buf, pos, np.uint64)
    node_offsets, pos = _rvec(buf, pos, np.uint64)
    entry_point, max_level, efc, efs, _ = struct.unpack_from("<5i", buf, pos); pos += 20
    storage_fourcc, = struct.unpack_from("<I", buf, pos); pos += 4
    neighbors, pos = _rvec(buf, pos, np.int32)
    vectors = None
    if storage_fourcc in (FOURCC_IXFI, FOURCC_IXF2):
        pos += 4 + 24 + 1 + 4
        (nwords,) = struct.unpack_from("<Q", buf, pos); pos += 8
        vectors = np.frombuffer(buf, np.float32, count=nwords, offset=pos).reshape(ntotal, d)
    return CSRGraph(d=d, ntotal=ntotal, metric_type=metric_type, levels=levels, level_ptr=level_ptr,
                    node_offsets=node_offsets, neighbors=neighbors, entry_point=entry_point,
                    max_level=max_level, ef_construction=efc, ef_search=efs, assign_probas=probas,
                    cum_nneighbor_per_level=cum, metric_arg=metric_arg,
                    storage_fourcc=storage_fourcc, vectors=vectors)
