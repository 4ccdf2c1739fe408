"""On-disk artefacts of the reference's DiskANN backend — writers and readers.

The CUDA Vamana searcher (csrc/vamana.cu) opens exactly the files
`PQFlashIndex::load_from_separate_paths` opens (DiskANN/src/pq_flash_index.cpp:1017-1467), so an index
built by the stock `leann-backend-diskann` loads unchanged.  These helpers exist for the other direction:
tests, the bench and `B200DiskannBuilder` need to *produce* such an index without the reference's
builder (not compilable here: MKL, libaio, Boost).  All layouts below follow the reference's readers:

  <p>_pq_pivots.bin        FixedChunkPQTable::load_pq_centroid_bin (DiskANN/src/pq.cpp:49-168):
                           bin<size_t>[4,1] of byte offsets at 0, then at those offsets
                           bin<float>[256, ndims] pivots, bin<float>[ndims, 1] centroid,
                           bin<u32>[n_chunks+1, 1] chunk offsets.   "bin" = i32 rows, i32 cols, data.
  <p>_pq_compressed.bin    bin<u8>[npts, n_chunks]                      (pq_flash_index.cpp:1056-1063)
  <p>_disk.index           sector 0: bin<u64> metadata {npts, ndims, medoid, max_node_len,
                           nnodes_per_sector, n_frozen, frozen_id, reorder_exists, file_size}
                           (pq_flash_index.cpp:1292-1362); node i lives in sector
                           1 + i / nnodes_per_sector at byte (i % nnodes_per_sector) * max_node_len
                           (:118-129) as {float coords[ndims]; u32 nnbrs; u32 nbrs[nnbrs]}.
  <p>_disk.index_medoids.bin          bin<u32>[m, 1]                    (pq_flash_index.cpp:1369-1384)
  <p>_disk.index_max_base_norm.bin    bin<float>[1, 1], MIPS only       (pq_flash_index.cpp:1432-1452)
  partition mode (what `is_recompute=True` builds leave behind, diskann_backend.py:268-291):
  <p>_partition.bin        u64 C, u64 n_partitions, u64 nd, then per partition {u32 size, u32 ids[size]},
                           then u32 id2partition[nd]                    (pq_flash_index.cpp:916-948)
  <p>_disk_graph.index     sector 0: i32 meta_n, i32 meta_dim, u64 meta[meta_n] with meta[0]=nd,
                           meta[1]=dim, meta[3]=max_node_len, meta[4]=C, meta[8]=file size (:951-1008);
                           partition p lives in sector p+1, its j-th node at byte j*graph_node_len as
                           {u32 nnbrs; u32 nbrs[...]} with graph_node_len = max_node_len - 4*dim
                           (pq_flash_index.cpp:2248-2290, 2509-2564).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

SECTOR_LEN = 4096
NUM_PQ_CENTROIDS = 256
METRIC_L2, METRIC_INNER_PRODUCT, METRIC_COSINE = 0, 1, 2
METRIC_BY_NAME = {"l2": METRIC_L2, "mips": METRIC_INNER_PRODUCT, "cosine": METRIC_COSINE}


# ------------------------------------------------------------------------------------------------ bin files
def write_bin(path, arr: np.ndarray, offset: int | None = None, fh=None) -> int:
    """DiskANN 'bin': i32 rows, i32 cols, row-major data.  Returns the number of bytes written."""
    a = np.ascontiguousarray(arr)
    if a.ndim == 1:
        a = a[:, None]
    hdr = struct.pack("<ii", a.shape[0], a.shape[1])
    if fh is not None:
        if offset is not None:
            fh.seek(offset)
        fh.write(hdr)
        fh.write(a.tobytes())
    else:
        with open(path, "wb") as f:
            f.write(hdr)
            f.write(a.tobytes())
    return 8 + a.nbytes


def read_bin(path, dtype, offset: int = 0) -> np.ndarray:
    with open(path, "rb") as f:
        f.seek(offset)
        r, c = struct.unpack("<ii", f.read(8))
        data = np.fromfile(f, dtype=dtype, count=r * c)
    if data.size != r * c:
        raise ValueError(f"{path}: truncated bin ({data.size} of {r * c} values)")
    return data.reshape(r, c)


# ------------------------------------------------------------------------------------------------ base prep
def prepare_base(emb: np.ndarray, metric: str) -> tuple[np.ndarray, float]:
    """What the reference's builder stores as node coordinates.  MIPS: every vector divided by the
    largest base norm plus the extra coordinate sqrt(1 - |e|^2 / max^2), which turns inner-product search into
    L2 search (the same arithmetic the searcher applies to freshly fetched embeddings,
    preprocess_fetched_embeddings, pq_flash_index.cpp:1723-1777); cosine: unit vectors; L2: unchanged.
    Returns (coords [n, data_dim], max_base_norm)."""
    e = np.ascontiguousarray(emb, np.float32)
    m = metric.lower()
    if m == "mips":
        norm_sq = (e * e).sum(1, dtype=np.float32)
        max_norm = np.float32(np.sqrt(norm_sq.max())) if len(e) else np.float32(1)
        res = np.float32(1) - norm_sq / (max_norm * max_norm)
        extra = np.sqrt(np.maximum(res, np.float32(0)), dtype=np.float32)
        return np.concatenate([e / max_norm, extra[:, None]], axis=1).astype(np.float32), float(max_norm)
    if m == "cosine":
        n = np.sqrt((e * e).sum(1, dtype=np.float32))
        n = np.where(n > 0, n, np.float32(1))
        return (e / n[:, None]).astype(np.float32), 0.0
    if m == "l2":
        return e.copy(), 0.0
    raise ValueError(f"unsupported distance_metric {metric!r}")


# ------------------------------------------------------------------------------------------------ PQ
def default_num_chunks(n: int, dim: int) -> int:
    """PQ bytes per vector the reference's default budget yields: search memory = embeddings / 10
    (diskann_backend.py:105-111), chunks = budget / npts clipped to [1, dim] (DiskANN/src/disk_utils.cpp)."""
    budget_gb = max(0.1, n * dim * 4 / (1024 ** 3) / 10)
    return int(max(1, min(dim, min(512, budget_gb * (1024 ** 3) // max(n, 1)))))


def even_chunk_offsets(ndims: int, n_chunks: int) -> np.ndarray:
    lo, rem = divmod(ndims, n_chunks)
    sizes = np.full(n_chunks, lo, np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)


@dataclass
class PQTable:
    pivots: np.ndarray         # [256, ndims] float32, centred
    centroid: np.ndarray       # [ndims] float32
    chunk_offsets: np.ndarray  # [n_chunks + 1] uint32

    @property
    def ndims(self) -> int:
        return int(self.pivots.shape[1])

    @property
    def n_chunks(self) -> int:
        return int(len(self.chunk_offsets) - 1)


def train_pq(coords: np.ndarray, n_chunks: int, zero_mean: bool, iters: int = 8, seed: int = 0,
             max_train: int = 262144, device: str | None = None) -> PQTable:
    """256-centroid k-means per chunk (Lloyd, k-means++-free random init) with torch.  The reference trains with
    MKL k-means (generate_pq_pivots, DiskANN/src/pq.cpp); any codebook is a valid index — search semantics do
    not depend on how it was trained.  zero_mean = subtract the global centroid first (L2/cosine; not MIPS)."""
    import torch
    n, ndims = coords.shape
    n_chunks = max(1, min(n_chunks, ndims))
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    rng = np.random.default_rng(seed)
    sample = coords if n <= max_train else coords[rng.choice(n, max_train, replace=False)]
    x = torch.from_numpy(np.ascontiguousarray(sample, np.float32)).to(dev)
    centroid = x.mean(0) if zero_mean else torch.zeros(ndims, device=dev)
    x = x - centroid
    offs = even_chunk_offsets(ndims, n_chunks)
    pivots = torch.zeros((NUM_PQ_CENTROIDS, ndims), dtype=torch.float32, device=dev)
    g = torch.Generator(device="cpu").manual_seed(seed)
    for c in range(n_chunks):
        xc = x[:, int(offs[c]):int(offs[c + 1])]
        pick = torch.randint(0, xc.shape[0], (NUM_PQ_CENTROIDS,), generator=g).to(dev)
        cen = xc[pick].clone()
        for _ in range(iters):
            d = (xc * xc).sum(1)[:, None] - 2 * xc @ cen.T + (cen * cen).sum(1)[None, :]
            a = d.argmin(1)
            sums = torch.zeros_like(cen).index_add_(0, a, xc)
            cnt = torch.bincount(a, minlength=NUM_PQ_CENTROIDS).float()
            cen = torch.where(cnt[:, None] > 0, sums / cnt.clamp(min=1)[:, None], cen)
        pivots[:, int(offs[c]):int(offs[c + 1])] = cen
    return PQTable(pivots.cpu().numpy(), centroid.cpu().numpy().astype(np.float32), offs)


def encode_pq(coords: np.ndarray, pq: PQTable, block: int = 65536, device: str | None = None) -> np.ndarray:
    """Nearest pivot per chunk -> codes [n, n_chunks] uint8 (generate_pq_data_from_pivots, DiskANN/src/pq.cpp)."""
    import torch
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    n = coords.shape[0]
    piv = torch.from_numpy(pq.pivots).to(dev)
    cen = torch.from_numpy(pq.centroid).to(dev)
    codes = np.empty((n, pq.n_chunks), np.uint8)
    for b0 in range(0, n, block):
        x = torch.from_numpy(np.ascontiguousarray(coords[b0:b0 + block], np.float32)).to(dev) - cen
        out = torch.empty((x.shape[0], pq.n_chunks), dtype=torch.uint8, device=dev)
        for c in range(pq.n_chunks):
            lo, hi = int(pq.chunk_offsets[c]), int(pq.chunk_offsets[c + 1])
            xc, pc = x[:, lo:hi], piv[:, lo:hi]
            d = (xc * xc).sum(1)[:, None] - 2 * xc @ pc.T + (pc * pc).sum(1)[None, :]
            out[:, c] = d.argmin(1).to(torch.uint8)
        codes[b0:b0 + block] = out.cpu().numpy()
    return codes


def write_pq_pivots(path, pq: PQTable) -> None:
    METADATA_SIZE = 4096
    with open(path, "wb") as f:
        f.write(b"\0" * METADATA_SIZE)
        offs = [METADATA_SIZE]
        offs.append(offs[-1] + write_bin(None, pq.pivots.astype(np.float32), offs[-1], f))
        offs.append(offs[-1] + write_bin(None, pq.centroid.astype(np.float32)[:, None], offs[-1], f))
        offs.append(offs[-1] + write_bin(None, pq.chunk_offsets.astype(np.uint32)[:, None], offs[-1], f))
        write_bin(None, np.asarray(offs, np.uint64)[:, None], 0, f)


def read_pq_pivots(path) -> PQTable:
    offs = read_bin(path, np.uint64).ravel()
    if len(offs) not in (4, 5):
        raise ValueError(f"{path}: expected 4 or 5 offsets, got {len(offs)}")
    piv = read_bin(path, np.float32, int(offs[0]))
    cen = read_bin(path, np.float32, int(offs[1])).ravel()
    co = read_bin(path, np.uint32, int(offs[3 if len(offs) == 5 else 2])).ravel()
    return PQTable(piv, cen, co)


# ------------------------------------------------------------------------------------------------ graph files
@dataclass
class VamanaGraph:
    """Flat Vamana graph: nbrs [n, R] int32 padded with -1, degree-prefix packed (valid ids first)."""
    nbrs: np.ndarray
    medoid: int

    @property
    def n(self) -> int:
        return int(self.nbrs.shape[0])

    @property
    def R(self) -> int:
        return int(self.nbrs.shape[1])

    def degrees(self) -> np.ndarray:
        return (self.nbrs >= 0).sum(1).astype(np.uint32)


def write_disk_index(path, coords: np.ndarray, g: VamanaGraph) -> None:
    n, ndims = coords.shape
    R = g.R
    max_node_len = ndims * 4 + 4 + R * 4
    nnodes_per_sector = SECTOR_LEN // max_node_len  # 0 -> multi-sector nodes
    sectors_per_node = -(-max_node_len // SECTOR_LEN)
    n_sectors = (-(-n // nnodes_per_sector)) if nnodes_per_sector > 0 else n * sectors_per_node
    file_size = (n_sectors + 1) * SECTOR_LEN
    deg = g.degrees()
    node = np.zeros((n, max_node_len), np.uint8)
    node[:, : ndims * 4] = np.ascontiguousarray(coords, np.float32).view(np.uint8).reshape(n, ndims * 4)
    node[:, ndims * 4: ndims * 4 + 4] = deg.astype("<u4").view(np.uint8).reshape(n, 4)
    nb = np.where(g.nbrs >= 0, g.nbrs, 0).astype("<u4")
    node[:, ndims * 4 + 4:] = nb.view(np.uint8).reshape(n, R * 4)
    meta = np.asarray([n, ndims, g.medoid, max_node_len, nnodes_per_sector, 0, 0, 0, file_size], np.uint64)
    with open(path, "wb") as f:
        first = bytearray(SECTOR_LEN)
        first[0:8] = struct.pack("<ii", len(meta), 1)
        first[8:8 + meta.nbytes] = meta.tobytes()
        f.write(first)
        if nnodes_per_sector > 0:
            pad = n_sectors * nnodes_per_sector - n
            body = np.concatenate([node, np.zeros((pad, max_node_len), np.uint8)]) if pad else node
            body = body.reshape(n_sectors, nnodes_per_sector * max_node_len)
            sect = np.zeros((n_sectors, SECTOR_LEN), np.uint8)
            sect[:, : body.shape[1]] = body
            f.write(sect.tobytes())
        else:
            sect = np.zeros((n, sectors_per_node * SECTOR_LEN), np.uint8)
            sect[:, :max_node_len] = node
            f.write(sect.tobytes())


def read_disk_index(path) -> tuple[np.ndarray, VamanaGraph]:
    with open(path, "rb") as f:
        nr, nc = struct.unpack("<ii", f.read(8))
        meta = np.frombuffer(f.read(8 * nr * nc), np.uint64)
    n, ndims, medoid, max_node_len, nps = (int(v) for v in meta[:5])
    R = (max_node_len - ndims * 4) // 4 - 1
    raw = np.fromfile(path, np.uint8, offset=SECTOR_LEN)
    if nps > 0:
        n_sectors = -(-n // nps)
        raw = raw[: n_sectors * SECTOR_LEN].reshape(n_sectors, SECTOR_LEN)[:, : nps * max_node_len]
        node = raw.reshape(n_sectors * nps, max_node_len)[:n]
    else:
        spn = -(-max_node_len // SECTOR_LEN)
        node = raw[: n * spn * SECTOR_LEN].reshape(n, spn * SECTOR_LEN)[:, :max_node_len]
    node = np.ascontiguousarray(node)
    coords = node[:, : ndims * 4].copy().view(np.float32).reshape(n, ndims)
    deg = node[:, ndims * 4: ndims * 4 + 4].copy().view("<u4").ravel()
    nb = node[:, ndims * 4 + 4:].copy().view("<u4").reshape(n, R).astype(np.int64)
    nb[np.arange(R)[None, :] >= deg[:, None]] = -1
    return coords, VamanaGraph(nb.astype(np.int32), medoid)


def write_partition_files(prefix: str, g: VamanaGraph, ndims: int, order: np.ndarray | None = None) -> None:
    """Adjacency-only sectors + the id -> partition map.  `order` = node ids in partition order (default:
    id order; the reference's partitioner groups graph neighbours, any grouping is a valid file)."""
    n, R = g.n, g.R
    graph_node_len = 4 + 4 * R
    max_node_len = graph_node_len + 4 * ndims
    C = SECTOR_LEN // graph_node_len
    order = np.arange(n, dtype=np.uint32) if order is None else np.asarray(order, np.uint32)
    n_part = -(-n // C)
    id2p = np.empty(n, np.uint32)
    deg = g.degrees()
    nb = np.where(g.nbrs >= 0, g.nbrs, 0).astype("<u4")
    file_size = (n_part + 1) * SECTOR_LEN
    with open(prefix + "_partition.bin", "wb") as pf, open(prefix + "_disk_graph.index", "wb") as gf:
        pf.write(struct.pack("<QQQ", C, n_part, n))
        meta = np.zeros(9, np.uint64)
        meta[0], meta[1], meta[2], meta[3], meta[4], meta[8] = n, ndims, g.medoid, max_node_len, C, file_size
        first = bytearray(SECTOR_LEN)
        first[0:8] = struct.pack("<ii", len(meta), 1)
        first[8:8 + meta.nbytes] = meta.tobytes()
        gf.write(first)
        for p in range(n_part):
            ids = order[p * C:(p + 1) * C]
            id2p[ids] = p
            pf.write(struct.pack("<I", len(ids)))
            pf.write(ids.astype("<u4").tobytes())
            sect = np.zeros(SECTOR_LEN, np.uint8)
            rec = np.zeros((len(ids), graph_node_len), np.uint8)
            rec[:, :4] = deg[ids].astype("<u4").view(np.uint8).reshape(-1, 4)
            rec[:, 4:] = nb[ids].view(np.uint8).reshape(len(ids), 4 * R)
            sect[: rec.size] = rec.ravel()
            gf.write(sect.tobytes())
        pf.write(id2p.astype("<u4").tobytes())


def write_diskann_index(dirpath, name: str, emb: np.ndarray, g: VamanaGraph, coords: np.ndarray, max_base_norm: float,
                        pq: PQTable, codes: np.ndarray, metric: str, partition: bool = False,
                        keep_disk_index: bool = True, partition_order: np.ndarray | None = None) -> str:
    """Writes the file set `DiskannSearcher` opens for prefix <dirpath>/<name> and returns that prefix."""
    d = Path(dirpath)
    d.mkdir(parents=True, exist_ok=True)
    prefix = str(d / name)
    write_pq_pivots(prefix + "_pq_pivots.bin", pq)
    write_bin(prefix + "_pq_compressed.bin", codes.astype(np.uint8))
    write_bin(prefix + "_disk.index_medoids.bin", np.asarray([[g.medoid]], np.uint32))
    if metric.lower() == "mips":
        write_bin(prefix + "_disk.index_max_base_norm.bin", np.asarray([[max_base_norm]], np.float32))
    if partition:
        write_partition_files(prefix, g, coords.shape[1], partition_order)
    if keep_disk_index or not partition:
        write_disk_index(prefix + "_disk.index", coords, g)
    return prefix


def index_files(prefix: str) -> dict[str, bool]:
    names = ["_pq_pivots.bin", "_pq_compressed.bin", "_disk.index", "_disk.index_medoids.bin",
             "_disk.index_max_base_norm.bin", "_partition.bin", "_disk_graph.index"]
    return {s: os.path.exists(prefix + s) for s in names}
