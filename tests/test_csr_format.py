"""CSR .index format: our reader against files written by the REFERENCE's converter, and
writer -> reader round trips."""
import numpy as np
import pytest

from leann_b200 import csr
from oracle.binding import Oracle


def test_reads_reference_written_file(golden_dir):
    g = csr.read_compact_index(str(golden_dir / "hnsw_small_ip.index"))
    assert (g.ntotal, g.d, g.metric_type) == (2000, 64, csr.METRIC_INNER_PRODUCT)
    assert g.storage_fourcc == csr.FOURCC_NULL and g.vectors is None
    assert g.node_offsets.size == g.ntotal + 1 and int(g.node_offsets[-1]) == g.level_ptr.size
    assert int(g.level_ptr[-1]) == g.neighbors.size
    assert np.array_equal(np.diff(g.node_offsets.astype(np.int64)) - 1, g.levels)
    assert 0 <= g.entry_point < g.ntotal and g.levels[g.entry_point] - 1 == g.max_level
    assert g.neighbors.min() >= 0 and g.neighbors.max() < g.ntotal


def test_reads_reference_file_with_flat_storage(golden_dir):
    g = csr.read_compact_index(str(golden_dir / "hnsw_small_ip_with_storage.index"))
    x = np.load(golden_dir / "hnsw_small_vectors.npy")
    assert g.vectors is not None and np.array_equal(g.vectors, x)
    g0 = csr.read_compact_index(str(golden_dir / "hnsw_small_ip.index"))
    assert np.array_equal(g.neighbors, g0.neighbors)


@pytest.mark.parametrize("with_vectors", [False, True])
def test_writer_reader_roundtrip_is_byte_identical_to_reference_layout(golden_dir, tmp_path, with_vectors):
    src = golden_dir / ("hnsw_small_ip_with_storage.index" if with_vectors else "hnsw_small_ip.index")
    g = csr.read_compact_index(str(src))
    out = tmp_path / "rt.index"
    csr.write_compact_index(str(out), g)
    assert out.read_bytes() == src.read_bytes()  # our writer == the reference's converter, byte for byte


def test_csr_from_padded_matches_list_constructor():
    rng = np.random.default_rng(0)
    n, M = 300, 4
    levels = np.minimum(rng.geometric(0.7, n), 3).astype(np.int32)
    level0 = np.full((n, 2 * M), -1, np.int32)
    adj = [dict() for _ in range(3)]
    for i in range(n):
        k = rng.integers(0, 2 * M + 1)
        nb = rng.choice(n, k, replace=False).astype(np.int32)
        level0[i, :k] = nb
        adj[0][i] = nb
    upper = {}
    for l in (1, 2):
        ids = np.nonzero(levels > l)[0]
        nbm = np.full((len(ids), M), -1, np.int32)
        for r, i in enumerate(ids):
            k = rng.integers(0, M + 1)
            nbm[r, :k] = rng.choice(n, k, replace=False)
            adj[l][i] = nbm[r, :k]
        upper[l] = (ids.astype(np.int64), nbm)
    ep = int(np.argmax(levels))
    a = csr.csr_from_padded(8, 0, levels, level0, upper, ep, M=M)
    b = csr.csr_from_level_lists(8, 0, levels, adj, ep, M=M)
    for f in ("level_ptr", "node_offsets", "neighbors", "levels"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for i in (0, 5, ep):
        for l in range(levels[i]):
            assert np.array_equal(a.neighbors_of(i, l), np.asarray(adj[l][i], np.int32))


def test_graph_builder_gives_navigable_graph():
    from leann_b200.graph_build import build_hnsw_graph
    from helpers import recall_at_k
    rng = np.random.default_rng(1)
    n, d = 6000, 32
    cent = rng.standard_normal((40, d))
    x = cent[rng.integers(0, 40, n)] + 0.5 * rng.standard_normal((n, d))
    x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    q = x[rng.choice(n, 50, replace=False)] + 0.05 * rng.standard_normal((50, d)).astype(np.float32)
    g = build_hnsw_graph(x, M=8, metric="mips", device="cpu")
    assert g.neighbors.max() < n and g.max_level >= 1
    deg0 = [len(g.neighbors_of(i, 0)) for i in range(0, n, 97)]
    assert max(deg0) <= 16
    D, I, _, _ = Oracle(g, x).search(q, 10, ef=64, nthreads=4)
    gt = np.argsort(-(q @ x.T), axis=1)[:, :10]
    assert recall_at_k(I, gt) >= 0.9
