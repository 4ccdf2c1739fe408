"""Plugin boundary, host side: registration and the error behaviour of the reference's
BaseSearcher / HNSWSearcher constructors (no GPU needed to reach these paths)."""
import json

import numpy as np
import pytest

import leann_b200
from leann_b200 import backend, interface


def test_backend_is_registered():  # shape of the reference's tests/test_ci_minimal.py:31-37
    assert "hnsw_b200" in leann_b200.BACKEND_REGISTRY
    fac = leann_b200.BACKEND_REGISTRY["hnsw_b200"]
    assert issubclass(fac, interface.LeannBackendFactoryInterface)
    assert callable(fac.searcher) and callable(fac.builder)


def test_missing_meta_raises_filenotfound(tmp_path):
    with pytest.raises(FileNotFoundError, match="metadata file not found"):
        backend.B200HnswSearcher(str(tmp_path / "idx.leann"))


def test_missing_dimensions_raises_valueerror(tmp_path):
    (tmp_path / "idx.leann.meta.json").write_text(json.dumps({"backend_name": "hnsw_b200"}))
    with pytest.raises(ValueError, match="Dimensions not found"):
        backend.B200HnswSearcher(str(tmp_path / "idx.leann"))


def test_missing_index_file_raises(tmp_path):
    (tmp_path / "idx.leann.meta.json").write_text(json.dumps({"dimensions": 64, "embedding_model": "m"}))
    with pytest.raises(FileNotFoundError, match="HNSW index file not found"):
        backend.B200HnswSearcher(str(tmp_path / "idx.leann"))


def test_unsupported_metric(tmp_path):
    (tmp_path / "idx.leann.meta.json").write_text(json.dumps({"dimensions": 64, "backend_kwargs": {"distance_metric": "hamming"}}))
    with pytest.raises(ValueError, match="Unsupported distance_metric"):
        backend.B200HnswSearcher(str(tmp_path / "idx.leann"))


def test_search_signature_matches_reference():
    import inspect
    sig = inspect.signature(backend.B200HnswSearcher.search)
    names = list(sig.parameters)
    for p in ("query", "top_k", "zmq_port", "complexity", "beam_width", "prune_ratio", "recompute_embeddings",
              "pruning_strategy", "batch_size"):
        assert p in names
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["complexity"], d["beam_width"], d["prune_ratio"], d["recompute_embeddings"], d["pruning_strategy"], d["batch_size"]) == (64, 1, 0.0, True, "global", 0)


def test_normalize_l2_zero_rows():
    x = np.array([[3.0, 4.0], [0.0, 0.0]], np.float32)
    y = backend.normalize_l2(x)
    assert np.allclose(y[0], [0.6, 0.8]) and np.all(y[1] == 0)


def test_builder_writes_reference_format(tmp_path):
    from leann_b200.csr import read_compact_index
    rng = np.random.default_rng(0)
    x = rng.standard_normal((500, 16)).astype(np.float32)
    b = leann_b200.BACKEND_REGISTRY["hnsw_b200"].builder(M=4, distance_metric="cosine", is_recompute=False)
    b.build(x, [str(i) for i in range(500)], str(tmp_path / "t.leann"))
    g = read_compact_index(str(tmp_path / "t.index"))
    assert g.ntotal == 500 and g.vectors is not None
    assert np.allclose(np.linalg.norm(g.vectors, axis=1), 1, atol=1e-5)
