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


def test_registers_into_the_real_leann_registry_when_leann_is_importable():
    """With the reference's own leann-core on the path, the backend lands in LEANN's BACKEND_REGISTRY and
    implements LEANN's interfaces (what LeannSearcher looks up by meta['backend_name'], CORE/api.py:625-637)."""
    import os
    import subprocess
    import sys
    core = "/root/reference/packages/leann-core/src"
    if not os.path.isdir(core):
        pytest.skip("reference checkout not present (GPU box)")
    code = (
        "import leann.registry as r, leann.interface as i\n"
        "import leann_b200.backend as b, leann_b200.interface as li\n"
        "assert li.HAVE_LEANN\n"
        "assert r.BACKEND_REGISTRY['hnsw_b200'] is b.B200HnswBackend\n"
        "assert issubclass(b.B200HnswBackend, i.LeannBackendFactoryInterface)\n"
        "assert issubclass(b.B200HnswSearcher, i.LeannBackendSearcherInterface)\n"
        "assert issubclass(b.B200HnswBuilder, i.LeannBackendBuilderInterface)\n"
        "import inspect\n"
        "ref = inspect.signature(i.LeannBackendSearcherInterface.search).parameters\n"
        "mine = inspect.signature(b.B200HnswSearcher.search).parameters\n"
        "assert all(k in mine for k in ref if k != 'kwargs'), (list(ref), list(mine))\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=core + os.pathsep + str(__import__('pathlib').Path(__file__).resolve().parents[1]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


# ---------------------------------------------------------------- DiskANN plugin (diskann_b200)
def test_diskann_backend_registered_and_constructor_errors(tmp_path):
    from leann_b200 import diskann_backend as db
    fac = leann_b200.BACKEND_REGISTRY["diskann_b200"]
    assert issubclass(fac, interface.LeannBackendFactoryInterface) and callable(fac.searcher) and callable(fac.builder)
    with pytest.raises(FileNotFoundError, match="metadata file not found"):
        db.B200DiskannSearcher(str(tmp_path / "idx.leann"))
    (tmp_path / "idx.leann.meta.json").write_text(json.dumps({"dimensions": 16, "backend_kwargs": {"distance_metric": "hamming"}}))
    with pytest.raises(ValueError, match="Unsupported distance_metric"):
        db.B200DiskannSearcher(str(tmp_path / "idx.leann"))
    (tmp_path / "idx.leann.meta.json").write_text(json.dumps({"dimensions": 16, "embedding_model": "m"}))
    with pytest.raises(FileNotFoundError, match="DiskANN index files not found"):
        db.B200DiskannSearcher(str(tmp_path / "idx.leann"))


def test_diskann_builder_writes_the_reference_file_set_and_search_fails_loudly_without_gpu(tmp_path):
    """is_recompute=True leaves the partition layout and drops _disk.index, like DiskannBuilder + _safe_cleanup_after_partition
    (diskann_backend.py:128-190, 268-291); without a GPU the searcher raises instead of falling back to anything."""
    import torch
    from leann_b200 import diskann_backend as db
    from leann_b200 import diskann_format as dfmt
    rng = np.random.default_rng(0)
    data = rng.standard_normal((600, 16)).astype(np.float32)
    ids = [str(i) for i in range(len(data))]
    b = db.B200DiskannBackend.builder(distance_metric="mips", graph_degree=8, pq_chunks=4)
    b.build(data, ids, str(tmp_path / "docs.leann"), is_recompute=True)
    files = dfmt.index_files(str(tmp_path / "docs"))
    assert files["_pq_pivots.bin"] and files["_pq_compressed.bin"] and files["_disk.index_medoids.bin"]
    assert files["_partition.bin"] and files["_disk_graph.index"] and files["_disk.index_max_base_norm.bin"]
    assert not files["_disk.index"]
    with pytest.raises(ValueError, match="decimal strings"):
        b.build(data, ["a"] * len(data), str(tmp_path / "bad.leann"))
    (tmp_path / "docs.leann.meta.json").write_text(json.dumps({"dimensions": 16, "embedding_model": "synthetic/tiny-bert",
                                                                "backend_kwargs": {"distance_metric": "mips"}}))
    s = db.B200DiskannBackend.searcher(str(tmp_path / "docs.leann"))
    assert s._partition_prefix.endswith("docs")
    q = rng.standard_normal((2, 16)).astype(np.float32)
    with pytest.raises(ValueError, match="zmq_port must be provided"):
        s.search(q, 5, recompute_embeddings=True)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU path"):
            s.search(q, 5, recompute_embeddings=False, skip_search_reorder=True)


def test_token_store_equals_the_reference_tokenizer_call(tmp_path):
    """the uint16 sidecars written once by tokenize_passages hold exactly the ids the reference's
    per-hop tokenizer call produces (embedding_compute.py:299-305: hf_tokenizer(batch, padding=True, truncation=True,
    max_length=...)) once padding is stripped by the attention mask — WordPiece, [CLS]/[SEP], truncation included.
    The vocabulary is built locally (no checkpoint on disk); the tokenizer classes are the real ones."""
    from tokenizers import BertWordPieceTokenizer
    from transformers import PreTrainedTokenizerFast

    from leann_b200.backend import tokenize_passages

    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "quick", "brown", "fox", "##es", "jump", "##s", "over", "lazy",
             "dog", ".", ",", "pride", "and", "prejudice", "un", "##believ", "##able", "it", "is", "a", "truth", "##ly"]
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(words))
    tok = PreTrainedTokenizerFast(tokenizer_object=BertWordPieceTokenizer(str(vf), lowercase=True)._tokenizer, unk_token="[UNK]",
                                  pad_token="[PAD]", cls_token="[CLS]", sep_token="[SEP]", mask_token="[MASK]")
    texts = ["The quick brown foxes jumps over the lazy dog.", "Pride and Prejudice", "unbelievable, truthly unbelievable zebra",
             "it is a truth " * 40, "."]
    max_len = 24
    tokenize_passages(str(tmp_path / "t.leann"), texts, tok, max_len)
    toks = np.load(tmp_path / "t.leann.tokens.npy")
    offs = np.load(tmp_path / "t.leann.tokoffsets.npy")
    assert toks.dtype == np.uint16 and offs.dtype == np.uint64 and len(offs) == len(texts) + 1
    ref = tok(texts, padding=True, truncation=True, max_length=max_len, return_tensors="np")  # the reference's call
    for i in range(len(texts)):
        want = ref["input_ids"][i][ref["attention_mask"][i] == 1]
        got = toks[int(offs[i]): int(offs[i + 1])]
        assert np.array_equal(got, want), (i, got, want)
        assert got[0] == 2 and got[-1] == 3 and len(got) <= max_len  # [CLS] ... [SEP], truncated like the reference
    with pytest.raises(ValueError):  # ids that do not fit the uint16 store are refused, never wrapped
        tokenize_passages(str(tmp_path / "u.leann"), ["x"], lambda t, truncation, max_length: {"input_ids": [101, 70000, 102]}, 8)
