"""Parity of the CUDA search path (through the C-ABI) with the CPU oracle / the committed
goldens produced by the compiled reference.  Bit-exact: ids, distances, ndis, nhops."""
import numpy as np
import pytest

from helpers import GOLDEN_CASES, golden_key, open_encoder_only, recall_at_k, stub_graph, write_leann_index
from leann_b200 import capi, csr, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
    return dict(x=np.load(golden_dir / "hnsw_small_vectors.npy"), q=np.load(golden_dir / "hnsw_small_queries.npy"),
                exp=np.load(golden_dir / "hnsw_small_expected.npz"), dir=golden_dir)


@pytest.mark.parametrize("tag", ["ip", "l2"])
def test_stored_vectors_match_reference_goldens(lib, cuda_ok, golden, tag):
    idx = capi.Index(str(golden["dir"] / f"hnsw_small_{tag}.index"))
    idx.set_vectors(golden["x"])
    for case in GOLDEN_CASES:
        ef, beam, batch, cr, k = case
        D, I = idx.search(golden["q"], k, capi.make_params(ef, beam, batch, bool(cr), recompute=False))
        ndis, nhops = idx.last_query_stats(len(golden["q"]))
        key = golden_key(tag, *case)
        assert np.array_equal(I, golden["exp"][key + "_I"]), key
        assert np.array_equal(D, golden["exp"][key + "_D"]), key
        assert np.array_equal(ndis, golden["exp"][key + "_ndis"]), key
        assert np.array_equal(nhops, golden["exp"][key + "_nhops"]), key


def test_index_with_flat_storage_needs_no_set_vectors(lib, cuda_ok, golden):
    idx = capi.Index(str(golden["dir"] / "hnsw_small_ip_with_storage.index"))
    assert idx.info.has_vectors == 1
    D, I = idx.search(golden["q"], 10, capi.make_params(64, recompute=False))
    assert np.array_equal(I, golden["exp"][golden_key("ip", 64, 1, 0, 1, 10) + "_I"])


def test_errors_and_edge_cases(lib, cuda_ok, golden, tmp_path):
    idx = capi.Index(str(golden["dir"] / "hnsw_small_ip.index"))
    with pytest.raises(capi.Lb2Error, match="stored vectors"):
        idx.search(golden["q"], 10, capi.make_params(recompute=False))
    with pytest.raises(capi.Lb2Error, match="lb2_set_passages"):
        idx.search(golden["q"], 10, capi.make_params(recompute=True))
    idx.set_vectors(golden["x"])
    # pruning parameters without loaded PQ tables are ignored, as in the reference (perform_pq_pruning needs an
    # initialised pq_data_loader, HNSW_search.cpp:442-445)
    D0, I0 = idx.search(golden["q"], 10, capi.make_params(recompute=False))
    D1, I1 = idx.search(golden["q"], 10, capi.make_params(prune_ratio=0.5, recompute=False))
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
    D, I = idx.search(golden["q"][:0], 10, capi.make_params(recompute=False))
    assert D.shape == (0, 10)
    # k larger than what a tiny graph can return: unfilled slots are (-1, -FLT_MAX) like faiss
    g = stub_graph(3, 64)
    f = tmp_path / "tiny.index"
    csr.write_compact_index(str(f), g)
    t = capi.Index(str(f))
    t.set_vectors(golden["x"][:3])
    D, I = t.search(golden["q"][:2], 5, capi.make_params(recompute=False))
    assert np.array_equal(I[:, 0], [0, 0]) and (I[:, 1:] == -1).all() and (D[:, 1:] == -np.finfo(np.float32).max).all()
    # empty index
    f0 = tmp_path / "empty.index"
    csr.write_compact_index(str(f0), stub_graph(0, 64))
    e = capi.Index(str(f0))
    e.set_vectors(np.zeros((0, 64), np.float32))
    D, I = e.search(golden["q"][:2], 3, capi.make_params(recompute=False))
    assert (I == -1).all()
    with pytest.raises(capi.Lb2Error, match="cannot open|not an IndexHNSWFlat"):
        capi.Index(str(tmp_path / "missing.index"))
    bad = tmp_path / "bad.index"
    bad.write_bytes((golden["dir"] / "hnsw_small_ip.index").read_bytes()[:5000])
    with pytest.raises(capi.Lb2Error, match="end of file|exceeds"):
        capi.Index(str(bad))


@pytest.fixture(scope="module")
def big(tmp_path_factory):
    """50k x 384 clustered vectors, graph from the torch builder (GPU), 400 queries."""
    from leann_b200.graph_build import build_hnsw_graph
    rng = np.random.default_rng(5)
    n, d, nq = 50000, 384, 400
    cent = rng.standard_normal((200, d)).astype(np.float32)
    x = cent[rng.integers(0, 200, n)] + 0.7 * rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = cent[rng.integers(0, 200, nq)] + 0.7 * rng.standard_normal((nq, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    g = build_hnsw_graph(x, M=16, metric="mips")
    f = tmp_path_factory.mktemp("big") / "big.index"
    csr.write_compact_index(str(f), g)
    return dict(x=x, q=q, g=g, f=f)


@pytest.mark.parametrize("case", [(64, 1, 0, 1, 10), (32, 4, 0, 1, 10), (48, 1, 64, 1, 5), (16, 2, 0, 0, 20), (200, 8, 0, 1, 100)])
def test_stored_vectors_match_oracle_at_scale(lib, cuda_ok, big, case):
    from oracle.binding import Oracle
    ef, beam, batch, cr, k = case
    idx = capi.Index(str(big["f"]))
    idx.set_vectors(big["x"])
    D, I = idx.search(big["q"], k, capi.make_params(ef, beam, batch, bool(cr), recompute=False))
    ndis, nhops = idx.last_query_stats(len(big["q"]))
    oD, oI, ond, onh = Oracle(big["g"], big["x"]).search(big["q"], k, ef=ef, beam=beam, batch_size=batch, check_rel=bool(cr), nthreads=16)
    assert np.array_equal(I, oI) and np.array_equal(D, oD)
    assert np.array_equal(ndis, ond) and np.array_equal(nhops, onh)
    if case == (64, 1, 0, 1, 10):
        gt = np.argsort(-(big["q"] @ big["x"].T), axis=1)[:, :10]
        assert recall_at_k(I, gt) > 0.9
    st = idx.last_stats
    assert st.ndis == ond.sum() and st.nhops == onh.sum() and st.n_kernel_launches >= 2


def test_l2_metric_and_repeat_calls_are_deterministic(lib, cuda_ok, big):
    idx = capi.Index(str(big["f"]))
    idx.set_vectors(big["x"])
    p = capi.make_params(40, 2, recompute=False)
    a = idx.search(big["q"], 10, p)
    b = idx.search(big["q"], 10, p)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    idx.configure(slots=7)  # far fewer slots than queries: slots are recycled
    c = idx.search(big["q"], 10, p)
    assert np.array_equal(a[1], c[1]) and np.array_equal(a[0], c[0])


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rc(tmp_path_factory):
    """Recompute fixture: 4000 synthetic passages, TINY encoder, graph built on the GPU embeddings."""
    from leann_b200.graph_build import build_hnsw_graph
    preset = synth.TINY
    w = synth.synthetic_weights(preset, 0)
    blob = synth.pack_weights(preset, w)
    tm, corpus = synth.make_corpus(4000, preset.vocab_size, seed=21, max_len=preset.max_pos)
    queries = synth.make_queries(tm, 64, seed=99)
    enc = open_encoder_only(preset, blob, corpus)
    E = enc.encode_ids(np.arange(corpus.n))
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    g = build_hnsw_graph(E, M=8, metric="mips")
    d = tmp_path_factory.mktemp("rc")
    index_path = write_leann_index(d, "rc", g, preset, corpus)
    return dict(preset=preset, w=w, blob=blob, corpus=corpus, E=E, Q=Q, g=g, index_path=index_path, dir=d)


@pytest.mark.parametrize("case", [(64, 1, 0, 1, 10), (32, 4, 0, 1, 10), (24, 1, 40, 1, 5)])
def test_recompute_search_is_exact_given_the_gpu_embeddings(lib, cuda_ok, rc, case):
    """Traversal + recompute + scoring on the GPU == oracle traversal over the embeddings the GPU
    encoder produces (the encoder's numerics are judged separately in test_gpu_encoder.py)."""
    from oracle.binding import Oracle
    ef, beam, batch, cr, k = case
    idx = capi.Index(str(rc["dir"] / "rc.index"))
    idx.set_passages(rc["corpus"].tokens, rc["corpus"].offsets)
    idx.set_encoder(rc["preset"].config(), rc["blob"])
    idx.configure(slots=24, passages_per_pass=256)  # several hops with recycled slots, several encoder passes per hop
    D, I = idx.search(rc["Q"], k, capi.make_params(ef, beam, batch, bool(cr), recompute=True))
    ndis, nhops = idx.last_query_stats(len(rc["Q"]))
    oD, oI, ond, onh = Oracle(rc["g"], rc["E"]).search(rc["Q"], k, ef=ef, beam=beam, batch_size=batch, check_rel=bool(cr), nthreads=8)
    assert np.array_equal(I, oI) and np.array_equal(D, oD)
    assert np.array_equal(ndis, ond) and np.array_equal(nhops, onh)
    st = idx.last_stats
    assert 0 < st.n_recomputed <= st.n_requested and st.n_tokens > 0 and st.encoder_ms > 0
    assert st.n_recomputed < st.n_requested  # cross-query de-duplication happened (shared entry point at least)


def test_recompute_search_bge_base_768d_cls_pooling(lib, cuda_ok, tmp_path):
    """Config-4 encoder family (bge-base: 12 layers, 768d, head_dim 64, CLS pooling) through the whole pipeline:
    GPU search == oracle traversal over the GPU's own embeddings, and the embeddings match the fp32 BertModel."""
    from leann_b200.graph_build import build_hnsw_graph
    from oracle.binding import Oracle
    from oracle.encoder_oracle import EncoderOracle
    preset = synth.BGE_BASE
    w = synth.synthetic_weights(preset, 3)
    blob = synth.pack_weights(preset, w)
    tm, corpus = synth.make_corpus(700, preset.vocab_size, seed=5, max_len=200)
    queries = synth.make_queries(tm, 24, seed=6)
    enc = open_encoder_only(preset, blob, corpus)
    E = enc.encode_ids(np.arange(corpus.n))
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    ref = EncoderOracle(preset, w).encode_store(corpus.tokens, corpus.offsets, ids=range(16))
    assert np.abs(E[:16] - ref).max() < 2e-3
    g = build_hnsw_graph(E, M=8, metric="mips")
    f = tmp_path / "bge.index"
    csr.write_compact_index(str(f), g)
    idx = capi.Index(str(f))
    idx.set_passages(corpus.tokens, corpus.offsets)
    idx.set_encoder(preset.config(), blob)
    idx.configure(slots=12, passages_per_pass=128)
    D, I = idx.search(Q, 10, capi.make_params(40, 2, recompute=True))
    oD, oI, ond, onh = Oracle(g, E).search(Q, 10, ef=40, beam=2, nthreads=8)
    assert np.array_equal(I, oI) and np.array_equal(D, oD)
    nd, nh = idx.last_query_stats(len(Q))
    assert np.array_equal(nd, ond) and np.array_equal(nh, onh)


def test_call_scope_dedup_gives_identical_results_with_fewer_recomputes(lib, cuda_ok, rc):
    idx = capi.Index(str(rc["dir"] / "rc.index"))
    idx.set_passages(rc["corpus"].tokens, rc["corpus"].offsets)
    idx.set_encoder(rc["preset"].config(), rc["blob"])
    idx.configure(slots=16, passages_per_pass=512)
    p = capi.make_params(48, 2, recompute=True)
    D0, I0 = idx.search(rc["Q"], 10, p)
    n_hop = idx.last_stats.n_recomputed
    idx.set_option("dedup_scope", 1)
    D1, I1 = idx.search(rc["Q"], 10, p)
    n_call = idx.last_stats.n_recomputed
    D2, I2 = idx.search(rc["Q"], 10, p)  # a second call starts from an empty table again
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1) and np.array_equal(I1, I2) and np.array_equal(D1, D2)
    assert n_call < n_hop and idx.last_stats.n_recomputed == n_call
    assert n_call <= rc["corpus"].n  # at most one encode per distinct passage per call
    idx.set_option("dedup_scope", 0)
    D3, I3 = idx.search(rc["Q"], 10, p)
    assert np.array_equal(I0, I3) and idx.last_stats.n_recomputed == n_hop
    with pytest.raises(capi.Lb2Error, match="unknown option"):
        idx.set_option("nope", 1)


def test_recompute_vs_fp32_reference_pipeline(lib, cuda_ok, rc):
    """Tier B: oracle traversal whose distances_batch is the fp32 BertModel forward
    (the reference's whole path on CPU) vs the GPU path: same ids where the gaps exceed fp16 noise,
    distances within 1e-3."""
    from oracle.binding import Oracle
    from oracle.encoder_oracle import EncoderOracle
    eo = EncoderOracle(rc["preset"], rc["w"])
    nq = 12
    q = eo.encode_store(*(lambda c: (c.tokens, c.offsets))(synth.make_queries(synth.make_corpus(4000, rc["preset"].vocab_size, seed=21, max_len=rc["preset"].max_pos)[0], 64, seed=99)))[:nq]
    oD, oI, _, _ = Oracle(rc["g"]).search(q, 10, ef=64, dist_fn=eo.distance_fn(rc["corpus"].tokens, rc["corpus"].offsets, True, 64))
    idx = capi.Index(str(rc["dir"] / "rc.index"))
    idx.set_passages(rc["corpus"].tokens, rc["corpus"].offsets)
    idx.set_encoder(rc["preset"].config(), rc["blob"])
    D, I = idx.search(q, 10, capi.make_params(64, recompute=True))
    overlap = recall_at_k(I, oI)
    assert overlap >= 0.9, overlap
    for i in range(nq):
        common = {int(v): j for j, v in enumerate(oI[i])}
        for j, v in enumerate(I[i]):
            if int(v) in common:
                assert abs(D[i, j] - oD[i, common[int(v)]]) < 1e-3


def test_plugin_searcher_end_to_end(lib, cuda_ok, rc):
    """Through the LEANN plugin API: factory -> searcher -> _ensure_server_running -> search."""
    import leann_b200
    from leann_b200 import backend  # noqa: F401  (registers)
    fac = leann_b200.BACKEND_REGISTRY["hnsw_b200"]
    s = fac.searcher(str(rc["index_path"]))
    with pytest.raises(ValueError, match="zmq_port must be provided"):
        s.search(rc["Q"][:2], 5, recompute_embeddings=True)
    with pytest.raises(RuntimeError, match="Recompute is required"):
        s.search(rc["Q"][:2], 5, recompute_embeddings=False)
    port = s._ensure_server_running(str(rc["index_path"]) + ".meta.json", 5557)
    out = s.search(rc["Q"][:5].astype(np.float64), 5, zmq_port=port, complexity=64, recompute_embeddings=True)
    assert set(out) == {"labels", "distances"}
    assert out["distances"].dtype == np.float32 and out["distances"].shape == (5, 5)
    assert all(isinstance(l, str) for row in out["labels"] for l in row) and len(out["labels"]) == 5
    assert (np.diff(out["distances"], axis=1) <= 0).all()  # +inner product, descending (IndexHNSW.cpp:433-438)
    qe = s.compute_query_embedding(rc["corpus"].passage(17))
    assert qe.shape == (1, rc["preset"].hidden) and np.array_equal(qe[0], rc["E"][17])
    top = s.search(qe, 1, zmq_port=port)["labels"][0][0]
    assert top == "17"
    # prune_ratio without PQ files is ignored like in the reference (no pq_data_loader)
    out2 = s.search(rc["Q"][:5], 5, zmq_port=port, prune_ratio=0.3)
    assert out2["labels"] == out["labels"]
    assert s.last_stats["ndis"] > 0
    # LeannSearcher.cleanup() -> embedding_server_manager.stop_server(): device handle released, next search re-opens it
    s.embedding_server_manager.stop_server()
    assert s._index is None and s.last_stats == {}
    out3 = s.search(rc["Q"][:5], 5, zmq_port=port, complexity=64, recompute_embeddings=True)
    assert out3["labels"] == out["labels"] and np.array_equal(out3["distances"], out["distances"])
    s.cleanup()
