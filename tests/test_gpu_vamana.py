"""Parity of the CUDA DiskANN/Vamana path (through the C-ABI) with the CPU oracle (oracle/vamana_oracle.c, whose queue and
PQ primitives are pinned to the reference's compiled code).  Bit-exact: expansion lists, ids, distances, counters."""
import numpy as np
import pytest

from helpers import open_encoder_only
from leann_b200 import capi, synth
from leann_b200.tooling import write_diskann_leann_index
from leann_b200.vamana_build import build_diskann_index
from oracle.vamana_binding import VamanaOracle

pytestmark = pytest.mark.gpu
FMAX = np.finfo(np.float32).max


def rows(n, d, seed, clusters=12, spread=0.45):
    rng = np.random.default_rng(seed)
    cen = rng.standard_normal((clusters, d)).astype(np.float32)
    return cen[rng.integers(0, clusters, n)] + spread * rng.standard_normal((n, d)).astype(np.float32)


def check_same(idx, o, q, k, L, beam, coords=None, emb=None, skip=False, io_limit=0):
    p = capi.make_diskann_params(L, beam, recompute_embeddings=emb is not None, skip_search_reorder=skip, io_limit=io_limit)
    D, I = idx.search(q, k, p)
    oD, oI, info = o.search(q, k, L=L, beam_width=beam, coords=coords, emb=emb, skip_search_reorder=skip,
                            io_limit=io_limit or 0xFFFFFFFF, nthreads=8)
    cap = info["full_ids"].shape[1]
    ids, n_full = idx.last_expansions(len(q), cap)
    assert np.array_equal(n_full, info["n_full"])
    for i in range(len(q)):
        assert np.array_equal(ids[i, :n_full[i]], info["full_ids"][i, :n_full[i]]), i
    cmps, hops = idx.last_query_stats(len(q))
    assert np.array_equal(cmps, info["cmps"]) and np.array_equal(hops, info["n_hops"])
    assert np.array_equal(I, oI)
    assert np.array_equal(D, oD)
    return D, I, info


@pytest.mark.parametrize("metric,partition,R,n_chunks", [("mips", False, 16, 12), ("mips", True, 32, 33), ("l2", False, 24, 16),
                                                         ("cosine", True, 16, 7), ("l2", True, 64, 20)])
def test_stored_coordinates_match_oracle_bit_for_bit(lib, cuda_ok, tmp_path, metric, partition, R, n_chunks):
    emb = rows(4000, 32, 3)
    if metric == "mips":
        emb *= np.random.default_rng(8).uniform(0.6, 1.4, (len(emb), 1)).astype(np.float32)
    prefix, g, coords, pq, codes, max_norm = build_diskann_index(tmp_path, "t", emb, metric=metric, R=R, n_chunks=n_chunks,
                                                                 partition=partition)
    q = rows(48, 32, 4)
    o = VamanaOracle(g, pq, codes, metric, max_norm)
    # partition mode reads adjacency from the partition pair and never opens _disk.index (no coordinates)
    idx = capi.DiskannIndex(prefix, metric, prefix if partition else None)
    di = idx.dinfo
    assert (di.npts, di.dim, di.data_dim, di.n_chunks, di.max_degree) == (4000, 32, 33 if metric == "mips" else 32, n_chunks, R)
    assert di.partitioned == int(partition) and di.has_coords == int(not partition) and di.n_edges == int(g.degrees().sum())
    if partition:
        with pytest.raises(capi.Lb2Error, match="full-precision coordinates"):
            idx.search(q, 10, capi.make_diskann_params(64, 1, recompute_embeddings=False))
        for L, beam, k in [(64, 1, 10), (32, 4, 5), (100, 2, 10)]:
            check_same(idx, o, q, k, L, beam, skip=True)
    else:
        for L, beam, k in [(64, 1, 10), (32, 4, 5), (100, 2, 10), (8, 1, 10), (64, 16, 1)]:
            check_same(idx, o, q, k, L, beam, coords=coords)
        check_same(idx, o, q, 10, 64, 2, skip=True)
        check_same(idx, o, q, 10, 64, 2, coords=coords, io_limit=20)
    idx.close()


@pytest.mark.parametrize("metric", ["mips", "l2"])
def test_committed_goldens(lib, cuda_ok, golden_dir, metric):
    """Standard and partition layouts of the committed fixture against the committed oracle outputs."""
    from helpers import VAMANA_GOLDEN_CASES, load_vamana_golden
    G = load_vamana_golden(golden_dir, metric)
    std = capi.DiskannIndex(G["prefix"], metric)
    part = capi.DiskannIndex(G["prefix"], metric, G["prefix"])
    for L, beam, k in VAMANA_GOLDEN_CASES:
        for idx, mode, p in ((std, "stored", capi.make_diskann_params(L, beam, recompute_embeddings=False)),
                             (std, "pq", capi.make_diskann_params(L, beam, recompute_embeddings=False, skip_search_reorder=True)),
                             (part, "pq", capi.make_diskann_params(L, beam, recompute_embeddings=False, skip_search_reorder=True))):
            key = f"{metric}_L{L}_b{beam}_k{k}_{mode}"
            D, I = idx.search(G["q"], k, p)
            assert np.array_equal(I, G["exp"][key + "_I"]) and np.array_equal(D, G["exp"][key + "_D"]), key
            w = G["exp"][key + "_full"].shape[1]
            ids, n_full = idx.last_expansions(len(G["q"]), w)
            assert np.array_equal(n_full, G["exp"][key + "_nfull"])
            m = np.arange(w)[None, :] < n_full[:, None]
            assert np.array_equal(ids[m], G["exp"][key + "_full"][m]), key
            cmps, hops = idx.last_query_stats(len(G["q"]))
            assert np.array_equal(cmps, G["exp"][key + "_cmps"]) and np.array_equal(hops, G["exp"][key + "_hops"])
    std.close(); part.close()


def test_two_waves_and_edge_cases(lib, cuda_ok, tmp_path):
    emb = rows(3000, 24, 5)
    prefix, g, coords, pq, codes, max_norm = build_diskann_index(tmp_path, "w", emb, metric="l2", R=16, n_chunks=8, partition=False)
    o = VamanaOracle(g, pq, codes, "l2", max_norm)
    idx = capi.DiskannIndex(prefix, "l2")
    q = rows(5000, 24, 6)  # > one wave of 4096 queries
    D, I = idx.search(q, 10, capi.make_diskann_params(32, 2, recompute_embeddings=False))
    oD, oI, _ = o.search(q, 10, L=32, beam_width=2, coords=coords, nthreads=8)
    assert np.array_equal(I, oI) and np.array_equal(D, oD)
    # same handle, second call: the visited bitsets were left clean
    D2, I2 = idx.search(q[:100], 10, capi.make_diskann_params(32, 2, recompute_embeddings=False))
    assert np.array_equal(I2, oI[:100])
    D0, I0 = idx.search(q[:0], 10, capi.make_diskann_params(recompute_embeddings=False))
    assert D0.shape == (0, 10)
    with pytest.raises(capi.Lb2Error, match="not implemented"):
        idx.search(q[:2], 10, capi.make_diskann_params(recompute_neighbors=True))
    with pytest.raises(capi.Lb2Error, match="lb2_set_passages"):
        idx.search(q[:2], 10, capi.make_diskann_params(recompute_embeddings=True))
    with pytest.raises(capi.Lb2Error, match="use lb2_diskann_search"):
        capi.Index.search(idx, q[:2], 10, capi.make_params(recompute=False))
    idx.close()
    # fewer expanded nodes than k: (-1, FLT_MAX) tail (the reference reads past the end of full_retset there)
    tiny = rows(300, 24, 7)
    prefix, g, coords, pq, codes, max_norm = build_diskann_index(tmp_path, "tiny", tiny, metric="l2", R=4, n_chunks=8, partition=False)
    o = VamanaOracle(g, pq, codes, "l2", max_norm)
    t = capi.DiskannIndex(prefix, "l2")
    D, I = t.search(q[:8], 20, capi.make_diskann_params(4, 1, recompute_embeddings=False))
    oD, oI, info = o.search(q[:8], 20, L=4, beam_width=1, coords=coords)
    assert np.array_equal(I, oI) and np.array_equal(D, oD)
    assert (info["n_full"] < 20).any() and (I == -1).any() and (D[I == -1] == FMAX).all()
    with pytest.raises(capi.Lb2Error):
        capi.DiskannIndex(str(tmp_path / "missing"), "l2")


@pytest.fixture(scope="module")
def rc(lib, cuda_ok, tmp_path_factory):
    """A small recompute-mode DiskANN index over encoder embeddings (MIPS, partition layout like is_recompute builds)."""
    preset = synth.TINY
    w = synth.synthetic_weights(preset, 2)
    blob = synth.pack_weights(preset, w)
    tm, corpus = synth.make_corpus(2500, preset.vocab_size, seed=21, max_len=preset.max_pos)
    queries = synth.make_queries(tm, 40, seed=22)
    enc = open_encoder_only(preset, blob, corpus)
    E = enc.encode_ids(np.arange(corpus.n))
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    enc.close()
    d = tmp_path_factory.mktemp("darc")
    index_path, art = write_diskann_leann_index(d, "rc", E, preset=preset, corpus=corpus, metric="mips", R=24, n_chunks=16,
                                                partition=True, weight_seed=2)
    prefix, g, coords, pq, codes, max_norm = art
    return dict(preset=preset, w=w, blob=blob, corpus=corpus, E=E, Q=Q, index_path=index_path, prefix=prefix, g=g, coords=coords,
                pq=pq, codes=codes, max_norm=max_norm)


def test_deferred_fetch_recompute_is_exact_given_the_gpu_embeddings(rc):
    """recompute_embeddings=True: PQ traversal, then ONE encoder pass over the de-duplicated expanded nodes of the whole
    batch, then the re-rank.  With the oracle re-ranking from the GPU's own embeddings everything is bit-identical."""
    idx = capi.DiskannIndex(rc["prefix"], "mips", rc["prefix"])
    idx.set_passages(rc["corpus"].tokens, rc["corpus"].offsets)
    idx.set_encoder(rc["preset"].config(), rc["blob"])
    o = VamanaOracle(rc["g"], rc["pq"], rc["codes"], "mips", rc["max_norm"])
    for L, beam in [(64, 1), (48, 4)]:
        D, I, info = check_same(idx, o, rc["Q"], 10, L, beam, emb=rc["E"])
        st = idx.last_stats
        total = int(info["n_full"].sum())
        uniq = len(np.unique(np.concatenate([info["full_ids"][i, :n] for i, n in enumerate(info["n_full"])])))
        assert st.n_requested == total and st.n_recomputed == uniq and uniq < total  # cross-query de-duplication
        assert st.n_tokens > 0 and st.encoder_ms > 0
    idx.configure(0, 64)  # several encoder passes per call
    check_same(idx, o, rc["Q"], 10, 64, 2, emb=rc["E"])
    # recall against exact inner-product search over the same embeddings
    D, I = idx.search(rc["Q"], 10, capi.make_diskann_params(128, 2))
    gt = np.argsort(-(rc["Q"] @ rc["E"].T), axis=1)[:, :10]
    rec = np.mean([len(set(I[i]) & set(gt[i])) / 10 for i in range(len(gt))])
    assert rec > 0.8, rec
    # distances are the reference's rescaled output: -(L2^2 in the extended space) * max_norm * |q|  ~  2 ip - const
    ip = np.take_along_axis(rc["Q"] @ rc["E"].T, I, axis=1)
    assert (np.diff(D, axis=1) <= 0).all() and (np.diff(ip, axis=1) <= 1e-4).all()
    idx.close()


def test_deferred_fetch_against_fp32_bert_oracle(rc):
    """Tier T-B: the oracle re-ranks from the fp32 BertModel embeddings; ids agree wherever the distance gap exceeds the
    fp16-encoder noise, distances within 1e-3 (relative to the rescale factor)."""
    from oracle.encoder_oracle import EncoderOracle
    eo = EncoderOracle(rc["preset"], rc["w"])
    E32 = eo.encode_store(rc["corpus"].tokens, rc["corpus"].offsets)
    assert np.abs(E32 - rc["E"]).max() < 2e-3
    idx = capi.DiskannIndex(rc["prefix"], "mips", rc["prefix"])
    idx.set_passages(rc["corpus"].tokens, rc["corpus"].offsets)
    idx.set_encoder(rc["preset"].config(), rc["blob"])
    o = VamanaOracle(rc["g"], rc["pq"], rc["codes"], "mips", rc["max_norm"])
    D, I = idx.search(rc["Q"], 10, capi.make_diskann_params(64, 2))
    oD, oI, info = o.search(rc["Q"], 10, L=64, beam_width=2, emb=E32, nthreads=8)
    ids, n_full = idx.last_expansions(len(rc["Q"]), info["full_ids"].shape[1])
    assert np.array_equal(n_full, info["n_full"])  # the traversal never touches embeddings: identical expansions
    agree = np.mean([len(set(I[i]) & set(oI[i])) / 10 for i in range(len(I))])
    assert agree > 0.97, agree
    same = I == oI
    assert np.abs(D[same] - oD[same]).max() < 1e-2 * max(1.0, rc["max_norm"])
    idx.close()


def test_plugin_searcher_end_to_end(rc):
    """Through the LEANN plugin API: factory -> searcher -> _ensure_server_running -> search."""
    import leann_b200
    from leann_b200 import diskann_backend  # noqa: F401  (registers)
    fac = leann_b200.BACKEND_REGISTRY["diskann_b200"]
    s = fac.searcher(str(rc["index_path"]))
    with pytest.raises(ValueError, match="zmq_port must be provided"):
        s.search(rc["Q"][:2], 5, recompute_embeddings=True)
    with pytest.raises(NotImplementedError, match="proportional"):
        s.search(rc["Q"][:2], 5, pruning_strategy="proportional", skip_search_reorder=True)
    with pytest.raises(RuntimeError, match="full-precision coordinates"):
        s.search(rc["Q"][:2], 5, recompute_embeddings=False)  # partition layout: no stored vectors
    port = s._ensure_server_running(str(rc["index_path"]) + ".meta.json", 5557)
    out = s.search(rc["Q"][:6].astype(np.float64), 5, zmq_port=port, complexity=64, beam_width=2, recompute_embeddings=True,
                   prune_ratio=0.3, pruning_strategy="local", batch_recompute=True, dedup_node_dis=True)
    assert set(out) == {"labels", "distances"} and out["distances"].shape == (6, 5) and out["distances"].dtype == np.float32
    assert all(isinstance(l, str) for row in out["labels"] for l in row)
    o = VamanaOracle(rc["g"], rc["pq"], rc["codes"], "mips", rc["max_norm"])
    oD, oI, _ = o.search(rc["Q"][:6], 5, L=64, beam_width=2, emb=rc["E"])
    assert [[int(x) for x in row] for row in out["labels"]] == oI.tolist() and np.array_equal(out["distances"], oD)
    qe = s.compute_query_embedding(rc["corpus"].passage(33))
    assert np.array_equal(qe[0], rc["E"][33])
    assert s.search(qe, 1, zmq_port=port, recompute_embeddings=True)["labels"][0][0] == "33"
    assert s.last_stats["n_recomputed"] > 0
    s.embedding_server_manager.stop_server()  # what LeannSearcher.cleanup() calls
    assert s._index is None
    again = s.search(rc["Q"][:6], 5, zmq_port=port, complexity=64, beam_width=2, recompute_embeddings=True)
    assert again["labels"] == out["labels"] and np.array_equal(again["distances"], out["distances"])
    s.cleanup()


def test_bge_base_768d_diskann_recompute(lib, cuda_ok, tmp_path):
    """Config C4's shape at test size: bge-base architecture (12 layers, 768d, CLS pooling) behind the DiskANN path — 769 stored
    dimensions (odd), default PQ budget rule, partition layout."""
    from leann_b200 import diskann_format as dfmt
    preset = synth.BGE_BASE
    w = synth.synthetic_weights(preset, 3)
    blob = synth.pack_weights(preset, w)
    tm, corpus = synth.make_corpus(700, preset.vocab_size, seed=5, max_len=200)
    queries = synth.make_queries(tm, 16, seed=6)
    enc = open_encoder_only(preset, blob, corpus)
    E = enc.encode_ids(np.arange(corpus.n))
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    enc.close()
    n_chunks = min(96, dfmt.default_num_chunks(len(E), 769))  # the rule gives dim-many chunks at this tiny size; keep the test light
    prefix, g, coords, pq, codes, mx = build_diskann_index(tmp_path, "bge", E, metric="mips", R=16, n_chunks=n_chunks, partition=True,
                                                           keep_disk_index=False)
    idx = capi.DiskannIndex(prefix, "mips", prefix)
    assert (idx.dinfo.dim, idx.dinfo.data_dim) == (768, 769)
    idx.set_passages(corpus.tokens, corpus.offsets)
    idx.set_encoder(preset.config(), blob)
    idx.configure(0, 128)
    o = VamanaOracle(g, pq, codes, "mips", mx)
    check_same(idx, o, Q, 10, 48, 2, emb=E)
    idx.close()
