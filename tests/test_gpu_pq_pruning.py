"""PQ-guided pruning on the GPU: lb2_load_pq_pruning + the three selection strategies of
search_from_candidates (faiss/impl/HNSW_search.cpp:442-465, 645-750) against the CPU restatement, which
tests/test_oracle_pq_pruning.py pins to the compiled reference — ids, distances, ndis, nhops bit for bit."""
import numpy as np
import pytest

from test_oracle_pq_pruning import PRUNE_CASES, make_world

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["mips", "l2"])
def test_gpu_pq_pruning_matches_the_oracle(lib, cuda_ok, tmp_path, metric):
    from leann_b200 import capi, csr
    from oracle.binding import Oracle

    W = make_world(tmp_path, metric)
    f = tmp_path / "g.index"
    csr.write_compact_index(str(f), W["g"])
    idx = capi.Index(str(f), 0)
    idx.set_vectors(W["x"])
    o = Oracle(W["g"], W["x"])
    # without PQ data the pruning knobs are ignored, exactly like the reference (HNSW_search.cpp:442-445)
    a = idx.search(W["q"], 10, capi.make_params(32, prune_ratio=0.5, local_prune=True, recompute=False))
    b = o.search(W["q"], 10, ef=32)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    idx.load_pq_pruning(W["prefix"] + "_pq_pivots.bin", W["prefix"] + "_pq_compressed.bin")
    o.set_pq(W["pq"], W["codes"])
    for pr, lp, sr, ef, beam, bs, cr, k in PRUNE_CASES:
        D, I = idx.search(W["q"], k, capi.make_params(ef, beam, bs, cr, pr, lp, sr, recompute=False))
        ndis, nhops = idx.last_query_stats(len(W["q"]))
        oD, oI, ond, onh = o.search(W["q"], k, ef=ef, beam=beam, batch_size=bs, check_rel=cr, prune_ratio=pr, local_prune=lp, send_ratio=sr)
        tag = (metric, pr, lp, sr, ef, beam, bs, cr, k)
        assert np.array_equal(I, oI), tag
        assert np.array_equal(D, oD), tag
        assert np.array_equal(ndis, ond) and np.array_equal(nhops, onh), tag
    idx.close()


def test_gpu_pq_pruning_in_recompute_mode_and_through_the_plugin(lib, cuda_ok, tmp_path):
    """prune_ratio through B200HnswSearcher.search with PQ sidecars next to the index: fewer passages are recomputed, results
    equal the oracle's traversal over the GPU's own embeddings."""
    from helpers import write_leann_index
    from leann_b200 import backend, capi, synth
    from leann_b200.graph_build import build_hnsw_graph
    from leann_b200.tooling import open_encoder_only
    from leann_b200.vamana_build import build_diskann_index
    from oracle.binding import Oracle

    preset = synth.TINY
    blob = synth.pack_weights(preset, synth.synthetic_weights(preset, 0))
    tm, corpus = synth.make_corpus(3000, preset.vocab_size, seed=5, max_len=preset.max_pos)
    queries = synth.make_queries(tm, 12, seed=6)
    enc = open_encoder_only(preset, blob, corpus, 0)
    E = enc.encode_ids(np.arange(corpus.n))
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    enc.close()
    g = build_hnsw_graph(E, M=8, metric="mips", device="cpu")
    index_path = write_leann_index(tmp_path, "pq", g, preset, corpus)
    prefix, _, _, pq, codes, _ = build_diskann_index(tmp_path, "pqfiles", E, metric="mips", R=8, n_chunks=24, device="cpu")
    for suffix in ("pq_pivots.bin", "pq_compressed.bin"):
        (tmp_path / f"pq.leann.{suffix}").write_bytes(open(f"{prefix}_{suffix}", "rb").read())
    s = backend.B200HnswBackend.searcher(str(index_path), device=0)
    port = s._ensure_server_running(str(index_path) + ".meta.json", 5557)
    o = Oracle(g, E)
    o.set_pq(pq, codes)
    full = s.search(Q, 10, zmq_port=port, complexity=32, recompute_embeddings=True)
    n_full = s.last_stats["n_recomputed"]
    for strategy, kw in (("global", {}), ("local", dict(local_prune=True)), ("proportional", dict(send_ratio=1.0))):
        out = s.search(Q, 10, zmq_port=port, complexity=32, prune_ratio=0.6, pruning_strategy=strategy, recompute_embeddings=True)
        oD, oI, ond, _ = o.search(Q, 10, ef=32, prune_ratio=0.6, **kw)
        assert [[int(x) for x in row] for row in out["labels"]] == oI.tolist(), strategy
        assert np.array_equal(out["distances"], oD), strategy
        assert s.last_stats["ndis"] == int(ond.sum())
        assert s.last_stats["n_recomputed"] < n_full
    assert len(full["labels"]) == len(Q)
    s.cleanup()
