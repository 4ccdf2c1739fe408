"""Id-level parity at benchmark scale (SURVEY 7 step 3, 8c tier T-A): the CUDA stored-vector search against the compiled
reference traversal (oracle/_ref) on a 1 M-passage world built exactly like bench.py's — GPU corpus sampler, GPU encoder,
GPU insertion-as-search graph — ids, distances, ndis and nhops bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_million_passages_stored_vector_search_is_bit_exact(lib, cuda_ok, tmp_path):
    import torch
    from leann_b200 import capi, csr, synth
    from leann_b200.graph_build import build_hnsw_graph_incremental
    from leann_b200.tooling import stub_graph
    from oracle.binding import Oracle, Reference, have_reference

    n, nq, k = 1_000_000, 64, 10
    preset = synth.MINILM_L6
    blob = synth.pack_weights(preset, synth.synthetic_weights(preset, 0))
    tm, corpus, queries = synth.make_bench_corpus(n, preset.vocab_size, preset.max_pos, "cuda:0", nq)
    stub = tmp_path / "stub.index"
    csr.write_compact_index(str(stub), stub_graph(n, preset.hidden))
    enc = capi.Index(str(stub), 0)
    enc.set_passages(corpus.tokens, corpus.offsets)
    enc.set_encoder(preset.config(), blob)
    E = torch.empty((n, preset.hidden), dtype=torch.float32, device="cuda:0")
    enc.encode_range_device(0, n, E.data_ptr())
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    enc.close()
    g = build_hnsw_graph_incremental(E, M=32, metric="mips", device="cuda:0", ef_construction=100)
    f = tmp_path / "m.index"
    csr.write_compact_index(str(f), g)
    idx = capi.Index(str(f), 0)
    idx.set_vectors_device(E.data_ptr())
    Eh = E.cpu().numpy()
    del E
    for ef, beam in ((64, 1), (32, 4)):
        D, I = idx.search(Q, k, capi.make_params(ef, beam, 0, True, recompute=False))
        ndis, nhops = idx.last_query_stats(nq)
        trav = Reference.from_csr(g, Eh, M=32) if have_reference() else Oracle(g, Eh)
        rD, rI, rnd, rnh = trav.search(Q, k, ef=ef, beam=beam, nthreads=8)
        assert np.array_equal(I, rI) and np.array_equal(D, rD), (ef, beam)
        assert np.array_equal(ndis, rnd) and np.array_equal(nhops, rnh), (ef, beam)
        gt = torch.topk(torch.from_numpy(Q).cuda() @ torch.from_numpy(Eh).cuda().T, k, dim=1).indices.cpu().numpy()
        rec = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(I, gt)]))
        print(f"1 M passages, ef={ef} beam={beam}: ids/D/ndis/nhops identical to the {'compiled reference' if have_reference() else 'C restatement'}; "
              f"recall@10 {rec:.3f}, mean ndis {ndis.mean():.0f}")
    idx.close()
