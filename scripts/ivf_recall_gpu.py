"""GPU experiment: does the partition-restricted candidate search (graph_build._knn_ivf, used above 2.5 M points)
cost recall?  Same 1 M corpus, exact-kNN candidates vs IVF candidates at several probe counts; stored-vector
search (the same traversal, no recompute).  usage: ivf_recall_gpu.py [chunks]"""
import sys, time, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import capi, csr, synth, build, graph_build
from leann_b200.tooling import stub_graph, recall_at_k
if build.needs_build(): build.build()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
preset = synth.MINILM_L6
w = synth.synthetic_weights(preset, 0); blob = synth.pack_weights(preset, w)
tm, corpus = synth.make_corpus(N, preset.vocab_size, seed=1234, max_len=preset.max_pos)
queries = synth.make_queries(tm, 2000, seed=4321)
work = Path(tempfile.mkdtemp())
csr.write_compact_index(str(work / "stub.index"), stub_graph(N, 384))
enc = capi.Index(str(work / "stub.index"), 0)
enc.set_passages(corpus.tokens, corpus.offsets); enc.set_encoder(preset.config(), blob)
E = torch.empty((N, 384), dtype=torch.float32, device="cuda")
enc.encode_range_device(0, N, E.data_ptr())
Q = enc.encode_tokens(queries.tokens, queries.offsets); enc.close()
gt = torch.topk(torch.from_numpy(Q).cuda() @ E.T, 10, dim=1).indices.cpu().numpy()
# candidate quality itself: recall of the IVF 48-NN lists against exact 48-NN on a sample
samp = torch.randperm(N, device="cuda")[:4096]
ex = torch.topk(E[samp] @ E.T, 49, dim=1).indices[:, 1:]
orig = graph_build._knn_ivf
for probes, rounds in ((12, 0), (12, 1), (12, 2), (12, 3), (6, 3)):
    if probes:
        graph_build._knn_ivf = (lambda x, k, ip, _p=probes: orig(x, k, ip, n_probe=_p))
        t = time.time()
        ci, cd = graph_build._knn_ivf(E, 48, True)
        torch.cuda.synchronize(); t1 = time.time()
        ci, cd = graph_build._refine_knn(E, ci, cd, True, rounds=rounds)
        torch.cuda.synchronize(); t2 = time.time()
        hit = (ci[samp][:, :, None] == ex[:, None, :]).any(2).float().mean().item()
        print(f"n_probe={probes} refine rounds={rounds}: candidate-list recall vs exact 48-NN = {hit:.4f}  (ivf {t1-t:.1f}s, refine {t2-t1:.1f}s)", flush=True)
        del ci, cd
    t = time.time()
    g = graph_build.build_hnsw_graph(E, M=32, metric="mips", ivf_threshold=(0 if probes else 1 << 40), ivf_refine_rounds=rounds)
    bt = time.time() - t
    f = work / "g.index"; csr.write_compact_index(str(f), g)
    idx = capi.Index(str(f), 0); idx.set_vectors_device(E.data_ptr())
    for ef in (64, 128):
        D, I = idx.search(Q, 10, capi.make_params(ef, recompute=False))
        nd, nh = idx.last_query_stats(len(Q))
        print(f"{'exact' if not probes else 'ivf n_probe=%d refine=%d' % (probes, rounds)} build {bt:.0f}s ef={ef}: recall {recall_at_k(I, gt):.4f} ndis {nd.mean():.0f} nhops {nh.mean():.0f}", flush=True)
    idx.close()
