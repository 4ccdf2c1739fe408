"""GPU experiment: recall@10 / search cost of the insertion-as-search builder's options at bench scale, measured with the
product's stored-vector search (same traversal as the recompute path, no encoder in the loop).
usage: graph_recall_10m.py [chunks] [variant ...]   variant = name:key=val,key=val"""
import sys, time, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import capi, csr, synth, build
from leann_b200.graph_build import build_hnsw_graph_incremental
from leann_b200.tooling import stub_graph, recall_at_k
import bench
if build.needs_build(): build.build()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
variants = sys.argv[2:] or ["base:", "sweep1:sweeps=1", "fill16:fill=16", "fill24:fill=24", "g10:growth=0.1", "efc400:ef_construction=400"]
preset = synth.MINILM_L6
blob = synth.pack_weights(preset, synth.synthetic_weights(preset, 0))
t0 = time.time()
import os
TS = int(os.environ.get("LB2_TOPIC_SIZE", 32)); PT = float(os.environ.get("LB2_P_TOPIC", 0.80)); PS = float(os.environ.get("LB2_P_SUPER", 0.10))
print(f"corpus: {N} passages, topic size {TS}, mix topic/super/background = {PT}/{PS}/{1-PT-PS:.2f}", flush=True)
tm, corpus = synth.make_corpus(N, preset.vocab_size, seed=1234, max_len=preset.max_pos, device="cuda:0", n_topics=max(4, N // TS), p_topic=PT, p_super=PS)
queries = synth.make_queries(tm, 4096, seed=4321, p_topic=PT, p_super=PS)
work = Path(tempfile.mkdtemp())
csr.write_compact_index(str(work / "stub.index"), stub_graph(N, 384))
enc = capi.Index(str(work / "stub.index"), 0)
enc.set_passages(corpus.tokens, corpus.offsets); enc.set_encoder(preset.config(), blob)
E = torch.empty((N, 384), dtype=torch.float32, device="cuda")
enc.encode_range_device(0, N, E.data_ptr())
Q = enc.encode_tokens(queries.tokens, queries.offsets); enc.close()
gt = bench.exact_ground_truth(Q, E)
same = np.mean([np.mean(corpus.topics[gt[i]] == queries.topics[i]) for i in range(len(gt))])
print(f"world: {N} passages in {time.time()-t0:.0f}s; fraction of exact top-10 inside the query's topic: {same:.3f}", flush=True)
for v in variants:
    name, _, kvs = v.partition(":")
    kw = {}
    for kv in filter(None, kvs.split(",")):
        k, _, val = kv.partition("=")
        kw[k] = float(val) if "." in val else int(val)
    t = time.time()
    g = build_hnsw_graph_incremental(E, M=32, metric="mips", device="cuda:0", verbose=bool(os.environ.get("LB2_BENCH_VERBOSE")), **kw)
    torch.cuda.synchronize(); bt = time.time() - t
    f = work / "g.index"; csr.write_compact_index(str(f), g)
    idx = capi.Index(str(f), 0); idx.set_vectors_device(E.data_ptr())
    for ef in (64, 96, 128):
        D, I = idx.search(Q, 10, capi.make_params(ef, recompute=False))
        nd, nh = idx.last_query_stats(len(Q))
        ov = np.array([len(set(a.tolist()) & set(b.tolist())) for a, b in zip(I, gt)])
        in_topic = np.mean([np.mean(corpus.topics[np.maximum(I[i], 0)] == queries.topics[i]) for i in range(len(I))])
        print(f"{name:10s} {kw} build {bt:.0f}s deg0 {g.neighbors.size/N:.1f} ef={ef}: recall {ov.mean()/10:.4f} (zero-overlap {np.mean(ov==0):.3f}, full {np.mean(ov==10):.3f}) "
              f"results-in-topic {in_topic:.3f} ndis {nd.mean():.0f} nhops {nh.mean():.0f}", flush=True)
    idx.close(); del g
