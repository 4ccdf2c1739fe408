"""GPU experiment: recall of the torch graph builder at bench scale (stored-vector search mode = the
same traversal, no recompute).  usage: graph_recall_gpu.py [chunks]"""
import sys, time, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import capi, csr, synth, build
from leann_b200.graph_build import build_hnsw_graph
from leann_b200.tooling import stub_graph, recall_at_k
if build.needs_build(): build.build()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
preset = synth.MINILM_L6
w = synth.synthetic_weights(preset, 0); blob = synth.pack_weights(preset, w)
for (pt, ps, qlen) in [(0.70, 0.15, 24), (0.80, 0.10, 32)]:
  print(f"=== corpus p_topic={pt} p_super={ps} query len {qlen}", flush=True)
  tm, corpus = synth.make_corpus(N, preset.vocab_size, seed=1234, max_len=preset.max_pos, p_topic=pt, p_super=ps)
  queries = synth.make_queries(tm, 2000, seed=4321, len_mean=qlen, p_topic=pt, p_super=ps)
  work = Path(tempfile.mkdtemp())
  csr.write_compact_index(str(work / "stub.index"), stub_graph(N, 384))
  enc = capi.Index(str(work / "stub.index"), 0)
  enc.set_passages(corpus.tokens, corpus.offsets); enc.set_encoder(preset.config(), blob)
  E = torch.empty((N, 384), dtype=torch.float32, device="cuda")
  enc.encode_range_device(0, N, E.data_ptr())
  Q = enc.encode_tokens(queries.tokens, queries.offsets); enc.close()
  Qt = torch.from_numpy(Q).cuda()
  S = Qt @ E.T
  gt = torch.topk(S, 10, dim=1).indices.cpu().numpy()
  top = torch.topk(S, 100, dim=1).values
  print(f"cos: mean {S.mean().item():.3f} top1 {top[:,0].mean().item():.3f} top10 {top[:,9].mean().item():.3f} top100 {top[:,99].mean().item():.3f}", flush=True)
  same = np.mean([np.mean(corpus.topics[gt[i]] == queries.topics[i]) for i in range(len(gt))])
  print("fraction of exact top-10 inside the query's topic:", same, flush=True)
  Eh = E.cpu().numpy()
  for kw in [dict(), dict(alpha=1.2), dict(alpha=1.1, knn_factor=2.0)]:
      M = kw.pop("M", 32)
      t = time.time(); g = build_hnsw_graph(E, M=M, metric="mips", **kw); bt = time.time() - t
      f = work / "g.index"; csr.write_compact_index(str(f), g)
      idx = capi.Index(str(f), 0); idx.set_vectors(Eh)
      for ef in (64, 128):
          D, I = idx.search(Q, 10, capi.make_params(ef, recompute=False))
          nd, nh = idx.last_query_stats(len(Q))
          print(f"M={M} {kw} build {bt:.0f}s deg0 {g.neighbors.size/N:.1f} ef={ef}: recall {recall_at_k(I, gt):.4f} ndis {nd.mean():.0f} nhops {nh.mean():.0f}  ({idx.last_stats.gpu_ms:.1f} ms for {len(Q)} q)", flush=True)
      idx.close()
