"""BASELINE configs[0] end to end on the GPU (kept in its own module, collected last)."""
import numpy as np
import pytest

from helpers import open_encoder_only, recall_at_k
from leann_b200 import capi, csr, synth

pytestmark = pytest.mark.gpu


def test_config_c1_pride_and_prejudice(lib, cuda_ok, golden_dir, tmp_path):
    """BASELINE configs[0] — the reference's own CPU-runnable case (1.25 k chunks of data/PrideandPrejudice.txt, MiniLM-L6
    architecture, HNSW M=32, ef=64): recompute search == oracle traversal over the GPU's embeddings, recall@10 vs exact search
    (fp32 calibration on the CPU oracle: 0.9975)."""
    from helpers import load_c1
    from leann_b200.graph_build import build_hnsw_graph
    from oracle.binding import Oracle
    corpus, queries = load_c1(golden_dir)
    preset = synth.MINILM_L6
    blob = synth.pack_weights(preset, synth.synthetic_weights(preset, 0))
    enc = open_encoder_only(preset, blob, corpus)
    E = enc.encode_ids(np.arange(corpus.n))
    Q = enc.encode_tokens(queries.tokens, queries.offsets)
    enc.close()
    g = build_hnsw_graph(E, M=32, metric="mips")
    f = tmp_path / "c1.index"
    csr.write_compact_index(str(f), g)
    idx = capi.Index(str(f))
    idx.set_passages(corpus.tokens, corpus.offsets)
    idx.set_encoder(preset.config(), blob)
    D, I = idx.search(Q, 10, capi.make_params(64, 1, recompute=True))
    oD, oI, ond, onh = Oracle(g, E).search(Q, 10, ef=64, nthreads=8)
    assert np.array_equal(I, oI) and np.array_equal(D, oD)
    gt = np.argsort(-(Q @ E.T), axis=1)[:, :10]
    assert recall_at_k(I, gt) >= 0.95
    idx.close()
