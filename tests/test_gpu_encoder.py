"""Recompute stage end to end (token ids -> pooled, normalised embedding) against the fp32
transformers.BertModel oracle on the same seeded weights."""
import numpy as np
import pytest

from helpers import open_encoder_only
from leann_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,n", [(synth.TINY, 300), (synth.MINILM_L6, 48), (synth.BGE_BASE, 12)])
def test_encoder_matches_fp32_bert(lib, cuda_ok, preset, n):
    from oracle.encoder_oracle import EncoderOracle
    w = synth.synthetic_weights(preset, 0)
    _, corpus = synth.make_corpus(n, preset.vocab_size, seed=11, max_len=min(256, preset.max_pos))
    idx = open_encoder_only(preset, synth.pack_weights(preset, w), corpus)
    E = idx.encode_ids(np.arange(n))
    ref = EncoderOracle(preset, w).encode_store(corpus.tokens, corpus.offsets)
    assert np.isfinite(E).all()
    assert np.allclose(np.linalg.norm(E, axis=1), 1, atol=1e-4)
    # what the traversal consumes: distances -E.q  (hnsw_embedding_server.py:195-200), tolerance of north_star
    q = ref[:8]
    dE, dD = float(np.abs(E - ref).max()), float(np.abs(E @ q.T - ref @ q.T).max())
    print(f"{preset.name}: max |dE| = {dE:.2e} (bound 2e-3), max |d distance| = {dD:.2e} (bound 1e-3)")
    assert dE < 2e-3, dE
    assert dD < 1e-3, dD


def test_encoding_is_batch_and_position_independent(lib, cuda_ok):
    preset = synth.TINY
    w = synth.synthetic_weights(preset, 1)
    _, corpus = synth.make_corpus(500, preset.vocab_size, seed=3, max_len=preset.max_pos)
    idx = open_encoder_only(preset, synth.pack_weights(preset, w), corpus)
    all_ = idx.encode_ids(np.arange(500))
    perm = np.random.default_rng(0).permutation(500)[:77]
    some = idx.encode_ids(perm)
    assert np.array_equal(some, all_[perm])  # bit-identical: a passage's embedding never depends on its batch
    one = idx.encode_ids(np.array([perm[5]]))
    assert np.array_equal(one[0], all_[perm[5]])
    idx.configure(0, 64)  # smaller encoder passes
    assert np.array_equal(idx.encode_ids(np.arange(500)), all_)


def test_encode_tokens_equals_encode_ids_and_truncates(lib, cuda_ok):
    preset = synth.TINY
    w = synth.synthetic_weights(preset, 2)
    _, corpus = synth.make_corpus(40, preset.vocab_size, seed=4, max_len=preset.max_pos)
    idx = open_encoder_only(preset, synth.pack_weights(preset, w), corpus)
    a = idx.encode_ids(np.arange(40))
    b = idx.encode_tokens(corpus.tokens, corpus.offsets)
    assert np.array_equal(a, b)
    long_seq = np.concatenate([corpus.passage(0), corpus.passage(1), corpus.passage(2)])[: 3 * preset.max_pos]
    t = idx.encode_tokens(long_seq.astype(np.uint16), np.array([0, long_seq.size], np.uint64))
    t2 = idx.encode_tokens(long_seq[: preset.max_pos].astype(np.uint16), np.array([0, preset.max_pos], np.uint64))
    assert np.array_equal(t, t2)  # truncation=True at max_seq_length


def test_embedding_server_shim_serves_gpu_embeddings(lib, cuda_ok):
    """The reference's msgpack/ZMQ protocol answered by the GPU encoder (what an unmodified LEANN client would receive)."""
    import threading
    zmq = pytest.importorskip("zmq")
    msgpack = pytest.importorskip("msgpack")
    from leann_b200.embedding_server import serve
    preset = synth.TINY
    blob = synth.pack_weights(preset, synth.synthetic_weights(preset, 4))
    tm, corpus = synth.make_corpus(300, preset.vocab_size, seed=31, max_len=preset.max_pos)
    idx = open_encoder_only(preset, blob, corpus)
    E = idx.encode_ids(np.arange(corpus.n))
    stop, ready = threading.Event(), threading.Event()
    port = 6200 + int(np.random.default_rng().integers(0, 300))
    th = threading.Thread(target=serve, args=(idx, port), kwargs=dict(distance_metric="mips", model_name=preset.name, shutdown=stop, ready=ready), daemon=True)
    th.start()
    assert ready.wait(10)
    ctx = zmq.Context()
    s = ctx.socket(zmq.REQ)
    s.setsockopt(zmq.RCVTIMEO, 20000)
    s.setsockopt(zmq.LINGER, 0)
    s.connect(f"tcp://127.0.0.1:{port}")
    try:
        ids = [5, 17, 299, 100000]
        s.send(msgpack.packb([ids, E[3].tolist()]))
        (dist,) = msgpack.unpackb(s.recv())
        assert dist[3] == 1e9 and np.allclose(dist[:3], -(E[[5, 17, 299]] @ E[3]), atol=1e-6)
        s.send(msgpack.packb([[7, 8]]))
        dims, flat = msgpack.unpackb(s.recv())
        assert dims == [2, preset.hidden] and np.array_equal(np.asarray(flat, np.float32).reshape(2, -1), E[[7, 8]])
    finally:
        stop.set()
        s.close()
        ctx.term()
        th.join(5)
        idx.close()


def _realistic_bert(preset, seed=0):
    """A transformers.BertModel with the statistics trained checkpoints have and N(0, 0.02) initialisation lacks (no
    checkpoint can be downloaded here): a few "massive activation" channels in the embeddings (x30), LayerNorm gains
    spread over 0.3..5 with large values on the outlier channels, non-zero LayerNorm biases, heavier-tailed linear weights."""
    import torch
    from transformers import BertConfig, BertModel

    cfg = BertConfig(vocab_size=preset.vocab_size, hidden_size=preset.hidden, num_hidden_layers=preset.layers,
                     num_attention_heads=preset.heads, intermediate_size=preset.ffn, hidden_act="gelu",
                     max_position_embeddings=preset.max_pos, type_vocab_size=preset.type_vocab, layer_norm_eps=preset.ln_eps,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    try:
        cfg._attn_implementation = "eager"
    except Exception:
        pass
    torch.manual_seed(seed)
    m = BertModel(cfg, add_pooling_layer=False).eval()
    g = torch.Generator().manual_seed(seed + 1)
    outliers = torch.randperm(preset.hidden, generator=g)[:4]
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "LayerNorm.weight" in name:
                p.copy_(torch.exp(torch.randn(p.shape, generator=g) * 0.45).clamp(0.3, 5.0))
                p[outliers] = 5.0
            elif "LayerNorm.bias" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif "word_embeddings" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
                p[:, outliers] *= 30.0
            elif name.endswith("dense.weight") or "self." in name and name.endswith("weight"):
                w = torch.randn(p.shape, generator=g) * 0.035
                w *= torch.exp(torch.randn(p.shape[0], 1, generator=g) * 0.3)  # per-output-row scale spread
                p.copy_(w)
            elif name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    return m


def test_hf_checkpoint_path_with_realistic_weight_statistics(lib, cuda_ok):
    """backend.weights_from_hf (the loader a real sentence-transformers checkpoint goes through) -> GPU encoder, against
    the SAME transformers model evaluated in fp32 (masked mean pool + L2 normalisation, embedding_compute.py:319-335)."""
    import torch
    from leann_b200.backend import weights_from_hf

    preset0 = synth.MINILM_L6
    hf = _realistic_bert(preset0)
    preset, blob = weights_from_hf(hf, preset0.name)
    assert (preset.hidden, preset.layers, preset.heads, preset.ffn, preset.max_pos, preset.pooling) == (384, 6, 12, 1536, 256, 0)
    n = 64
    _, corpus = synth.make_corpus(n, preset.vocab_size, seed=21, max_len=preset.max_pos)
    idx = open_encoder_only(preset, blob, corpus)
    E = idx.encode_ids(np.arange(n))
    ref = []
    with torch.inference_mode():
        for i in range(n):
            ids = torch.from_numpy(corpus.passage(i)[: preset.max_pos].astype(np.int64))[None]
            h = hf(input_ids=ids, attention_mask=torch.ones_like(ids)).last_hidden_state[0]
            ref.append(torch.nn.functional.normalize(h.mean(0), dim=0).numpy())
    ref = np.stack(ref)
    dE = float(np.abs(E - ref).max())
    dD = float(np.abs(E @ ref[:16].T - ref @ ref[:16].T).max())
    print(f"realistic-statistics weights: max |dE| = {dE:.2e}, max |d distance| = {dD:.2e}, cos >= {float((E * ref).sum(1).min()):.6f}")
    assert np.isfinite(E).all()
    assert dD < 1e-3, (dE, dD)  # north_star: distances within 1e-3
    assert dE < 2e-3, dE
