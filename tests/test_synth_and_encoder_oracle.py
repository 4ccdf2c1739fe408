import numpy as np

from leann_b200 import synth
from oracle.encoder_oracle import EncoderOracle


def test_corpus_is_deterministic_and_well_formed():
    tm, c = synth.make_corpus(300, 2048, seed=5, max_len=64)
    tm2, c2 = synth.make_corpus(300, 2048, seed=5, max_len=64)
    assert np.array_equal(c.tokens, c2.tokens) and np.array_equal(c.offsets, c2.offsets)
    lens = np.diff(c.offsets.astype(np.int64))
    assert lens.min() >= 16 and lens.max() <= 64
    starts = c.offsets[:-1].astype(np.int64)
    assert np.all(c.tokens[starts] == synth.CLS_ID) and np.all(c.tokens[starts + lens - 1] == synth.SEP_ID)
    assert c.tokens.max() < 2048
    q = synth.make_queries(tm, 20)
    assert np.diff(q.offsets.astype(np.int64)).max() <= 64


def test_weight_blob_layout():
    p = synth.TINY
    w = synth.synthetic_weights(p, 3)
    blob = synth.pack_weights(p, w)
    assert blob.dtype == np.float32 and blob.size == sum(int(np.prod(s)) for _, s in synth.weight_layout(p))
    # matrices survive an fp16 round trip unchanged: oracle and device hold the same values
    assert np.array_equal(w["l0.w_1"], w["l0.w_1"].astype(np.float16).astype(np.float32))


def test_encoder_oracle_padding_invariance_and_norm():
    p = synth.TINY
    w = synth.synthetic_weights(p, 0)
    _, c = synth.make_corpus(12, p.vocab_size, seed=2, max_len=p.max_pos)
    o = EncoderOracle(p, w)
    a = o.encode_store(c.tokens, c.offsets, batch_size=12)  # padded to the longest of 12
    b = o.encode_store(c.tokens, c.offsets, batch_size=1)   # no padding at all
    assert np.abs(a - b).max() < 2e-6
    assert np.allclose(np.linalg.norm(a, axis=1), 1, atol=1e-5)


def test_flops_formula_matches_survey():
    assert abs(synth.MINILM_L6.flops_per_chunk(128) / 1e9 - 2.869) < 0.005
    assert abs(synth.MINILM_L6.flops_per_chunk(256) / 1e9 - 6.039) < 0.005
    assert abs(synth.BGE_BASE.flops_per_chunk(128) / 1e9 - 22.35) < 0.02


def test_config_c1_fixture(golden_dir):
    """BASELINE configs[0]: the reference's sample document as a committed token stream (tests/golden/make_c1_fixture.py).
    With the fp32 BertModel oracle (MiniLM-L6 architecture, synthetic weights) + the torch graph builder + the C HNSW oracle this
    corpus gives recall@10 = 0.9975 at M=32, ef=64 (calibration run, 36 s of CPU encoding — too slow for this suite; the GPU suite
    runs the same case end to end in tests/test_gpu_z_config_c1.py)."""
    from helpers import load_c1
    z = np.load(golden_dir / "c1_pride_tokens.npz")
    assert int(z["source_bytes"]) == 772389 and z["tokens"].dtype == np.uint16 and len(z["tokens"]) == 159175
    assert 1000 <= z["tokens"].min() and z["tokens"].max() < 30000
    corpus, queries = load_c1(golden_dir)
    lens = np.diff(corpus.offsets.astype(np.int64))
    assert corpus.n == 1253 and lens.max() == 256 and lens.min() >= 100 and queries.n == 40
    assert (corpus.tokens[corpus.offsets[:-1].astype(np.int64)] == 101).all()
    assert (corpus.tokens[corpus.offsets[1:].astype(np.int64) - 1] == 102).all()
    # consecutive chunks overlap by half (the reference's 256/128 split)
    assert np.array_equal(corpus.passage(0)[128:255], corpus.passage(1)[1:128])
