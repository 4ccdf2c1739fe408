"""Synthetic inputs for tests and the benchmark (there is no network for real weights / corpora).

* encoder presets with the architecture constants of the two models BASELINE.json names
  (all-MiniLM-L6-v2, bge-base-en-v1.5; public model cards);
* seeded synthetic weights, packed in the blob order lb2_set_encoder() documents;
* topic-structured token corpora (after SURVEY.md §8d, with a second level: N/32 topics grouped
  into super-topics that share a sub-vocabulary) so that mean-pooled embeddings have a
  non-degenerate neighbourhood structure: a flat K = sqrt(N) topic model gives ~1000 nearly
  equidistant passages per topic at 1 M and no ANN index reaches recall 0.9 on it at ef = 64.

Everything here is deterministic in (seed, sizes) and uses numpy only.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .capi import EncoderConfig

CLS_ID, SEP_ID = 101, 102
FIRST_WORD_ID = 1000


@dataclass(frozen=True)
class ModelPreset:
    name: str
    vocab_size: int
    hidden: int
    layers: int
    heads: int
    ffn: int
    max_pos: int  # max_seq_length the sentence-transformers wrapper truncates to
    type_vocab: int
    ln_eps: float
    pooling: int  # 0 mean, 1 cls
    normalize: int

    def config(self) -> EncoderConfig:
        return EncoderConfig(self.vocab_size, self.hidden, self.layers, self.heads, self.ffn, self.max_pos,
                             self.type_vocab, self.ln_eps, self.pooling, self.normalize)

    def flops_per_chunk(self, L: int) -> float:
        """F(L) of SURVEY.md §8d: layers * L * (2*(4h^2 + 2hf) + 4*L*h)."""
        h, f = self.hidden, self.ffn
        return self.layers * L * (2.0 * (4 * h * h + 2 * h * f) + 4.0 * L * h)


MINILM_L6 = ModelPreset("sentence-transformers/all-MiniLM-L6-v2", 30522, 384, 6, 12, 1536, 256, 2, 1e-12, 0, 1)
BGE_BASE = ModelPreset("BAAI/bge-base-en-v1.5", 30522, 768, 12, 12, 3072, 512, 2, 1e-12, 1, 1)
TINY = ModelPreset("synthetic/tiny-bert", 2048, 384, 2, 12, 768, 64, 2, 1e-12, 0, 1)  # fast CPU-oracle tests
PRESETS = {p.name: p for p in (MINILM_L6, BGE_BASE, TINY)}


def weight_layout(p: ModelPreset) -> list[tuple[str, tuple[int, ...]]]:
    """(name, shape) in blob order — the order documented in include/leann_b200.h."""
    H, F = p.hidden, p.ffn
    out = [("word_emb", (p.vocab_size, H)), ("pos_emb", (p.max_pos, H)), ("type_emb", (p.type_vocab, H)),
           ("emb_ln_g", (H,)), ("emb_ln_b", (H,))]
    for l in range(p.layers):
        out += [(f"l{l}.w_qkv", (3 * H, H)), (f"l{l}.b_qkv", (3 * H,)), (f"l{l}.w_o", (H, H)), (f"l{l}.b_o", (H,)),
                (f"l{l}.ln1_g", (H,)), (f"l{l}.ln1_b", (H,)), (f"l{l}.w_1", (F, H)), (f"l{l}.b_1", (F,)),
                (f"l{l}.w_2", (H, F)), (f"l{l}.b_2", (H,)), (f"l{l}.ln2_g", (H,)), (f"l{l}.ln2_b", (H,))]
    return out


def synthetic_weights(p: ModelPreset, seed: int = 0) -> dict[str, np.ndarray]:
    """Seeded fp32 tensors.  Matrices are rounded through fp16 so the fp32 oracle and the fp16
    device copy hold the same values (the reference loads fp16 weights, embedding_compute.py:157-158)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in weight_layout(p):
        base = name.split(".")[-1]
        if base == "word_emb":
            a = rng.normal(0, 0.10, shape)
        elif base in ("pos_emb", "type_emb"):
            a = rng.normal(0, 0.02, shape)
        elif base.endswith("_g"):
            a = 1.0 + rng.normal(0, 0.05, shape)
        elif base.startswith("b_") or base.endswith("_b"):
            a = rng.normal(0, 0.02, shape)
        else:  # linear weights
            a = rng.normal(0, 0.04, shape)
        a = a.astype(np.float32)
        if len(shape) == 2:
            a = a.astype(np.float16).astype(np.float32)
        w[name] = a
    return w


def pack_weights(p: ModelPreset, w: dict[str, np.ndarray]) -> np.ndarray:
    return np.concatenate([np.ascontiguousarray(w[name], np.float32).reshape(-1) for name, _ in weight_layout(p)])


# ------------------------------------------------------------------------------------------------
@dataclass
class Corpus:
    tokens: np.ndarray   # uint16 [total]
    offsets: np.ndarray  # uint64 [n + 1]
    topics: np.ndarray   # int32 [n]

    @property
    def n(self) -> int:
        return self.offsets.size - 1

    def passage(self, i: int) -> np.ndarray:
        return self.tokens[int(self.offsets[i]): int(self.offsets[i + 1])]


def _searchsorted_right(sorted_f64: np.ndarray, values_f64: np.ndarray) -> np.ndarray:
    """np.searchsorted(side='right'); large inputs go through torch on the GPU when there is one
    (identical result: exact float64 comparisons), because 10^8 cache-missing binary searches
    take a minute on the host."""
    if values_f64.size > (1 << 22):
        try:
            import torch

            if torch.cuda.is_available():
                a = torch.from_numpy(np.ascontiguousarray(sorted_f64)).cuda()
                out = np.empty(values_f64.size, np.int64)
                step = 1 << 26
                for b0 in range(0, values_f64.size, step):
                    v = torch.from_numpy(np.ascontiguousarray(values_f64[b0:b0 + step])).cuda()
                    out[b0:b0 + step] = torch.searchsorted(a, v, right=True).cpu().numpy()
                return out
        except Exception:
            pass
    return np.searchsorted(sorted_f64, values_f64, side="right")


class TopicModel:
    """Two-level topic model shared by the corpus and the query generator.

    S super-topics own a private sub-vocabulary of `sub_vocab` words; each of the T topics belongs to
    one super-topic and is a peaked distribution over `topic_words` words drawn from it.  A passage
    picks a topic and draws tokens 80 % from the topic, 10 % uniformly from the super-topic vocabulary,
    10 % from a global Zipf background (measured at 1 M passages through the 6-layer encoder: 90 % of a
    query's exact top-10 lie in its own topic and HNSW M=32 reaches recall@10 0.94 at efSearch 64;
    with a 70/15/15 mix the figures are 71 % and 0.85 — gpurun_out/graph_recall2.log).  With ~32 passages per topic the exact top-10 of a query are
    concentrated in its topic while neighbouring topics stay closer than unrelated ones — the local
    structure real text embeddings have and i.i.d. bags of words lack."""

    def __init__(self, vocab_size: int, n_topics: int, seed: int = 1234, sub_vocab: int = 2048, topic_words: int = 64,
                 alpha: float = 0.3):
        rng = np.random.default_rng(seed)
        self.vocab_size = vocab_size
        self.n_topics = n_topics
        nwords = vocab_size - FIRST_WORD_ID
        sub_vocab = min(sub_vocab, nwords)
        topic_words = min(topic_words, sub_vocab)
        self.sub_vocab, self.tw = sub_vocab, topic_words
        # super-topics: sqrt(T) of them up to one million passages (T = 31 250 topics -> 177 super-topics of ~5 600 passages);
        # beyond that the super-topic SIZE is held there (T / 176.8), like the topic size.  Measured at 10 M passages this
        # does not by itself move graph recall (recall@10 at efSearch 64: 0.834 with sqrt(T), 0.838 with the fixed size —
        # profiles/r02_graph_recall_10m_sqrt_corpus.log, profiles/r02_graph_recall_10m_variants.log); what does is the topic
        # size and the topic share of the token mix, see BENCH_CORPUS below.
        self.n_super = max(2, int(round(np.sqrt(n_topics)))) if n_topics <= 31250 else int(round(n_topics / 176.8))
        self.super_words = (np.stack([rng.choice(nwords, sub_vocab, replace=False) for _ in range(self.n_super)])
                            .astype(np.int32) + FIRST_WORD_ID)                       # [S, sub_vocab]
        self.topic_super = rng.integers(0, self.n_super, n_topics).astype(np.int32)  # [T]
        # topic_words columns of the super-topic vocabulary per topic (a rare duplicate just merges two weights);
        # drawn directly so that 3*10^5 topics (10 M passages) cost seconds, not an argsort of 6*10^8 numbers
        pick = rng.integers(0, sub_vocab, (n_topics, topic_words))
        self.topic_word_ids = self.super_words[self.topic_super[:, None], pick]  # [T, tw]
        g = rng.gamma(alpha, 1.0, (n_topics, topic_words)) + 1e-9
        g /= g.sum(1, keepdims=True)
        cdf = np.cumsum(g, 1)
        cdf[:, -1] = 1.0
        self.flat_cdf = (cdf + np.arange(n_topics)[:, None]).reshape(-1)  # row t lives in [t, t + 1)
        zipf = 1.0 / np.arange(1, nwords + 1) ** 1.1
        zipf /= zipf.sum()
        self.zipf_cdf = np.cumsum(zipf)
        self.zipf_cdf[-1] = 1.0
        self.zipf_perm = rng.permutation(nwords).astype(np.int32) + FIRST_WORD_ID

    def sample_gpu(self, n: int, seed: int, len_mean: float, len_std: float, len_min: int, len_max: int,
                   p_topic: float = 0.80, p_super: float = 0.10, device: str = "cuda", chunk: int = 1 << 20) -> Corpus:
        """The same generative model as sample(), drawn with torch on the GPU (10^9 tokens in seconds instead of a
        minute of single-threaded numpy).  A different random stream than sample(): deterministic in (seed, sizes,
        GPU generator), not interchangeable with the numpy corpus of the same seed."""
        import torch

        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        topics = torch.randint(0, self.n_topics, (n,), generator=g, device=dev, dtype=torch.int64)
        lens = torch.clamp(torch.round(torch.randn(n, generator=g, device=dev) * len_std + len_mean), len_min, len_max).long()
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(lens, 0)
        total = int(offsets[-1])
        flat_cdf = torch.from_numpy(self.flat_cdf).to(dev)
        word_ids = torch.from_numpy(self.topic_word_ids.reshape(-1).astype(np.int64)).to(dev)
        super_words = torch.from_numpy(self.super_words.astype(np.int64)).to(dev)
        topic_super = torch.from_numpy(self.topic_super.astype(np.int64)).to(dev)
        zipf_cdf = torch.from_numpy(self.zipf_cdf).to(dev)
        zipf_perm = torch.from_numpy(self.zipf_perm.astype(np.int64)).to(dev)
        tokens = np.empty(total, np.uint16)
        for p0 in range(0, n, chunk):
            p1 = min(n, p0 + chunk)
            ln = lens[p0:p1]
            tt = torch.repeat_interleave(topics[p0:p1], ln)
            m = tt.numel()
            u = torch.rand(m, generator=g, device=dev, dtype=torch.float64)
            src = torch.rand(m, generator=g, device=dev)
            idx = torch.searchsorted(flat_cdf, u + tt.double(), right=True)
            idx = torch.minimum(idx, (tt + 1) * self.tw - 1)
            tok = word_ids[idx]
            is_super = (src >= p_topic) & (src < p_topic + p_super)
            col = torch.clamp((u * self.sub_vocab).long(), max=self.sub_vocab - 1)
            tok = torch.where(is_super, super_words[topic_super[tt], col], tok)
            zi = torch.clamp(torch.searchsorted(zipf_cdf, u, right=True), max=zipf_perm.numel() - 1)
            tok = torch.where(src >= p_topic + p_super, zipf_perm[zi], tok)
            st = offsets[p0:p1] - offsets[p0]
            tok[st] = CLS_ID
            tok[st + ln - 1] = SEP_ID
            tokens[int(offsets[p0]):int(offsets[p1])] = tok.to(torch.int16).cpu().numpy().view(np.uint16)  # ids < 32768
        return Corpus(tokens, offsets.cpu().numpy().astype(np.uint64), topics.cpu().numpy().astype(np.int32))

    def sample(self, n: int, seed: int, len_mean: float, len_std: float, len_min: int, len_max: int,
               p_topic: float = 0.80, p_super: float = 0.10) -> Corpus:
        rng = np.random.default_rng(seed)
        topics = rng.integers(0, self.n_topics, n).astype(np.int32)
        lens = np.clip(np.rint(rng.normal(len_mean, len_std, n)), len_min, len_max).astype(np.int64)
        offsets = np.zeros(n + 1, np.uint64)
        offsets[1:] = np.cumsum(lens)
        total = int(offsets[-1])
        tok_topic = np.repeat(topics, lens)
        u = rng.random(total)
        src = rng.random(total)
        m_topic, m_super = src < p_topic, (src >= p_topic) & (src < p_topic + p_super)
        m_bg = src >= p_topic + p_super
        tokens = np.empty(total, np.uint16)
        tt = tok_topic[m_topic]
        idx = _searchsorted_right(self.flat_cdf, u[m_topic] + tt)  # one search over the concatenated CDFs
        idx = np.minimum(idx, (tt.astype(np.int64) + 1) * self.tw - 1)
        tokens[m_topic] = self.topic_word_ids.reshape(-1)[idx]
        col = np.minimum((u[m_super] * self.sub_vocab).astype(np.int64), self.sub_vocab - 1)
        tokens[m_super] = self.super_words[self.topic_super[tok_topic[m_super]], col]
        zi = _searchsorted_right(self.zipf_cdf, u[m_bg])
        tokens[m_bg] = self.zipf_perm[np.minimum(zi, self.zipf_perm.size - 1)]
        starts = offsets[:-1].astype(np.int64)
        tokens[starts] = CLS_ID
        tokens[starts + lens - 1] = SEP_ID
        return Corpus(tokens, offsets, topics)


# The corpus the benchmark and the at-scale parity test search.  Same generator, tighter clusters than the defaults
# (64 passages per topic, 90 / 5 / 5 % topic / super-topic / background tokens; queries drawn with the same mix): measured at
# 10 M passages through the 6-layer encoder, 94 % of a query's exact top-10 lie in its own topic and the HNSW graph
# (M 32, efConstruction 200) reaches recall@10 0.910 at efSearch 64.  The defaults (32 per topic, 80 / 10 / 10) give 0.93 at
# 1 M but 0.84 at 10 M, where no builder setting tried (second sweep, fill, level-1 coverage, efConstruction 256) got past
# 0.855 — the variants and their numbers are in profiles/r02_graph_recall_10m_variants.log and DESIGN.md section 5.
BENCH_CORPUS = {"topic_size": 64, "p_topic": 0.90, "p_super": 0.05}


def make_bench_corpus(n: int, vocab_size: int, max_len: int, device: str | None, n_queries: int):
    """(topic model, passages, query pool) of the benchmark world."""
    c = BENCH_CORPUS
    tm, corpus = make_corpus(n, vocab_size, seed=1234, max_len=max_len, n_topics=max(4, n // c["topic_size"]),
                             p_topic=c["p_topic"], p_super=c["p_super"], device=device)
    queries = make_queries(tm, n_queries, seed=4321, p_topic=c["p_topic"], p_super=c["p_super"])
    return tm, corpus, queries


def make_corpus(n: int, vocab_size: int = 30522, seed: int = 1234, max_len: int = 256, n_topics: int | None = None,
                p_topic: float = 0.80, p_super: float = 0.10, device: str | None = None):
    """device="cuda[:i]" draws large corpora (>= 2^20 passages) with the GPU sampler; small ones always use numpy so
    that tests and fixtures do not depend on the GPU's random stream."""
    tm = TopicModel(vocab_size, n_topics or max(4, n // 32), seed)
    if device is not None and str(device).startswith("cuda") and n >= (1 << 20) and vocab_size <= 32768:
        corpus = tm.sample_gpu(n, seed + 1, 128, 48, 16, max_len, p_topic, p_super, device=device)
    else:
        corpus = tm.sample(n, seed + 1, 128, 48, 16, max_len, p_topic, p_super)
    return tm, corpus


def make_queries(tm: TopicModel, nq: int, seed: int = 4321, len_mean: float = 32, p_topic: float = 0.80,
                 p_super: float = 0.10) -> Corpus:
    return tm.sample(nq, seed, len_mean, len_mean / 3, 4, 64, p_topic, p_super)
