"""Shared helpers of the test-suite (test infrastructure; may use oracle/)."""
from __future__ import annotations

import json
import tempfile
from pathlib import Path

import numpy as np

from leann_b200 import csr, synth

GOLDEN_CASES = [  # (ef, beam, batch, check_rel, k) — keep in sync with tests/golden/make_golden.py
    (64, 1, 0, 1, 10), (16, 1, 0, 1, 10), (8, 1, 0, 1, 20), (32, 4, 0, 1, 10), (32, 16, 0, 1, 5),
    (32, 1, 24, 1, 10), (24, 2, 0, 0, 10), (128, 1, 0, 1, 1), (64, 3, 40, 0, 7),
]


def golden_key(tag, ef, beam, batch, cr, k):
    return f"{tag}_ef{ef}_b{beam}_bs{batch}_cr{cr}_k{k}"


from leann_b200.tooling import open_encoder_only, recall_at_k, stub_graph, write_leann_index  # noqa: E402,F401


VAMANA_GOLDEN_CASES = [(64, 1, 10), (32, 4, 5), (16, 2, 10), (100, 2, 10)]  # keep in sync with tests/golden/make_vamana_golden.py


def load_vamana_golden(golden_dir, metric):
    """The committed DiskANN-format fixture, read back with the format readers (not the builder)."""
    from leann_b200 import diskann_format as dfmt

    prefix = str(golden_dir / f"vamana_small_{metric}")
    coords, g = dfmt.read_disk_index(prefix + "_disk.index")
    pq = dfmt.read_pq_pivots(prefix + "_pq_pivots.bin")
    codes = dfmt.read_bin(prefix + "_pq_compressed.bin", np.uint8)
    mx = float(dfmt.read_bin(prefix + "_disk.index_max_base_norm.bin", np.float32)[0, 0]) if metric == "mips" else 0.0
    g.medoid = int(dfmt.read_bin(prefix + "_disk.index_medoids.bin", np.uint32)[0, 0])
    return dict(prefix=prefix, g=g, coords=coords, pq=pq, codes=codes, max_norm=mx,
                emb=np.load(golden_dir / "vamana_small_emb.npy"), q=np.load(golden_dir / "vamana_small_queries.npy"),
                exp=np.load(golden_dir / "vamana_small_expected.npz"))


def load_c1(golden_dir, n_queries=40, seed=5):
    """BASELINE config C1 (the reference's sample document) as a synth.Corpus: 254-token chunks at stride 127, each wrapped in
    [CLS]=101 / [SEP]=102, plus queries = random 12..30-token spans of the same text (wrapped the same way)."""
    stream = np.load(golden_dir / "c1_pride_tokens.npz")["tokens"]
    starts = np.arange(0, max(1, len(stream) - 127), 127)
    toks, offs = [], [0]
    for s in starts:
        body = stream[s:s + 254]
        toks.append(np.concatenate([[101], body, [102]]).astype(np.uint16))
        offs.append(offs[-1] + len(toks[-1]))
    corpus = synth.Corpus(np.concatenate(toks), np.asarray(offs, np.uint64), np.zeros(len(starts), np.int32))
    rng = np.random.default_rng(seed)
    qt, qo = [], [0]
    for _ in range(n_queries):
        a = int(rng.integers(0, len(stream) - 40))
        body = stream[a:a + int(rng.integers(12, 31))]
        qt.append(np.concatenate([[101], body, [102]]).astype(np.uint16))
        qo.append(qo[-1] + len(qt[-1]))
    queries = synth.Corpus(np.concatenate(qt), np.asarray(qo, np.uint64), np.zeros(n_queries, np.int32))
    return corpus, queries
