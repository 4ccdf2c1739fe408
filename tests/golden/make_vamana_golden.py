"""Generates the committed DiskANN/Vamana fixtures:  a small index in the reference's file formats + the outputs of the CPU oracle
(oracle/vamana_oracle.c — queue and PQ primitives pinned to the reference's compiled neighbor.h / pq.cpp; the search loop itself
is a restatement, see its header).  The reference's own searcher cannot be built here, so these goldens pin *regressions* of the
oracle + CUDA pair, not the reference binary.
    python tests/golden/make_vamana_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from leann_b200.vamana_build import build_diskann_index  # noqa: E402
from oracle.vamana_binding import VamanaOracle  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = [(64, 1, 10), (32, 4, 5), (16, 2, 10), (100, 2, 10)]  # (L, beam_width, k) — keep in sync with tests/test_vamana_oracle.py


def main():
    rng = np.random.default_rng(77)
    cen = rng.standard_normal((10, 16)).astype(np.float32)
    emb = cen[rng.integers(0, 10, 1200)] + 0.4 * rng.standard_normal((1200, 16)).astype(np.float32)
    emb *= rng.uniform(0.7, 1.3, (1200, 1)).astype(np.float32)
    q = cen[rng.integers(0, 10, 24)] + 0.4 * rng.standard_normal((24, 16)).astype(np.float32)
    exp = {}
    for metric in ("mips", "l2"):
        prefix, g, coords, pq, codes, mx = build_diskann_index(OUT, f"vamana_small_{metric}", emb, metric=metric, R=12, n_chunks=6,
                                                               partition=True, keep_disk_index=True, device="cpu")
        o = VamanaOracle(g, pq, codes, metric, mx)
        for L, beam, k in CASES:
            for mode, kw in (("stored", dict(coords=coords)), ("deferred", dict(emb=emb)), ("pq", dict(skip_search_reorder=True))):
                D, I, info = o.search(q, k, L=L, beam_width=beam, **kw)
                key = f"{metric}_L{L}_b{beam}_k{k}_{mode}"
                exp[key + "_D"], exp[key + "_I"] = D, I
                exp[key + "_nfull"], exp[key + "_cmps"], exp[key + "_hops"] = info["n_full"], info["cmps"], info["n_hops"]
                w = int(info["n_full"].max())
                exp[key + "_full"] = info["full_ids"][:, :w]
    np.save(OUT / "vamana_small_emb.npy", emb)
    np.save(OUT / "vamana_small_queries.npy", q)
    np.savez_compressed(OUT / "vamana_small_expected.npz", **exp)
    print("wrote", sorted(p.name for p in OUT.glob("vamana_small_*")))


if __name__ == "__main__":
    main()
