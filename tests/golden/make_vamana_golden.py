"""Generates the committed DiskANN/Vamana fixtures: a small index in the reference's file formats, the outputs of the CPU oracle
(oracle/vamana_oracle.c) and — keys *_refI / *_refD / *_refcmps / *_refhops / *_refios — the outputs of the REFERENCE ITSELF:
PQFlashIndex::load + cached_beam_search compiled from /root/reference (oracle/_ref/libleann_ref_diskann_flash.so, built by
`make -C oracle ref` in the dev container) reading the very files written here, in both layouts (<prefix>_disk.index and
<prefix>_partition.bin + _disk_graph.index).  The oracle must reproduce the reference's ids exactly and its distances within
2e-5 (canonical summation order, see vamana_oracle.c); tests/test_vamana_oracle.py checks both live and from this file.
    python tests/golden/make_vamana_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from leann_b200.vamana_build import build_diskann_index  # noqa: E402
from oracle.vamana_binding import DiskannFlash, VamanaOracle, have_diskann_flash  # noqa: E402

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
        if not have_diskann_flash():
            raise SystemExit("oracle/_ref/libleann_ref_diskann_flash.so missing: run `make -C oracle ref` (dev container) first")
        ref_std = DiskannFlash(prefix, metric)                     # <prefix>_disk.index
        ref_part = DiskannFlash(prefix, metric, partition_prefix=prefix)  # partition layout (what LEANN's recompute mode leaves)
        ref_std.set_embeddings(emb)
        ref_part.set_embeddings(emb)
        for L, beam, k in CASES:
            for mode, kw in (("stored", dict(coords=coords)), ("deferred", dict(emb=emb)), ("pq", dict(skip_search_reorder=True))):
                D, I, info = o.search(q, k, L=L, beam_width=beam, **kw)
                key = f"{metric}_L{L}_b{beam}_k{k}_{mode}"
                exp[key + "_D"], exp[key + "_I"] = D, I
                exp[key + "_nfull"], exp[key + "_cmps"], exp[key + "_hops"] = info["n_full"], info["cmps"], info["n_hops"]
                w = int(info["n_full"].max())
                exp[key + "_full"] = info["full_ids"][:, :w]
                rkw = dict(deferred_fetch=(mode == "deferred"), skip_search_reorder=(mode == "pq"))
                rD, rI, rs = ref_std.search(q, k, L=L, beam_width=beam, **rkw)
                exp[key + "_refD"], exp[key + "_refI"] = rD, rI
                exp[key + "_refcmps"], exp[key + "_refhops"], exp[key + "_refios"] = rs["n_cmps"], rs["n_hops"], rs["n_ios"]
                if mode != "stored":  # the partition layout holds no coordinates
                    pD, pI, ps = ref_part.search(q, k, L=L, beam_width=beam, **rkw)
                    assert np.array_equal(pI, rI) and np.array_equal(pD, rD) and all(np.array_equal(ps[x], rs[x]) for x in rs), key
    np.save(OUT / "vamana_small_emb.npy", emb)
    np.save(OUT / "vamana_small_queries.npy", q)
    np.savez_compressed(OUT / "vamana_small_expected.npz", **exp)
    print("wrote", sorted(p.name for p in OUT.glob("vamana_small_*")))


if __name__ == "__main__":
    main()
