"""Generates the committed golden fixtures of tests/golden/ FROM THE REFERENCE ITSELF.

Run in the dev container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

What comes from where:
  * graph        : the reference's HNSW::add_with_locks (faiss fork, impl/HNSW.cpp:839-890) driven
                   single-threaded by oracle/ref_harness.cpp -> deterministic links;
  * hnsw_small.index : that graph written in the fork's NON-compact IndexHNSWFlat layout by this
                   script, then converted by the reference's OWN converter
                   leann_backend_hnsw/convert_to_csr.py (imported from /root/reference, pure Python)
                   -> pins our CSR reader against the reference's writer;
  * expected I / D / ndis / nhops : the reference's HNSW::search + search_from_candidates compiled
                   from /root/reference (oracle/_ref/libleann_ref.so), for a grid of
                   (efSearch, beam_size, batch_size, check_relative_distance, k, metric).
Fixture shape follows the vendored faiss test (faiss/tests/test_hnsw.cpp:215-243): d=64, nb=2000, nq=10.
"""
from __future__ import annotations

import io
import struct
import sys
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.binding import Reference  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = [  # (ef, beam, batch, check_rel, k)
    (64, 1, 0, 1, 10), (16, 1, 0, 1, 10), (8, 1, 0, 1, 20), (32, 4, 0, 1, 10), (32, 16, 0, 1, 5),
    (32, 1, 24, 1, 10), (24, 2, 0, 0, 10), (128, 1, 0, 1, 1), (64, 3, 40, 0, 7),
]


def write_noncompact(path, ex, vectors):
    """IndexHNSWFlat as the fork's write_index lays it out (impl/index_write.cpp write_HNSW +
    the 'false' compact flag byte the fork inserts before `offsets`, cf. index_read.cpp:559-600)."""
    wv = lambda f, a, dt: (f.write(struct.pack("<Q", a.size)), f.write(np.ascontiguousarray(a, dt).tobytes()))
    metric = 0 if ex["metric_ip"] else 1
    with open(path, "wb") as f:
        f.write(b"IHNf")
        f.write(struct.pack("<iqqq?i", ex["d"], ex["ntotal"], 1 << 20, 1 << 20, True, metric))
        wv(f, ex["probas"], np.float64)
        wv(f, ex["cum"], np.int32)
        wv(f, ex["levels"], np.int32)
        f.write(struct.pack("<?", False))
        wv(f, ex["offsets"], np.uint64)
        wv(f, ex["neighbors"], np.int32)
        f.write(struct.pack("<5i", ex["entry_point"], ex["max_level"], ex["ef_construction"], ex["ef_search"], 1))
        f.write(b"IxFI" if ex["metric_ip"] else b"IxF2")
        f.write(struct.pack("<iqqq?i", ex["d"], ex["ntotal"], 1 << 20, 1 << 20, True, metric))
        f.write(struct.pack("<Q", vectors.size))
        f.write(np.ascontiguousarray(vectors, np.float32).tobytes())


def main():
    # the reference's converter, loaded by path (importing the package would pull in the SWIG faiss module)
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_convert_to_csr", "/root/reference/packages/leann-backend-hnsw/leann_backend_hnsw/convert_to_csr.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    convert_hnsw_graph_to_csr = mod.convert_hnsw_graph_to_csr

    rng = np.random.default_rng(12345)
    n, d, nq = 2000, 64, 10
    x = rng.random((n, d), dtype=np.float32)
    q = rng.random((nq, d), dtype=np.float32)
    # a few exact duplicates and a duplicated query: exercises distance ties (pop_min / heap tie rules)
    x[100] = x[7]; x[101] = x[7]; x[1500] = x[300]
    q[3] = x[7]
    np.save(OUT / "hnsw_small_vectors.npy", x)
    np.save(OUT / "hnsw_small_queries.npy", q)
    expected = {}
    for metric_ip in (True, False):
        tag = "ip" if metric_ip else "l2"
        R = Reference(d, M=8, metric_ip=metric_ip)
        R.build(x, ef_construction=40, nthreads=1)
        ex = R.export()
        nc = OUT / f"_tmp_{tag}.noncompact"
        write_noncompact(nc, ex, x)
        with redirect_stdout(io.StringIO()):
            ok_pruned = convert_hnsw_graph_to_csr(str(nc), str(OUT / f"hnsw_small_{tag}.index"), prune_embeddings=True)
        assert ok_pruned
        if metric_ip:
            with redirect_stdout(io.StringIO()):
                assert convert_hnsw_graph_to_csr(str(nc), str(OUT / "hnsw_small_ip_with_storage.index"), prune_embeddings=False)
        nc.unlink()
        for (ef, beam, batch, cr, k) in CASES:
            D, I, ndis, nhops = R.search(q, k, ef=ef, beam=beam, batch_size=batch, check_rel=bool(cr))
            key = f"{tag}_ef{ef}_b{beam}_bs{batch}_cr{cr}_k{k}"
            expected[key + "_D"] = D
            expected[key + "_I"] = I
            expected[key + "_ndis"] = ndis
            expected[key + "_nhops"] = nhops
    np.savez_compressed(OUT / "hnsw_small_expected.npz", **expected)
    print("wrote", sorted(p.name for p in OUT.glob("hnsw_small*")))


if __name__ == "__main__":
    main()
