"""Plugin boundary driven by the REFERENCE's own LeannSearcher (leann-core imported from /root/reference, nothing copied):
registry lookup by meta['backend_name'], constructor kwargs, _ensure_server_running / compute_query_embedding / search call
protocol (leann-core/src/leann/api.py:606-780), label -> passage mapping.  No GPU here, so the device handle (capi.Index /
capi.DiskannIndex) is replaced by a recording fake: what is under test is every line of host glue between LEANN and the C ABI.
Skipped on the GPU box, where /root/reference does not exist (tests/test_gpu_search.py covers the plugin API there)."""
import json
import os
import pickle
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

CORE = "/root/reference/packages/leann-core/src"
ROOT = Path(__file__).resolve().parents[1]

DRIVER = r'''
import json, sys
import numpy as np
import leann, leann.api as api, leann.registry as reg
import leann_b200.backend as hb, leann_b200.diskann_backend as db
from leann_b200 import capi

calls = []

class FakeHandle:
    """Stands in for the ctypes handle: returns ids 3,1,2 with descending scores and records how it was called."""
    def __init__(self, *a, **kw):
        calls.append(("open", type(self).__name__, [str(x) for x in a]))
        self.info = type("I", (), dict(d=8, ntotal=5, has_vectors=0))()
        self.dinfo = type("D", (), dict(dim=8))()
        self.last_stats = capi.SearchStats()
    def set_passages(self, tokens, offsets): calls.append(("set_passages", int(tokens.size), int(offsets.size)))
    def set_encoder(self, cfg, blob): calls.append(("set_encoder", int(cfg.hidden), int(blob.size)))
    def configure(self, *a): calls.append(("configure",) + a)
    def set_option(self, *a): calls.append(("set_option",) + a)
    def encode_tokens(self, toks, offs):
        calls.append(("encode_tokens", toks.tolist()))
        return np.full((1, 8), 0.5, np.float32)
    def search(self, q, k, params):
        calls.append(("search", q.shape, q.dtype.name, k, {f: getattr(params, f) for f, _ in params._fields_}))
        return np.array([[0.9, 0.8, 0.7][:k]], np.float32), np.array([[3, 1, 2][:k]], np.int64)
    def close(self): calls.append(("close",))

capi.Index = FakeHandle
class FakeDiskann(FakeHandle): pass
capi.DiskannIndex = FakeDiskann

out = {}
for name in ("hnsw", "diskann"):
    calls.clear()
    s = api.LeannSearcher(sys.argv[1] + f"/{name}.leann")
    assert isinstance(s.backend_impl, (hb.B200HnswSearcher, db.B200DiskannSearcher))
    s.backend_impl._tokenizer = lambda text, truncation=True, max_length=0: {"input_ids": [101, 7, 8, 102]}
    res = s.search("where is the answer", top_k=3, complexity=48, beam_width=2, recompute_embeddings=True)
    out[name] = dict(texts=[r.text for r in res], ids=[r.id for r in res], scores=[float(r.score) for r in res])
    s.cleanup()                      # LeannSearcher.cleanup -> embedding_server_manager.stop_server() -> handle released
    assert s.backend_impl._index is None
    res2 = s.search("again", top_k=2, recompute_embeddings=True)   # ... and transparently re-opened, like a re-spawned server
    assert [r.id for r in res2] == ["3", "1"]
    out[name]["calls"] = [list(map(lambda x: x if not isinstance(x, tuple) else list(x), c)) for c in calls]
    del s, res, res2                 # LeannSearcher.__del__ -> cleanup() again: must not leak into the next backend's record
print("RESULT " + json.dumps(out))
'''


def _make_index_dir(d: Path, name: str, backend: str):
    texts = [f"passage number {i}" for i in range(5)]
    offs = {}
    with open(d / f"{name}.leann.passages.jsonl", "w") as f:
        for i, t in enumerate(texts):
            offs[str(i)] = f.tell()
            f.write(json.dumps({"id": str(i), "text": t, "metadata": {}}) + "\n")
    with open(d / f"{name}.leann.passages.idx", "wb") as f:
        pickle.dump(offs, f)
    meta = {"version": "1.0", "backend_name": backend, "embedding_model": "synthetic/tiny-bert", "dimensions": 8,
            "backend_kwargs": {"distance_metric": "mips"}, "embedding_mode": "sentence-transformers",
            "passage_sources": [{"type": "jsonl", "path": f"{name}.leann.passages.jsonl", "index_path": f"{name}.leann.passages.idx"}],
            "is_compact": True, "is_pruned": True, "b200_synthetic_weights": True}
    (d / f"{name}.leann.meta.json").write_text(json.dumps(meta))
    np.save(d / f"{name}.leann.tokens.npy", np.arange(20, dtype=np.uint16))
    np.save(d / f"{name}.leann.tokoffsets.npy", np.arange(0, 24, 4, dtype=np.uint64))
    if backend == "hnsw_b200":
        (d / f"{name}.index").write_bytes(b"placeholder: the fake handle never parses it")
    else:
        (d / f"{name}_pq_compressed.bin").write_bytes(b"placeholder")


@pytest.mark.skipif(not os.path.isdir(CORE), reason="reference checkout not present (GPU box)")
def test_reference_leannsearcher_drives_both_backends(tmp_path):
    _make_index_dir(tmp_path, "hnsw", "hnsw_b200")
    _make_index_dir(tmp_path, "diskann", "diskann_b200")
    env = dict(os.environ, PYTHONPATH=CORE + os.pathsep + str(ROOT))
    p = subprocess.run([sys.executable, "-c", DRIVER, str(tmp_path)], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    out = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    for name, handle in (("hnsw", "FakeHandle"), ("diskann", "FakeDiskann")):
        r = out[name]
        assert r["ids"] == ["3", "1", "2"] and r["texts"] == ["passage number 3", "passage number 1", "passage number 2"]
        assert r["scores"] == pytest.approx([0.9, 0.8, 0.7])
        kinds = [c[0] for c in r["calls"]]
        assert kinds[0] == "open" and r["calls"][0][1] == handle
        # recompute stage attached once (passages + encoder), query embedded on the same encoder, then one search call
        assert kinds == ["open", "set_passages", "set_encoder", "encode_tokens", "search", "close",
                         "open", "set_passages", "set_encoder", "encode_tokens", "search"]
        assert ["encode_tokens", [101, 7, 8, 102]] in r["calls"]
        search = [c for c in r["calls"] if c[0] == "search"][0]
        assert search[1] == [1, 8] and search[2] == "float32" and search[3] == 3
        prm = search[4]
        if name == "hnsw":
            assert (prm["efSearch"], prm["beam_size"], prm["recompute"], prm["batch_size"]) == (48, 2, 1, 0)
        else:
            assert (prm["complexity"], prm["beam_width"], prm["deferred_fetch"], prm["recompute_neighbors"]) == (48, 2, 1, 0)
        assert [c for c in r["calls"] if c[0] == "search"][1][3] == 2
