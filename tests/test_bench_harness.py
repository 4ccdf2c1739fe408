"""Host logic of bench.py that the headline number rests on (no GPU): the exact ground truth used for recall@10, the
result comparison behind the `parity` object, the cached-world key and the file rendezvous of the ranks."""
import argparse
import importlib.util
import sys
import threading
import time
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("lb2_bench_module", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lb2_bench_module"] = mod
    spec.loader.exec_module(mod)  # the __main__ guard keeps the run out
    return mod


def _args(**kw):
    d = dict(chunks=1000, efc=200, sweeps=0, pool=64, cache="/tmp/lb2_cache_test", rebuild=False)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("n,d,nq", [(1000, 32, 17), (777, 48, 5), (64, 16, 3)])
def test_exact_ground_truth_is_the_brute_force_top_k(bench, n, d, nq):
    """Two-stage ground truth (fp16 shortlist of 16 groups of 64 columns, fp32 re-score) == fp32 brute force
    (benchmarks/run_evaluation.py:358-367 with k = 10), including n that is not a multiple of the group size."""
    g = torch.Generator().manual_seed(n + d)
    E = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(nq, d, generator=g), dim=1).numpy()
    gt = bench.exact_ground_truth(Q, E, k=10)
    ref = torch.topk(torch.from_numpy(Q) @ E.T, 10, dim=1).indices.numpy()
    assert gt.shape == (nq, 10)
    for a, b in zip(gt, ref):
        assert set(a.tolist()) == set(b.tolist())
        assert len(set(a.tolist())) == 10 and a.max() < n


def test_compare_results_counts_overlap_identity_and_distance_gap(bench):
    I_ref = np.array([[1, 2, 3], [4, 5, 6]], np.int64)
    D_ref = np.array([[0.9, 0.8, 0.7], [0.6, 0.5, 0.4]], np.float32)
    I_gpu = np.array([[3, 2, 1], [4, 5, 9]], np.int64)          # first query: same set, other order; second: one id differs
    D_gpu = np.array([[0.7, 0.8001, 0.9], [0.6, 0.5, 0.3]], np.float32)
    r = bench.compare_results(I_gpu, D_gpu, I_ref, D_ref)
    assert r["queries"] == 2 and r["identical_id_sets"] == 1
    assert abs(r["topk_overlap"] - (1.0 + 2 / 3) / 2) < 1e-9
    assert abs(r["max_abs_dD"] - 1e-4) < 1e-6                   # compared id by id, not slot by slot
    # unfilled slots (-1) are not ids
    r = bench.compare_results(np.array([[7, -1]]), np.zeros((1, 2), np.float32), np.array([[7, -1]]), np.zeros((1, 2), np.float32))
    assert r["identical_id_sets"] == 1 and r["topk_overlap"] == 1.0


def test_world_key_depends_on_everything_a_cached_world_holds(bench, tmp_path):
    a = bench.world_dir(_args(cache=str(tmp_path)))
    assert a == bench.world_dir(_args(cache=str(tmp_path)))
    for change in (dict(chunks=1001), dict(efc=100), dict(sweeps=1), dict(pool=128)):
        assert bench.world_dir(_args(cache=str(tmp_path), **change)) != a
    assert a.parent == tmp_path and a.name.startswith("c1000_")


def test_ranks_wait_for_the_builders_done_marker(bench, tmp_path, monkeypatch):
    """ensure_world: the builder rank builds, every other rank polls for DONE and never builds (bench.py run under torchrun:
    this rendezvous happens BEFORE the process group exists)."""
    args = _args(cache=str(tmp_path))
    built = []

    def fake_build(a, wd, device):
        wd.mkdir(parents=True, exist_ok=True)
        time.sleep(0.5)
        built.append(device)
        (wd / "DONE").write_text("ok")

    monkeypatch.setattr(bench, "build_world", fake_build)
    got = {}
    t = threading.Thread(target=lambda: got.setdefault("waiter", bench.ensure_world(args, builder=False, device=1)))
    t.start()
    got["builder"] = bench.ensure_world(args, builder=True, device=0)
    t.join(10)
    assert not t.is_alive()
    assert built == [0] and got["waiter"] == got["builder"] and (got["builder"] / "DONE").exists()
    # a cached world is reused by both kinds of rank; --rebuild drops it on the builder only
    assert bench.ensure_world(args, builder=True, device=0) == got["builder"] and built == [0]
    args.rebuild = True
    bench.ensure_world(args, builder=True, device=0)
    assert built == [0, 0]


def test_other_ranks_wait_until_rank0_has_the_library_in_place(bench, tmp_path, monkeypatch):
    """library_rendezvous: with sources newer than the .so, rank 0 rebuilds; ranks > 0 neither build nor return before it is done."""
    monkeypatch.setenv("MASTER_PORT", "29999")
    events = []

    class FakeBuild:
        @staticmethod
        def needs_build():
            return True

        @staticmethod
        def build():
            events.append("build-start")
            time.sleep(0.6)
            events.append("build-end")

    def other():
        bench.library_rendezvous(FakeBuild, 1, 2, str(tmp_path), timeout_s=10)
        events.append("rank1-go")

    t = threading.Thread(target=other)
    t.start()
    time.sleep(0.1)
    bench.library_rendezvous(FakeBuild, 0, 2, str(tmp_path))
    t.join(10)
    assert events == ["build-start", "build-end", "rank1-go"]
    # a lone process just builds
    events.clear()
    bench.library_rendezvous(FakeBuild, 0, 1, str(tmp_path))
    assert events == ["build-start", "build-end"]
