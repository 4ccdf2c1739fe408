"""A/B of the tail-block rotation of the tcgen05 attention kernel (LB2_ATTN_ROTATE, attention_tc.cu) on the bench corpus'
length mix and on two other mixes; CUDA-event timing of the kernel call through the C-ABI test hook, interleaved runs."""
import os, sys
from pathlib import Path
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import build, capi

if build.needs_build():
    build.build()
lib = capi.load()
H, heads = 384, 12
dev = "cuda"


def run(mu, sd, n_seq, reps=6):
    gl = torch.Generator().manual_seed(1)
    lens = torch.clamp((torch.randn(n_seq, generator=gl) * sd + mu).round(), 16, 256).to(torch.int32)
    T = int(lens.sum())
    qkv = torch.randn(heads, T, 96, device=dev).half()
    dl = lens.to(dev)
    ds = (torch.cumsum(dl, 0) - dl).to(torch.int32)
    outs, times = {}, {"0": [], "1": []}
    for rep in range(reps):
        for rot in ("0", "1"):
            os.environ["LB2_ATTN_ROTATE"] = rot
            ctx = torch.full((T, H), float("nan"), device=dev, dtype=torch.float16)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.lb2_test_attention_f16(qkv.data_ptr(), ds.data_ptr(), dl.data_ptr(), n_seq, T, H, heads, 256, ctx.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, lib.lb2_last_error()
            if rep > 0:
                times[rot].append(e0.elapsed_time(e1))
            outs[rot] = ctx
    same = bool(torch.equal(outs["0"], outs["1"])) and bool(torch.isfinite(outs["1"]).all())
    m0, m1 = min(times["0"]), min(times["1"])
    print(f"lengths N({mu},{sd}) n_seq={n_seq} tokens={T}: rotate=0 {m0:.3f} ms, rotate=1 {m1:.3f} ms ({(m0 / m1 - 1) * 100:+.1f} %), "
          f"bit-identical outputs: {same}", flush=True)


run(128, 48, 4096)
run(160, 40, 3300)
run(96, 40, 5400)
run(140, 8, 3700)   # almost every passage has a tail of <= 32 rows
