"""Diagnostics for the tcgen05 attention kernel (csrc/attention_tc.cu): structured inputs that separate the stages
(S = Q.K^T operand layout, P in tensor memory, V as MN-major operand, O read-out) and print where the first errors are."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import capi  # noqa: E402

lib = capi.load()
H, heads, hd = 384, 12, 32


def run(q, k, v, lens):
    """q, k, v: [T, heads, hd] fp16 cuda.  Returns ctx [T, H] fp16 and the fp32 reference."""
    T = q.shape[0]
    qkvh = torch.stack([q, k, v], 2).permute(1, 0, 2, 3).contiguous()  # [heads][T][3][hd]
    ctx = torch.full((T, H), float("nan"), device="cuda", dtype=torch.float16)
    dl = torch.tensor(lens, dtype=torch.int32, device="cuda")
    ds = (torch.cumsum(dl, 0) - dl).to(torch.int32)
    rc = lib.lb2_test_attention_f16(qkvh.data_ptr(), ds.data_ptr(), dl.data_ptr(), len(lens), T, H, heads, 256, ctx.data_ptr())
    assert rc == 0, lib.lb2_last_error()
    ref = torch.empty(T, H, device="cuda")
    off = 0
    for L in lens:
        qq, kk, vv = (t[off:off + L].float().transpose(0, 1) for t in (q, k, v))
        p = torch.softmax(qq @ kk.transpose(1, 2) / hd ** 0.5, dim=-1)
        ref[off:off + L] = (p @ vv).transpose(0, 1).reshape(L, H)
        off += L
    return ctx, ref


def report(name, ctx, ref, lens):
    err = (ctx.float() - ref).abs()
    bad = ~torch.isfinite(ctx)
    print(f"== {name}: max err {err[~bad].max().item() if (~bad).any() else float('nan'):.4g}, non-finite {int(bad.sum())} / {ctx.numel()}")
    off = 0
    for L in lens:
        e = err[off:off + L]
        e = torch.where(torch.isfinite(e), e, torch.full_like(e, 9.0))
        if e.max() > 4e-3:
            r, c = divmod(int(e.argmax()), H)
            print(f"   L={L}: max {e.max().item():.4g} at row {r} head {c // hd} col {c % hd}; rows>tol {int((e.max(1).values > 4e-3).sum())}/{L}; "
                  f"got {ctx[off + r, c].item():.4f} ref {ref[off + r, c].item():.4f}")
            print("     got row0 head0:", np.round(ctx[off, :8].float().cpu().numpy(), 3), " ref:", np.round(ref[off, :8].cpu().numpy(), 3))
        off += L


g = torch.Generator(device="cuda").manual_seed(0)
for lens in ([16], [32], [64], [100], [128], [129], [200], [256], [1, 2, 15, 17, 63, 65, 100, 128, 200, 255, 256]):
    T = sum(lens)
    z = torch.zeros(T, heads, hd, device="cuda", dtype=torch.float16)
    cols = torch.arange(hd, device="cuda", dtype=torch.float16)[None, None, :].expand(T, heads, hd).contiguous()
    rows = (torch.arange(T, device="cuda", dtype=torch.float16) % 64)[:, None, None].expand(T, heads, hd).contiguous() / 8
    rnd = lambda s=1.0: (torch.randn(T, heads, hd, device="cuda", generator=g) * s).half()  # noqa: E731
    print(f"######## lens {lens}")
    report("uniform P, V = column index (V column mapping, O read-out)", *run(z, z, cols, lens), lens)
    report("uniform P, V = key index / 8 (every key row is used once)", *run(z, z, rows, lens), lens)
    report("uniform P, V random (MN-major descriptor)", *run(z, z, rnd(), lens), lens)
    report("random Q K, V = column index (S operands, softmax, P layout)", *run(rnd(), rnd(), cols, lens), lens)
    report("random Q K V", *run(rnd(1.5), rnd(1.5), rnd(1.5), lens), lens)
