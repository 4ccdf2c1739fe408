"""Micro-benchmarks of the recompute-stage kernels through the C-ABI test hooks (1 GPU).
Prints one line per kernel/shape: time, algorithmic TFLOP/s (GEMM, attention) or GB/s (LayerNorm)."""
import sys, time
from pathlib import Path
import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import build, capi

if build.needs_build():
    build.build()
lib = capi.load()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 128
dev = "cuda"
torch.manual_seed(0)


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


for (N, K, epi, name) in [(1152, 384, 0, "qkv"), (384, 384, 2, "attn-out+res"), (1536, 384, 1, "ffn-up+gelu"), (384, 1536, 2, "ffn-down+res")]:
    A = (torch.randn(T, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev) * 0.1
    res = torch.randn(T, N, device=dev).half()
    C = torch.empty(T, N, device=dev, dtype=torch.float16)
    dt = timeit(lambda: lib.lb2_test_gemm_f16(A.data_ptr(), W.data_ptr(), b.data_ptr(), res.data_ptr(), C.data_ptr(), T, N, K, epi))
    t2 = timeit(lambda: torch.matmul(A, W.T))
    print(f"gemm {name:14s} M={T} N={N} K={K}: {dt*1e3:8.3f} ms  {2.0*T*N*K/dt/1e12:7.1f} TFLOP/s   (torch.matmul fp16 no epilogue: {2.0*T*N*K/t2/1e12:7.1f})")
    del A, W, res, C

for K, name in ((384, "attn-out+res+LN fused"), (1536, "ffn-down+res+LN fused")):
    N = 384
    A = (torch.randn(T, K, device=dev) * 0.5).half(); W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev) * 0.1; res = torch.randn(T, N, device=dev).half(); gm = torch.ones(N, device=dev); bt = torch.zeros(N, device=dev)
    C = torch.empty(T, N, device=dev, dtype=torch.float16)
    dt = timeit(lambda: lib.lb2_test_gemm_res_ln_f16(A.data_ptr(), W.data_ptr(), b.data_ptr(), res.data_ptr(), gm.data_ptr(), bt.data_ptr(), 1e-12, C.data_ptr(), T, N, K))
    print(f"gemm {name:22s} M={T} N={N} K={K}: {dt*1e3:8.3f} ms  {2.0*T*N*K/dt/1e12:7.1f} TFLOP/s")
    del A, W, res, C

H, heads = 384, 12
for L in (64, 128, 256):
    n_seq = T // L
    qkv = (torch.randn(n_seq * L, 3 * H, device=dev)).half()
    ctx = torch.empty(n_seq * L, H, device=dev, dtype=torch.float16)
    dl = torch.full((n_seq,), L, dtype=torch.int32, device=dev)
    ds = (torch.cumsum(dl, 0) - dl).to(torch.int32)
    dt = timeit(lambda: lib.lb2_test_attention_f16(qkv.data_ptr(), ds.data_ptr(), dl.data_ptr(), n_seq, n_seq * L, H, heads, 256, ctx.data_ptr()), 5)
    fl = 4.0 * L * H * n_seq * L
    print(f"attention L={L:3d} n_seq={n_seq}: {dt*1e3:8.3f} ms  {fl/dt/1e12:7.1f} TFLOP/s  ({n_seq*L/dt/1e6:7.1f} Mtok/s)")
    del qkv, ctx
# the bench corpus' length distribution: N(128, 48) clipped to [16, 256]
gl = torch.Generator().manual_seed(1)
lens = torch.clamp((torch.randn(T // 128, generator=gl) * 48 + 128).round(), 16, 256).to(torch.int32)
Tm = int(lens.sum())
qkv = (torch.randn(Tm, 3 * H, device=dev)).half()
ctx = torch.empty(Tm, H, device=dev, dtype=torch.float16)
dl = lens.to(dev)
ds = (torch.cumsum(dl, 0) - dl).to(torch.int32)
dt = timeit(lambda: lib.lb2_test_attention_f16(qkv.data_ptr(), ds.data_ptr(), dl.data_ptr(), len(lens), Tm, H, heads, 256, ctx.data_ptr()), 5)
fl = 4.0 * H * float((lens.double() ** 2).sum())
print(f"attention mixed lengths (mean {Tm/len(lens):.0f}) n_seq={len(lens)} tokens={Tm}: {dt*1e3:8.3f} ms  {fl/dt/1e12:7.1f} TFLOP/s  ({Tm/dt/1e6:7.1f} Mtok/s)")
del qkv, ctx

x = torch.randn(T, H, device=dev).half(); g = torch.randn(H, device=dev); bb = torch.randn(H, device=dev); o = torch.empty_like(x)
dt = timeit(lambda: lib.lb2_test_layernorm_f16(x.data_ptr(), g.data_ptr(), bb.data_ptr(), o.data_ptr(), T, H, 1e-12))
print(f"layernorm rows={T}: {dt*1e3:8.3f} ms  {2*T*H*2/dt/1e9:7.0f} GB/s")
