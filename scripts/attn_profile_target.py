"""ncu target: the tcgen05 attention kernel at bench size (the bench corpus' length mix N(128, 48), ~524 k packed tokens, 12 heads)
through the C-ABI test hook, a few launches.  Numbers printed under ncu are not bench values."""
import sys
from pathlib import Path
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from leann_b200 import build, capi

if build.needs_build():
    build.build()
lib = capi.load()
H, heads, n_seq = 384, 12, 4096
gl = torch.Generator().manual_seed(1)
lens = torch.clamp((torch.randn(n_seq, generator=gl) * 48 + 128).round(), 16, 256).to(torch.int32)
T = int(lens.sum())
qkv = torch.randn(heads, T, 96, device="cuda").half()
dl = lens.cuda()
ds = (torch.cumsum(dl, 0) - dl).to(torch.int32)
ctx = torch.empty((T, H), device="cuda", dtype=torch.float16)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    assert lib.lb2_test_attention_f16(qkv.data_ptr(), ds.data_ptr(), dl.data_ptr(), n_seq, T, H, heads, 256, ctx.data_ptr()) == 0
torch.cuda.synchronize()
print("tokens", T)
