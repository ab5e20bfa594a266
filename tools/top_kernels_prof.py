"""Warm launches of the three dominant kernels at cfg2 shapes, for one `ncu --set full` capture:
fa_fwd / fa_bwd (attn1, 6144 x 6144, 32 heads), the CTA-pair GEMM (qkv projection and FF2 dgrad)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops

dev = "cuda"
B, H, N, D, F = 1, 32, 6144, 2048, 8192
g = torch.Generator(device="cpu").manual_seed(0)
q, k, v, do = (torch.randn(B * N, D, generator=g).to(dev, torch.bfloat16) for _ in range(4))
x = (torch.randn(N, D, generator=g) * 0.05).to(dev, torch.bfloat16)
w3 = (torch.randn(3 * D, D, generator=g) * 0.05).to(dev, torch.bfloat16)
xf = (torch.randn(N, F, generator=g) * 0.05).to(dev, torch.bfloat16)
wfd = (torch.randn(F, D, generator=g) * 0.05).to(dev, torch.bfloat16)
for _ in range(2):
    o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    ops.fa_bwd(q, k, v, o, do, lse, B, H, N, N, dk, dv, None, 0.125)
    ops.gemm(x, w3)                           # qkv forward: [6144, 6144, 2048]
    ops.gemm(xf, wfd, b_rows_are_k=True)      # FF1 dgrad: [6144, 2048, 8192], B read MN-major
torch.cuda.synchronize()
print("ok")
