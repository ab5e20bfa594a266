"""Warm launches of the dominant kernels at cfg2 shapes, for one `ncu --set full` capture:
fa_fwd / fa_bwd (attn1, 6144 x 6144, 32 heads), the CTA-pair GEMM (qkv projection and FF1 dgrad), and the two
heaviest row kernels (norm_mod_bwd with the residual gradient, qknorm_rope_bwd on the q/k pair)."""
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
dy, dres = (torch.randn(N, D, generator=g).to(dev, torch.bfloat16) for _ in range(2))
scale = (torch.randn(1, D, generator=g) * 0.1).to(dev, torch.bfloat16)
qkv = torch.randn(N, 3 * D, generator=g).to(dev, torch.bfloat16)
wq, wk = (torch.randn(D, generator=g).to(dev, torch.bfloat16) for _ in range(2))
cos, sin = (torch.randn(N, D, generator=g).to(dev, torch.bfloat16) for _ in range(2))
dq32 = torch.randn(N, D, generator=g).to(dev)
dqkv = torch.empty(N, 3 * D, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    ops.fa_bwd(q, k, v, o, do, lse, B, H, N, N, dk, dv, None, 0.125)
    ops.gemm(x, w3)                           # qkv forward: [6144, 6144, 2048]
    ops.gemm(xf, wfd, b_rows_are_k=True)      # FF1 dgrad: [6144, 2048, 8192], B read MN-major
    ops.norm_mod_bwd(dy, x, scale, N, 1e-6, False, dres=dres)
    ops.qknorm_rope_bwd(dq32, dk, qkv[:, :D], qkv[:, D:2 * D], wq, wk, cos, sin, dqkv[:, :D], dqkv[:, D:2 * D])
torch.cuda.synchronize()
print("ok")
