"""One warm launch of fa_fwd and fa_bwd at the cfg2 attn1 shape (for ncu captures)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops

dev = "cuda"
B, H, N, D = 1, 32, int(os.environ.get("N", 6144)), 2048
g = torch.Generator(device="cpu").manual_seed(0)
q, k, v, do = (torch.randn(B * N, D, generator=g).to(dev, torch.bfloat16) for _ in range(4))
for _ in range(2):
    o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    dq = ops.fa_bwd(q, k, v, o, do, lse, B, H, N, N, dk, dv, None, 0.125)
torch.cuda.synchronize()
print("ok")
