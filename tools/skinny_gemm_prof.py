"""The LoRA side GEMM shapes (t = x A^T: [6144, 64, 2048]; dA = dt^T x: [64, 2048, 6144] split-K) for one ncu capture."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops

M, D = 6144, 2048
x = (torch.randn(M, D, device="cuda") * 0.1).bfloat16()
a = (torch.randn(64, D, device="cuda") * 0.1).bfloat16()
dt = (torch.randn(M, 64, device="cuda") * 0.1).bfloat16()
for _ in range(3):
    t = ops.gemm(x, a, block_n=64)
    dA = ops.gemm(dt[:, :32], x, a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32, split_k=0)
torch.cuda.synchronize()
print("ok")
