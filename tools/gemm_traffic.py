"""The two K = 8192 GEMMs of a block (FF2 forward with gate + residual; FF1 dgrad) for an ncu DRAM-traffic capture:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_kernel python tools/gemm_traffic.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops
M, D, F = 6144, 2048, 8192
r = lambda *s: (torch.randn(*s, device="cuda") * 0.05).bfloat16()
xf, W_df, W_fd, bias_d, gate, res = r(M, F), r(D, F), r(F, D), r(D), r(1, D), r(M, D)
flush = torch.empty(512 << 20, device="cuda", dtype=torch.uint8)
for _ in range(2):
    flush.zero_()
    ops.gemm(xf, W_df, bias=bias_d, gate=gate, rows_per_gate=M, res=res)       # FF2 forward  [6144 x 2048 x 8192]
    flush.zero_()
    ops.gemm(xf, W_fd, b_rows_are_k=True)                                      # FF1 dgrad    [6144 x 2048 x 8192], B MN-major
torch.cuda.synchronize()
print("ok")
