import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops
B, H, N, D = 1, 32, 6144, 2048
g = torch.Generator(device="cpu").manual_seed(0)
q, k, v, do = (torch.randn(B * N, D, generator=g).to("cuda", torch.bfloat16) for _ in range(4))
o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
dk, dv = torch.empty_like(k), torch.empty_like(v)
delta = ops.attn_delta(o, do, B, H, N)
dq = torch.zeros(B * N, D, device="cuda")
f = lambda: ops.fa_bwd(q, k, v, o, do, lse, B, H, N, N, dk, dv, None, 0.125, delta=delta, dq_accum=dq)
for _ in range(3): f()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(os.environ.get("B200LTX_LIB", "default"), f"bwd {ts[5]*1e3:.1f} us  {8.0*B*H*N*N*64/ts[5]/1e9:.1f} TF/s  (min {ts[0]*1e3:.1f})")
