import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops
B, H, N, D = 1, 32, 6144, 2048
g = torch.Generator(device="cpu").manual_seed(0)
q, k, v = (torch.randn(B * N, D, generator=g).to("cuda", torch.bfloat16) for _ in range(3))
f = lambda: ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
o, lse = f()
# accuracy vs fp32 torch on one head slice
qh, kh, vh = q[:1024, :64].float(), k[:, :64].float(), v[:, :64].float()
ref = torch.softmax(qh @ kh.T * 0.125, -1) @ vh
err = float((o[:1024, :64].float() - ref).norm() / ref.norm())
for _ in range(3): f()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(os.path.basename(os.environ.get("B200LTX_LIB", "default")), f"fwd {ts[5]*1e3:.1f} us  {4.0*B*H*N*N*64/ts[5]/1e9:.1f} TF/s  rel err {err:.2e}")
