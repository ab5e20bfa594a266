"""Debug: per-event clock64 timeline of one fa_bwd CTA (library built with -DB200_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops, lib
B, H, N, D = 1, 32, 6144, 2048
g = torch.Generator(device="cpu").manual_seed(0)
q, k, v, do = (torch.randn(B * N, D, generator=g).to("cuda", torch.bfloat16) for _ in range(4))
for _ in range(2):
    o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    dq = ops.fa_bwd(q, k, v, o, do, lse, B, H, N, N, dk, dv, None, 0.125)
torch.cuda.synchronize()
L = lib.load()
buf = (ctypes.c_ulonglong * 8192)()
L.b200_debug_bwd_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", L.b200_debug_bwd_trace(buf, 8192))
t = list(buf)
base = min(x for x in t if x)
r = lambda x: x - base if x else -1
print("MMA warp: g  wake  issued_dV+scores  end_step | compute w2: enter wake ld_done arrived | drain: wake freed tma")
for g_ in range(20, 36):
    m = [r(t[g_ * 4 + j]) for j in range(3)]
    c = [r(t[2048 + g_ * 4 + j]) for j in range(4)]
    d = [r(t[4096 + (g_ // 2) * 4 + j]) for j in range(3)] if g_ % 2 else []
    print(g_, m, c, d)
