"""Debug: per-event clock64 timeline of one fa_fwd CTA (library built with -DB200_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops, lib
B, H, N, D = 1, 32, 6144, 2048
g = torch.Generator(device="cpu").manual_seed(0)
q, k, v = (torch.randn(B * N, D, generator=g).to("cuda", torch.bfloat16) for _ in range(3))
for _ in range(2):
    o, lse = ops.fa_fwd(q, k, v, B, H, N, N, None, 0.125)
torch.cuda.synchronize()
L = lib.load()
buf = (ctypes.c_ulonglong * 8192)()
L.b200_debug_fwd_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", L.b200_debug_fwd_trace(buf, 8192))
t = list(buf)
base = min(x for x in t if x)
r = lambda x: x - base if x else -1
print("j | MMA: loop_top qk_issued pa_wake pb_wake | softmax w2: enter s_full_wake ld_done+s_free max_done pva_ok pa_arrived pv_ok pb_arrived")
for j in range(20, 32):
    print(j, [r(t[j * 4 + i]) for i in range(4)], [r(t[2048 + j * 8 + i]) for i in range(8)])
