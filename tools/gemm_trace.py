"""Debug: per-CTA phase timeline (globaltimer) of one GEMM launch; library built with -DB200_TRACE
(python video-generation-for-human-avatars_b200/build.py variant trace B200_TRACE  +  B200LTX_LIB=...libb200ltx_trace.so).
Phases: 0 entry, 1 prologue done, 2 first operands of the last tile landed, 3 last MMA committed, 4 epilogue sees the
accumulator, 5 epilogue stores issued, 6 staging read back, 7 exit."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops, lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1584
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
a = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = None if "nobias" in sys.argv else torch.zeros(N, device="cuda", dtype=torch.bfloat16)
hot = "hot" in sys.argv   # do not flush L2 before the traced launch
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
kw = dict(bias=bias)
if "gateres" in sys.argv:   # the out-projection epilogue: gate * (acc + bias) + residual
    kw.update(gate=(torch.randn(1, N, device="cuda") * 0.1).bfloat16(), rows_per_gate=M,
              res=(torch.randn(M, N, device="cuda") * 0.1).bfloat16())
for _ in range(3):
    ops.gemm(a, w, **kw)
if not hot:
    flush.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(); ops.gemm(a, w, **kw); e1.record()
torch.cuda.synchronize()
L = lib.load()
buf = (ctypes.c_ulonglong * (160 * 32))()
L.b200_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", L.b200_debug_gemm_trace(buf, 160 * 32), f"event time {e0.elapsed_time(e1) * 1e3:.1f} us (eager, cold L2)")
t = [list(buf[i * 32:(i + 1) * 32]) for i in range(160)]
t = [r for r in t if r[0]]
base = min(r[0] for r in t)
print(f"{len(t)} CTAs; entry spread {max(r[0] for r in t) - base} ns; last exit at {max(r[7] for r in t) - base} ns")
names = ["entry", "prologue", "1st operands", "MMA done", "epi sees acc", "stores issued", "staging read", "exit",
         "slab0 start", "slab0 tmem", "slab0 packed", "slab0 stored", "slab1 start", "slab1 tmem", "slab1 packed", "slab1 stored"]
for i in (0, 1, len(t) // 2, len(t) - 1):
    print(f"cta {i:3d}: " + "  ".join(f"{n} {r - base}" for n, r in zip(names, t[i])))
import statistics
for j in range(1, 8):
    d = [r[j] - r[j - 1] for r in t if r[j] and r[j - 1]]
    if d:
        print(f"phase {names[j-1]} -> {names[j]}: median {statistics.median(d)} ns, max {max(d)} ns")

print("epilogue warp 4 of cta 0 (ns after it saw the accumulator):",
      "  ".join(f"{n} {t[0][8 + k] - t[0][4]}" for k, n in enumerate(names[8:])))

print("  inside the slabs:", "  ".join(f"{n} {t[0][16 + k] - t[0][4]}" for k, n in enumerate(
    ["s0 hc0 tmem", "s0 hc0 math+sts", "s0 hc1 tmem", "s0 hc1 math+sts", "s1 hc0 tmem", "s1 hc0 math+sts", "s1 hc1 tmem", "s1 hc1 math+sts"])))
