"""Micro-benchmark of the row / element-wise kernels at the cfg2 shapes (6144 rows x 2048).
Each timing brackets K launches on K different buffer sets (K x 100-300 MB, far more than the 126 MB L2), so the
inputs are cold and the ~2 us CUDA-event resolution is amortised.  GB/s = algorithmic bytes / time."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops

dev = "cuda"
M, D, H, K = 6144, 2048, 32, 6
BF = torch.bfloat16


def r(*s, dt=BF):
    return torch.randn(*s, device=dev).to(dt)


def timeit(fns, n=7):
    """The K launches are captured in a CUDA graph and replayed: the Python / ctypes call costs ~15 us, more than some
    of these kernels take."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / len(fns))
    ts.sort()
    return ts[len(ts) // 2]


sets = []
for _ in range(K):
    sets.append(dict(x=r(M, D), dy=r(M, D), dres=r(M, D), out=r(M, D), qkv=r(M, 3 * D), cos=r(M, D), sin=r(M, D),
                     qk=r(M, 2 * D), dq32=r(M, D, dt=torch.float32), dk=r(M, D), dqkv=r(M, 3 * D), o=r(M, D), do=r(M, D)))
scale, shift, gate, wq, wk = r(1, D), r(1, D), r(1, D), r(D), r(D)
cases = [
    ("norm_mod_fwd", lambda s: ops.norm_mod_fwd(s["x"], scale, shift, M, 1e-6, False, s["out"]), 4.0 * M * D),
    ("norm_mod_bwd (+dres)", lambda s: ops.norm_mod_bwd(s["dy"], s["x"], scale, M, 1e-6, False, dres=s["dres"]), 8.0 * M * D),
    ("qknorm_rope_fwd", lambda s: ops.qknorm_rope_fwd(s["qkv"][:, :D], s["qkv"][:, D:2 * D], wq, wk, s["cos"], s["sin"],
                                                      s["qk"][:, :D], s["qk"][:, D:]), 12.0 * M * D),
    ("qknorm_rope_bwd", lambda s: ops.qknorm_rope_bwd(s["dq32"], s["dk"], s["qkv"][:, :D], s["qkv"][:, D:2 * D], wq, wk,
                                                      s["cos"], s["sin"], s["dqkv"][:, :D], s["dqkv"][:, D:2 * D]),
     18.0 * M * D),
    ("rowscale", lambda s: ops.rowscale(s["dy"], gate, M), 4.0 * M * D),
    ("attn_delta", lambda s: ops.attn_delta(s["o"], s["do"], 1, H, M), 4.0 * M * D),
]
for name, fn, nbytes in cases:
    t = timeit([(lambda s=s: fn(s)) for s in sets])
    print(f"{name:24s} {t * 1e3:7.1f} us  {nbytes / t / 1e6:7.0f} GB/s", flush=True)
