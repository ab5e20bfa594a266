"""Micro-benchmark of the GEMM shapes of one LTXV-2B block at cfg2 (M = 6144), warm, CUDA events, L2 flushed; the last
column is cuBLASLt (torch.matmul / F.linear, no epilogue work beyond the bias) on the same operands, same box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
from b200_ltx import ops

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=6144, help="rows (tokens): 6144 = cfg2, 1584 = cfg5 on 8 sequence-parallel ranks")
ap.add_argument("--bn", default="128,256", help="block_n values to time (0 = the library's own choice); 256 means CTA "
                                                "pairs unless B200_GEMM_NO_PAIR=1")
ap.add_argument("--no-lib", action="store_true")
ap.add_argument("--graph", action="store_true", help="time 12 launches replayed from one CUDA graph instead of single "
                                                     "eager launches (small shapes: an eager launch is host-bound)")
args = ap.parse_args()
M, D, F = args.m, 2048, 8192
BNS = [int(v) for v in args.bn.split(",")]
print(f"M = {M}, pair kernel {'off' if os.environ.get('B200_GEMM_NO_PAIR') == '1' else 'on'}", flush=True)
dev = "cuda"
def r(*s): return (torch.randn(*s, device=dev) * 0.05).bfloat16()
x, xf = r(M, D), r(M, F)
W_dd, W_fd, W_df, W_3d = r(D, D), r(F, D), r(D, F), r(3 * D, D)
bias_d, bias_f, bias_3d = r(D), r(F), r(3 * D)
gate, res, pre = r(1, D), r(M, D), r(M, F)
big3 = r(M, 3 * D)
import torch.nn.functional as Fnn
lib_cases = {
    "fwd N=2048 K=2048 bias": lambda: Fnn.linear(x, W_dd, bias_d),
    "fwd N=2048 K=2048 bias+gate+res": lambda: Fnn.linear(x, W_dd, bias_d),
    "fwd N=6144 K=2048 bias (qkv)": lambda: Fnn.linear(x, W_3d, bias_3d),
    "fwd N=8192 K=2048 gelu+aux": lambda: Fnn.linear(x, W_fd, bias_f),
    "fwd N=2048 K=8192 gate+res": lambda: Fnn.linear(xf, W_df, bias_d),
    "dgrad N=2048 K=2048": lambda: torch.matmul(x, W_dd),
    "dgrad N=2048 K=6144 (qkv)": lambda: torch.matmul(big3, W_3d),
    "dgrad N=8192 K=2048 gelu'": lambda: torch.matmul(x, W_df),
    "dgrad N=2048 K=8192": lambda: torch.matmul(xf, W_fd),
}
cases = [
    ("fwd N=2048 K=2048 bias", lambda bn: ops.gemm(x, W_dd, bias=bias_d, block_n=bn), 2 * M * D * D),
    ("fwd N=2048 K=2048 bias+gate+res", lambda bn: ops.gemm(x, W_dd, bias=bias_d, gate=gate, rows_per_gate=M, res=res, block_n=bn), 2 * M * D * D),
    ("fwd N=6144 K=2048 bias (qkv)", lambda bn: ops.gemm(x, W_3d, bias=bias_3d, block_n=bn), 2 * M * 3 * D * D),
    ("fwd N=8192 K=2048 gelu+aux", lambda bn: ops.gemm(x, W_fd, bias=bias_f, epilogue=ops.EPI_GELU, aux=pre, block_n=bn), 2 * M * F * D),
    ("fwd N=2048 K=8192 gate+res", lambda bn: ops.gemm(xf, W_df, bias=bias_d, gate=gate, rows_per_gate=M, res=res, block_n=bn), 2 * M * F * D),
    ("dgrad N=2048 K=2048", lambda bn: ops.gemm(x, W_dd, b_rows_are_k=True, block_n=bn), 2 * M * D * D),
    ("dgrad N=2048 K=6144 (qkv)", lambda bn: ops.gemm(big3, W_3d, b_rows_are_k=True, block_n=bn), 2 * M * 3 * D * D),
    ("dgrad N=8192 K=2048 gelu'", lambda bn: ops.gemm(x, W_df, b_rows_are_k=True, epilogue=ops.EPI_GELU_GRAD, aux=pre, block_n=bn), 2 * M * F * D),
    ("dgrad N=2048 K=8192", lambda bn: ops.gemm(xf, W_fd, b_rows_are_k=True, block_n=bn), 2 * M * F * D),
]
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def graph_time(fn, reps=12):
    """GPU-only time per launch: `reps` launches captured in one CUDA graph and replayed (no host launch cost between
    them -- what a captured train step sees; the activation operand stays in L2 as it does behind its producer)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return ts[len(ts) // 2]


if args.graph:
    for name, fn, flops in cases:
        if not args.no_lib:
            t = graph_time(lib_cases[name])
            print(f"{name:36s} cuBLASLt {t*1e3:8.1f} us {flops / t / 1e9:8.1f} TF/s (graph-replayed, plain GEMM)", flush=True)
        for bn in BNS:
            t = graph_time(lambda: fn(bn))
            print(f"{name:36s} bn={bn:3d} {t*1e3:8.1f} us {flops / t / 1e9:8.1f} TF/s (graph-replayed)", flush=True)
    sys.exit(0)

for name, fn, flops in cases:
    lf = lib_cases[name]
    if not args.no_lib:
        for _ in range(3): lf()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lf(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"{name:36s} cuBLASLt {ts[len(ts) // 2]*1e3:8.1f} us {flops / ts[len(ts) // 2] / 1e9:8.1f} TF/s (plain GEMM, no fused epilogue)", flush=True)
    for bn in BNS:
        for _ in range(3): fn(bn)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(bn); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        t = ts[len(ts) // 2]
        print(f"{name:36s} bn={bn:3d} {t*1e3:8.1f} us {flops / t / 1e9:8.1f} TF/s", flush=True)
