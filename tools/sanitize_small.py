"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck are slow: tiny sizes)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from b200_ltx import ops

BF16 = torch.bfloat16
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.3).to("cuda", BF16)
# GEMMs: single CTA tiles, CTA pair, second operand pair, batched, split-K fp32
x, w = r(640, 256), r(512, 256)
ops.gemm(x, w, bias=r(512))
ops.gemm(x, w, a2=r(640, 64), b2=r(512, 64), res=r(640, 512), gate=r(1, 512), rows_per_gate=640)
ops.gemm(r(256, 640), r(256, 512), a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32, split_k=0)
# every compiled epilogue variant of the TMA-store path, ragged M and N (block_n 128 / 256, CTA pair at M >= 512)
xr, wr = r(600, 192), r(456, 192)
pre = torch.empty(600, 456, device="cuda", dtype=BF16)
for bn in (128, 256):
    ops.gemm(xr, wr, block_n=bn)                                                     # plain
    ops.gemm(xr, wr, bias=r(456), block_n=bn)                                        # bias only
    ops.gemm(xr, wr, res=r(600, 456), block_n=bn)                                    # general, no bias / gate
    ops.gemm(xr, wr, bias=r(456), gate=r(2, 456), rows_per_gate=300, res=r(600, 456), block_n=bn)
    ops.gemm(xr, wr, bias=r(456), epilogue=ops.EPI_GELU, aux=pre, block_n=bn)
    ops.gemm(xr, wr, epilogue=ops.EPI_GELU_GRAD, aux=pre, block_n=bn)
    ops.gemm(xr, wr, bias=r(456), gate=r(2, 456), rows_per_gate=300, res=r(600, 456), epilogue=ops.EPI_STASH, aux=pre, block_n=bn)
out = torch.empty(256, 3 * 256, device="cuda", dtype=BF16)
ops.gemm_batched(r(256, 128), r(3 * 256, 128), out, 256, 256, 128, 3, {"b": (256, 0), "c": (0, 256), "bias": 256}, bias=r(768))
# attention fwd / bwd, ragged + bias + q-split
for (B, H, Nq, Nk, bias) in [(1, 2, 200, 200, False), (2, 2, 130, 70, True), (1, 4, 1100, 256, True)]:
    q, k, v, do = r(B * Nq, H * 64), r(B * Nk, H * 64), r(B * Nk, H * 64), r(B * Nq, H * 64)
    kb = None
    if bias:
        kb = torch.zeros(B, Nk, device="cuda")
        kb[:, Nk // 2:] = -10000.0
    o, lse = ops.fa_fwd(q, k, v, B, H, Nq, Nk, kb, 0.125)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    ops.fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, kb, 0.125)
# element-wise families
xx = r(200, 512)
y = ops.norm_mod_fwd(xx, r(1, 512), r(1, 512), 200, 1e-6)
ops.norm_mod_bwd(y, xx, r(1, 512), 200, 1e-6, dres=r(200, 512))
ops.rowscale(xx, r(1, 512), 200)
ops.colsum(xx)
ops.rf_loss(r(2, 96, 128), r(2, 96, 128))
ops.rf_noise(r(2, 96, 128), r(2, 96, 128), torch.tensor([0.3, 0.8]))
torch.cuda.synchronize()
print("sanitize_small ok")
