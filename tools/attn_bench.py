"""Micro-benchmark of the flash-attention kernels at the LTXV shapes (warm, CUDA events, L2 flushed), with the
library comparator on the same box and clocks: F.scaled_dot_product_attention (what the reference calls,
attention.py:1057) through every backend this torch build offers for the shape, forward and forward+backward."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200_ltx import ops

dev = "cuda"
H, D = 32, 2048
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=8):
    if os.environ.get("B200_ATTN_GRAPH"):      # small shapes: GPU-only time from a replayed graph of 6 launches
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(6):
                fn()
        g.replay()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 6)
        ts.sort()
        return ts[len(ts) // 2]
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


CASES = [("attn1 cfg2", 1, 6144, 6144, False, 32), ("attn1 cfg3", 4, 3328, 3328, False, 32),
         ("attn1 cfg5", 1, 12672, 12672, False, 32), ("attn2 cfg2 (15/256 keys)", 1, 6144, 256, True, 32),
         # sequence-parallel shapes of cfg5: head exchange (all tokens, 32/P heads) and K/V gather (N/P queries, all keys)
         ("attn1 cfg5 heads P=8", 1, 12672, 12672, False, 4), ("attn1 cfg5 heads P=4", 1, 12672, 12672, False, 8),
         ("attn1 cfg5 gather P=8", 1, 1584, 12672, False, 32)]
only = os.environ.get("B200_ATTN_CASES")
for (name, B, Nq, Nk, bias, H) in CASES:
    if only and not any(tok in name for tok in only.split(",")):
        continue
    D = H * 64
    g = torch.Generator(device="cpu").manual_seed(0)
    q = torch.randn(B * Nq, D, generator=g).to(dev, torch.bfloat16)
    k = torch.randn(B * Nk, D, generator=g).to(dev, torch.bfloat16)
    v = torch.randn(B * Nk, D, generator=g).to(dev, torch.bfloat16)
    do = torch.randn(B * Nq, D, generator=g).to(dev, torch.bfloat16)
    kb = None
    if bias:
        kb = torch.zeros(B, Nk, device=dev)
        kb[:, 15:] = -10000.0
    o, lse = ops.fa_fwd(q, k, v, B, H, Nq, Nk, kb, 0.125)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    delta = ops.attn_delta(o, do, B, H, Nq)
    dq = torch.zeros(B * Nq, D, device=dev)
    tf = timeit(lambda: ops.fa_fwd(q, k, v, B, H, Nq, Nk, kb, 0.125))
    tb = timeit(lambda: ops.fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, kb, 0.125, delta=delta, dq_accum=dq))
    fl = 4.0 * B * H * Nq * Nk * 64
    print(f"{name:28s} fwd {tf*1e3:8.1f} us {fl/tf/1e9:7.1f} TF/s   bwd {tb*1e3:8.1f} us {2*fl/tb/1e9:7.1f} TF/s", flush=True)
    if os.environ.get("B200_NO_SDPA"):
        continue
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    qh = q.view(B, Nq, H, 64).transpose(1, 2).detach().requires_grad_(True)
    kh = k.view(B, Nk, H, 64).transpose(1, 2).detach().requires_grad_(True)
    vh = v.view(B, Nk, H, 64).transpose(1, 2).detach().requires_grad_(True)
    doh = do.view(B, Nq, H, 64).transpose(1, 2)
    am = kb.to(torch.bfloat16)[:, None, None, :].expand(B, H, Nq, Nk) if kb is not None else None
    for bname, be in (("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
                      ("mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel([be]):
                def fwd():
                    with torch.no_grad():
                        return F.scaled_dot_product_attention(qh, kh, vh, attn_mask=am)

                def fwdbwd():
                    qh.grad = kh.grad = vh.grad = None
                    F.scaled_dot_product_attention(qh, kh, vh, attn_mask=am).backward(doh)
                t1 = timeit(fwd, 5)
                t2 = timeit(fwdbwd, 5)
            print(f"    torch SDPA {bname:14s} fwd {t1*1e3:8.1f} us {fl/t1/1e9:7.1f} TF/s   bwd (fwd+bwd - fwd) "
                  f"{(t2-t1)*1e3:8.1f} us {2*fl/max(t2-t1, 1e-6)/1e9:7.1f} TF/s", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"    torch SDPA {bname:14s} not available for this shape: {str(e).splitlines()[0][:90]}", flush=True)
