"""Host logic of the LoRA-fused linears for every legal `config.lora_rank` (config.py:22 default 8,
train-avatars.yaml:36 sets 32; peft accepts any positive rank), on CPU.

ops.LinearFn / ops.CtxKVFn / lora.LoraLinear.merge run over the plain-torch stand-ins of the raw kernels
(tests/torch_kernels.py), which also enforce the argument contract of the C ABI (extents in multiples of 8, 16-byte
aligned pointers and pitches, whole tiles per batch group): a rank that the device library would refuse fails here.
Reference semantics: peft lora.Linear, y = W x + b + (alpha / r) B (A x) (training.py:61-68)."""
import pytest
import torch

import torch_kernels as tk

BF16 = torch.bfloat16
RANKS = [1, 8, 12, 32, 64, 72, 96, 128]


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def _randn(*shape, seed, scale=1.0, dtype=BF16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("r", RANKS)
def test_linear_fn_any_rank(r):
    from b200_ltx import ops
    M, K, N, s = 192, 128, 256, 0.5
    x = _randn(M, K, seed=1).requires_grad_(True)
    W = _randn(N, K, seed=2, scale=0.08)
    b = _randn(N, seed=3, scale=0.1)
    A = _randn(r, K, seed=4, scale=0.08, dtype=torch.float32).requires_grad_(True)
    B = _randn(N, r, seed=5, scale=0.08, dtype=torch.float32).requires_grad_(True)
    res = _randn(M, N, seed=6).requires_grad_(True)
    dy = _randn(M, N, seed=7)
    with tk.patched():
        y = ops.LinearFn.apply(x, W, b, A, B, s, None, 0, res)
        y.backward(dy)
    xr, Ar, Br, rr = [t.detach().float().requires_grad_(True) for t in (x, A, B, res)]
    yr = xr @ W.float().t() + b.float() + s * (xr @ Ar.t()) @ Br.t() + rr
    yr.backward(dy.float())
    assert _rel(y, yr) < 1e-2
    assert A.grad.shape == A.shape and B.grad.shape == B.shape and A.grad.dtype == torch.float32
    for got, want in ((x.grad, xr.grad), (A.grad, Ar.grad), (B.grad, Br.grad), (res.grad, rr.grad)):
        assert _rel(got, want) < 2e-2, (r, _rel(got, want))


@pytest.mark.parametrize("r", RANKS)
def test_ctx_kv_fn_any_rank(r):
    """attn2 keys / values of all blocks from one strided-batched projection (G = 2 blocks x (k | v))."""
    from b200_ltx import ops
    M, Dc, D, G, s = 128, 128, 256, 4, 2.0
    x = _randn(M, Dc, seed=1).requires_grad_(True)
    Wkv = _randn(G * D, Dc, seed=2, scale=0.08)
    bkv = _randn(G * D, seed=3, scale=0.1)
    As = [_randn(r, Dc, seed=10 + g, scale=0.08, dtype=torch.float32).requires_grad_(True) for g in range(G)]
    Bs = [_randn(D, r, seed=20 + g, scale=0.08, dtype=torch.float32).requires_grad_(True) for g in range(G)]
    dys = [_randn(M, D, seed=30 + g) for g in range(G)]
    with tk.patched():
        outs = ops.CtxKVFn.apply(x, Wkv, bkv, G, s, r, *As, *Bs)
        torch.autograd.backward(outs, dys)
    xr = x.detach().float().requires_grad_(True)
    Ar = [a.detach().clone().requires_grad_(True) for a in As]
    Br = [b.detach().clone().requires_grad_(True) for b in Bs]
    refs = [xr @ Wkv[g * D:(g + 1) * D].float().t() + bkv[g * D:(g + 1) * D].float() + s * (xr @ Ar[g].t()) @ Br[g].t()
            for g in range(G)]
    torch.autograd.backward(refs, [d.float() for d in dys])
    for g in range(G):
        assert _rel(outs[g], refs[g]) < 1e-2
        assert As[g].grad.shape == As[g].shape and Bs[g].grad.shape == Bs[g].shape
        assert _rel(As[g].grad, Ar[g].grad) < 2e-2 and _rel(Bs[g].grad, Br[g].grad) < 2e-2
    assert _rel(x.grad, xr.grad) < 2e-2


@pytest.mark.parametrize("r", [8, 12, 32])
def test_prestaged_adapters_match_per_call_staging(r):
    """ops.prestage_lora (all adapters of a model in a handful of launches) hands LinearFn the same padded operands as
    the per-call path, and the staged copies never outlive clear_lora_stage()."""
    from b200_ltx import ops
    A = _randn(r, 64, seed=1, dtype=torch.float32)
    B = _randn(48, r, seed=2, dtype=torch.float32)
    a0, b0 = ops.stage_lora(A, B, 0.5)
    ops.prestage_lora([(A, B, 0.5), (A.clone(), B.clone(), 1.0)])
    a1, b1 = ops.stage_lora(A, B, 0.5)
    assert a1.shape == (ops.lora_pad(r), 64) and b1.shape == (48, ops.lora_pad(r))
    assert torch.equal(a0, a1) and torch.equal(b0, b1)
    assert a1.data_ptr() != a0.data_ptr()            # came from the staged stack
    ops.clear_lora_stage()
    assert ops.stage_lora(A, B, 0.5)[0].data_ptr() != a1.data_ptr()
    with pytest.raises(Exception):
        ops.lora_pad(0)
