"""Tensor-level wrappers over the C ABI (lib.py) and the autograd Functions built from them.

Every function here launches hand-written sm_100a kernels from libb200ltx.so on the current CUDA
stream.  PyTorch supplies device memory, streams and the autograd graph only."""
import os
from typing import Optional, Tuple

import torch

from . import lib as _lib

EPI_NONE, EPI_GELU, EPI_GELU_GRAD, EPI_STASH = 0, 1, 2, 3
LORA_PAD = 64  # LoRA ranks are zero-padded to whole 64-wide k-blocks of the GEMM


def lora_pad(r: int) -> int:
    """Padded rank: the adapters enter the GEMMs as whole 64-wide k blocks (config.lora_rank is free, config.py:22)."""
    if r < 1:
        raise _lib.B200Error(f"LoRA rank {r}")
    return (r + LORA_PAD - 1) // LORA_PAD * LORA_PAD


BF16 = torch.bfloat16

# counts kernels launched through the C ABI (bench.py reports it as gpu_launches)
launch_count = 0


def _L():
    return _lib.load()


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk2d(t: torch.Tensor, name: str, dtype=BF16):
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != dtype or not t.is_cuda:
        raise _lib.B200Error(f"{name}: expected a CUDA {dtype} matrix with unit inner stride, got "
                             f"{tuple(t.shape)} {t.dtype} strides {t.stride()} on {t.device}")


def _count(n=1):
    global launch_count
    launch_count += n


class KernelTimer:
    """CUDA-event timing of kernel families on the launching stream (bench.py roofline numbers).
    `families=None` records everything; otherwise only the named families."""

    def __init__(self, families=None, every=1):
        self.families = set(families) if families else None
        self.every = max(1, every)  # time one launch in `every` (event records cost ~2 us each)
        self._seen = 0
        self.records = []  # (family, work, unit, start_event, end_event, detail)

    def want(self, family):
        if self.families is not None and family not in self.families:
            return False
        self._seen += 1
        return self._seen % self.every == 0

    def summary(self, by_kernel=False):
        """Totals per family; `by_kernel=True`: per (family, shape) -- one entry per distinct kernel launch shape, so
        that a family of many different GEMM shapes is not averaged into one "kernel"."""
        torch.cuda.synchronize()
        out = {}
        for fam, work, unit, e0, e1, detail in self.records:
            key = fam if not by_kernel or not detail else f"{fam} {detail}"
            d = out.setdefault(key, {"ms": 0.0, "work": 0.0, "launches": 0, "unit": unit, "family": fam})
            d["ms"] += e0.elapsed_time(e1)
            d["work"] += work
            d["launches"] += 1
        return out


timer: Optional[KernelTimer] = None


def _call(family, work, unit, cfn, *args, launches=1, detail=None):
    """Invoke one C-ABI entry point; optionally bracket it with CUDA events on the current stream."""
    t = timer
    if t is not None and t.want(family):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = cfn(*args)
        e1.record()
        t.records.append((family, work, unit, e0, e1, detail))
    else:
        rc = cfn(*args)
    _lib.check(rc, cfn.__name__)
    _count(launches)


# ---------------------------------------------------------------------------------------------
# raw ops
# ---------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_rows_are_k=False, b_rows_are_k=False,
         a2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, out_dtype=BF16, bias=None, gate=None, rows_per_gate=0,
         res=None, aux=None, epilogue=EPI_NONE, block_n=0, split_k=1) -> torch.Tensor:
    """out[M,N] = epi(a @ b^T (+ a2 @ b2^T)); see b200_gemm_bf16 in include/b200ltx.h."""
    _chk2d(a, "gemm a")
    _chk2d(b, "gemm b")
    M, K = (a.shape[1], a.shape[0]) if a_rows_are_k else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_rows_are_k else b.shape
    if K != Kb:
        raise _lib.B200Error(f"gemm: reduction mismatch {K} vs {Kb}")
    K2 = 0
    if a2 is not None:
        _chk2d(a2, "gemm a2")
        _chk2d(b2, "gemm b2")
        M2, K2 = (a2.shape[1], a2.shape[0]) if a_rows_are_k else a2.shape
        N2, K2b = (b2.shape[1], b2.shape[0]) if b_rows_are_k else b2.shape
        if (M2, N2, K2) != (M, N, K2b):
            raise _lib.B200Error("gemm: second operand pair shape mismatch")
    if split_k == 0:  # auto: split skinny fp32 reductions (LoRA / caption wgrad) over idle SMs
        split_k = 1
        plain = out_dtype == torch.float32 and out is None and bias is None and gate is None and res is None \
            and aux is None and epilogue == EPI_NONE
        bn = block_n or (64 if N <= 64 else 128)
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
        kb = (K + 63) // 64 + (K2 + 63) // 64
        if plain and tiles <= 48 and kb >= 16:
            split_k = max(1, min(148 // tiles, kb // 4))
    if out is None:
        out = (torch.zeros if split_k > 1 else torch.empty)((M, N), device=a.device, dtype=out_dtype)
    _chk2d(out, "gemm out", out.dtype)
    if out.dtype not in (BF16, torch.float32) or tuple(out.shape) != (M, N):
        raise _lib.B200Error("gemm: bad output tensor")
    for t, nm in ((res, "res"), (aux, "aux")):
        if t is not None:
            _chk2d(t, "gemm " + nm)
            if tuple(t.shape) != (M, N):
                raise _lib.B200Error(f"gemm: {nm} shape mismatch")
    gate_stride = 0
    if gate is not None:
        _chk2d(gate, "gemm gate")
        gate_stride = gate.stride(0)
        if gate.shape[1] != N or rows_per_gate <= 0 or gate.shape[0] * rows_per_gate < M:
            raise _lib.B200Error("gemm: gate shape / rows_per_gate mismatch")
    if bias is not None and (bias.dtype != BF16 or bias.numel() != N or not bias.is_contiguous()):
        raise _lib.B200Error("gemm: bias must be a contiguous bf16 vector of length N")
    # Stream-K for the last partial wave is built and parity-tested but OFF by default: on this power-capped part the
    # idle SMs of a partial wave hand their power budget to the busy ones, so evening out the wave measured within
    # noise at K = 8192 and slower at K = 2048 ("what measurements changed").  B200_GEMM_STREAMK=1 enables.
    ws = _gemm_workspace(a.device) if (M * N >= (1 << 22) and os.environ.get("B200_GEMM_STREAMK") == "1") else None
    _call("gemm", 2.0 * M * N * (K + K2), "flop", _L().b200_gemm_bf16_ws,
        _p(a), a.stride(0), int(a_rows_are_k), _p(b), b.stride(0), int(b_rows_are_k),
        _p(a2), a2.stride(0) if a2 is not None else 0, _p(b2), b2.stride(0) if b2 is not None else 0, K2,
        _p(out), out.stride(0), int(out.dtype == torch.float32), M, N, K, epilogue,
        _p(bias), _p(gate), gate_stride, rows_per_gate, _p(res), res.stride(0) if res is not None else 0,
        _p(aux), aux.stride(0) if aux is not None else 0, block_n, split_k, _p(ws), ws.numel() if ws is not None else 0,
        _s(), detail=f"{M}x{N}x{K + K2}" + ("t" if b_rows_are_k else "") + ("T" if a_rows_are_k else ""))
    return out


_gemm_ws = {}


def _gemm_workspace(device):
    """Stream-K workspace of the current stream (b200_gemm_bf16_ws: one per stream, flags zeroed once)."""
    key = (device.index, _s())
    ws = _gemm_ws.get(key)
    if ws is None:
        n = _L().b200_gemm_workspace_bytes()
        ws = torch.empty(n, device=device, dtype=torch.uint8)
        ws[:16384].zero_()
        _gemm_ws[key] = ws
    return ws


def gemm_batched(a, b, out, M, N, K, groups, offs, *, a_rows_are_k=False, b_rows_are_k=False, a2=None, b2=None,
                 K2=0, bias=None, block_n=0):
    """`groups` GEMMs of shape [M,N,K(+K2)] in one launch; operand g = sub-block displaced by g * offs[name]
    (rows, cols) inside the given tensor (offs keys: a, b, a2, b2, c -> (rows, cols); bias -> elements).
    See b200_gemm_bf16_batched in include/b200ltx.h."""
    import ctypes
    for t, nm in ((a, "a"), (b, "b")):
        _chk2d(t, "gemm_batched " + nm)
    _chk2d(out, "gemm_batched out", out.dtype)
    if a2 is not None:
        _chk2d(a2, "gemm_batched a2")
        _chk2d(b2, "gemm_batched b2")
    z = (0, 0)
    o = [*offs.get("a", z), *offs.get("b", z), *offs.get("a2", z), *offs.get("b2", z), *offs.get("c", z),
         offs.get("bias", 0)]
    arr = (ctypes.c_int32 * 11)(*[int(x) for x in o])
    _call("gemm", 2.0 * groups * M * N * (K + K2), "flop", _L().b200_gemm_bf16_batched,
          _p(a), a.stride(0), int(a_rows_are_k), _p(b), b.stride(0), int(b_rows_are_k),
          _p(a2), a2.stride(0) if a2 is not None else 0, _p(b2), b2.stride(0) if b2 is not None else 0, K2,
          _p(out), out.stride(0), int(out.dtype == torch.float32), M, N, K, _p(bias), block_n, groups,
          ctypes.cast(arr, ctypes.c_void_p), _s())
    return out


def norm_mod_fwd(x, scale, shift, rows_per_mod, eps, layernorm=False, out=None):
    _chk2d(x, "norm_mod x")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), device=x.device, dtype=BF16)
    mod_stride = 0
    for t in (scale, shift):
        if t is not None:
            _chk2d(t, "norm_mod scale/shift")
            mod_stride = t.stride(0)
    if scale is not None and shift is not None and scale.stride(0) != shift.stride(0):
        raise _lib.B200Error("norm_mod: scale and shift must share a row stride")
    _call("norm_mod_fwd", 4.0 * rows * D, "byte", _L().b200_norm_mod_fwd, _p(x), x.stride(0), _p(out), out.stride(0), _p(scale), _p(shift),
                                mod_stride, rows, D, rows_per_mod, eps, int(layernorm), _s())
    return out


def norm_mod_bwd(dy, x, scale, rows_per_mod, eps, layernorm=False, dres=None, want_prod=False):
    """-> dx, or (dx, prod) with prod = dy * xhat (bf16 [rows, D]) when the AdaLN scale trains."""
    _chk2d(dy, "norm_mod_bwd dy")
    _chk2d(x, "norm_mod_bwd x")
    rows, D = x.shape
    dx = torch.empty((rows, D), device=x.device, dtype=BF16)
    prod = torch.empty((rows, D), device=x.device, dtype=BF16) if want_prod else None
    if dres is not None:
        _chk2d(dres, "norm_mod_bwd dres")
    _call("norm_mod_bwd", (8.0 if dres is not None else 6.0) * rows * D, "byte", _L().b200_norm_mod_bwd, _p(dy), dy.stride(0), _p(x), x.stride(0), _p(scale),
                                scale.stride(0) if scale is not None else 0, _p(dres),
                                dres.stride(0) if dres is not None else 0, _p(dx), dx.stride(0), _p(prod),
                                prod.stride(0) if prod is not None else 0, rows, D,
                                rows_per_mod, eps, int(layernorm), _s())
    return (dx, prod) if want_prod else dx


def colsum_groups(a, b=None, rows_per_group=0):
    """out[g, n] = sum over the rows of group g of a[r, n] * (b[r, n] if b is given else 1): fp32 [groups, N]."""
    _chk2d(a, "colsum_groups a")
    rows, N = a.shape
    rpg = rows_per_group or rows
    if b is not None:
        _chk2d(b, "colsum_groups b")
    out = torch.empty((max(rows // max(rpg, 1), 0), N), device=a.device, dtype=torch.float32)
    if rows == 0:
        return out
    _call("colsum_groups", (2.0 + (2.0 if b is not None else 0.0)) * rows * N, "byte", _L().b200_colsum_groups, _p(a),
          a.stride(0), _p(b), b.stride(0) if b is not None else 0, _p(out), rows, N, rpg, None, 0, _s())
    return out


def qknorm_rope_fwd(xq, xk, wq, wk, cos, sin, oq, ok, eps=1e-5):
    rows_q = xq.shape[0] if xq is not None else 0
    rows_k = xk.shape[0] if xk is not None else 0
    D = (xq if xq is not None else xk).shape[1]
    _call("qknorm_rope_fwd", (4.0 + (4.0 if cos is not None else 0.0)) * (rows_q + rows_k) * D, "byte", _L().b200_qknorm_rope_fwd,
        _p(xq), xq.stride(0) if xq is not None else 0, _p(xk), xk.stride(0) if xk is not None else 0,
        _p(wq), _p(wk), _p(cos), _p(sin), cos.stride(0) if cos is not None else 0,
        _p(oq), oq.stride(0) if oq is not None else 0, _p(ok), ok.stride(0) if ok is not None else 0,
        rows_q, rows_k, D, eps, _s())


def qknorm_rope_bwd(dq, dk, xq, xk, wq, wk, cos, sin, oq, ok, eps=1e-5, prod_q=None, prod_k=None):
    """prod_q / prod_k: optional bf16 outputs (RoPE^T dq) * xhat_q, (RoPE^T dk) * xhat_k (qk-norm weight gradients)."""
    rows_q = xq.shape[0] if xq is not None else 0
    rows_k = xk.shape[0] if xk is not None else 0
    D = (xq if xq is not None else xk).shape[1]
    _call("qknorm_rope_bwd", 10.0 * (rows_q + rows_k) * D, "byte", _L().b200_qknorm_rope_bwd,
        _p(dq), dq.stride(0) if dq is not None else 0, int(dq is not None and dq.dtype == torch.float32),
        _p(dk), dk.stride(0) if dk is not None else 0, int(dk is not None and dk.dtype == torch.float32),
        _p(xq), xq.stride(0) if xq is not None else 0, _p(xk), xk.stride(0) if xk is not None else 0,
        _p(wq), _p(wk), _p(cos), _p(sin), cos.stride(0) if cos is not None else 0,
        _p(oq), oq.stride(0) if oq is not None else 0, _p(ok), ok.stride(0) if ok is not None else 0,
        _p(prod_q), prod_q.stride(0) if prod_q is not None else 0, _p(prod_k),
        prod_k.stride(0) if prod_k is not None else 0, rows_q, rows_k, D, eps, _s())


def _fa_family(base, Nq, Nk, key_bias, attn1):
    """attn1 (latent self-attention, also its sequence-sharded form where Nq is the local shard and Nk the shard or all
    of the clip) vs attn2 (cross-attention to the caption tokens: a key bias, or a few hundred keys)."""
    if attn1 is None:
        attn1 = key_bias is None and Nk >= 1024
    return base if attn1 else base + "_attn2"


def fa_fwd(q, k, v, B, H, Nq, Nk, key_bias=None, scale=0.125, need_lse=True, attn1=None, batch_keep=None,
           pass_src=None):
    """q [B*Nq, >=H*64], k/v [B*Nk, >=H*64] (row-strided views allowed) -> o [B*Nq, H*64], lse.
    batch_keep: optional fp32 [B] of 0 / 1; entries with 0 get rows of `pass_src` [B*Nq, >=H*64] (default: the value
    rows) as output and cost no attention work (the STG skips of attention.py:1071-1086)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _chk2d(t, "fa_fwd " + nm)
    o = torch.empty((B * Nq, H * 64), device=q.device, dtype=BF16)
    lse = torch.empty((B, H, Nq), device=q.device, dtype=torch.float32) if need_lse else None
    if key_bias is not None and (key_bias.dtype != torch.float32 or tuple(key_bias.shape) != (B, Nk)
                                 or not key_bias.is_contiguous()):
        raise _lib.B200Error("fa_fwd: key_bias must be contiguous fp32 [B, Nk]")
    if batch_keep is not None and (batch_keep.dtype != torch.float32 or batch_keep.numel() != B
                                   or not batch_keep.is_contiguous() or (pass_src is None and Nq != Nk)):
        raise _lib.B200Error("fa_fwd: batch_keep must be contiguous fp32 [B] (and Nq == Nk unless pass_src is given)")
    if pass_src is not None:
        _chk2d(pass_src, "fa_fwd pass_src")
    ws_bytes = _L().b200_fa_fwd_workspace_bytes(B, H, Nq, Nk)   # > 0: the last wave of CTAs is split along the keys
    ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8) if ws_bytes else None
    _call(_fa_family("fa_fwd", Nq, Nk, key_bias, attn1), 4.0 * B * H * Nq * Nk * 64, "flop", _L().b200_fa_fwd_ws,
          _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0),
          _p(lse), _p(key_bias), _p(batch_keep), _p(pass_src), pass_src.stride(0) if pass_src is not None else 0,
          B, H, Nq, Nk, 64, scale, _p(ws), ws_bytes, _s(),
          launches=2 if ws_bytes else 1,
          detail=f"{B}x{H}x{Nq}x{Nk}")
    return o, lse


def attn_delta(o, do, B, H, Nq, zero=None):
    """delta [B, H, Nq] fp32; `zero` (fp32 [B*Nq, >= H*64]): the dQ accumulator of the backward, cleared in the same pass."""
    delta = torch.empty((B, H, Nq), device=o.device, dtype=torch.float32)
    if zero is not None and (zero.dtype != torch.float32 or zero.stride(1) != 1 or zero.shape[0] != B * Nq):
        raise _lib.B200Error("attn_delta: the accumulator to clear must be fp32 [B*Nq, >= H*64], unit column stride")
    _call("attn_delta", (4.0 + (4.0 if zero is not None else 0.0)) * B * Nq * H * 64, "byte", _L().b200_attn_delta_zero,
          _p(o), o.stride(0), _p(do), do.stride(0), _p(delta), _p(zero), zero.stride(0) if zero is not None else 0,
          B, H, Nq, _s())
    return delta


def attn_merge(o_acc, lse_acc, o_i, lse_i, B, H, N, first, out=None):
    """(o_acc, lse_acc) <- online-softmax merge with the partial result (o_i, lse_i); see b200_attn_merge."""
    _call("attn_merge", (4.0 + 2.0 + 4.0) * B * N * H * 64, "byte", _L().b200_attn_merge, _p(o_acc), o_acc.stride(0),
          _p(lse_acc), _p(o_i), o_i.stride(0), _p(lse_i), _p(out), out.stride(0) if out is not None else 0, B, H, N,
          int(first), _s())


def fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, key_bias=None, scale=0.125, delta=None, dq_accum=None,
           attn1=None):
    """Returns dq as fp32 [B*Nq, H*64]; writes bf16 dk/dv into the given (possibly strided) views.
    `delta` / `dq_accum` may be supplied to accumulate dq over several key shards (ring attention)."""
    _chk2d(do, "fa_bwd do")
    dq = dq_accum
    if delta is None:
        if dq is None:   # the pre-pass clears the accumulator on its way: no separate fill launch
            dq = torch.empty((B * Nq, H * 64), device=q.device, dtype=torch.float32)
            delta = attn_delta(o, do, B, H, Nq, zero=dq)
        else:
            delta = attn_delta(o, do, B, H, Nq)
    elif dq is None:
        dq = torch.zeros((B * Nq, H * 64), device=q.device, dtype=torch.float32)
    ws_bytes = _L().b200_fa_bwd_workspace_bytes(B, H, Nq, Nk)
    ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8) if ws_bytes else None
    _call(_fa_family("fa_bwd", Nq, Nk, key_bias, attn1), 8.0 * B * H * Nq * Nk * 64, "flop", _L().b200_fa_bwd,
          _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(do), do.stride(0),
          _p(lse), _p(delta), _p(key_bias), _p(dq), dq.stride(0), _p(dk), dk.stride(0),
          _p(dv), dv.stride(0), B, H, Nq, Nk, 64, scale, _p(ws), ws_bytes, _s(),
          launches=2 if ws_bytes else 1, detail=f"{B}x{H}x{Nq}x{Nk}")
    return dq


def rf_noise(x0, noise, t, want_xt=True, want_v=True):
    if x0.dtype != BF16 or noise.dtype != BF16 or not x0.is_contiguous() or not noise.is_contiguous():
        raise _lib.B200Error("rf_noise: contiguous bf16 tensors required")
    t = t.to(device=x0.device, dtype=torch.float32).contiguous()
    xt = torch.empty_like(x0) if want_xt else None
    v = torch.empty_like(x0) if want_v else None
    B = x0.shape[0]
    _call("rf_noise", 2.0 * x0.numel() * (2 + int(want_xt) + int(want_v)), "byte", _L().b200_rf_noise, _p(x0), _p(noise), _p(t), _p(xt), _p(v), B, x0.numel() // max(B, 1), _s())
    return xt, v


def rf_loss(out, target, grad_scale=1.0, want_grad=True):
    if out.dtype != BF16 or target.dtype != BF16 or not out.is_contiguous() or not target.is_contiguous():
        raise _lib.B200Error("rf_loss: contiguous bf16 tensors required")
    dout = torch.empty_like(out) if want_grad else None
    loss = torch.empty((), device=out.device, dtype=torch.float32)
    nbytes = _L().b200_rf_loss_workspace_bytes()
    ws = torch.empty(nbytes, device=out.device, dtype=torch.uint8)
    _call("rf_loss", 2.0 * out.numel() * (2 + int(want_grad)), "byte", _L().b200_rf_loss, _p(out), _p(target), _p(dout), _p(loss), out.numel(), grad_scale, _p(ws), nbytes, _s(), launches=2)
    return loss, dout


def lerp_condition_(tokens, ref, pose, w_ref=0.85, w_pose=0.5, token_offset=0):
    """In-place on tokens [B,N,C] (contiguous bf16); ref [B,C,1,H,W], pose [B,C,F,H,W].  `tokens` may be
    the contiguous shard [token_offset, token_offset+N) of the F*H*W tokens (sequence sharding)."""
    B, N, C = tokens.shape
    HW = ref.shape[3] * ref.shape[4]
    N_total = pose.shape[2] * HW
    for t in (tokens, ref, pose):
        if t.dtype != BF16 or not t.is_contiguous():
            raise _lib.B200Error("lerp_condition: contiguous bf16 tensors required")
    if token_offset + N > N_total or ref.shape[2] != 1:
        raise _lib.B200Error("lerp_condition: pose/ref shapes do not match the token count")
    _call("lerp_condition", 6.0 * tokens.numel(), "byte", _L().b200_lerp_condition, _p(tokens), _p(ref), _p(pose), B, N, C, HW, w_ref, w_pose, token_offset, N_total, _s())
    return tokens


def guidance_step_(v, x, x_next, dt, noise_level, scalars, has_cfg, has_stg, cfg_star=False, rescale=False,
                   workspace=None):
    """One sampling step's element-wise tail, in place on the fp32 latents `x` [B,N,C]; see b200_guidance_step.
    v: [conds*B, N, C] bf16 model output; x_next: None or [n_next*B, N, C] bf16 (next model input, one copy per
    condition); dt: fp32 [1] or [N]; noise_level: None or fp32 [B,N]; scalars: CUDA fp32 [4]."""
    B, N, C = x.shape
    conds = 1 + int(bool(has_cfg)) + int(bool(has_stg))
    if (v.dtype != BF16 or x.dtype != torch.float32 or not v.is_contiguous() or not x.is_contiguous()
            or tuple(v.shape) != (conds * B, N, C)):
        raise _lib.B200Error(f"guidance_step: v must be contiguous bf16 [{conds * B},{N},{C}], x contiguous fp32 [B,N,C]")
    n_next = 0
    if x_next is not None:
        if x_next.dtype != BF16 or not x_next.is_contiguous() or x_next.shape[1:] != x.shape[1:] or x_next.shape[0] % B:
            raise _lib.B200Error("guidance_step: x_next must be contiguous bf16 [n_next*B, N, C]")
        n_next = x_next.shape[0] // B
    if dt.dtype != torch.float32 or not dt.is_contiguous() or dt.numel() not in (1, N):
        raise _lib.B200Error("guidance_step: dt must be fp32 with 1 or N elements")
    if noise_level is not None and (noise_level.dtype != torch.float32 or not noise_level.is_contiguous()
                                    or tuple(noise_level.shape) != (B, N)):
        raise _lib.B200Error("guidance_step: noise_level must be contiguous fp32 [B,N]")
    if scalars.dtype != torch.float32 or not scalars.is_cuda or scalars.numel() < 4 or not scalars.is_contiguous():
        raise _lib.B200Error("guidance_step: scalars must be a CUDA fp32 tensor {guidance, stg, rescale, t}")
    need = _L().b200_guidance_step_workspace_bytes(B)
    if workspace is None and ((has_cfg and cfg_star) or rescale):
        workspace = torch.empty(need, device=x.device, dtype=torch.uint8)
    n_k = 1 + int(bool(has_cfg and cfg_star)) + int(bool(rescale))
    _call("guidance_step", (2.0 * conds * n_k + 8.0 + 2.0 * n_next) * x.numel(), "byte", _L().b200_guidance_step,
          _p(v), _p(x), _p(x_next) if x_next is not None else None, n_next, _p(dt), int(dt.numel() == N and N > 1),
          _p(noise_level) if noise_level is not None else None, _p(scalars), B, N, C, int(bool(has_cfg)),
          int(bool(has_stg)), int(bool(cfg_star)), int(bool(rescale)), _p(workspace) if workspace is not None else None,
          workspace.numel() if workspace is not None else 0, _s(), launches=n_k)
    return x


def rowscale(x, g, rows_per_mod):
    _chk2d(x, "rowscale x")
    _chk2d(g, "rowscale g")
    out = torch.empty((x.shape[0], x.shape[1]), device=x.device, dtype=BF16)
    _call("rowscale", 4.0 * x.numel(), "byte", _L().b200_rowscale, _p(x), x.stride(0), _p(g), g.stride(0), _p(out), out.stride(0), x.shape[0],
                            x.shape[1], rows_per_mod, _s())
    return out


def colsum(x):
    _chk2d(x, "colsum x")
    out = torch.empty((x.shape[1],), device=x.device, dtype=torch.float32)
    _call("colsum", 2.0 * x.numel(), "byte", _L().b200_colsum, _p(x), x.stride(0), _p(out), x.shape[0], x.shape[1], _s())
    return out


# ---------------------------------------------------------------------------------------------
# LoRA staging: fp32 peft adapters -> zero-padded bf16 GEMM operands
# ---------------------------------------------------------------------------------------------
# id(A) -> (a_pad, b_pad) staged once per forward for every adapter of the model (prestage_lora)
_lora_stage_cache = {}


def prestage_lora(adapters) -> None:
    """Stage all LoRA adapters of a model with a handful of launches instead of ~5 per adapter.

    `adapters`: list of (A [r,K] fp32, B [N,r] fp32, scaling).  Adapters of equal shape are stacked,
    cast to bf16 and zero-padded to a multiple of 64 together; LinearFn.forward picks its views from the cache."""
    _lora_stage_cache.clear()
    groups = {}
    for A, B, s in adapters:
        groups.setdefault((tuple(A.shape), tuple(B.shape), A.device), []).append((A, B, s))
    for (ashape, bshape, dev), items in groups.items():
        r, K = ashape
        N = bshape[0]
        pad = lora_pad(r)
        n = len(items)
        a_all = torch.zeros((n, pad, K), device=dev, dtype=BF16)
        b_all = torch.zeros((n, N, pad), device=dev, dtype=BF16)
        a_all[:, :r] = torch.stack([A.detach() for A, _, _ in items])
        bs = torch.stack([B.detach() for _, B, _ in items])
        scal = [float(s) for _, _, s in items]
        if any(x != 1.0 for x in scal):
            bs = torch.stack([b * x for b, x in zip(bs.unbind(0), scal)])  # python scalars: graph-capturable
        b_all[:, :, :r] = bs
        for i, (A, _, _) in enumerate(items):
            _lora_stage_cache[id(A)] = (a_all[i], b_all[i], A._version)


def clear_lora_stage() -> None:
    """Drop the staged copies (modules.transformer_forward calls this when a forward ends): FusedAdamW updates the
    adapters through raw pointers, which does not move `_version`, so a staged copy must never outlive its forward."""
    _lora_stage_cache.clear()


def stage_lora(A: torch.Tensor, B: torch.Tensor, scaling: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """A [r, K] -> A_pad [pad, K] ; B [N, r] -> (scaling * B)_pad [N, pad], both bf16, pad = lora_pad(r)."""
    hit = _lora_stage_cache.get(id(A))
    if hit is not None and hit[2] == A._version:
        return hit[0], hit[1]
    r = A.shape[0]
    pad = lora_pad(r)
    a_pad = torch.zeros((pad, A.shape[1]), device=A.device, dtype=BF16)
    a_pad[:r] = A.detach()
    b_pad = torch.zeros((B.shape[0], pad), device=B.device, dtype=BF16)
    b_pad[:, :r] = B.detach() * scaling
    return a_pad, b_pad


# ---------------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------------
# A low-priority stream for work whose result only the optimizer reads (the rank-r LoRA weight gradients: 2 launches of
# 16-144 CTAs per adapted linear, ~1.9 ms per step).  When set (train.GraphedTrainStep sets it for the capture), that
# work forks from the main stream and runs in the SMs the big persistent kernels leave idle (tail waves, the gaps
# between kernels); the caller joins the stream before the optimizer update.  None: everything stays on one stream.
side_stream = None
side_stream_used = False   # set when work was forked onto side_stream since the caller last cleared it


class _OnSideStream:
    """with _OnSideStream(*inputs): ... -- runs the body on `side_stream` (after everything queued so far on the
    current stream) and tells the caching allocator that `inputs` are still in use there."""

    def __init__(self, *inputs, enabled=True):
        self.inputs = inputs
        self.ctx = None
        self.enabled = enabled

    def __enter__(self):
        s = side_stream if self.enabled else None
        if s is None:
            return self
        global side_stream_used
        side_stream_used = True
        s.wait_stream(torch.cuda.current_stream())
        for t in self.inputs:
            if t is not None:
                t.record_stream(s)
        self.ctx = torch.cuda.stream(s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _group_grad(t, other, rows_per_group, like):
    """Gradient of a per-group row vector that was broadcast over `rows_per_group` consecutive rows:
    sum over the group's rows of t (* other).  One row per group: the product itself."""
    if rows_per_group == 1:
        g = t if other is None else (t.float() * other.float())
        return g.to(like.dtype)
    return colsum_groups(t, other, rows_per_group).to(like.dtype)


class GradJoin:
    """Side channel for `y = f(x) + x` when f starts and ends in a LinearFn (attn2: to_q ... to_out with the residual
    in its epilogue).  The closing node ("send") parks the residual-branch gradient here instead of returning it, the
    opening node ("recv", which always runs later in the backward) adds it in the epilogue of its dgrad GEMM -- the two
    gradients of x meet inside a GEMM instead of in an autograd add kernel (55 launches of 25 MB per step)."""
    __slots__ = ("grad",)

    def __init__(self):
        self.grad = None


class LinearFn(torch.autograd.Function):
    """y = gate * (x W^T + b + s (x A^T) B^T) + res    (every piece after x W^T optional).

    One GEMM forward (LoRA up-projection, bias, AdaLN gate and residual fused in the epilogue)
    plus a 64-wide GEMM for the LoRA down-projection.  Backward: dgrad reads W as an MN-major B
    operand (no transposed copy), the LoRA dt*A term rides in the same GEMM; wgrad / LoRA grads use
    the [K,M]-layout A operand."""

    @staticmethod
    def forward(ctx, x, W, b, A, B, scaling, gate, rows_per_gate, res, join=None, join_role=None):
        has_lora = A is not None
        t = a_pad = b_pad = None
        if has_lora:
            a_pad, b_pad = stage_lora(A, B, scaling)
            t = gemm(x, a_pad, block_n=64)
        u = None
        if gate is not None and ctx.needs_input_grad[6]:
            # trainable AdaLN gate: the pre-gate output leaves through the epilogue's second store (d(gate) = sum dy * u)
            u = torch.empty((x.shape[0], W.shape[0]), device=x.device, dtype=BF16)
        y = gemm(x, W, a2=t, b2=b_pad, bias=b, gate=gate, rows_per_gate=rows_per_gate, res=res, aux=u,
                 epilogue=EPI_STASH if u is not None else EPI_NONE)
        ctx.save_for_backward(x, W, t, a_pad, b_pad, gate, u)
        ctx.meta = (has_lora, scaling, rows_per_gate, A.shape[0] if has_lora else 0,
                    b is not None, res is not None)
        ctx.join = (join, join_role)
        ctx.adapters = (A, B)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, t, a_pad, b_pad, gate, u = ctx.saved_tensors
        has_lora, scaling, rpg, r, has_bias, has_res = ctx.meta
        need = ctx.needs_input_grad
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        g = rowscale(dy, gate, rpg) if gate is not None else dy
        dgate = _group_grad(dy, u, rpg, gate) if (need[6] and u is not None) else None
        dx = dW = db = dA = dB = None
        dt = None
        if has_lora and (need[0] or need[3]):
            dt = gemm(g, b_pad, b_rows_are_k=True, block_n=64)          # [M,64] = g (sB)
        join, role = ctx.join
        if need[0]:
            extra = None
            if join is not None and role == "recv":
                extra, join.grad = join.grad, None      # the residual branch's gradient of x, parked by the sender
            dx = gemm(g, W, b_rows_are_k=True, a2=dt, b2=a_pad if has_lora else None, res=extra)
        # the rank-r column views make the GEMMs write exactly [r, K] / [N, r] tensors: autograd can take them as
        # .grad without the copy it makes for a slice of a padded buffer.  (A rank that is not a multiple of 8 cannot be
        # a GEMM extent: the next multiple is computed -- the extra rows / columns come out zero -- and sliced.)
        rr = r if r % 8 == 0 else (r + 7) // 8 * 8
        if has_lora and (need[3] or need[4]):
            # The side stream is only safe when AccumulateGrad takes the returned tensors as `.grad` without touching
            # them (it then launches nothing): with an existing `.grad` (accumulation, bucket views) autograd adds on
            # the main stream, unordered against the side stream -- stay on the main stream in that case.
            pA, pB = ctx.adapters
            steal = (pA is None or pA.grad is None) and (pB is None or pB.grad is None)
            with _OnSideStream(g, dt, x, t, enabled=steal):   # only the optimizer reads these: off the critical path
                if need[3]:
                    dA = gemm(dt[:, :rr], x, a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32, split_k=0)
                    dA = dA if rr == r else dA[:r]
                if need[4]:
                    dB = gemm(g, t[:, :rr], a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32, block_n=64,
                              split_k=0)
                    dB = dB * scaling if scaling != 1.0 else dB
                    dB = dB if rr == r else dB[:, :r]
        if need[1]:
            dW = gemm(g, x, a_rows_are_k=True, b_rows_are_k=True)
        if has_bias and need[2]:
            db = (colsum_groups(g)[0] if g.shape[0] >= 1024 and g.shape[1] % 8 == 0 else colsum(g)).to(BF16)
        dres = dy if (has_res and need[8]) else None
        if dres is not None and join is not None and role == "send":
            join.grad, dres = dres, None                # delivered through the receiving node's dgrad epilogue
        return dx, dW, db, dA, dB, None, dgate, None, dres, None, None


class FeedForwardFn(torch.autograd.Function):
    """y = gate * (gelu_tanh(x W1^T + b1) W2^T + b2) + res.   Saves only the bf16 pre-activation;
    GELU' is applied in the epilogue of the W2 dgrad GEMM."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, gate, rows_per_gate, res):
        M = x.shape[0]
        pre = torch.empty((M, W1.shape[0]), device=x.device, dtype=BF16)
        act = gemm(x, W1, bias=b1, epilogue=EPI_GELU, aux=pre)
        u = None
        if gate is not None and ctx.needs_input_grad[5]:   # trainable AdaLN gate: stash the pre-gate output (see LinearFn)
            u = torch.empty((M, W2.shape[0]), device=x.device, dtype=BF16)
        y = gemm(act, W2, bias=b2, gate=gate, rows_per_gate=rows_per_gate, res=res, aux=u,
                 epilogue=EPI_STASH if u is not None else EPI_NONE)
        train_w = W1.requires_grad or W2.requires_grad
        ctx.save_for_backward(x, W1, W2, pre, gate, act if train_w else None, u)
        ctx.meta = (rows_per_gate, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W1, W2, pre, gate, act, u = ctx.saved_tensors
        rpg, has_res = ctx.meta
        need = ctx.needs_input_grad
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        g = rowscale(dy, gate, rpg) if gate is not None else dy
        dgate = _group_grad(dy, u, rpg, gate) if (need[5] and u is not None) else None
        dh = gemm(g, W2, b_rows_are_k=True, epilogue=EPI_GELU_GRAD, aux=pre)   # [M, Dff]
        dx = gemm(dh, W1, b_rows_are_k=True) if need[0] else None
        dW1 = db1 = dW2 = db2 = None
        if need[1]:
            dW1 = gemm(dh, x, a_rows_are_k=True, b_rows_are_k=True)
        if need[2]:
            db1 = colsum(dh).to(BF16)
        if need[3]:
            dW2 = gemm(g, act, a_rows_are_k=True, b_rows_are_k=True)
        if need[4]:
            db2 = colsum(g).to(BF16)
        dres = dy if (has_res and need[7]) else None
        return dx, dW1, db1, dW2, db2, dgate, None, dres


def _norm_mod_backward(ctx, dy, dres, x, scale, rpm, eps, ln):
    """dx (+ residual gradient) and, when the AdaLN tables train (training.py:75-91), d(scale) = sum_rows dy * xhat and
    d(shift) = sum_rows dy per modulation group -- the product leaves the same kernel, the sums are colsum_groups."""
    need_scale, need_shift = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
    dscale = dshift = None
    if need_scale:
        dx, prod = norm_mod_bwd(dy, x, scale, rpm, eps, ln, dres=dres, want_prod=True)
        dscale = _group_grad(prod, None, rpm, dy)
    else:
        dx = norm_mod_bwd(dy, x, scale, rpm, eps, ln, dres=dres)
    if need_shift:
        dshift = _group_grad(dy, None, rpm, dy)
    return dx, dscale, dshift, None, None, None


class NormModFn(torch.autograd.Function):
    """y = norm(x) * (1 + scale) + shift  with RMSNorm / LayerNorm (no affine)."""

    @staticmethod
    def forward(ctx, x, scale, shift, rows_per_mod, eps, layernorm):
        y = norm_mod_fwd(x, scale, shift, rows_per_mod, eps, layernorm)
        ctx.save_for_backward(x, scale)
        ctx.meta = (rows_per_mod, eps, layernorm)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        rpm, eps, ln = ctx.meta
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        return _norm_mod_backward(ctx, dy, None, x, scale, rpm, eps, ln)


class NormModResFn(torch.autograd.Function):
    """(y, x_res) = (norm(x) * (1 + scale) + shift, x): the residual branch leaves through the same node, so the
    backward adds its gradient inside the norm kernel (`dres`) instead of autograd launching an add at the join."""

    @staticmethod
    def forward(ctx, x, scale, shift, rows_per_mod, eps, layernorm):
        y = norm_mod_fwd(x, scale, shift, rows_per_mod, eps, layernorm)
        ctx.save_for_backward(x, scale)
        ctx.meta = (rows_per_mod, eps, layernorm)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x, scale = ctx.saved_tensors
        rpm, eps, ln = ctx.meta
        if dy is None:
            return dres, None, None, None, None, None
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        if dres is not None and dres.stride(1) != 1:
            dres = dres.contiguous()
        return _norm_mod_backward(ctx, dy, dres, x, scale, rpm, eps, ln)


class AttnCoreFn(torch.autograd.Function):
    """o = softmax(rope(qnorm(q)) rope(knorm(k))^T / 8 + key_bias) v  on token-major tensors.

    q_pre [B*Nq, D], k_pre / v [B*Nk, D] may be column slices of packed projection buffers.  The
    backward keeps dq in fp32 between the attention and the qk-norm kernels and returns gradients
    for the pre-norm projections."""

    @staticmethod
    def forward(ctx, q_pre, k_pre, v, wq, wk, cos, sin, key_bias, B, H, Nq, Nk, scale):
        D = H * 64
        qk = torch.empty((B * Nq + B * Nk, D), device=q_pre.device, dtype=BF16)
        q, k = qk[:B * Nq], qk[B * Nq:]
        qknorm_rope_fwd(q_pre, k_pre, wq, wk, cos, sin, q, k)
        o, lse = fa_fwd(q, k, v, B, H, Nq, Nk, key_bias, scale)
        ctx.save_for_backward(q_pre, k_pre, v, wq, wk, cos, sin, key_bias, qk, o, lse)
        ctx.meta = (B, H, Nq, Nk, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        q_pre, k_pre, v, wq, wk, cos, sin, key_bias, qk, o, lse = ctx.saved_tensors
        B, H, Nq, Nk, scale = ctx.meta
        train_w = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        D = H * 64
        q, k = qk[:B * Nq], qk[B * Nq:]
        do = do if do.stride(1) == 1 else do.contiguous()
        dk = torch.empty((B * Nk, D), device=do.device, dtype=BF16)
        dv = torch.empty((B * Nk, D), device=do.device, dtype=BF16)
        dq32 = fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, key_bias, scale)
        dq_pre = torch.empty((B * Nq, D), device=do.device, dtype=BF16)
        dk_pre = torch.empty((B * Nk, D), device=do.device, dtype=BF16)
        pq = torch.empty((B * Nq, D), device=do.device, dtype=BF16) if train_w else None
        pk = torch.empty((B * Nk, D), device=do.device, dtype=BF16) if train_w else None
        qknorm_rope_bwd(dq32, dk, q_pre, k_pre, wq, wk, cos, sin, dq_pre, dk_pre, prod_q=pq, prod_k=pk)
        dwq = colsum_groups(pq)[0].to(wq.dtype) if ctx.needs_input_grad[3] else None
        dwk = colsum_groups(pk)[0].to(wk.dtype) if ctx.needs_input_grad[4] else None
        return dq_pre, dk_pre, dv, dwq, dwk, None, None, None, None, None, None, None, None


class FlashAttnFn(torch.autograd.Function):
    """o = softmax(q k^T * scale + key_bias) v on already normalised / rotated token-major q, k (the attention core
    alone; used when the qk-norm weights are trainable and the norm + RoPE run as autograd-visible ops)."""

    @staticmethod
    def forward(ctx, q, k, v, key_bias, B, H, Nq, Nk, scale):
        q, k, v = (t if t.stride(1) == 1 else t.contiguous() for t in (q, k, v))
        o, lse = fa_fwd(q, k, v, B, H, Nq, Nk, key_bias, scale)
        ctx.save_for_backward(q, k, v, key_bias, o, lse)
        ctx.meta = (B, H, Nq, Nk, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, key_bias, o, lse = ctx.saved_tensors
        B, H, Nq, Nk, scale = ctx.meta
        do = do if do.stride(1) == 1 else do.contiguous()
        dk = torch.empty((B * Nk, H * 64), device=do.device, dtype=BF16)
        dv = torch.empty((B * Nk, H * 64), device=do.device, dtype=BF16)
        dq32 = fa_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, dk, dv, key_bias, scale)
        return dq32.to(BF16), dk, dv, None, None, None, None, None, None


def wants_grad(*tensors) -> bool:
    """True when autograd is recording and one of the tensors needs a gradient: the fused kernels treat AdaLN
    modulations, gates and qk-norm weights as constants, so such a piece then runs un-fused (train_mode='full')."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def norm_mod(x, scale, shift, rows_per_mod, eps, layernorm, with_res: bool):
    """norm(x) * (1 + scale) + shift (+ the residual view of x): one fused kernel forward, one backward (the residual
    gradient folded in); when the modulation trains (train_mode='full', training.py:75-91) the backward also emits
    dy * xhat and two grouped column sums give d(scale) / d(shift)."""
    if with_res:
        return NormModResFn.apply(x, scale, shift, rows_per_mod, eps, layernorm)
    return NormModFn.apply(x, scale, shift, rows_per_mod, eps, layernorm), None


def gate_residual(u, gate, rows_per_gate, res):
    """res + gate * u as torch ops (the form autograd needs when the gate is trainable)."""
    rows, D = u.shape
    if gate is not None:
        g = rows // rows_per_gate
        u = (u.view(g, rows_per_gate, D) * gate.reshape(g, 1, D)).reshape(rows, D)
    return u + res if res is not None else u


def rope_torch(x, cos, sin):
    """apply_rotary_emb (attention.py:917-932): x cos + rot(x) sin with rot pairs (-x2, x1); bf16 op by op."""
    xr = x.reshape(x.shape[0], -1, 2)
    rot = torch.stack((-xr[..., 1], xr[..., 0]), dim=-1).reshape(x.shape)
    return x * cos + rot * sin


class CtxKVFn(torch.autograd.Function):
    """attn2 key / value projections (+ LoRA) of ALL blocks from the projected caption tokens, which are the
    same for every block: 2 launches forward (LoRA down-projection, strided-batched projection with the LoRA
    up-projection as second operand pair) and 4 backward, instead of 4 forward + 10 backward PER BLOCK of
    latency-bound M = 256 GEMMs on 2 SMs.

        kv_g = ctx W_g^T + b_g + (ctx A_g^T)(s B_g)^T        g = (block, k | v)

    Returns G tensors [M, D]: column blocks of one [M, G*D] buffer.  Wkv [G*D, Dc] / bkv [G*D] are cached
    concatenations of the frozen weights; As / Bs are the G fp32 adapters (or empty lists)."""

    @staticmethod
    def forward(ctx, x, Wkv, bkv, G, scaling, r, *adapters):
        M, Dc = x.shape
        D = Wkv.shape[0] // G
        has_lora = len(adapters) > 0
        kv = torch.empty((M, G * D), device=x.device, dtype=BF16)
        a_st = b_st = t = None
        if has_lora:
            As, Bs = adapters[:G], adapters[G:]
            pad = lora_pad(r)
            a_st = torch.zeros((G, pad, Dc), device=x.device, dtype=BF16)
            a_st[:, :r] = torch.stack([A.detach() for A in As])
            b_st = torch.zeros((G, D, pad), device=x.device, dtype=BF16)
            bs = torch.stack([B.detach() for B in Bs])
            b_st[:, :, :r] = bs * scaling if scaling != 1.0 else bs
            a_st, b_st = a_st.view(G * pad, Dc), b_st.view(G * D, pad)
            t = gemm(x, a_st)  # [M, G*pad]
        pad = lora_pad(r) if has_lora else 0
        gemm_batched(x, Wkv, kv, M, D, Dc, G, {"b": (D, 0), "a2": (0, pad), "b2": (D, 0), "c": (0, D), "bias": D},
                     a2=t, b2=b_st, K2=pad, bias=bkv)
        ctx.save_for_backward(x, Wkv, a_st, b_st, t)
        ctx.meta = (G, D, scaling, r, has_lora)
        return tuple(kv[:, g * D:(g + 1) * D] for g in range(G))

    @staticmethod
    def backward(ctx, *grads):
        x, Wkv, a_st, b_st, t = ctx.saved_tensors
        G, D, scaling, r, has_lora = ctx.meta
        M, Dc = x.shape
        zero = None
        cols = []
        for g in grads:
            if g is None:
                if zero is None:
                    zero = torch.zeros((M, D), device=x.device, dtype=BF16)
                g = zero
            cols.append(g)
        dkv = torch.cat(cols, dim=1)  # [M, G*D]
        dt = None
        pad = lora_pad(r) if has_lora else 0
        if has_lora:
            dt = torch.empty((M, G * pad), device=x.device, dtype=BF16)
            gemm_batched(dkv, b_st, dt, M, pad, D, G, {"a": (0, D), "b": (D, 0), "c": (0, pad)},
                         b_rows_are_k=True, block_n=64)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dkv, Wkv, b_rows_are_k=True, a2=dt, b2=a_st, out_dtype=torch.float32, split_k=0).to(BF16)
        dAs = dBs = ()
        if has_lora:
            dA = gemm(dt, x, a_rows_are_k=True, b_rows_are_k=True, out_dtype=torch.float32, split_k=0)
            dA = dA.view(G, pad, Dc)
            dB = torch.empty((G * D, pad), device=x.device, dtype=torch.float32)
            gemm_batched(dkv, t, dB, D, pad, M, G, {"a": (0, D), "b": (0, pad), "c": (D, 0)},
                         a_rows_are_k=True, b_rows_are_k=True, block_n=64)
            dB = dB.view(G, D, pad)
            if scaling != 1.0:
                dB = dB * scaling
            need = ctx.needs_input_grad
            dAs = tuple(dA[g, :r] if need[6 + g] else None for g in range(G))
            dBs = tuple(dB[g, :, :r] if need[6 + G + g] else None for g in range(G))
        return (dx, None, None, None, None, None) + dAs + dBs


def linear(x, W, b=None, lora=None, gate=None, rows_per_gate=0, res=None, join=None, join_role=None):
    A, B, s = lora if lora is not None else (None, None, 1.0)
    return LinearFn.apply(x, W, b, A, B, s, gate, rows_per_gate, res, join, join_role)


class SelfAttnFn(torch.autograd.Function):
    """attn1 of a block with frozen projections, as one autograd node:

        y = gate * (FA(rope(qnorm(x Wq^T)), rope(knorm(x Wk^T)), x Wv^T) Wo^T + bo) + res

    Forward: one projection GEMM into a packed [M,3D] buffer, qk-norm+RoPE, flash attention, output
    GEMM with the AdaLN gate and the residual in its epilogue.  Backward: one dgrad GEMM over the
    packed [M,3D] gradient against the cached [3D,D] concatenation of the frozen weights."""

    @staticmethod
    def forward(ctx, x, Wqkv, bqkv, wqn, wkn, cos, sin, Wo, bo, gate, rows_per_gate, res, key_bias, B, H, N,
                scale, sp=None, batch_keep=None, pass_input=False):
        M, D = x.shape[0], H * 64
        qkv = gemm(x, Wqkv, bias=bqkv)  # one [M,3D] GEMM against the [3D,D] weight concatenation
        qk = torch.empty((M, 2 * D), device=x.device, dtype=BF16)
        qknorm_rope_fwd(qkv[:, :D], qkv[:, D:2 * D], wqn, wkn, cos, sin, qk[:, :D], qk[:, D:])
        kv = o_h = None
        if sp is None:
            # batch_keep (inference only): STG skip of some batch entries inside the attention launch -- their output is
            # the value rows ("attention values") or the attention input x ("attention skip"), attention.py:1071-1086
            o, lse = fa_fwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], B, H, N, N, key_bias, scale, attn1=True,
                            batch_keep=batch_keep, pass_src=x if (batch_keep is not None and pass_input) else None)
        else:  # sequence-sharded: the local queries meet every rank's keys/values around the ring
            if key_bias is not None:
                raise _lib.B200Error("ring attn1: a key mask on the sharded self-attention is not built")
            from . import ring
            if sp.mode == "gather":
                o, lse, k_all, v_all = ring.gather_fwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], sp.group, B, H, N, scale)
                kv = torch.stack([k_all, v_all])  # [2, P*n, D] kept for the backward
            elif sp.mode == "heads":
                o, lse, kv, o_h = ring.heads_fwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], sp.group, B, H, N, scale, sp.impl)
            else:
                o, lse, kv = ring.ring_fwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], sp.group, B, H, N, scale, sp.impl)
        u = None
        if gate is not None and ctx.needs_input_grad[9]:   # trainable gate: stash the pre-gate output (see LinearFn)
            u = torch.empty((M, Wo.shape[0]), device=x.device, dtype=BF16)
        y = gemm(o, Wo, bias=bo, gate=gate, rows_per_gate=rows_per_gate, res=res, aux=u,
                 epilogue=EPI_STASH if u is not None else EPI_NONE)
        train_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        heads = o_h is not None   # head-exchange mode: the backward reads q / k / v / o in the head-sharded layout
        ctx.save_for_backward(qkv, None if heads else qk, o if (not heads or ctx.needs_input_grad[7]) else None, lse, Wqkv,
                              wqn, wkn, cos, sin, Wo, gate, key_bias, kv, u, x if train_w else None, o_h)
        ctx.meta = (B, H, N, scale, rows_per_gate, res is not None, sp)
        return y

    @staticmethod
    def backward(ctx, dy):
        qkv, qk, o, lse, Wqkv, wqn, wkn, cos, sin, Wo, gate, key_bias, kv, u, x, o_h = ctx.saved_tensors
        B, H, N, scale, rpg, has_res, sp = ctx.meta
        need = ctx.needs_input_grad
        D = H * 64
        M = qkv.shape[0]
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        g = rowscale(dy, gate, rpg) if gate is not None else dy
        dgate = _group_grad(dy, u, rpg, gate) if (need[9] and u is not None) else None
        dWo = gemm(g, o, a_rows_are_k=True, b_rows_are_k=True) if need[7] else None
        dbo = colsum_groups(g)[0].to(BF16) if need[8] else None
        do = gemm(g, Wo, b_rows_are_k=True)
        dqkv = torch.empty((M, 3 * D), device=dy.device, dtype=BF16)
        if sp is None:
            dk_post = torch.empty((M, D), device=dy.device, dtype=BF16)
            dq32 = fa_bwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], o, do, lse, B, H, N, N, dk_post, dqkv[:, 2 * D:],
                          key_bias, scale, attn1=True)
        else:
            from . import ring
            if sp.mode == "gather":
                dq32, dk_post, dv_loc = ring.gather_bwd(qk[:, :D], kv[0], kv[1], o, do, lse, sp.group, B, H, N, scale)
                dqkv[:, 2 * D:].copy_(dv_loc)
            elif sp.mode == "heads":
                dq32, dk_post, dv_loc = ring.heads_bwd(kv, o_h, do, lse, sp.group, B, H, N, scale, sp.impl)
                dqkv[:, 2 * D:].copy_(dv_loc)
            else:
                dq32, dkv = ring.ring_bwd(qk[:, :D], kv, o, do, lse, sp.group, B, H, N, scale, sp.impl)
                dk_post = dkv[0]  # fp32, fully reduced over the ring
                dqkv[:, 2 * D:].copy_(dkv[1])
        train_n = need[3] or need[4]
        pq = torch.empty((M, D), device=dy.device, dtype=BF16) if train_n else None
        pk = torch.empty((M, D), device=dy.device, dtype=BF16) if train_n else None
        qknorm_rope_bwd(dq32, dk_post, qkv[:, :D], qkv[:, D:2 * D], wqn, wkn, cos, sin, dqkv[:, :D], dqkv[:, D:2 * D],
                        prod_q=pq, prod_k=pk)
        dwqn = colsum_groups(pq)[0].to(wqn.dtype) if need[3] else None
        dwkn = colsum_groups(pk)[0].to(wkn.dtype) if need[4] else None
        dx = gemm(dqkv, Wqkv, b_rows_are_k=True) if need[0] else None
        dWqkv = gemm(dqkv, x, a_rows_are_k=True, b_rows_are_k=True) if need[1] else None   # [3D, D]: q | k | v rows
        dbqkv = colsum_groups(dqkv)[0].to(BF16) if need[2] else None
        dres = dy if (has_res and need[11]) else None
        return (dx, dWqkv, dbqkv, dwqn, dwkn, None, None, dWo, dbo, dgate, None, dres) + (None,) * 8
