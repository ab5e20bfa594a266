"""Sequence-sharded ring variant of attn1 for long clips (BASELINE config 5).

Every rank owns a contiguous shard of N/P latent tokens.  All token-local ops (norms, projections,
FFN, attn2 against the replicated caption tokens) need no communication; attn1 needs every key/value.
K/V shards travel around a ring of P ranks (NCCL P2P over NVLink/NVSwitch through torch.distributed):
at hop h a rank attends its local queries to the shard of rank (r - h) mod P while the next shard is in
flight on the communicator's stream, and folds the partial (O, LSE) into fp32 accumulators with the
online-softmax merge kernel.  The backward walks the ring again: dQ accumulates locally in fp32 over
the hops (the flash backward reduces into a caller-owned buffer), while the fp32 dK/dV accumulator of
a shard travels one hop behind its shard and arrives, fully reduced, back at the owner after P hops;
its transfer overlaps the next hop's backward kernel.

Why copy-then-compute and not peer loads inside the flash kernel: every 128-query CTA re-reads all
keys/values of its head, i.e. N/128 passes over the shard; served from the local 126 MB L2 that is
free, served over NVLink it would be ~48x the shard size per layer (2.5 GB at cfg5/P=2, ~2.8 ms at
900 GB/s against 0.6 ms of math).  One DMA of the shard per hop (52 MB at P=2, 13 MB at P=8; 58 us at
900 GB/s) hidden behind the ~0.3-1.2 ms hop kernel is the right shape.

The reference has no counterpart (no sequence parallelism anywhere); numerics are
checked against the un-sharded path."""
from typing import Optional

import torch
import torch.distributed as dist

from . import ops

BF16 = torch.bfloat16


def _peers(group):
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    if group is not None:
        nxt, prv = dist.get_global_rank(group, nxt), dist.get_global_rank(group, prv)
    return nxt, prv


def _ring_exchange(send: torch.Tensor, recv: torch.Tensor, group) -> list:
    """Post send-to-next / recv-from-previous of one buffer; returns the work handles."""
    nxt, prv = _peers(group)
    return dist.batch_isend_irecv([dist.P2POp(dist.isend, send, nxt, group), dist.P2POp(dist.irecv, recv, prv, group)])


def _wait(works):
    for w in works:
        w.wait()


class LocalAttention:
    """The per-hop kernels.  The default runs the b200 flash kernels; CPU tests inject a plain torch
    implementation with the same methods to exercise the ring logic over gloo."""

    def fwd(self, q, k, v, B, H, nq, nk, scale):
        return ops.fa_fwd(q, k, v, B, H, nq, nk, None, scale, attn1=True)

    def merge(self, o_acc, lse_acc, o_i, lse_i, B, H, n, first, out):
        ops.attn_merge(o_acc, lse_acc, o_i, lse_i, B, H, n, first, out)

    def delta(self, o, do, B, H, nq):
        return ops.attn_delta(o, do, B, H, nq)

    def bwd(self, q, k, v, o, do, lse, delta, dq_accum, B, H, nq, nk, scale):
        dk = torch.empty_like(k)
        dv = torch.empty_like(v)
        ops.fa_bwd(q, k, v, o, do, lse, B, H, nq, nk, dk, dv, None, scale, delta=delta, dq_accum=dq_accum, attn1=True)
        return dk, dv


def ring_fwd(q, k, v, group, B, H, n_local, scale, impl: Optional[LocalAttention] = None):
    """q/k/v: [B*n_local, >= H*64] (row-strided views allowed).  Returns (out bf16 [B*n_local, H*64],
    lse fp32 [B,H,n_local] over ALL keys, kv = the contiguous [2, B*n_local, H*64] message buffer)."""
    impl = impl or LocalAttention()
    world = dist.get_world_size(group)
    D = H * 64
    kv = torch.stack([k, v])  # contiguous copy: one message per hop
    o_acc = torch.empty((B * n_local, D), device=q.device, dtype=torch.float32)
    lse_acc = torch.empty((B, H, n_local), device=q.device, dtype=torch.float32)
    out = torch.empty((B * n_local, D), device=q.device, dtype=q.dtype)
    cur = kv
    for hop in range(world):
        works, nxt = [], None
        if hop + 1 < world:
            nxt = torch.empty_like(kv)
            works = _ring_exchange(cur, nxt, group)  # in flight during this hop's attention
        o_i, lse_i = impl.fwd(q, cur[0], cur[1], B, H, n_local, n_local, scale)
        impl.merge(o_acc, lse_acc, o_i, lse_i, B, H, n_local, hop == 0, out if hop + 1 == world else None)
        _wait(works)
        if nxt is not None:
            cur = nxt
    return out, lse_acc, kv


def ring_bwd(q, kv, out, do, lse, group, B, H, n_local, scale, impl: Optional[LocalAttention] = None):
    """Returns (dq fp32 [B*n_local, H*64], dkv fp32 [2, B*n_local, H*64]) for the local shard."""
    impl = impl or LocalAttention()
    world = dist.get_world_size(group)
    delta = impl.delta(out, do, B, H, n_local)
    dq = torch.zeros((B * n_local, H * 64), device=q.device, dtype=torch.float32)
    cur_kv = kv
    g_works, g_in = [], None
    for hop in range(world):
        kv_works, nxt_kv = [], None
        if hop + 1 < world:
            nxt_kv = torch.empty_like(kv)
            kv_works = _ring_exchange(cur_kv, nxt_kv, group)
        dk, dv = impl.bwd(q, cur_kv[0], cur_kv[1], out, do, lse, delta, dq, B, H, n_local, n_local, scale)
        # the accumulator of the shard processed here was sent by the previous rank after its last hop
        _wait(g_works)
        if g_in is None:
            g = torch.stack([dk, dv]).float()
        else:
            g = g_in
            g[0] += dk
            g[1] += dv
        _wait(kv_works)
        # ... and follows its shard to the next rank (after the last hop: back home), overlapping
        # with the next hop's backward kernel
        g_in = torch.empty_like(g)
        g_works = _ring_exchange(g, g_in, group)
        if nxt_kv is not None:
            cur_kv = nxt_kv
    _wait(g_works)
    return dq, g_in


class RingAttnFn(torch.autograd.Function):
    """o = softmax(q [k_0 .. k_{P-1}]^T * scale) [v_0 .. v_{P-1}] for the local query shard."""

    @staticmethod
    def forward(ctx, q, k, v, group, B, H, n_local, scale, impl):
        out, lse, kv = ring_fwd(q, k, v, group, B, H, n_local, scale, impl)
        ctx.save_for_backward(q, kv, out, lse)
        ctx.meta = (group, B, H, n_local, scale, impl)
        return out

    @staticmethod
    def backward(ctx, do):
        q, kv, out, lse = ctx.saved_tensors
        group, B, H, n_local, scale, impl = ctx.meta
        dq, dkv = ring_bwd(q, kv, out, do.contiguous(), lse, group, B, H, n_local, scale, impl)
        return dq.to(q.dtype), dkv[0].to(q.dtype), dkv[1].to(q.dtype), None, None, None, None, None, None


def ring_attention(q, k, v, group, B, H, n_local, scale, impl: Optional[LocalAttention] = None):
    return RingAttnFn.apply(q, k, v, group, B, H, n_local, scale, impl)


def gather_fwd(q, k, v, group, B, H, n_local, scale):
    """All-gather variant: every rank collects all keys / values once (two NCCL all-gathers per layer) and runs ONE
    flash-attention launch of its n_local queries against all N keys -- ~6 host operations per layer and direction
    instead of ~16 P for the hop loop, which is what matters once the shards are small (P >= 4: the hop kernels take
    0.15-0.25 ms each and the ring becomes launch bound).  No overlap of the transfer with the math (104 MB of K/V per
    layer at cfg5, ~0.3 ms at NVLink all-gather rates, against ~1.5 ms of attention at P = 4).
    Returns (out, lse, k_all, v_all)."""
    world = dist.get_world_size(group)
    D = H * 64
    if B != 1:
        raise ops._lib.B200Error("gather-mode sequence parallelism is built for batch 1 (long single clips)")
    k_all = torch.empty((world * n_local, D), device=q.device, dtype=q.dtype)
    v_all = torch.empty((world * n_local, D), device=q.device, dtype=q.dtype)
    dist.all_gather_into_tensor(k_all, k.contiguous(), group=group)
    dist.all_gather_into_tensor(v_all, v.contiguous(), group=group)
    out, lse = ops.fa_fwd(q, k_all, v_all, B, H, n_local, world * n_local, None, scale, attn1=True)
    return out, lse, k_all, v_all


def gather_bwd(q, k_all, v_all, out, do, lse, group, B, H, n_local, scale):
    """Backward of gather_fwd: one flash backward against all keys, then reduce-scatter of the bf16 dK / dV partials
    (rank r keeps the sum for its own shard).  Returns (dq fp32, dk bf16, dv bf16) for the local shard."""
    world = dist.get_world_size(group)
    D = H * 64
    dk_all = torch.empty_like(k_all)
    dv_all = torch.empty_like(v_all)
    dq = ops.fa_bwd(q, k_all, v_all, out, do, lse, B, H, n_local, world * n_local, dk_all, dv_all, None, scale, attn1=True)
    dk = torch.empty((n_local, D), device=q.device, dtype=k_all.dtype)
    dv = torch.empty((n_local, D), device=q.device, dtype=k_all.dtype)
    dist.reduce_scatter_tensor(dk, dk_all, op=dist.ReduceOp.SUM, group=group)
    dist.reduce_scatter_tensor(dv, dv_all, op=dist.ReduceOp.SUM, group=group)
    return dq, dk, dv


# ---- head exchange: tokens <-> heads with one all-to-all each way ------------------------------------------------------
# A rank hands every peer the slice of ITS tokens that belongs to the PEER's H/P heads and receives all N tokens of its
# own heads: attention then runs un-sharded over N x N for H/P heads -- no partial softmax, no partial dK/dV, exactly the
# arithmetic of the single-GPU kernel -- and a second all-to-all brings the output rows home.  Per layer and rank the
# forward moves 4 (q, k, v, o) x (P-1)/P x n_local x D bf16 (23 MB at cfg5 / P = 8, against 91 MB of gathered K/V) and the
# backward the same plus fp32 dQ; nothing is reduced across ranks.  Needs H % P == 0 (32 heads: P in 2, 4, 8, 16, 32).

def _all_to_all(x: torch.Tensor, group) -> torch.Tensor:
    """x [P, ...] contiguous: slice r goes to rank r; returns [P, ...] with slice r = what rank r sent here."""
    out = torch.empty_like(x)
    dist.all_to_all_single(out, x, group=group)
    return out


def tokens_to_heads(parts, group, B, n_local):
    """parts: k tensors [B*n_local, D] of one dtype (row-strided views allowed), this rank's tokens x all heads.
    Returns [k, B*P*n_local, D/P]: all tokens (batch-major, sequence order) x this rank's heads."""
    P = dist.get_world_size(group)
    k, D = len(parts), parts[0].shape[1]
    cp = D // P
    send = torch.empty((P, k, B * n_local, cp), device=parts[0].device, dtype=parts[0].dtype)
    for i, t in enumerate(parts):
        send[:, i].copy_(t.unflatten(1, (P, cp)).transpose(0, 1))
    recv = _all_to_all(send, group)                     # [P = source rank = token chunk, k, B*n, cp]
    return recv.view(P, k, B, n_local, cp).permute(1, 2, 0, 3, 4).reshape(k, B * P * n_local, cp)


def heads_to_tokens(parts, group, B, n_local):
    """Inverse of tokens_to_heads: k tensors [B*P*n_local, D/P] -> [k, B*n_local, D]."""
    P = dist.get_world_size(group)
    k, cp = len(parts), parts[0].shape[1]
    send = torch.empty((P, k, B, n_local, cp), device=parts[0].device, dtype=parts[0].dtype)
    for i, t in enumerate(parts):
        send[:, i].copy_(t.view(B, P, n_local, cp).transpose(0, 1))
    recv = _all_to_all(send, group)                     # [P = source rank = head chunk, k, B, n, cp]
    return recv.view(P, k, B * n_local, cp).permute(1, 2, 0, 3).reshape(k, B * n_local, P * cp)


def heads_fwd(q, k, v, group, B, H, n_local, scale, impl: Optional[LocalAttention] = None):
    """Returns (out [B*n_local, H*64] for the local tokens, lse [B, H/P, N], qkv_h [3, B*N, D/P], o_h [B*N, D/P]): the
    last three are what the backward needs, in the head-sharded layout."""
    impl = impl or LocalAttention()
    P = dist.get_world_size(group)
    if H % P:
        raise ops._lib.B200Error(f"head-exchange sequence parallelism needs the head count ({H}) to divide by the group "
                                 f"size ({P})")
    N = P * n_local
    qkv_h = tokens_to_heads([q, k, v], group, B, n_local)
    o_h, lse = impl.fwd(qkv_h[0], qkv_h[1], qkv_h[2], B, H // P, N, N, scale)
    out = heads_to_tokens([o_h], group, B, n_local)[0]
    return out, lse, qkv_h, o_h


def heads_bwd(qkv_h, o_h, do, lse, group, B, H, n_local, scale, impl: Optional[LocalAttention] = None):
    """Returns (dq fp32, dk, dv) [B*n_local, H*64] for the local tokens; every one is a complete sum (the attention ran
    over all tokens of its heads), so nothing is reduced."""
    impl = impl or LocalAttention()
    P = dist.get_world_size(group)
    hp, N = H // P, P * n_local
    do_h = tokens_to_heads([do], group, B, n_local)[0]
    delta = impl.delta(o_h, do_h, B, hp, N)
    dq_h = torch.zeros((B * N, hp * 64), device=do.device, dtype=torch.float32)
    dk_h, dv_h = impl.bwd(qkv_h[0], qkv_h[1], qkv_h[2], o_h, do_h, lse, delta, dq_h, B, hp, N, N, scale)
    # one message for the three gradients: per destination rank [dq fp32 | dk | dv] as raw bytes
    cp = hp * 64
    rows = B * n_local
    seg = rows * cp                                                   # elements per destination and tensor
    parts = (dq_h, dk_h, dv_h)
    offs = [0]
    for t in parts:
        offs.append(offs[-1] + seg * t.element_size())
    send = torch.empty((P, offs[-1]), device=do.device, dtype=torch.uint8)
    for i, t in enumerate(parts):
        send[:, offs[i]:offs[i + 1]].view(t.dtype).view(P, B, n_local, cp).copy_(
            t.view(B, P, n_local, cp).transpose(0, 1))
    recv = _all_to_all(send, group)                                    # [P = source rank = head chunk, bytes]
    out = [recv[:, offs[i]:offs[i + 1]].view(t.dtype).view(P, rows, cp).transpose(0, 1).reshape(rows, P * cp)
           for i, t in enumerate(parts)]
    return out[0], out[1], out[2]


class HeadsAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, group, B, H, n_local, scale, impl):
        out, lse, qkv_h, o_h = heads_fwd(q, k, v, group, B, H, n_local, scale, impl)
        ctx.save_for_backward(qkv_h, o_h, lse)
        ctx.meta = (group, B, H, n_local, scale, impl)
        return out

    @staticmethod
    def backward(ctx, do):
        qkv_h, o_h, lse = ctx.saved_tensors
        group, B, H, n_local, scale, impl = ctx.meta
        dq, dk, dv = heads_bwd(qkv_h, o_h, do.contiguous(), lse, group, B, H, n_local, scale, impl)
        return dq.to(do.dtype), dk, dv, None, None, None, None, None, None


def heads_attention(q, k, v, group, B, H, n_local, scale, impl: Optional[LocalAttention] = None):
    return HeadsAttnFn.apply(q, k, v, group, B, H, n_local, scale, impl)


class SequenceParallel:
    """Per-model sequence-sharding state (set by api.enable_sequence_parallel).  mode: "ring" (K/V hops overlapped
    with the per-hop attention, online-softmax merge), "gather" (all-gather K/V, one attention launch) or "heads"
    (all-to-all: every rank attends ALL tokens for H/P heads)."""

    def __init__(self, group=None, impl: Optional[LocalAttention] = None, mode: str = "ring"):
        if mode not in ("ring", "gather", "heads"):
            raise ValueError("sequence-parallel mode must be 'ring', 'gather' or 'heads'")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.impl = impl
        self.mode = mode

    def shard(self, n_total: int):
        """(offset, length) of this rank's contiguous token shard."""
        if n_total % self.world:
            raise ops._lib.B200Error(f"sequence sharding needs the token count ({n_total}) to divide by the "
                                     f"group size ({self.world})")
        n = n_total // self.world
        return self.rank * n, n
