"""world_size-2/3 gloo tests of the sequence-sharded ring attn1 logic (b200_ltx.ring) on CPU.

The per-hop kernels are injected as plain torch fp32 (ring.LocalAttention protocol), so what is checked
here is the ring itself: hop order, the online-softmax merge, dQ accumulation over hops, and the dK/dV
accumulator that travels one hop behind its shard and comes home after P hops."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class TorchLocalAttention:
    """ring.LocalAttention protocol in plain torch (token-major [B*n, H*64] tensors)."""

    @staticmethod
    def _heads(x, B, H, n):
        return x.reshape(B, n, H, 64).permute(0, 2, 1, 3).float()

    def fwd(self, q, k, v, B, H, nq, nk, scale):
        s = self._heads(q, B, H, nq) @ self._heads(k, B, H, nk).transpose(-1, -2) * scale
        lse = torch.logsumexp(s, dim=-1)
        o = torch.softmax(s, dim=-1) @ self._heads(v, B, H, nk)
        return o.permute(0, 2, 1, 3).reshape(B * nq, H * 64).to(q.dtype), lse

    def merge(self, o_acc, lse_acc, o_i, lse_i, B, H, n, first, out):
        li = lse_i
        if first:
            o_acc.copy_(o_i.float())
            lse_acc.copy_(li)
        else:
            new = torch.logaddexp(lse_acc, li)
            wa = torch.exp(lse_acc - new).permute(0, 2, 1).reshape(B * n, H, 1)
            wb = torch.exp(li - new).permute(0, 2, 1).reshape(B * n, H, 1)
            o_acc.copy_((o_acc.view(B * n, H, 64) * wa + o_i.float().view(B * n, H, 64) * wb).view(B * n, H * 64))
            lse_acc.copy_(new)
        if out is not None:
            out.copy_(o_acc.to(out.dtype))

    def delta(self, o, do, B, H, nq):
        return (o.float() * do.float()).view(B, nq, H, 64).sum(-1).permute(0, 2, 1).contiguous()

    def bwd(self, q, k, v, o, do, lse, delta, dq_accum, B, H, nq, nk, scale):
        qh, kh, vh, doh = (self._heads(t, B, H, n) for t, n in ((q, nq), (k, nk), (v, nk), (do, nq)))
        p = torch.exp(qh @ kh.transpose(-1, -2) * scale - lse[..., None])
        dv = p.transpose(-1, -2) @ doh
        ds = p * (doh @ vh.transpose(-1, -2) - delta[..., None]) * scale
        dq_accum += (ds @ kh).permute(0, 2, 1, 3).reshape(B * nq, H * 64)
        dk = ds.transpose(-1, -2) @ qh
        back = lambda t: t.permute(0, 2, 1, 3).reshape(B * nk, H * 64).to(k.dtype)
        return back(dk), back(dv)


def _worker(rank, world, port, out, mode="ring"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_ltx import ring
    B, H, n = 2, (2 * world if mode == "heads" else 3), 10
    D, N = H * 64, n * world
    g = torch.Generator().manual_seed(7)
    q, k, v, do = (torch.randn(B, N, D, generator=g) for _ in range(4))

    def shard(t):
        return t[:, rank * n:(rank + 1) * n].reshape(B * n, D).clone()
    ql, kl, vl = (shard(t).requires_grad_(True) for t in (q, k, v))
    attend = ring.heads_attention if mode == "heads" else ring.ring_attention
    o = attend(ql, kl, vl, None, B, H, n, 0.125, TorchLocalAttention())
    o.backward(shard(do))
    # un-sharded reference
    qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
    hd = lambda t: t.view(B, N, H, 64).permute(0, 2, 1, 3)
    oref = (torch.softmax(hd(qf) @ hd(kf).transpose(-1, -2) * 0.125, -1) @ hd(vf)).permute(0, 2, 1, 3).reshape(B, N, D)
    oref.backward(do)
    err = max(float((o.detach() - shard(oref.detach())).abs().max()),
              float((ql.grad - shard(qf.grad)).abs().max()),
              float((kl.grad - shard(kf.grad)).abs().max()),
              float((vl.grad - shard(vf.grad)).abs().max()))
    out[rank] = err
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_attention_matches_unsharded(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and all(v < 2e-5 for v in out.values()), dict(out)


@pytest.mark.parametrize("world", [2, 3])
def test_head_exchange_attention_matches_unsharded(world):
    """mode="heads": tokens -> heads all-to-all, un-sharded attention over H/P heads, heads -> tokens all-to-all (batch 2:
    the exchange also has to keep the batch-major token order)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, "heads"), nprocs=world, join=True)
    assert len(out) == world and all(v < 2e-5 for v in out.values()), dict(out)
