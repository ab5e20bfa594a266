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


def _model_worker(rank, world, port, out, mode):
    """Whole train step of the mirror model, sequence-sharded over `world` gloo ranks (tokens, coordinates and noise
    sharded inside train_step, timestep broadcast, conditioning lerp at the shard's token offset, attn1 through the ring /
    the head exchange, everything else token-local) against the same step un-sharded: the mean of the per-shard losses
    is the loss, the average of the per-shard gradients is the gradient."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.join(os.path.dirname(here), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import model_checks as mc
    import ref_block as rb
    import torch_kernels as tk
    from b200_ltx import api
    torch.set_num_threads(2)
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=2, cross_attention_dim=128, caption_channels=64)
    P = rb.init_params(cfg, 8, seed=5)
    b = rb.synthetic_batch(cfg, 1, 3, 2, 2 * world, 16, 77, 9)      # 3 frames x 2 x 2P tokens: frame 0 ends inside rank 0's shard

    class Cfg:
        rf_log_normal_mu, rf_log_normal_sigma = -0.5, 1.0
        rf_quantile_min, rf_quantile_max = 0.005, 0.999
        transformer_loss_weight = 1.0

    def step(model):
        for p in model.parameters():
            p.grad = None
        t = torch.tensor([0.37]) if rank == 0 else torch.tensor([0.99])     # rank 0's value must win (broadcast)
        loss = api.train_step(model, {n: b[n] for n in ("latents", "ref_image_latents", "pose_latents")},
                              api.RectifiedFlowScheduler(), api.SymmetricPatchifier(1), Cfg(), b["prompt_embeds"],
                              b["prompt_mask"], device=torch.device("cpu"), t=t, noise=b["noise"].to(torch.bfloat16))[0]
        loss.backward()
        return loss.detach().float(), {n: p.grad.detach().float().clone() for n, p in model.named_parameters()
                                       if p.grad is not None}

    with tk.patched():
        model = mc.build_b200_model(cfg, P, 8, device="cpu").train()
        sp = api.enable_sequence_parallel(model, None, TorchLocalAttention(), mode)
        assert sp is not None and sp.world == world
        loss_s, grads_s = step(model)
        dist.all_reduce(loss_s)
        for g in grads_s.values():
            dist.all_reduce(g)
        api.enable_sequence_parallel(model, disable=True)
        loss_f, grads_f = step(model) if rank == 0 else (None, None)
    if rank == 0:
        worst = abs(float(loss_s) / world - float(loss_f)) / float(loss_f)
        assert set(grads_s) == set(grads_f) and len(grads_f) == 20
        for n in grads_f:
            worst = max(worst, float((grads_s[n] / world - grads_f[n]).norm()) / (float(grads_f[n].norm()) + 1e-20))
        out[0] = worst
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ring", "heads"])
def test_whole_model_sequence_parallel_train_step_world2(mode):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_model_worker, args=(2, _free_port(), out, mode), nprocs=2, join=True)
    assert len(out) == 1 and out[0] < 2e-2, dict(out)
