"""world_size-2 gloo test of the bucketed gradient all-reduce (b200_ltx.dp.GradBucketer) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, deferred=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_ltx.dp import GradBucketer
    torch.manual_seed(0)
    # stand-in for "LoRA fp32 + caption projection bf16->fp32" parameter groups of different dtypes
    lin = [torch.nn.Linear(16, 16) for _ in range(6)]
    extra = torch.nn.Linear(16, 4).to(torch.float64)
    params = [(f"l{i}.{n}", p) for i, m in enumerate(lin) for n, p in m.named_parameters()]
    params += [(f"e.{n}", p) for n, p in extra.named_parameters()]
    bk = GradBucketer(params, bucket_bytes=2048)
    assert len(bk.buckets) >= 3
    bk.overlap = not deferred
    for step in range(2):
        bk.zero_grad()
        x = torch.full((4, 16), float(rank + 1 + step))
        h = x
        for m in lin:
            h = torch.tanh(m(h))
        loss = extra(h.double()).sum()
        loss.backward()
        if deferred:
            bk.reduce_now()   # the graph-replayed step: hooks only counted, one reduction after the backward
        else:
            bk.finish()
    flat = torch.cat([p.grad.flatten().double() for _, p in params])
    # reference: average of the per-rank grads computed without the bucketer
    bk.close()
    ref = []
    for r in range(world):
        for _, p in params:
            p.grad = None
        x = torch.full((4, 16), float(r + 1 + 1))
        h = x
        for m in lin:
            h = torch.tanh(m(h))
        extra(h.double()).sum().backward()
        ref.append(torch.cat([p.grad.flatten().double() for _, p in params]))
    ref = sum(ref) / world
    out[rank] = float((flat - ref).abs().max())
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert len(out) == 2 and all(v < 1e-6 for v in out.values()), dict(out)


def test_deferred_allreduce_world2():
    """overlap = False + reduce_now(): the mode train.GraphedTrainStep uses between its two CUDA graphs."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out, True), nprocs=2, join=True)
    assert len(out) == 2 and all(v < 1e-6 for v in out.values()), dict(out)


def _accum_worker(rank, world, port, out):
    """Two accumulation windows of two micro-steps each; between them the 'optimizer' drops `.grad`
    (zero_grad(set_to_none=True), training.py:207) instead of calling the bucketer's zero_grad."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_ltx.dp import GradBucketer
    torch.manual_seed(0)
    lin = [torch.nn.Linear(8, 8) for _ in range(4)]
    params = [(f"l{i}.{n}", p) for i, m in enumerate(lin) for n, p in m.named_parameters()]
    bk = GradBucketer(params, bucket_bytes=512)

    def run(x):
        h = x
        for m in lin:
            h = torch.tanh(m(h))
        h.sum().backward()

    def inputs(r, window):
        return [torch.full((2, 8), float(r + 1 + 3 * window + k)) for k in range(2)]

    worst = 0.0
    for window in range(2):
        if window == 0:
            bk.zero_grad()
        else:
            for _, p in params:
                p.grad = None          # what torch.optim's zero_grad(set_to_none=True) leaves behind
            bk.zero_grad()             # clears the flat buckets; the hooks re-attach the views
            for _, p in params[::2]:
                p.grad = None          # and a caller that drops some of them again after that
        xs = inputs(rank, window)
        with bk.no_sync():
            run(xs[0])
        run(xs[1])
        bk.finish()
        got = torch.cat([p.grad.flatten() for _, p in params]).clone()
        # a second synchronising backward inside the same window must raise, not double-reduce
        if window == 0:
            try:
                run(xs[1])
                raised = False
            except RuntimeError:
                raised = True
            assert raised
        # reference: plain autograd on detached copies of the layers
        ref = 0
        for r in range(world):
            import copy
            lin2 = copy.deepcopy(lin)
            for m in lin2:
                m.zero_grad(set_to_none=True)
            for x in inputs(r, window):
                h = x
                for m in lin2:
                    h = torch.tanh(m(h))
                h.sum().backward()
            ref = ref + torch.cat([p.grad.flatten() for m in lin2 for p in m.parameters()])
        worst = max(worst, float((got - ref / world).abs().max()))
    out[rank] = worst
    dist.destroy_process_group()


def test_gradient_accumulation_and_set_to_none_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_accum_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert len(out) == 2 and all(v < 1e-6 for v in out.values()), dict(out)


def test_optimizer_state_partition_is_balanced_and_deterministic():
    from b200_ltx.optim import partition_by_size
    # train_mode="full" of LTXV-2B: per block 8 attention matrices [2048, 2048] + small vectors, plus a few large ones
    sizes = []
    for _ in range(28):
        sizes += [2048 * 2048] * 8 + [2048] * 10 + [6 * 2048]
    sizes += [4096 * 2048, 2048 * 2048, 2048 * 6 * 2048, 128 * 2048, 2 * 2048]
    for world in (2, 4, 8):
        owner = partition_by_size(sizes, world)
        assert owner == partition_by_size(sizes, world)
        load = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(world)]
        assert max(load) - min(load) <= max(sizes)            # within one largest tensor of each other
        assert max(load) <= 1.05 * sum(sizes) / world


def _model_worker(rank, world, port, out, accumulate):
    """The product's own mirror model and train_step (real autograd Functions, plain-torch stand-ins for the kernels)
    under the GradBucketer: rank r trains on batch r; the bucketed result must equal the average of the per-batch
    gradients computed without any bucketer.  LoRA gradients reach `.grad` as the Functions' own fp32 outputs and the
    caption projection's as bf16 -- two bucket dtypes, hooks firing in the backward's order, the LoRA weight-gradient
    GEMMs meeting an existing `.grad` (bucket view) instead of a stolen tensor."""
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
    from b200_ltx.dp import GradBucketer
    torch.set_num_threads(2)
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=2, cross_attention_dim=128, caption_channels=64)
    P = rb.init_params(cfg, 8, seed=1)

    class Cfg:
        rf_log_normal_mu, rf_log_normal_sigma = -0.5, 1.0
        rf_quantile_min, rf_quantile_max = 0.005, 0.999
        transformer_loss_weight = 1.0

    def batch_of(r, k):
        b = rb.synthetic_batch(cfg, 1, 2, 4, 4, 16, 100 + 10 * r + k, 9)
        return b, torch.tensor([0.2 + 0.3 * r + 0.1 * k])

    def backward(model, r, k):
        b, t = batch_of(r, k)
        loss = api.train_step(model, {n: b[n] for n in ("latents", "ref_image_latents", "pose_latents")},
                              api.RectifiedFlowScheduler(), api.SymmetricPatchifier(1), Cfg(), b["prompt_embeds"],
                              b["prompt_mask"], device=torch.device("cpu"), t=t, noise=b["noise"].to(torch.bfloat16))[0]
        loss.backward()

    micro = 2 if accumulate else 1
    with tk.patched():
        model = mc.build_b200_model(cfg, P, 8, device="cpu").train()
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        assert {p.dtype for _, p in named} == {torch.float32, torch.bfloat16}
        bk = GradBucketer(named, bucket_bytes=16 << 10)
        assert len(bk.buckets) >= 3
        bk.zero_grad()
        for k in range(micro):
            if k + 1 < micro:
                with bk.no_sync():
                    backward(model, rank, k)
            else:
                backward(model, rank, k)
        bk.finish()
        got = {n: p.grad.detach().float().clone() for n, p in named}
        bk.close()
        ref = {n: torch.zeros_like(g) for n, g in got.items()}
        for r in range(world):
            for _, p in named:
                p.grad = None
            for k in range(micro):
                backward(model, r, k)
            for n, p in named:
                ref[n] += p.grad.detach().float() / world
    worst = 0.0
    for n in got:
        denom = float(ref[n].norm()) + 1e-20
        worst = max(worst, float((got[n] - ref[n]).norm()) / denom)
    out[rank] = worst
    dist.destroy_process_group()


def _run_model_workers(accumulate):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_model_worker, args=(2, _free_port(), out, accumulate), nprocs=2, join=True)
    # bf16 buckets: the sum over ranks is rounded in bf16 where the reference sums fp32 copies
    assert len(out) == 2 and all(v < 1e-2 for v in out.values()), dict(out)
    assert abs(out[0] - out[1]) < 1e-12     # both ranks hold the same averaged gradients


def test_whole_model_train_step_bucketed_world2():
    _run_model_workers(accumulate=False)


def test_whole_model_gradient_accumulation_bucketed_world2():
    _run_model_workers(accumulate=True)
