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
