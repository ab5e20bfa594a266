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
