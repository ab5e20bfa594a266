"""2-GPU check of optimizer-state sharding (optim.FusedAdamW(shard_group=...)): every rank must end up with exactly the
parameters a replicated FusedAdamW produces from the same gradients, while holding only its share of the moments; the
sharded step must also replay from a CUDA graph.  Run by tests/test_zero_gpu.py (spawned, NCCL) or directly:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/zero_checks.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401,E402  (puts the package on sys.path)


def _params(dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = [(2048, 128), (128, 2048), (2048,), (6, 2048), (512, 512), (4096, 64), (64,), (1000, 3), (2048, 128), (7,),
              (512, 512), (2048, 128)]
    out = []
    for i, s in enumerate(shapes):
        dt = torch.bfloat16 if i % 3 == 2 else torch.float32
        out.append(torch.nn.Parameter(torch.randn(s, generator=g).to(dt).to(dev)))
    return out


def _grads(params, step, dev):
    g = torch.Generator(device="cpu").manual_seed(100 + step)
    return [torch.randn(p.shape, generator=g).to(p.dtype).to(dev) for p in params]


def check(rank, world):
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("ZERO_CHECKS_DUMP_AFTER", "100")), exit=True)  # a hang names itself
    try:
        _check(rank, world)
    finally:
        faulthandler.cancel_dump_traceback_later()


def _check(rank, world):
    from b200_ltx.optim import FusedAdamW
    dev = torch.device("cuda", torch.cuda.current_device())
    ref_p, shard_p = _params(dev), _params(dev)
    ref = FusedAdamW(ref_p, lr=1e-2, weight_decay=0.05)
    sh = FusedAdamW(shard_p, lr=1e-2, weight_decay=0.05, shard_group=dist.group.WORLD)
    owned = [sh.owned(p) for p in shard_p]
    assert any(owned) and not all(owned), owned
    for step in range(3):
        gs = _grads(ref_p, step, dev)
        for p, q, g in zip(ref_p, shard_p, gs):
            p.grad, q.grad = g.clone(), g.clone()
        ref.step()
        sh.step()
        for i, (p, q) in enumerate(zip(ref_p, shard_p)):
            assert torch.equal(p.data, q.data), (rank, step, i)
    # only the owned tensors carry moments, and the shares add up to the replicated state
    mine = torch.tensor([float(sh.state_bytes())], device=dev)
    dist.all_reduce(mine)
    assert int(mine.item()) == ref.state_bytes(), (mine.item(), ref.state_bytes())
    assert 0 < sh.state_bytes() < 0.7 * ref.state_bytes(), (sh.state_bytes(), ref.state_bytes())
    assert all((("exp_avg" in sh.state[p]) == o) for p, o in zip(shard_p, owned))

    # the sharded step (update + coalesced broadcasts) replays from a CUDA graph: static gradient buffers
    gs = _grads(ref_p, 10, dev)
    for p, q, g in zip(ref_p, shard_p, gs):
        p.grad.copy_(g)
        q.grad.copy_(g)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sh.step()
        ref.step()          # keep the two in lock step through the warm-up
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        sh.step()
    for step in range(11, 14):
        gs = _grads(ref_p, step, dev)
        for p, q, g in zip(ref_p, shard_p, gs):
            p.grad.copy_(g)
            q.grad.copy_(g)
        ref.step()
        graph.replay()
    torch.cuda.synchronize()
    # capture does not execute: the replicated optimizer has done 3 + 1 (warm-up) + 3 steps, the sharded 3 + 1 + 3 replays
    for i, (p, q) in enumerate(zip(ref_p, shard_p)):
        assert torch.equal(p.data, q.data), (rank, "graph", i, (p.data.float() - q.data.float()).abs().max().item())
    del graph   # before the communicator goes: a live graph with captured NCCL kernels blocks ncclCommDestroy
    torch.cuda.synchronize()
    if rank == 0:
        print(f"zero_checks: {len(shard_p)} tensors over {world} ranks, state {sh.state_bytes()} of {ref.state_bytes()} "
              f"bytes on rank 0, eager and graph-replayed updates bit-identical to the replicated optimizer")


def _spawned(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        check(rank, world)
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)     # the peer may be parked in a collective: do not wait for it in destroy_process_group
    dist.destroy_process_group()


if __name__ == "__main__":
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    check(rank, world)
    dist.destroy_process_group()
