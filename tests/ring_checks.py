"""Sequence-sharded (ring attn1) train step on P GPUs against the un-sharded b200 path and the fp32
oracle, on identical weights and inputs.  Run by tests/test_ring_gpu.py (spawned, NCCL) or directly:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/ring_checks.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run(rank, world, verbose=True, mode="ring"):
    import model_checks as mc
    import ref_block as rb
    from b200_ltx import api
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    # 6 frames x 4 x 7 = 168 tokens: 84 per rank at P=2 (ragged against the 128-row tiles), 5 frames do not
    # split on frame boundaries at P=4 either
    f, h, w = 6, 4, 7 * (world // 2 if world > 2 else 1)
    P = rb.init_params(cfg, 32, seed=0)
    nb = 1 if mode == "gather" else 2  # the all-gather mode is built for single long clips
    batch = rb.synthetic_batch(cfg, nb, f, h, w, 24, 1234, 15)
    for k in ("latents", "pose_latents", "ref_image_latents", "prompt_embeds", "noise"):
        batch[k] = batch[k].to(torch.bfloat16).float()
    P = {k: (v if "lora_" in k else v.to(torch.bfloat16).float()) for k, v in P.items()}
    t = torch.tensor([0.4, 0.73][:nb])
    l32, o32, g32 = mc.oracle_loss_grads(P, cfg, batch, t, torch.float32, dev)

    def one(sharded):
        model = mc.build_b200_model(cfg, P, 32, dev)
        if sharded:
            assert api.enable_sequence_parallel(model, mode=mode) is not None
        holder = {}
        root = model.base_model.model
        hk = root.register_forward_hook(lambda m, a, o: holder.__setitem__("out", o.sample.detach()))
        loss, grads = mc.b200_loss_grads(model, batch, t, device=dev)
        hk.remove()
        return loss, holder["out"], grads
    l_full, o_full, g_full = one(False)
    l_sp, o_sp, g_sp = one(True)
    # per-shard partial means -> average over the group (what dp.GradBucketer.finish does)
    for v in g_sp.values():
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        v /= world
    lsum = l_sp.clone().float()
    dist.all_reduce(lsum)
    lsum /= world
    n = o_full.shape[1] // world
    sl = slice(rank * n, (rank + 1) * n)
    worst = 0.0
    e_out = mc.rel(o_sp, o_full[:, sl])
    e_out32, e_full32 = mc.rel(o_sp, o32[:, sl]), mc.rel(o_full[:, sl], o32[:, sl])
    if verbose and rank == 0:
        print(f"  {mode} P={world}: velocity shard vs un-sharded rel {e_out:.3e}; vs fp32 oracle {e_out32:.3e} "
              f"(un-sharded {e_full32:.3e}); loss {float(lsum):.6f} vs {float(l_full):.6f} vs oracle {float(l32):.6f}")
    assert e_out32 <= max(2 * e_full32, mc.OUT_FLOOR), (e_out32, e_full32)
    assert abs(float(lsum) - float(l32)) / float(l32) < mc.OUT_FLOOR
    for k in sorted(g32):
        e_s, e_f = mc.rel(g_sp[k], g32[k]), mc.rel(g_full[k], g32[k])
        worst = max(worst, e_s)
        if verbose and rank == 0:
            print(f"  grad {k:56s} E_ring={e_s:.3e} E_unsharded={e_f:.3e}")
        assert e_s <= max(2 * e_f, mc.GRAD_FLOOR), (k, e_s, e_f)
    return worst


def _spawned(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        run(rank, world)
        run(rank, world, verbose=False, mode="gather")
        run(rank, world, verbose=False, mode="heads")
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)     # the peer may be parked in a collective: do not wait for it in destroy_process_group
    dist.destroy_process_group()


if __name__ == "__main__":
    r, w = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", r))))
    for m in ("ring", "gather", "heads"):
        run(int(os.environ.get("LOCAL_RANK", r)), w, mode=m)
    dist.destroy_process_group()
    if r == 0:
        print("ring_checks ok")
