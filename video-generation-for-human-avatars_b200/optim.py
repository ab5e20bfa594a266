"""AdamW for the trainable tensors of the LoRA fine-tune as ONE kernel launch.

Drop-in for the reference's `torch.optim.AdamW(params, lr=config.learning_rate)` (training.py:271; stepped at
training.py:206): same defaults (betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2, decoupled decay), state in the
parameter's dtype, `state_dict()` keys `step / exp_avg / exp_avg_sq` per parameter.  torch's fused implementation
walks the 116 tensors in ~8 multi-tensor launches (0.34 ms per step in the profile); here a device table of
(param, grad, exp_avg, exp_avg_sq) entries drives a single launch of `b200_adamw_step`.  The step counter and the
hyper-parameters are device tensors, so `step()` can be captured in a CUDA graph (train.GraphedTrainStep); after
changing `param_groups[i]["lr"]` (an LR schedule) under graph replay call `sync_hyperparameters()`.

Optimizer-state sharding for data-parallel runs (the reference's configs/ds_config_zero2.json:7-15 asks DeepSpeed for it;
`train_mode="full"` has 983 M trainable parameters = 7.9 GB of fp32 moments per replica): `FusedAdamW(params, ...,
shard_group=group)` deals the parameter tensors out to the ranks of `group` (largest first, onto the least loaded rank),
keeps `exp_avg` / `exp_avg_sq` only for the tensors a rank owns, updates those with the one launch, and broadcasts every
updated tensor from its owner (one NCCL broadcast per tensor: launch cost only matters eagerly -- under
train.GraphedTrainStep they are captured with the step).  Gradients are expected to be already averaged over the group
(dp.GradBucketer): every rank then applies the identical update a replicated optimizer would.  `state_dict()` of a sharded
optimizer holds the moments of the tensors THIS rank owns (a checkpoint is one file per rank, as with ZeRO); parameters
themselves are identical on every rank after each step."""
import struct
from typing import List

import torch

from . import lib as _lib
from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 shard_group=None, shard: bool = False):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}      # group index -> launch plan
        self._keepalive: List[torch.Tensor] = []  # pinned tables a captured graph re-reads on replay
        # optimizer-state sharding (ZeRO-1): parameter -> owning rank of `shard_group`
        self._shard_group, self._owner, self._rank, self._world = None, {}, 0, 1
        if (shard or shard_group is not None) and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(shard_group)
            if world > 1:
                self._shard_group, self._world = shard_group, world
                self._rank = torch.distributed.get_rank(shard_group)
                all_params = [p for g in self.param_groups for p in g["params"]]
                self._owner = {id(p): r for p, r in zip(all_params, partition_by_size([p.numel() for p in all_params], world))}

    def owned(self, p) -> bool:
        return self._world == 1 or self._owner[id(p)] == self._rank

    def state_bytes(self) -> int:
        """Bytes of exp_avg + exp_avg_sq held by THIS rank."""
        return sum(st[k].numel() * st[k].element_size() for st in self.state.values() for k in ("exp_avg", "exp_avg_sq")
                   if k in st)

    # ---- state ----
    def _group_state(self, gi: int, group):
        plan = self._plans.get(gi)
        if plan is None:
            first = next((p for p in group["params"]), None)
            if first is None:
                return None
            dev = first.device
            if dev.type != "cuda":
                raise _lib.B200Error("FusedAdamW: parameters must live on a CUDA device (there is no CPU path)")
            plan = self._plans[gi] = dict(dev=dev, ptrs=None, n_blocks=0, table=None, block_map=None,
                                          step=torch.zeros(1, device=dev, dtype=torch.float32),
                                          hyper=torch.empty(5, device=dev, dtype=torch.float32), hyper_host=None,
                                          done=torch.zeros(1, device=dev, dtype=torch.int32))
        return plan

    def _hyper_tuple(self, group):
        lr = group["lr"]
        return (float(lr), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                float(group["weight_decay"]))

    def sync_hyperparameters(self):
        """Push the param_groups' current lr / betas / eps / weight_decay to the device (needed by hand only when the
        step runs as a graph replay; an eager `step()` does it itself)."""
        for gi, group in enumerate(self.param_groups):
            plan = self._group_state(gi, group)
            if plan is None:
                continue
            h = self._hyper_tuple(group)
            if h != plan["hyper_host"]:
                plan["hyper"].copy_(torch.tensor(h, dtype=torch.float32), non_blocking=False)
                plan["hyper_host"] = h

    def _pinned(self, n_entries: int, n_blocks: int):
        return (torch.empty(48 * n_entries, dtype=torch.uint8).pin_memory(),
                torch.empty(2 * n_blocks, dtype=torch.int32).pin_memory())

    def _build(self, plan, entries):
        """entries: list of (p, g, m, v).  Packs the 48-byte table and the block map and ships them to the device.
        Host staging is pinned memory allocated OUTSIDE any stream capture: an eager rebuild reuses one slot with a
        blocking copy; a rebuild during capture (gradients re-allocated from the graph's pool) takes the spare slot
        prepared by an earlier eager step and never touches it again, because the captured copy re-reads it on every
        replay."""
        chunk = _lib.load().b200_adamw_chunk_elems()
        raw = bytearray()
        bmap = []
        for i, (p, g, m, v) in enumerate(entries):
            raw += struct.pack("<QQQQqii", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                               int(p.dtype == torch.bfloat16), 0)
            for c in range((p.numel() + chunk - 1) // chunk):
                bmap += [i, c]
        n_e, n_b = len(entries), len(bmap) // 2
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            spare = plan.get("spare")
            if spare is None or spare[0].numel() != 48 * n_e or spare[1].numel() != 2 * n_b:
                raise _lib.B200Error("FusedAdamW: run one eager step() with the same parameters before capturing "
                                     "(the pinned staging for the capture-time table is prepared there)")
            host_t, host_m = spare
            plan["spare"] = None
            self._keepalive += [host_t, host_m]
        else:
            slot = plan.get("slot")
            if slot is None or slot[0].numel() != 48 * n_e or slot[1].numel() != 2 * n_b:
                slot = plan["slot"] = self._pinned(n_e, n_b)
            if plan.get("spare") is None or plan["spare"][0].numel() != 48 * n_e or plan["spare"][1].numel() != 2 * n_b:
                plan["spare"] = self._pinned(n_e, n_b)
            host_t, host_m = slot
        host_t.copy_(torch.frombuffer(raw, dtype=torch.uint8))
        host_m.copy_(torch.tensor(bmap, dtype=torch.int32))
        # fresh device tensors on every rebuild: a captured graph keeps reading the ones it was captured with
        plan["table"] = torch.empty(host_t.numel(), device=plan["dev"], dtype=torch.uint8)
        plan["block_map"] = torch.empty(host_m.numel(), device=plan["dev"], dtype=torch.int32)
        plan["table"].copy_(host_t, non_blocking=capturing)
        plan["block_map"].copy_(host_m, non_blocking=capturing)
        plan["n_blocks"] = n_b

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plan = self._group_state(gi, group)
            if plan is None:
                continue
            entries = []
            for p in group["params"]:
                if p.grad is None or not self.owned(p):
                    continue
                if p.dtype not in (torch.float32, torch.bfloat16) or p.grad.dtype != p.dtype:
                    raise _lib.B200Error("FusedAdamW: fp32 or bf16 parameters with gradients of the same dtype")
                if not p.is_contiguous() or not p.grad.is_contiguous() or p.grad.is_sparse:
                    raise _lib.B200Error("FusedAdamW: parameters and gradients must be dense and contiguous")
                st = self.state[p]
                if not st:
                    st["step"] = plan["step"]           # shared by the group (torch keeps one per parameter)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                elif st["step"] is not plan["step"]:    # after load_state_dict: adopt the loaded count
                    plan["step"].copy_(torch.as_tensor(st["step"], dtype=torch.float32).reshape(1))
                    st["step"] = plan["step"]
                entries.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            if not entries:
                continue
            ptrs = tuple(t.data_ptr() for e in entries for t in e)
            if ptrs != plan["ptrs"]:     # first step, or gradients were re-allocated (zero_grad(set_to_none=True))
                self._build(plan, entries)
                plan["ptrs"] = ptrs
            h = self._hyper_tuple(group)
            if h != plan["hyper_host"]:
                if torch.cuda.is_current_stream_capturing():
                    raise _lib.B200Error("FusedAdamW: hyper-parameters changed during graph capture; call "
                                         "sync_hyperparameters() before capturing")
                plan["hyper"].copy_(torch.tensor(h, dtype=torch.float32))
                plan["hyper_host"] = h
            n = sum(e[0].numel() for e in entries)
            ops._call("adamw", 28.0 * n, "byte", _lib.load().b200_adamw_step, plan["table"].data_ptr(),
                      plan["block_map"].data_ptr(), plan["n_blocks"], plan["step"].data_ptr(), plan["hyper"].data_ptr(),
                      plan["done"].data_ptr(), ops._s())
        if self._world > 1:
            # every tensor travels from the rank that updated it (same order on every rank)
            dist = torch.distributed
            todo = []
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is None:
                        continue
                    src = self._owner[id(p)]
                    if self._shard_group is not None:
                        src = dist.get_global_rank(self._shard_group, src)
                    todo.append((p.data, src))
            if todo:
                # one NCCL group call for all of them (ncclGroupStart/End): a single launch instead of one per tensor
                with dist._coalescing_manager(group=self._shard_group, device=todo[0][0].device, async_ops=False):
                    for t, src in todo:
                        dist.broadcast(t, src, group=self._shard_group)
        return loss


def partition_by_size(sizes, world: int):
    """Owner rank per tensor: largest tensors first, each onto the currently least loaded rank (deterministic, identical
    on every rank).  Returns a list aligned with `sizes`."""
    load = [0] * world
    owner = [0] * len(sizes)
    for i in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += sizes[i]
    return owner
