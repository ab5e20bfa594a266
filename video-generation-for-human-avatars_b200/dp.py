"""Data-parallel gradient exchange for the LoRA + caption-projection parameters.

The reference has no runnable multi-GPU path (its DeepSpeed loop is stale); the block
path shards by batch, so the only collective is one all-reduce of ~84 MB of gradients per step.
`GradBucketer` keeps every trainable parameter's `.grad` as a view into a few flat buckets (reverse
block order, so a bucket completes while earlier blocks are still in backward), and launches an
asynchronous NCCL all-reduce for a bucket from the post-accumulate-grad hook of its last parameter.
`finish()` waits for the collectives and rescales by 1/world_size.  Works with gloo on CPU for tests."""
from typing import List

import torch
import torch.distributed as dist


class GradBucketer:
    def __init__(self, named_params, process_group=None, bucket_bytes: int = 24 << 20):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [(n, p) for n, p in named_params if p.requires_grad]
        # autograd finishes the last block first: order buckets by reverse registration order
        params = list(reversed(params))
        self.buckets: List[dict] = []
        by_dtype = {}
        for n, p in params:
            cur = by_dtype.get(p.dtype)
            nbytes = p.numel() * p.element_size()
            if cur is None or cur["bytes"] + nbytes > bucket_bytes:
                cur = {"params": [], "bytes": 0, "dtype": p.dtype, "device": p.device}
                by_dtype[p.dtype] = cur
                self.buckets.append(cur)
            cur["params"].append(p)
            cur["bytes"] += nbytes
        self._handles = []
        self.overlap = True  # False: hooks only count (GraphedTrainStep reduces after the captured backward)
        for b in self.buckets:
            total = sum(p.numel() for p in b["params"])
            b["flat"] = torch.zeros(total, dtype=b["dtype"], device=b["device"])
            off = 0
            for p in b["params"]:
                p.grad = b["flat"][off:off + p.numel()].view_as(p)
                off += p.numel()
            b["pending"] = len(b["params"])
        self._bucket_of = {id(p): b for b in self.buckets for p in b["params"]}
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for b in self.buckets for p in b["params"]]

    def _on_grad(self, p):
        b = self._bucket_of[id(p)]
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1 and self.overlap:
            self._handles.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero_grad(self):
        """Keep `.grad` as bucket views (never set_to_none) and re-arm the ready counters."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
        self._handles = []

    def finish(self):
        """Wait for the bucket all-reduces (launch any bucket whose hooks did not all fire) and average."""
        if self.world > 1:
            for b in self.buckets:
                if b["pending"] != 0:
                    self._handles.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group,
                                                         async_op=True))
            for h in self._handles:
                h.wait()
            for b in self.buckets:
                b["flat"].div_(self.world)
        self._handles = []

    def reduce_now(self):
        """All-reduce and average every bucket at once (no overlap): used between the two CUDA graphs of a
        graph-replayed step, where the Python hooks do not run."""
        if self.world > 1:
            hs = [dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for b in self.buckets]
            for h in hs:
                h.wait()
            for b in self.buckets:
                b["flat"].div_(self.world)

    def bytes_per_step(self) -> int:
        return sum(b["bytes"] for b in self.buckets)
