"""Data-parallel gradient exchange for the LoRA + caption-projection parameters.

The reference has no runnable multi-GPU path (its DeepSpeed loop is stale); the block
path shards by batch, so the only collective is one all-reduce of ~84 MB of gradients per step.
`GradBucketer` keeps every trainable parameter's `.grad` as a view into a few flat buckets (reverse
block order, so a bucket completes while earlier blocks are still in backward), and launches an
asynchronous NCCL all-reduce for a bucket from the post-accumulate-grad hook of its last parameter.
`finish()` waits for the collectives and rescales by 1/world_size.  Works with gloo on CPU for tests.

Gradient accumulation (training.py:199-206, `gradient_accumulation_steps > 1`): run every micro-step but the last
under `with bucketer.no_sync():` -- the hooks then only keep `.grad` attached to the buckets and nothing is
reduced; the last micro-step (outside the context) reduces the accumulated sums once.  A second synchronising
backward without `zero_grad()` in between raises instead of reducing a bucket twice.  An optimizer that calls
`zero_grad(set_to_none=True)` (training.py:207) detaches `.grad` from the buckets: the hook of the next backward
notices (address check), copies the fresh gradient into its bucket slot and re-attaches the view."""
import contextlib
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
        self._sync = True    # False inside no_sync(): accumulate locally, reduce nothing
        self._views = {}     # id(param) -> its slot in the flat bucket
        for b in self.buckets:
            total = sum(p.numel() for p in b["params"])
            b["flat"] = torch.zeros(total, dtype=b["dtype"], device=b["device"])
            off = 0
            for p in b["params"]:
                self._views[id(p)] = b["flat"][off:off + p.numel()].view_as(p)
                p.grad = self._views[id(p)]
                off += p.numel()
            b["pending"] = len(b["params"])
            b["reduced"] = False
        self._bucket_of = {id(p): b for b in self.buckets for p in b["params"]}
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for b in self.buckets for p in b["params"]]

    @contextlib.contextmanager
    def no_sync(self):
        """Micro-steps of a gradient-accumulation window except the last: gradients add up in the buckets locally."""
        prev, self._sync = self._sync, False
        try:
            yield self
        finally:
            self._sync = prev

    def _attach(self, p):
        """`.grad` must live in its bucket slot.  An optimizer's zero_grad(set_to_none=True) or an assignment to
        `.grad` breaks that: move the freshly accumulated gradient into the slot and re-attach the view."""
        view = self._views[id(p)]
        g = p.grad
        if g is None:
            p.grad = view
        elif g.data_ptr() != view.data_ptr():
            view.add_(g.reshape(view.shape).to(view.dtype))   # the slot may already hold earlier micro-steps
            p.grad = view

    def _on_grad(self, p):
        self._attach(p)
        if not self._sync:
            return
        b = self._bucket_of[id(p)]
        if b["reduced"]:
            raise RuntimeError("GradBucketer: a second synchronising backward before zero_grad(); run the earlier "
                               "micro-steps of an accumulation window under `with bucketer.no_sync():`")
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1 and self.overlap:
            b["reduced"] = True
            self._handles.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero_grad(self):
        """Keep `.grad` as bucket views (never set_to_none) and re-arm the ready counters."""
        for h in self._handles:   # never zero a bucket under a collective that is still in flight
            h.wait()
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["reduced"] = False
            for p in b["params"]:
                if p.grad is None or p.grad.data_ptr() != self._views[id(p)].data_ptr():
                    p.grad = self._views[id(p)]
        self._handles = []

    def finish(self):
        """Wait for the bucket all-reduces (launch any bucket whose hooks did not all fire) and average."""
        for b in self.buckets:
            for p in b["params"]:
                self._attach(p)
        if self.world > 1:
            for b in self.buckets:
                if not b["reduced"]:
                    b["reduced"] = True
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

    def close(self):
        """Remove the hooks (the parameters keep their current `.grad`)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def bytes_per_step(self) -> int:
        return sum(b["bytes"] for b in self.buckets)
