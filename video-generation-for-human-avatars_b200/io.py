"""Latent I/O for the block path: the reference trains from pre-encoded latent triples on disk
(ltx_video/dataset.py:46-97 -- `<encoder_dir>/<stem>.pt`, `<condition_dir>/<stem>.pt` (pose), `<condition_dir>/<stem>_ref.pt`,
each a dict with a "latents" tensor) collated into `{"latents", "pose_latents", "ref_image_latents", "stem"}` batches,
which `train_step` then moves to the device one tensor at a time on the compute stream (training.py:109-117).

Here the loader side keeps the reference's file layout and batch dict, and the device side is a feeder that keeps the
copies off the critical path: batches are cast to the model dtype on the host, staged in pinned memory and copied on a
dedicated copy stream into a small ring of device buffers while the previous step computes; the consumer gets device
tensors that are already resident (or, with `train.GraphedTrainStep`, has them copied straight into the graph's static
input buffers).  The reference has no sampler sharding, so data-parallel ranks shard the sample list themselves."""
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional

import torch

KEYS = ("latents", "pose_latents", "ref_image_latents")


def shard_indices(n_items: int, rank: int, world: int, epoch: int = 0, shuffle: bool = True, seed: int = 0) -> List[int]:
    """This rank's sample indices of one epoch: a seeded permutation dealt round-robin, truncated to equal length."""
    order = list(range(n_items))
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        order = torch.randperm(n_items, generator=g).tolist()
    per_rank = n_items // world
    return order[rank:per_rank * world:world]


class LatentTripleDataset(torch.utils.data.Dataset):
    """Same on-disk contract and item dict as the reference's LatentPairDataset (dataset.py:46-97)."""

    def __init__(self, condition_latents_dir: str, encoder_latents_dir: str):
        self.cond, self.enc = Path(condition_latents_dir), Path(encoder_latents_dir)
        stems = sorted(p.stem for p in self.enc.glob("*.pt") if not p.stem.endswith("_ref"))
        self.items = [s for s in stems if (self.cond / f"{s}.pt").exists() and (self.cond / f"{s}_ref.pt").exists()]

    def __len__(self):
        return len(self.items)

    @staticmethod
    def _load(path):
        return torch.load(path, map_location="cpu")["latents"].squeeze()

    def __getitem__(self, idx):
        stem = self.items[idx]
        ref = self._load(self.cond / f"{stem}_ref.pt")
        if ref.ndim == 3:               # [C, H, W] -> [C, 1, H, W]
            ref = ref.unsqueeze(1)
        return {"latents": self._load(self.enc / f"{stem}.pt"), "pose_latents": self._load(self.cond / f"{stem}.pt"),
                "ref_image_latents": ref, "stem": stem}


def collate_latent_triples(items: List[dict]) -> dict:
    out = {k: torch.stack([it[k] for it in items], dim=0) for k in KEYS}
    out["stem"] = [it["stem"] for it in items]
    return out


class DeviceFeeder:
    """Iterates `batches` (dicts as above, e.g. a torch DataLoader over either dataset class) and yields them as
    device-resident tensors of `dtype`, `depth` batches ahead of the consumer.

    Per batch: host cast -> pinned staging buffer -> `copy_(non_blocking=True)` on a private copy stream into device
    buffer `i % (depth + 1)` -> event.  `__next__` makes the current stream wait on that event (no host sync).  There
    is one buffer more than the look-ahead: the batch just handed out keeps its buffer while `depth` later ones are in
    flight, and that buffer is refilled only behind the event the consumer records with `batch["_release"]()` after the
    step that read it (a consumer that never releases costs a stream synchronize per batch instead of a race).  On a CPU
    device it degrades to a cast."""

    def __init__(self, batches: Iterable[dict], device, dtype=torch.bfloat16, depth: int = 2):
        self.src, self.device, self.dtype, self.depth = iter(batches), torch.device(device), dtype, max(1, depth)
        self.cuda = self.device.type == "cuda"
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.queue: List[tuple] = []
        self.n_slots = self.depth + 1
        self.slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.n_slots
        self.pinned: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.n_slots
        self.done: List[Optional[torch.cuda.Event]] = [None] * self.n_slots   # compute finished reading slot i
        self.released: List[bool] = [True] * self.n_slots                      # ... and said so (`_release`)
        self.users: List[Optional[torch.cuda.Stream]] = [None] * self.n_slots
        self.copied: List[Optional[torch.cuda.Event]] = [None] * self.n_slots  # the H2D copies out of pinned slot i finished
        self.sources: List[Optional[dict]] = [None] * self.n_slots
        self.n = 0
        for _ in range(self.depth):
            self._stage()

    def reset(self, batches: Iterable[dict]) -> "DeviceFeeder":
        """Start over on a new source (the next epoch) with the SAME pinned and device buffers: pinning host memory is
        a device-synchronising allocation of milliseconds, not something to repeat per epoch."""
        if self.queue:
            raise RuntimeError("DeviceFeeder.reset: the previous source has staged batches that were never consumed")
        self.src = iter(batches)
        for _ in range(self.depth):
            self._stage()
        return self

    def _buffers(self, store, slot, batch, pin):
        cur = store[slot]
        if cur is None or any(cur[k].shape != batch[k].shape for k in KEYS):
            cur = {k: (torch.empty(batch[k].shape, dtype=self.dtype).pin_memory() if pin else
                       torch.empty(batch[k].shape, dtype=self.dtype, device=self.device)) for k in KEYS}
            store[slot] = cur
        return cur

    def _stage(self):
        try:
            batch = next(self.src)
        except StopIteration:
            return
        slot = self.n % self.n_slots
        self.n += 1
        if not self.cuda:
            self.queue.append(({k: batch[k].to(self.dtype) for k in KEYS}, None, batch.get("stem"), slot))
            return
        # a batch that already sits in pinned memory in the target dtype is copied from where it is (the caller keeps it
        # unchanged until the batch has been yielded); anything else is cast into this slot's pinned staging buffers
        direct = all(batch[k].is_pinned() and batch[k].dtype == self.dtype and batch[k].is_contiguous() for k in KEYS)
        host = batch if direct else self._buffers(self.pinned, slot, batch, True)
        dev = self._buffers(self.slots, slot, batch, False)
        if self.copied[slot] is not None and not direct:
            self.copied[slot].synchronize()   # the pinned slot is about to be overwritten by the host: its DMA must be done
        if not self.released[slot] and self.users[slot] is not None:
            self.users[slot].synchronize()                        # handed out and never released: wait for its reader
            self.released[slot] = True
        with torch.cuda.stream(self.copy_stream):
            if self.done[slot] is not None:
                self.copy_stream.wait_event(self.done[slot])      # the step that used this slot has finished with it
            for k in KEYS:
                if not direct:
                    host[k].copy_(batch[k])                        # cast on the host, into pinned memory
                dev[k].copy_(host[k], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        self.copied[slot] = ready
        self.sources[slot] = batch if direct else None   # keeps a directly copied source alive while its DMA is in flight
        self.queue.append((dev, ready, batch.get("stem"), slot))

    def __iter__(self) -> Iterator[dict]:
        return self

    def __next__(self) -> dict:
        if not self.queue:
            raise StopIteration
        dev, ready, stem, slot = self.queue.pop(0)
        if ready is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ready)
            ev = torch.cuda.Event()
            self.done[slot], self.released[slot], self.users[slot] = ev, False, cur

            def release(ev=ev, cur=cur, slot=slot):                    # call after the step that consumed the batch
                ev.record(cur)
                self.released[slot] = True
            out = dict(dev)
            out["stem"] = stem
            out["_release"] = release
        else:
            out = dict(dev)
            out["stem"] = stem
            out["_release"] = lambda: None
        self._stage()
        return out
