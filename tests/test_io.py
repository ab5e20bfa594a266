"""Latent I/O: on-disk triples -> batch dicts -> device feeder; against the reference's own dataset
class where /root/reference is mounted."""
import os

import pytest
import torch

import ref_import


def _write_triples(tmp, n=5, c=8, f=3, h=4, w=4):
    enc, cond = tmp / "enc", tmp / "cond"
    enc.mkdir()
    cond.mkdir()
    g = torch.Generator().manual_seed(0)
    for i in range(n):
        torch.save({"latents": torch.randn(1, c, f, h, w, generator=g)}, enc / f"clip{i:02d}.pt")
        torch.save({"latents": torch.randn(1, c, f, h, w, generator=g)}, cond / f"clip{i:02d}.pt")
        ref = torch.randn(1, c, 1, h, w, generator=g) if i % 2 else torch.randn(c, h, w, generator=g)
        torch.save({"latents": ref}, cond / f"clip{i:02d}_ref.pt")
    torch.save({"latents": torch.randn(1, c, f, h, w)}, enc / "orphan.pt")        # no condition files: skipped
    torch.save({"latents": torch.randn(1, c, 1, h, w)}, enc / "clip00_ref.pt")    # *_ref in the encoder dir: skipped
    return str(cond), str(enc)


def test_dataset_collate_and_feeder_cpu(tmp_path):
    from b200_ltx import api
    cond, enc = _write_triples(tmp_path)
    ds = api.LatentTripleDataset(cond, enc)
    assert len(ds) == 5 and ds.items[0] == "clip00"
    it = ds[0]
    assert it["latents"].shape == (8, 3, 4, 4) and it["ref_image_latents"].shape == (8, 1, 4, 4)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, collate_fn=api.collate_latent_triples, drop_last=True)
    batches = list(loader)
    assert batches[0]["latents"].shape == (2, 8, 3, 4, 4) and batches[0]["stem"] == ["clip00", "clip01"]
    fed = list(api.DeviceFeeder(loader, "cpu", dtype=torch.bfloat16, depth=2))
    assert len(fed) == len(batches) == 2
    for a, b in zip(fed, batches):
        for k in ("latents", "pose_latents", "ref_image_latents"):
            assert a[k].dtype == torch.bfloat16 and torch.equal(a[k], b[k].to(torch.bfloat16))
        assert a["stem"] == b["stem"]
        a["_release"]()
    # data-parallel sharding: disjoint, equal-sized, covering
    parts = [api.shard_indices(11, r, 4, epoch=3) for r in range(4)]
    assert all(len(p) == 2 for p in parts) and len(set(sum(parts, []))) == 8
    assert api.shard_indices(11, 1, 4, epoch=3) == parts[1] and api.shard_indices(11, 1, 4, epoch=4) != parts[1]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_dataset_matches_reference_dataset(tmp_path):
    """Same items, same tensors, same collated batch as ltx_video/dataset.py:5-97."""
    ref_import._prepare()
    from ltx_video.dataset import LatentPairDataset, collate_latent_pairs
    from b200_ltx import api
    cond, enc = _write_triples(tmp_path)
    ours, ref = api.LatentTripleDataset(cond, enc), LatentPairDataset(cond, enc)
    assert ours.items == ref.items
    a = api.collate_latent_triples([ours[i] for i in range(len(ours))])
    b = collate_latent_pairs([ref[i] for i in range(len(ref))])
    assert set(a) == set(b)
    for k in b:
        assert (a[k] == b[k]) if k == "stem" else torch.equal(a[k], b[k])


@pytest.mark.gpu
def test_device_feeder_overlapped_copies(tmp_path):
    from b200_ltx import api
    cond, enc = _write_triples(tmp_path, n=8, c=128, f=3, h=8, w=8)
    ds = api.LatentTripleDataset(cond, enc)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, collate_fn=api.collate_latent_triples)
    want = list(loader)
    acc = []
    for got, ref in zip(api.DeviceFeeder(loader, "cuda", depth=2), want):
        assert got["latents"].is_cuda and got["latents"].dtype == torch.bfloat16
        # a "step" that reads the batch after some queued work, then releases the slot
        x = torch.zeros(1 << 22, device="cuda").add_(1.0)
        acc.append((got["latents"].float().sum() + x[0], got["pose_latents"].float().clone(), ref))
        got["_release"]()
    torch.cuda.synchronize()
    assert len(acc) == 4
    for s, pose, ref in acc:
        assert torch.equal(pose.cpu(), ref["pose_latents"].to(torch.bfloat16).float())
        assert abs(float(s) - 1.0 - float(ref["latents"].to(torch.bfloat16).float().sum())) < 0.5


def test_device_feeder_reset_reuses_buffers_cpu():
    from b200_ltx import api
    mk = lambda v: {"latents": torch.full((1, 2, 1, 2, 2), float(v)), "pose_latents": torch.zeros(1, 2, 1, 2, 2),
                    "ref_image_latents": torch.zeros(1, 2, 1, 2, 2)}
    f = api.DeviceFeeder([mk(1), mk(2)], "cpu", depth=2)
    assert [float(b["latents"].flatten()[0]) for b in f] == [1.0, 2.0]
    assert [float(b["latents"].flatten()[0]) for b in f.reset([mk(3), mk(4), mk(5)])] == [3.0, 4.0, 5.0]
    f.reset([mk(6)])
    import pytest
    with pytest.raises(RuntimeError):
        f.reset([mk(7)])
