"""Boundary contract on the reference's OWN classes, on CPU.

`api.install()` is applied to an instance of the reference's unmodified `Transformer3DModel` (plain and wrapped by
the reference's own `apply_training_strategy(..., "lora_audio")`), and the product's functional forwards
(modules.transformer_forward / block_forward / attention_forward) then run over plain-torch stand-ins of the raw
kernels (tests/torch_kernels.py) -- so every attribute the host logic reads from the reference's modules, the peft
attribute layout, the skip-layer paths and the gradient-checkpoint branch are exercised against the reference's own
forward.  Survival contract (torch_utils.py:66-102): deepcopy -> merge_and_unload -> state_dict on an installed
model.  Needs /root/reference (build container); the GPU box runs the mirror-class tests instead."""
import copy
import os
import sys

import pytest
import torch

import ref_block as rb
import ref_import
import torch_kernels as tk

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import make_golden as mg  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")

BF16 = torch.bfloat16


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def _reference_model(lora_rank, seed=0, layers=2):
    ns = ref_import.load()
    cfg = dict(mg.TINY, num_layers=layers)
    P = rb.init_params(cfg, lora_rank, seed=seed)
    model = mg.build_reference_model(ns, cfg, lora_rank, P).to(BF16)
    for n, p in model.named_parameters():           # peft keeps adapters in fp32 (autocast_adapter_dtype)
        if "lora_" in n:
            p.data = p.data.float()
    return ns, cfg, model.eval()


def _inputs(cfg, B=2, f=2, h=4, w=4, n_ctx=24, valid=15, seed=5, frac_coords=False):
    b = rb.synthetic_batch(cfg, B, f, h, w, n_ctx, seed, valid)
    tokens, coords = rb.patchify(b["latents"])
    if frac_coords:
        coords = coords.float()
        coords[:, 0] = coords[:, 0] * (1.0 / 25)
    return dict(hidden_states=tokens.to(BF16).contiguous(), indices_grid=coords,
                ref_image_hidden_states=b["ref_image_latents"].to(BF16),
                pose_hidden_states=b["pose_latents"].to(BF16),
                encoder_hidden_states=b["prompt_embeds"].expand(B, -1, -1).to(BF16).contiguous(),
                timestep=torch.tensor([0.4, 0.73][:B]),
                encoder_attention_mask=b["prompt_mask"].expand(B, -1))


def _run(model, inp, **kw):
    x = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inp.items()}
    with torch.no_grad():
        out = model(**x, return_dict=False, **kw)[0]
    return out, x["hidden_states"]


@pytest.mark.parametrize("lora_rank", [0, 32, 12, 96])
def test_install_on_reference_model_matches_reference_forward(lora_rank):
    from b200_ltx import api, modules
    ns, cfg, model = _reference_model(lora_rank)
    inp = _inputs(cfg)
    want, tok_ref = _run(model, inp)
    root = model.base_model.model if hasattr(model, "base_model") else model
    with tk.patched():
        api.install(model)
        assert isinstance(root.transformer_blocks[0].attn1.processor, modules.B200AttnProcessor)
        assert isinstance(root.transformer_blocks[1].attn2.get_processor(), modules.B200AttnProcessor)
        got, tok = _run(model, inp)
        # return_dict=True keeps the reference's output protocol (.sample and [0])
        x = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inp.items()}
        with torch.no_grad():
            o = model(**x)
        assert torch.equal(o.sample, got) and torch.equal(o[0], got)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert _rel(got, want) < 2e-2, _rel(got, want)
    assert torch.equal(tok, tok_ref) or _rel(tok, tok_ref) < 4e-3     # in-place conditioning lerp
    assert not torch.equal(tok, inp["hidden_states"])
    # uninstall gives the reference's own forwards back, bit for bit
    api.uninstall(model)
    for blk in root.transformer_blocks:
        blk.attn1.set_processor(ns.AttnProcessor2_0())
        blk.attn2.set_processor(ns.AttnProcessor2_0())
    again, _ = _run(model, inp)
    assert torch.equal(again, want)


@pytest.mark.parametrize("strategy", ["AttentionSkip", "AttentionValues", "TransformerBlock"])
def test_skip_layer_strategies_on_reference_model(strategy):
    """STG paths (attention.py:1071-1086, 312-319) through the host logic, sampling-style inputs: fractional
    coordinates, per-token timesteps [B, N], two conditions with the second one perturbed in block 1."""
    from b200_ltx import api
    ns, cfg, model = _reference_model(0, seed=3)
    inp = _inputs(cfg, frac_coords=True)
    N = inp["hidden_states"].shape[1]
    inp["timestep"] = torch.full((2, N), 0.9)
    inp["timestep"][:, :16] = 0.0                      # conditioned first frame
    skip = model.create_skip_layer_mask(1, 2, 1, [1])
    strat = getattr(ns.SkipLayerStrategy, strategy)
    want, _ = _run(model, inp, skip_layer_mask=skip, skip_layer_strategy=strat)
    plain, _ = _run(model, inp)
    assert not torch.equal(want[1], plain[1]) and _rel(want[0], plain[0]) < 1e-6   # only condition 1 is perturbed
    with tk.patched():
        api.install(model)
        got, _ = _run(model, inp, skip_layer_mask=skip, skip_layer_strategy=strat)
        # the mask the reference's own create_skip_layer_mask built carries no host-side note: every block applies it
        assert _rel(got, want) < 2e-2, _rel(got, want)
        assert _rel(got[1], want[1]) < 2e-2


def test_gradient_checkpoint_branch_on_reference_model():
    """transformer3d.py:503-534: training + gradient_checkpointing passes the block arguments positionally through
    torch.utils.checkpoint; the rebound block_forward must accept exactly that order."""
    from b200_ltx import api
    ns, cfg, model = _reference_model(32)
    root = model.base_model.model
    inp = _inputs(cfg)
    want, _ = _run(model, inp)
    with tk.patched():
        api.install(model)
        model.train()
        root.gradient_checkpointing = True
        got, _ = _run(model, inp)
        model.eval()
    assert _rel(got, want) < 2e-2


def test_installed_model_survives_deepcopy_merge_state_dict():
    """torch_utils.py:66-102 (save_lora_merged_weights): deepcopy(model) -> merge_and_unload() -> state_dict()."""
    from b200_ltx import api, modules
    ns, cfg, model = _reference_model(32)
    root = model.base_model.model
    inp = _inputs(cfg)
    with tk.patched():
        api.install(model)
        before, _ = _run(model, inp)                               # fills the weight / rope caches
        assert modules.side(root).get("wkv_all") is None or modules.side(root)["wkv_all"][1].numel() > 0
        assert not any(k.startswith("_b200") for k in root.__dict__)          # nothing cached on the modules
        assert not any(k.startswith("_b200") for k in root.transformer_blocks[0].attn1.__dict__)
        clone = copy.deepcopy(model)
        assert not modules._side_tables.get(clone.base_model.model)            # derived state is never copied
        merged = clone.merge_and_unload()
        sd = merged.state_dict()
        plain_keys = set(rb.param_shapes(cfg, 0))
        assert set(sd) == plain_keys, set(sd) ^ plain_keys                    # no lora_ / base_layer / cache keys
        assert all(v.dtype == BF16 for v in sd.values())
        # the merged copy still runs on the installed forwards and reproduces the adapted model
        out_merged, _ = _run(merged, inp)
        assert _rel(out_merged, before) < 1.5e-2
        # ... and the live model is untouched: same output, adapters still separate
        after, _ = _run(model, inp)
        assert torch.equal(after, before)
        assert any("lora_A" in k for k in model.state_dict())


def test_merge_then_forward_uses_merged_weights():
    """forward -> merge -> forward on the SAME object must not reuse cached pre-merge concatenations."""
    from b200_ltx import api, lora, modules
    cfg = dict(mg.TINY, num_layers=2)
    P = rb.init_params(cfg, 32, seed=1)
    with tk.patched():
        m = api.build_model(dict(api.LTXV_2B_CONFIG, **cfg), device="cpu")
        m = lora.apply_training_strategy(m, 32, 32)
        sd = {k: P[k.replace("base_model.model.", "").replace(".base_layer.", ".")].to(v.dtype)
              for k, v in m.state_dict().items()}
        m.load_state_dict(sd)
        m.eval()
        # caption tokens: B * L a multiple of 128 so that the batched K/V cache path is taken
        inp = _inputs(cfg, B=2, n_ctx=64, valid=40)
        with_adapters, _ = _run(m, inp)
        assert modules.side(m.base_model.model).get("wkv_all") is not None
        base = m.merge_and_unload()
        merged, _ = _run(base, inp)
        assert _rel(merged, with_adapters) < 1.5e-2, _rel(merged, with_adapters)
        # control: dropping the adapters WITHOUT merging changes the output by much more than that
        m2 = api.build_model(dict(api.LTXV_2B_CONFIG, **cfg), device="cpu")
        m2.load_state_dict({k: v for k, v in P.items() if "lora_" not in k}, strict=True)
        no_lora, _ = _run(m2.to(BF16).eval(), inp)
        assert _rel(no_lora, with_adapters) > 3 * _rel(merged, with_adapters)


@pytest.mark.parametrize("strategy", ["AttentionSkip", "AttentionValues"])
def test_fused_stg_skip_host_logic_matches_reference(strategy):
    """The mirror model's own create_skip_layer_mask marks its rows as 0 / 1, so the attention-level STG skips ride
    inside the attention launch (batch_keep / pass-through source): host logic against the reference's forward."""
    from b200_ltx import api, modules
    ns, cfg, ref_model = _reference_model(0, seed=3, layers=3)
    inp = _inputs(cfg, frac_coords=True)
    want, _ = _run(ref_model, inp, skip_layer_mask=ref_model.create_skip_layer_mask(1, 2, 1, [0, 2]),
                   skip_layer_strategy=getattr(ns.SkipLayerStrategy, strategy))
    with tk.patched():
        m = api.build_model(dict(api.LTXV_2B_CONFIG, **cfg), device="cpu")
        m.load_state_dict({k: v.to(BF16) for k, v in ref_model.state_dict().items()}, strict=True)
        m = m.eval()
        calls = []
        orig = tk.fa_fwd

        def spy(*a, **k):
            calls.append(k.get("batch_keep") is not None)
            return orig(*a, **k)
        from b200_ltx import ops
        saved = ops.fa_fwd
        ops.fa_fwd = spy
        try:
            got, _ = _run(m, inp, skip_layer_mask=m.create_skip_layer_mask(1, 2, 1, [0, 2]),
                          skip_layer_strategy=getattr(modules.SkipLayerStrategy, strategy))
        finally:
            ops.fa_fwd = saved
    assert _rel(got, want) < 2e-2, _rel(got, want)
    # attn1 of blocks 0 and 2 took the in-launch skip, block 1 (all-ones row) and every attn2 did not
    assert sum(calls) == 2 and len(calls) == 6


@pytest.mark.parametrize("B", [1, 2])
def test_product_train_step_on_reference_model_matches_reference_train_step(B):
    """The product's train_step (train.py) driving an INSTALLED instance of the reference's own peft-wrapped model, with
    the product's scheduler and patchifier, against the reference's own train_step (training.py:94-166) on the same
    model, batch and generator state: same four return values (loss, rel_mse, nrmse, {"transformer_mse"}) within the
    bf16 tolerance -- timestep draw, noise draw, noising, velocity target, the in-place conditioning of the noisy
    tokens and the metrics all follow the reference's order."""
    from b200_ltx import api
    from b200_ltx.modules import SymmetricPatchifier
    from b200_ltx.scheduler import RectifiedFlowScheduler
    from b200_ltx.train import train_step
    ns, cfg, model = _reference_model(32)
    tr = ref_import.load_training()
    b = rb.synthetic_batch(cfg, B, 2, 4, 4, 24, 9, 15)
    batch = {k: b[k] for k in ("latents", "ref_image_latents", "pose_latents")}

    class Cfg:
        rf_log_normal_mu, rf_log_normal_sigma = -0.5, 1.0
        rf_quantile_min, rf_quantile_max = 0.005, 0.999
        transformer_loss_weight = 0.75
    model.train()
    torch.manual_seed(21)
    want = tr.train_step(model, batch, ns.RectifiedFlowScheduler(), ns.SymmetricPatchifier(patch_size=1), Cfg(),
                         b["prompt_embeds"], b["prompt_mask"], device=torch.device("cpu"))
    with tk.patched():
        api.install(model)
        torch.manual_seed(21)
        got = train_step(model, batch, RectifiedFlowScheduler(), SymmetricPatchifier(1), Cfg(), b["prompt_embeds"],
                         b["prompt_mask"], device=torch.device("cpu"))
        api.uninstall(model)
    assert len(got) == len(want) == 4 and set(got[3]) == set(want[3]) == {"transformer_mse"}
    assert all(x.dim() == 0 for x in got[:3])
    for name, g, w in (("loss", got[0], want[0]), ("rel_mse", got[1], want[1]), ("nrmse", got[2], want[2]),
                       ("transformer_mse", got[3]["transformer_mse"], want[3]["transformer_mse"])):
        # (the reference's dict value is a python float -- an .item() host sync per step, training.py:162; the product
        # keeps a 0-d tensor, which the caller's `float(v)` at logging time, training.py:218-219, accepts)
        g, w = [float(x.detach()) if torch.is_tensor(x) else float(x) for x in (g, w)]
        assert abs(g - w) <= 2e-2 * abs(w), (name, g, w)


@pytest.mark.parametrize("train_mode,lora_rank", [("lora_audio", 32), ("lora_audio", 12), ("full", 0)])
def test_product_backward_on_reference_model_matches_reference_autograd(train_mode, lora_rank):
    """Forward AND backward of the installed path on the reference's own model instance (the reference's own
    apply_training_strategy decides what trains): every gradient the reference's autograd produces is produced by the
    product's autograd Functions (LinearFn / CtxKVFn / SelfAttnFn / FeedForwardFn / NormMod*Fn with their side channels)
    with the same name, shape and dtype, and agrees within the bf16 tolerance.  Kernels = plain-torch stand-ins."""
    from b200_ltx import api
    from b200_ltx.train import rf_mse_loss
    ns = ref_import.load()
    cfg = dict(mg.TINY, num_layers=2)
    P = rb.init_params(cfg, lora_rank, seed=2)
    model = mg.build_reference_model(ns, cfg, lora_rank, P, train_mode=train_mode).to(BF16)
    for n, p in model.named_parameters():
        if "lora_" in n:
            p.data = p.data.float()
    model.train()
    inp = _inputs(cfg, seed=8)
    target = torch.randn(inp["hidden_states"].shape, generator=torch.Generator().manual_seed(3)).to(BF16)

    def grads(loss_fn):
        model.zero_grad(set_to_none=True)
        x = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inp.items()}
        out = model(**x, return_dict=False)[0]
        loss_fn(out, target).backward()
        return out.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    want_out, want = grads(torch.nn.functional.mse_loss)
    with tk.patched():
        api.install(model)
        got_out, got = grads(rf_mse_loss)
        api.uninstall(model)
    assert _rel(got_out, want_out) < 2e-2
    assert set(got) == set(want) and len(want) >= 8
    worst = ("", 0.0)
    for n in want:
        assert got[n].shape == want[n].shape and got[n].dtype == want[n].dtype, n
        wn = float(want[n].float().norm())
        if wn == 0.0:
            assert float(got[n].float().norm()) == 0.0, n
            continue
        e = _rel(got[n], want[n])
        worst = max(worst, (n, e), key=lambda t: t[1])
    print(f"worst gradient: {worst[0]} rel err {worst[1]:.3e} over {len(want)} gradients")
    assert worst[1] < 5e-2, worst


def test_gradient_checkpoint_backward_on_reference_model_equals_plain_backward():
    """transformer3d.py:503-534 with training + gradient_checkpointing: torch.utils.checkpoint(use_reentrant=False)
    re-runs every block's forward during the backward.  The fused path hands gradients around outside autograd's view
    (GradJoin parks the attn2 residual-branch gradient between two LinearFn nodes; NormModResFn returns the residual as
    an aliased second output): the recomputation must not double-count or lose any of it.  Same gradients, to the last
    bit, with and without checkpointing, on the reference's own model class."""
    from b200_ltx import api
    from b200_ltx.train import rf_mse_loss
    ns, cfg, model = _reference_model(32, seed=4)
    root = model.base_model.model
    model.train()
    inp = _inputs(cfg, seed=6)
    target = torch.randn(inp["hidden_states"].shape, generator=torch.Generator().manual_seed(1)).to(BF16)

    def grads():
        model.zero_grad(set_to_none=True)
        x = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inp.items()}
        out = model(**x, return_dict=False)[0]
        rf_mse_loss(out, target).backward()
        return out.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    with tk.patched():
        api.install(model)
        out_plain, plain = grads()
        root.gradient_checkpointing = True
        out_ckpt, ckpt = grads()
        root.gradient_checkpointing = False
        api.uninstall(model)
    assert torch.equal(out_plain, out_ckpt)
    assert set(plain) == set(ckpt) and len(plain) == 20
    for n in plain:
        assert torch.equal(plain[n], ckpt[n]), (n, _rel(ckpt[n], plain[n]))


@pytest.mark.parametrize("train_mode,lora_rank", [("lora_audio", 32), ("lora_audio", 8), ("full", 0)])
def test_mirror_training_strategy_names_shapes_and_trainable_set_equal_the_references(train_mode, lora_rank):
    """What a checkpoint, an optimizer and a `"lora_" in name` filter see: the mirror model after the product's
    apply_training_strategy (lora.py) against the reference's model after its own apply_training_strategy
    (training.py:42-91, peft stand-in): identical parameter names (incl. the `base_model.model.` prefix and
    `.base_layer.` / `.lora_A.default.` parts), shapes, dtypes after `.to(bfloat16)` (adapters stay fp32), the same
    trainable set, the same state_dict keys, and strict loading of each other's state dict."""
    from b200_ltx import api, lora
    ns = ref_import.load()
    cfg = dict(mg.TINY, num_layers=2)
    P = rb.init_params(cfg, lora_rank, seed=0)
    ref_model = mg.build_reference_model(ns, cfg, lora_rank, P, train_mode=train_mode).to(BF16)
    for n, p in ref_model.named_parameters():
        if "lora_" in n:
            p.data = p.data.float()
    mirror = api.build_model(dict(api.LTXV_2B_CONFIG, **cfg), device="cpu")
    mirror = lora.apply_training_strategy(mirror, lora_rank, lora_rank, train_mode=train_mode)
    want = {n: (tuple(p.shape), p.dtype, p.requires_grad) for n, p in ref_model.named_parameters()}
    got = {n: (tuple(p.shape), p.dtype, p.requires_grad) for n, p in mirror.named_parameters()}
    assert set(got) == set(want), sorted(set(got) ^ set(want))[:6]
    for n in want:
        assert got[n] == want[n], (n, got[n], want[n])
    assert list(dict(mirror.named_parameters())) == list(dict(ref_model.named_parameters()))     # same ORDER too
    assert set(mirror.state_dict()) == set(ref_model.state_dict())
    mirror.load_state_dict(ref_model.state_dict(), strict=True)
    ref_model.load_state_dict(mirror.state_dict(), strict=True)
    n_train = sum(r for _, _, r in want.values())
    assert n_train == (20 if train_mode == "lora_audio" else 55)
