"""GPU parity of the whole path (train step: forward, loss, backward) against the oracle and the
committed golden vectors."""
import os

import pytest
import torch

import model_checks as mc
import ref_block as rb

TINY_CASE = dict(b=2, f=3, h=4, w=8, n_ctx=24, valid_ctx=15, lora_rank=32, seed_w=0, seed_x=1234, t=[0.4, 0.73])


@pytest.mark.gpu
def test_tiny_train_step_vs_oracle_and_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_train_fp32.pt"))
    assert g["case"] == TINY_CASE
    res = mc.run_parity(g["cfg"], TINY_CASE)
    # and directly against the committed reference output (generated from the reference's own modules)
    P = rb.init_params(g["cfg"], 32, seed=0)
    batch = rb.synthetic_batch(g["cfg"], 2, 3, 4, 8, 24, 1234, 15)
    model = mc.build_b200_model(g["cfg"], P, 32)
    loss, grads = mc.b200_loss_grads(model, batch, torch.tensor(TINY_CASE["t"]))
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 2e-2
    for k, v in g["grads"].items():
        assert mc.rel(grads[k], v.cuda()) < 6e-2, k
    assert res["velocity output"][0] < 3e-2


@pytest.mark.gpu
def test_full_width_two_blocks():
    """D=2048 / 32 heads / FF 8192 / caption 4096, 2 blocks, 512 + 128 ragged tokens, 15 of 256 keys valid."""
    cfg = dict(rb.LTXV_2B, num_layers=2)
    case = dict(b=1, f=5, h=8, w=16, n_ctx=256, valid_ctx=15, lora_rank=32, seed_w=1, seed_x=7, t=[0.4])
    mc.run_parity(cfg, case)


@pytest.mark.gpu
def test_cfg2_token_count_full_width_two_blocks():
    """BASELINE config 2 geometry (121x512x768 -> 16x16x24 = 6144 latent tokens, 256 caption tokens with 15 valid,
    D = 2048 / 32 heads / FF 8192) on 2 blocks: the CTA-pair GEMMs, the multi-wave attention grids and the batched
    caption K/V path at their real sizes, against the fp32 oracle on the same GPU."""
    cfg = dict(rb.LTXV_2B, num_layers=2)
    case = dict(b=1, f=16, h=16, w=24, n_ctx=256, valid_ctx=15, lora_rank=32, seed_w=6, seed_x=17, t=[0.35])
    mc.run_parity(cfg, case)


@pytest.mark.gpu
@pytest.mark.parametrize("lora_rank", [32, 8, 12, 96])
def test_batched_caption_kv_path(lora_rank):
    """(B * caption tokens) % 128 == 0 and D % 256 == 0: the attn2 keys / values of all blocks come from one
    strided-batched projection (ops.CtxKVFn) -- 3 blocks, 2 x 64 caption tokens with 40 valid, ragged latents.
    `config.lora_rank` is free (config.py:22 defaults to 8, train-avatars.yaml:36 sets 32): also a rank that is not a
    multiple of 8 and one wider than a 64-wide k block, through the whole train step."""
    cfg = dict(rb.LTXV_2B, num_layers=3, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    case = dict(b=2, f=3, h=5, w=7, n_ctx=64, valid_ctx=40, lora_rank=lora_rank, seed_w=4, seed_x=11, t=[0.3, 0.8])
    from b200_ltx import ops
    before = ops.launch_count
    mc.run_parity(cfg, case)
    assert ops.launch_count > before


@pytest.mark.gpu
def test_zero_init_lora_b_matches_step0_state():
    """peft's real step-0 state: B = 0 => every lora_A grad is exactly 0, lora_B grads are not."""
    cfg = dict(rb.LTXV_2B, num_layers=1, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    P = rb.init_params(cfg, 32, seed=2, lora_b_std=0.0)
    batch = rb.synthetic_batch(cfg, 1, 2, 4, 8, 24, 5, None)
    model = mc.build_b200_model(cfg, P, 32)
    _, grads = mc.b200_loss_grads(model, batch, torch.tensor([0.5]))
    for k, v in grads.items():
        if "lora_A" in k:
            assert float(v.abs().max()) == 0.0, k
        if "lora_B" in k:
            assert float(v.abs().max()) > 0.0, k


@pytest.mark.gpu
def test_forward_mutates_input_like_reference_and_processor_protocol():
    from b200_ltx import api, modules
    cfg = dict(rb.LTXV_2B, num_layers=1, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    P = rb.init_params(cfg, 0, seed=3)
    model = mc.build_b200_model(cfg, P, 0).eval()
    b = rb.synthetic_batch(cfg, 2, 2, 4, 4, 24, 99, None)
    tokens, coords = rb.patchify(b["latents"])
    x = tokens.contiguous().cuda().to(torch.bfloat16)
    before = x.clone()
    ref, pose = b["ref_image_latents"].cuda().bfloat16(), b["pose_latents"].cuda().bfloat16()
    with torch.no_grad():
        out = model(x, coords.cuda(), ref, pose, b["prompt_embeds"].expand(2, -1, -1).cuda().bfloat16(),
                    torch.tensor([0.9, 0.9], device="cuda"), encoder_attention_mask=b["prompt_mask"].expand(2, -1).cuda(),
                    return_dict=False)[0]
    assert not torch.equal(x, before)                       # conditioning lerp is in place
    want = before.clone()
    rb.condition_tokens_(want, ref, pose)
    assert mc.maxabs(x, want) <= 2 ** -6
    assert out.shape == (2, 32, 128) and torch.isfinite(out.float()).all()
    # processor protocol: install() on an instance whose forwards were reset keeps results identical
    api.uninstall(model)
    for blk in model.transformer_blocks:
        blk.attn1.set_processor(None)
    api.install(model)
    assert isinstance(model.transformer_blocks[0].attn1.processor, modules.B200AttnProcessor)
    with torch.no_grad():
        out2 = model(before.clone(), coords.cuda(), ref, pose, b["prompt_embeds"].expand(2, -1, -1).cuda().bfloat16(),
                     torch.tensor([0.9, 0.9], device="cuda"),
                     encoder_attention_mask=b["prompt_mask"].expand(2, -1).cuda(), return_dict=False)[0]
    assert torch.equal(out, out2)


def _sampling_case(steps=6, B=2):
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    P = rb.init_params(cfg, 0, seed=5)
    P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
    model = mc.build_b200_model(cfg, P, 0).eval()
    b = rb.synthetic_batch(cfg, B, 3, 4, 6, 24, 21, 15)
    tokens, coords = rb.patchify(b["noise"].transpose(1, 2).reshape(B, -1, 3, 4, 6))
    fc = coords.float()
    fc[:, 0] = fc[:, 0] * (1.0 / 25)
    x0 = tokens.to(torch.bfloat16)
    ref, pose = b["ref_image_latents"].to(torch.bfloat16), b["pose_latents"].to(torch.bfloat16)
    enc, msk = b["prompt_embeds"].to(torch.bfloat16).expand(B, -1, -1), b["prompt_mask"].expand(B, -1)
    return cfg, P, model, x0, fc, ref, pose, enc, msk


@pytest.mark.gpu
def test_sampling_loop_matches_oracle_loop():
    """6 Euler steps of the rectified-flow sampler with one condition (inference-avatars.yaml): forward-only kernels,
    fp32 running latents, the conditioning lerp reaching the caller's latents on the first step only -- against the
    oracle's restatement of the pipeline loop over the fp32 oracle forward."""
    import ref_sampling as rs
    from b200_ltx import api
    cfg, P, model, x0, fc, ref, pose, enc, msk = _sampling_case()
    steps = 6
    sched = api.RectifiedFlowScheduler()
    lat = x0.clone().cuda().contiguous()
    x = api.denoise(model, lat, fc.cuda(), ref.cuda(), pose.cuda(), enc.cuda(), msk.cuda(), sched,
                    num_inference_steps=steps)
    assert x.dtype == torch.float32 and x.data_ptr() != lat.data_ptr()
    lat_o = x0.float().clone()
    xr = rs.denoise_loop(P, cfg, lat_o, fc, ref.float(), pose.float(), enc.float(), msk, rb.uniform_timesteps(steps))
    err = mc.rel(x.cpu(), xr)
    print(f"  sampling loop ({steps} steps): rel err vs fp32 oracle loop {err:.3e}")
    assert err < 3e-2
    # the first step conditioned the caller's tensor in place, exactly as the oracle's (reference's) aliasing does
    assert not torch.equal(lat.cpu(), x0) and mc.rel(lat.cpu(), lat_o) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_guided_sampling_loop_matches_oracle_loop(use_graph):
    """CFG (with the CFG* projection) + STG (attention-values skip of block 1) + std rescale, a negative prompt, a
    conditioning mask with hard (1.0) and soft tokens, and guidance switched off for the last two steps -- eager and
    with the captured step; against oracle/ref_sampling.py (pipeline_ltx_video.py:1089-1288)."""
    import ref_sampling as rs
    from b200_ltx import api, modules
    cfg, P, model, x0, fc, ref, pose, enc, msk = _sampling_case()
    steps = 7
    B, N, _ = x0.shape
    g = torch.Generator().manual_seed(3)
    neg = torch.randn(enc.shape, generator=g).to(torch.bfloat16)
    neg_m = torch.zeros_like(msk)
    neg_m[:, :9] = 1
    cm = torch.zeros(B, N)
    cm[:, :24] = 1.0          # first latent frame: hard conditioning
    cm[0, 24:40] = 0.5
    gs = [3.0] * 5 + [1.0] * 2
    stg = [1.0] * 5 + [0.0] * 2
    rsc = [0.7] * steps
    sampler = api.Denoiser(model, api.RectifiedFlowScheduler(), graph=use_graph)
    kw = dict(num_inference_steps=steps, negative_prompt_embeds=neg.cuda(), negative_prompt_attention_mask=neg_m.cuda(),
              guidance_scale=gs, stg_scale=stg, rescaling_scale=rsc, cfg_star_rescale=True, skip_block_list=[1],
              skip_layer_strategy=modules.SkipLayerStrategy.AttentionValues, conditioning_mask=cm.cuda())
    lat = x0.clone().cuda().contiguous()
    x = sampler(lat, fc.cuda(), ref.cuda(), pose.cuda(), enc.cuda(), msk.cuda(), **kw)
    # a second run through the same object (with graph=True: every guided step is now a replay, from step 0 on)
    x_again = sampler(x0.clone().cuda().contiguous(), fc.cuda(), ref.cuda(), pose.cuda(), enc.cuda(), msk.cuda(), **kw)
    assert mc.rel(x_again, x) < 1e-6
    assert torch.equal(lat.cpu(), x0)   # three conditions on the first step: the caller's latents are left alone
    xr = rs.denoise_loop(P, cfg, x0.float().clone(), fc, ref.float(), pose.float(), enc.float(), msk,
                         rb.uniform_timesteps(steps), neg.float(), neg_m, gs, stg, rsc, True, [1],
                         rb.STG_ATTENTION_VALUES, cm)
    err = mc.rel(x.cpu(), xr)
    print(f"  guided sampling loop ({steps} steps, graph={use_graph}): rel err vs fp32 oracle loop {err:.3e}")
    assert err < 3e-2
    assert torch.equal(x[:, :24].cpu(), x0[:, :24].float())   # hard-conditioned tokens never move


@pytest.mark.gpu
@pytest.mark.parametrize("optimizer", ["torch_fused", "b200_fused"])
def test_graphed_train_step_learns_and_matches_eager_shapes(optimizer):
    """train.GraphedTrainStep: the captured micro-step replays with fresh timestep / noise draws, updates the
    LoRA + caption parameters and drives the loss down on a fixed tiny batch (functional check of the graph path),
    with torch's fused AdamW and with the single-launch b200_ltx.optim.FusedAdamW."""
    from b200_ltx import api, optim, train
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    P = rb.init_params(cfg, 32, seed=8)
    P = {k: (v if "lora_" in k else v.to(torch.bfloat16).float()) for k, v in P.items()}
    model = mc.build_b200_model(cfg, P, 32).train()
    b = rb.synthetic_batch(cfg, 2, 3, 4, 8, 64, 31, 40)
    batch = {k: b[k].to(torch.bfloat16).cuda() for k in ("latents", "pose_latents", "ref_image_latents")}
    prompt, mask = b["prompt_embeds"].to(torch.bfloat16).cuda(), b["prompt_mask"].cuda()
    params = [p for p in model.parameters() if p.requires_grad]
    before = [p.detach().clone() for p in params]
    if optimizer == "torch_fused":
        opt = torch.optim.AdamW(params, lr=2e-3, fused=True, capturable=True)
    else:
        opt = optim.FusedAdamW(params, lr=2e-3)

    class Cfg:
        rf_log_normal_mu, rf_log_normal_sigma, rf_quantile_min, rf_quantile_max = -0.5, 1.0, 0.005, 0.999
        transformer_loss_weight = 1.0
    step = train.GraphedTrainStep(model, opt, api.RectifiedFlowScheduler(), api.SymmetricPatchifier(1), Cfg, prompt, mask,
                                  batch, warmup=2)
    assert step.launches > 50
    losses = [float(step(batch)) for _ in range(40)]
    assert all(torch.isfinite(torch.tensor(losses)))
    assert len(set(round(x, 6) for x in losses[:5])) > 1           # fresh t / noise on every replay
    assert sum(losses[-10:]) / 10 < sum(losses[:10]) / 10          # it learns
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, params))


@pytest.mark.gpu
@pytest.mark.parametrize("rank", [32, 12])
def test_lora_merge_matches_matmul(rank):
    """LoraLinear.merge on the device (one rank-r GEMM accumulating into the bf16 weight in place) against the fp32
    matmul of peft's merge_and_unload (torch_utils.py:66-102)."""
    from b200_ltx import lora
    torch.manual_seed(0)
    base = torch.nn.Linear(2048, 2048, device="cuda", dtype=torch.bfloat16)
    layer = lora.LoraLinear(base, rank, 2 * rank)
    layer.lora_B["default"].weight.data.normal_(0, 0.02)
    w0 = base.weight.detach().float().clone()
    want = w0 + 2.0 * (layer.lora_B["default"].weight.float() @ layer.lora_A["default"].weight.float())
    layer.merge()
    assert mc.rel(base.weight.float(), want) < 3e-3
    assert mc.rel(base.weight.float() - w0, want - w0) < 0.25   # the update itself (bf16 rounding of W + dW dominates)


@pytest.mark.gpu
def test_full_train_mode_gradients_vs_oracle():
    """train_mode='full' (training.py:75-91: proj_out, scale_shift_tables, adaln_single, caption_projection and every
    attention parameter train; no adapters): output, loss and ALL those gradients -- base weights and biases, qk-norm
    weights, the per-block and final scale/shift/gate tables, the timestep MLP -- against the fp32 oracle, with the
    same two-sided bf16 tolerance as the LoRA mode."""
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    case = dict(b=2, f=3, h=4, w=8, n_ctx=24, valid_ctx=15, lora_rank=0, seed_w=3, seed_x=77, t=[0.35, 0.8],
                train_mode="full")
    res = mc.run_parity(cfg, case)
    names = " ".join(res)
    for needle in ("scale_shift_table", "adaln_single.linear.weight", "attn1.q_norm.weight", "attn2.to_k.bias",
                   "proj_out.weight", "caption_projection.linear_1.weight"):
        assert needle in names, needle
    assert "ff.net" not in names and "patchify_proj" not in names


# ---------------------------------------------------------------------------------------------------------------
# round 2: parity at the depths / geometries BASELINE.json names, and the branches that had no GPU test
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cfg2_full_depth_28_blocks_6144_tokens():
    """THE headline configuration (BASELINE config 2): LTXV-2B, all 28 blocks (transformer3d.py:501-551), 6144 latent
    tokens, 256 caption tokens with 15 valid, LoRA r=32 on attn2 + trainable caption projection; velocity output, loss
    and all 228 trainable gradients against the fp32 oracle run on the same GPU, two-sided bf16 tolerance."""
    cfg = dict(rb.LTXV_2B)
    case = dict(b=1, f=16, h=16, w=24, n_ctx=256, valid_ctx=15, lora_rank=32, seed_w=9, seed_x=23, t=[0.45])
    res = mc.run_parity(cfg, case, verbose=False)
    e_o, e_r, m_o, m_r = res["velocity output"]
    grads = {k: v for k, v in res.items() if k.startswith("grad ")}
    worst = max(grads, key=lambda k: grads[k][0] / max(2 * grads[k][1], mc.GRAD_FLOOR))
    print(f"  28 blocks x 6144 tokens: velocity E_ours={e_o:.3e} E_ref_bf16={e_r:.3e} (maxabs {m_o:.2e} / {m_r:.2e}); "
          f"loss E_ours={res['loss'][0]:.3e}; {len(grads)} gradients, worst ratio to its bound: {worst} "
          f"E_ours={grads[worst][0]:.3e} E_ref={grads[worst][1]:.3e}")
    assert len(grads) == 28 * 8 + 4


@pytest.mark.gpu
def test_cfg3_geometry_batch4_full_width():
    """BASELINE config 3 per-GPU shape: batch 4 x (97x512x512 -> 13x16x16 = 3328 tokens), full width, 2 blocks."""
    cfg = dict(rb.LTXV_2B, num_layers=2)
    case = dict(b=4, f=13, h=16, w=16, n_ctx=256, valid_ctx=15, lora_rank=32, seed_w=10, seed_x=29,
                t=[0.2, 0.45, 0.7, 0.93])
    mc.run_parity(cfg, case, verbose=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name,f,h,w", [("cfg4", 16, 15, 22), ("cfg5", 33, 16, 24)])
def test_sampling_and_long_clip_geometries_forward(name, f, h, w):
    """BASELINE config 4 (121x480x704 -> 16x15x22 = 5280 tokens: ragged against the 128-row tiles, fractional
    coordinates, per-token timesteps with a conditioned first frame) and config 5 (257x512x768 -> 12672 tokens),
    forward only, full width, 2 blocks, against the fp32 oracle."""
    cfg = dict(rb.LTXV_2B, num_layers=2)
    case = dict(b=1, f=f, h=h, w=w, n_ctx=256, valid_ctx=15, seed_w=11, seed_x=31, t=[0.6], per_token_t=(name == "cfg4"))
    mc.run_forward_parity(cfg, case)


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", [rb.STG_ATTENTION_SKIP, rb.STG_ATTENTION_VALUES, rb.STG_TRANSFORMER_BLOCK])
def test_stg_skip_layer_strategies_vs_oracle(strategy):
    """attention.py:1071-1086 / 312-319: two conditions (text, perturbed), the perturbed one skipping blocks 0 and 2
    (block 1 keeps the fused attn1 node: create_skip_layer_mask notes host-side which rows differ from all-ones)."""
    cfg = dict(rb.LTXV_2B, num_layers=3, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    case = dict(b=2, f=3, h=5, w=7, n_ctx=24, valid_ctx=15, seed_w=12, seed_x=37, t=[0.8, 0.8])
    res = mc.run_forward_parity(cfg, case, skip_layer_strategy=strategy, skip_blocks=[0, 2])
    plain = mc.run_forward_parity(cfg, case, verbose=False)
    ours, _ = res["_tensors"]
    ours_plain, _ = plain["_tensors"]
    assert mc.rel(ours[0], ours_plain[0]) < 1e-6          # the text condition is untouched
    assert mc.rel(ours[1], ours_plain[1]) > 1e-3          # the perturbed one is not


@pytest.mark.gpu
def test_gradient_checkpoint_branch_matches_plain_backward():
    """transformer3d.py:503-534: `model.training and gradient_checkpointing` runs every block under
    torch.utils.checkpoint(use_reentrant=False).  The recomputed forward goes through the same autograd Functions
    (GradJoin side channel, NormModResFn's aliased residual output): loss and all gradients must equal the
    non-checkpointed step bit for bit, and stay inside the oracle tolerance."""
    cfg = dict(rb.LTXV_2B, num_layers=3, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    case = dict(b=2, f=3, h=4, w=8, n_ctx=64, valid_ctx=40, lora_rank=32, seed_w=13, seed_x=41, t=[0.3, 0.6])
    mc.run_parity(cfg, dict(case, gradient_checkpointing=True), verbose=False)
    P = rb.init_params(cfg, 32, seed=13)
    P = {k: (v if "lora_" in k else v.to(torch.bfloat16).float()) for k, v in P.items()}
    batch = rb.synthetic_batch(cfg, 2, 3, 4, 8, 64, 41, 40)
    t = torch.tensor(case["t"])
    model = mc.build_b200_model(cfg, P, 32).train()
    l0, g0 = mc.b200_loss_grads(model, batch, t)
    g0 = {k: v.clone() for k, v in g0.items()}
    model.base_model.model.gradient_checkpointing = True
    l1, g1 = mc.b200_loss_grads(model, batch, t)
    assert torch.equal(l0, l1)
    assert set(g0) == set(g1)
    for k in g0:
        # not bit-equal by design: without checkpointing the caption keys / values of all blocks come from one batched
        # projection whose dgrad reduces over the blocks in fp32 (ops.CtxKVFn), under checkpointing every block
        # projects for itself and autograd sums bf16 partials; split-K weight gradients are fp32 atomics
        assert mc.rel(g1[k], g0[k]) < 1.5e-2, (k, mc.rel(g1[k], g0[k]))


@pytest.mark.gpu
def test_installed_model_survives_deepcopy_merge_state_dict_and_remerge_forward():
    """torch_utils.py:66-102: deepcopy(model) -> merge_and_unload() -> state_dict() on a model that has already run
    (weight-concatenation / RoPE caches filled); and ADVICE r1: forward -> merge -> forward on the same object must
    see the merged attn2 K/V weights (batched caption-K/V cache re-keyed)."""
    import copy
    from b200_ltx import modules
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=4, cross_attention_dim=256, caption_channels=64)
    P = rb.init_params(cfg, 32, seed=14)
    P = {k: (v if "lora_" in k else v.to(torch.bfloat16).float()) for k, v in P.items()}
    model = mc.build_b200_model(cfg, P, 32).eval()
    b = rb.synthetic_batch(cfg, 2, 3, 4, 8, 64, 43, 40)
    tokens, coords = rb.patchify(b["latents"])
    args = lambda: (tokens.contiguous().cuda().bfloat16(), coords.cuda(), b["ref_image_latents"].cuda().bfloat16(),  # noqa: E731
                    b["pose_latents"].cuda().bfloat16(), b["prompt_embeds"].expand(2, -1, -1).cuda().bfloat16(),
                    torch.tensor([0.5, 0.5], device="cuda"))
    kw = dict(encoder_attention_mask=b["prompt_mask"].expand(2, -1).cuda(), return_dict=False)
    with torch.no_grad():
        with_adapters = model(*args(), **kw)[0]
    root = model.base_model.model
    assert modules.side(root).get("wkv_all") is not None and modules.side(root).get("rope") is not None
    mem0 = torch.cuda.memory_allocated()
    clone = copy.deepcopy(model)
    n_param_bytes = sum(p.numel() * p.element_size() for p in model.parameters())
    assert torch.cuda.memory_allocated() - mem0 < 1.25 * n_param_bytes + (8 << 20)   # parameters only, no caches
    merged = clone.merge_and_unload()
    sd = merged.state_dict()
    assert set(sd) == set(rb.param_shapes(cfg, 0))
    with torch.no_grad():
        out_merged = merged(*args(), **kw)[0]
        again = model(*args(), **kw)[0]
    assert torch.equal(again, with_adapters)
    assert mc.rel(out_merged, with_adapters) < 1.5e-2
    # same object: forward (caches filled above) -> merge in place -> forward
    base = model.merge_and_unload()
    with torch.no_grad():
        out2 = base(*args(), **kw)[0]
    assert mc.rel(out2, with_adapters) < 1.5e-2, mc.rel(out2, with_adapters)
    assert torch.equal(out2, out_merged)
