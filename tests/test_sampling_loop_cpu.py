"""Host logic of the sampling loop object (sampling.Denoiser) on CPU, over plain-torch stand-ins of the kernels, against
the oracle's restatement of the reference's pipeline loop (oracle/ref_sampling.py -- itself bit-equal to the reference's
own LTXVideoPipeline.__call__, tests/test_oracle.py).

What runs here is everything the Denoiser does around the kernels: the per-step guidance configuration (1-3 conditions),
the prompt batch order [negative, positive, positive] and its slices, the replicated conditioning inputs, skip-layer masks
per step, per-step guidance lists, the timestep rows / Euler steps from the device tables, the first-step aliasing of the
caller's latents, the hand-over of the next step's model input between configurations, and reuse of one Denoiser."""
import pytest
import torch

import model_checks as mc
import ref_block as rb
import ref_sampling as rs
import torch_kernels as tk

BF16 = torch.bfloat16


def _case(B=2):
    cfg = dict(rb.LTXV_2B, num_layers=2, num_attention_heads=2, cross_attention_dim=128, caption_channels=64)
    P = rb.init_params(cfg, 0, seed=5)
    P = {k: v.to(BF16).float() for k, v in P.items()}
    b = rb.synthetic_batch(cfg, B, 3, 4, 6, 24, 21, 15)
    tokens, coords = rb.patchify(b["noise"].transpose(1, 2).reshape(B, -1, 3, 4, 6))
    fc = coords.float()
    fc[:, 0] = fc[:, 0] * (1.0 / 25)
    x0 = tokens.to(BF16).contiguous()
    ref, pose = b["ref_image_latents"].to(BF16), b["pose_latents"].to(BF16)
    enc, msk = b["prompt_embeds"].to(BF16).expand(B, -1, -1).contiguous(), b["prompt_mask"].expand(B, -1).contiguous()
    return cfg, P, x0, fc, ref, pose, enc, msk


def test_single_condition_loop_and_first_step_aliasing():
    from b200_ltx import api
    cfg, P, x0, fc, ref, pose, enc, msk = _case()
    steps = 5
    with tk.patched():
        model = mc.build_b200_model(cfg, P, 0, device="cpu").eval()
        lat = x0.clone().as_subclass(tk.AsIfOnDevice)
        seen = []
        x = api.denoise(model, lat, fc, ref, pose, enc, msk, api.RectifiedFlowScheduler(), num_inference_steps=steps,
                        callback=lambda i, cur: seen.append((i, cur.clone())))
    lat_o = x0.float().clone()
    xr = rs.denoise_loop(P, cfg, lat_o, fc, ref.float(), pose.float(), enc.float(), msk, rb.uniform_timesteps(steps))
    assert x.dtype == torch.float32 and x.data_ptr() != lat.data_ptr()
    assert mc.rel(x, xr) < 3e-2
    # one condition on the first step: the caller's latents come back conditioned, as the reference's aliasing does
    assert not torch.equal(torch.Tensor(lat), x0) and mc.rel(torch.Tensor(lat), lat_o) < 1e-2
    assert [i for i, _ in seen] == list(range(steps)) and torch.equal(seen[-1][1], x)


@pytest.mark.parametrize("strategy", [rb.STG_ATTENTION_VALUES, rb.STG_ATTENTION_SKIP, rb.STG_TRANSFORMER_BLOCK])
def test_guided_loop_with_mask_lists_and_reuse(strategy):
    """CFG* + STG + std rescale for five steps, then no guidance for two; negative prompt; hard and soft conditioning."""
    from b200_ltx import api, modules
    cfg, P, x0, fc, ref, pose, enc, msk = _case()
    steps = 7
    B, N, _ = x0.shape
    g = torch.Generator().manual_seed(3)
    neg = torch.randn(enc.shape, generator=g).to(BF16)
    neg_m = torch.zeros_like(msk)
    neg_m[:, :9] = 1
    cm = torch.zeros(B, N)
    cm[:, :24] = 1.0
    cm[0, 24:40] = 0.5
    gs, stg, rsc = [3.0] * 5 + [1.0] * 2, [1.0] * 5 + [0.0] * 2, [0.7] * steps
    skips = [[1]] * 3 + [[0, 1]] * 4
    strat_p = {rb.STG_ATTENTION_SKIP: modules.SkipLayerStrategy.AttentionSkip,
               rb.STG_ATTENTION_VALUES: modules.SkipLayerStrategy.AttentionValues,
               rb.STG_TRANSFORMER_BLOCK: modules.SkipLayerStrategy.TransformerBlock}[strategy]
    kw = dict(num_inference_steps=steps, negative_prompt_embeds=neg, negative_prompt_attention_mask=neg_m,
              guidance_scale=gs, stg_scale=stg, rescaling_scale=rsc, cfg_star_rescale=True, skip_block_list=skips,
              skip_layer_strategy=strat_p, conditioning_mask=cm)
    with tk.patched():
        model = mc.build_b200_model(cfg, P, 0, device="cpu").eval()
        sampler = api.Denoiser(model, api.RectifiedFlowScheduler(), graph=False)
        lat = x0.clone().as_subclass(tk.AsIfOnDevice)
        x = sampler(lat, fc, ref, pose, enc, msk, **kw)
        x_again = sampler(x0.clone().as_subclass(tk.AsIfOnDevice), fc, ref, pose, enc, msk, **kw)
        assert len(sampler._cfgs) == 3          # (cfg + stg, skip [1]), (cfg + stg, skip [0, 1]), (one condition)
    assert torch.equal(x_again, x)
    assert torch.equal(torch.Tensor(lat), x0)   # three conditions on the first step: the caller's latents are left alone
    xr = rs.denoise_loop(P, cfg, x0.float().clone(), fc, ref.float(), pose.float(), enc.float(), msk,
                         rb.uniform_timesteps(steps), neg.float(), neg_m, gs, stg, rsc, True, skips, strategy, cm)
    assert mc.rel(x, xr) < 3e-2, mc.rel(x, xr)
    assert torch.equal(x[:, :24], x0[:, :24].float())     # hard-conditioned tokens never move


def test_denoiser_argument_checks():
    from b200_ltx import api
    from b200_ltx.lib import B200Error
    cfg, P, x0, fc, ref, pose, enc, msk = _case()
    with tk.patched():
        model = mc.build_b200_model(cfg, P, 0, device="cpu").eval()
        sampler = api.Denoiser(model, api.RectifiedFlowScheduler())
        lat = x0.clone().as_subclass(tk.AsIfOnDevice)
        with pytest.raises(B200Error, match="CUDA bfloat16"):
            sampler(x0.clone(), fc, ref, pose, enc, msk, num_inference_steps=2)          # a plain CPU tensor
        with pytest.raises(B200Error, match="guidance_scale has 3 entries for 2 steps"):
            sampler(lat, fc, ref, pose, enc, msk, num_inference_steps=2, guidance_scale=[1.0, 2.0, 3.0])
        with pytest.raises(B200Error, match="conditioning_mask"):
            sampler(lat, fc, ref, pose, enc, msk, num_inference_steps=2, conditioning_mask=torch.zeros(1, 3))
        with pytest.raises(B200Error, match="stochastic_sampling"):
            sampler(lat, fc, ref, pose, enc, msk, num_inference_steps=2, stochastic_sampling=True)
