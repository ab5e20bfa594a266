"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's denoising loop with guidance.

Follows pipelines/pipeline_ltx_video.py:1089-1288 (the loop), :1346-1379 (`denoising_step`) and
transformer3d.py:187-203 (`create_skip_layer_mask`) on top of ref_block.transformer_forward / rf_step.  Only tests,
`__graft_entry__.smoke()` and bench.py's cpu_baseline leg may import this; the product never does.

Parity status: PINNED.  The reference's unmodified `LTXVideoPipeline.__call__` is driven on CPU by
oracle/ref_pipeline.py (VAE / diffusers-pipeline stand-ins, the loop itself as written) and this restatement equals
its output bit for bit -- single condition, CFG, CFG* + STG + std rescale with per-step guidance lists, all four
skip-layer strategies -- and `denoising_step` with a conditioning mask (tests/test_oracle.py, live where
/root/reference is mounted; tests/golden/tiny_guided_sampling_fp32.pt, written by oracle/make_golden.py, elsewhere).
One deliberate difference: the reference's CFG* projection multiplies a [B, 1] alpha into a [B, N, C] prediction
(:1238), which only broadcasts for batch size 1 (it raises for B > 1, test_oracle pins that); this file reshapes alpha
to [B, 1, 1], identical at B = 1 and defined for B > 1.

Reference behaviours kept on purpose:
* the prompt batch order is [negative, positive, positive] and a step uses the slice its guidance flags select
  (:1026-1036, :1100-1108);
* `num_conds == 1` hands the transformer the running latents themselves, and its conditioning lerp is in place
  (transformer3d.py:447-466): with a bf16 model this aliases only while the latents still have the model's dtype,
  i.e. on the first step -- `scheduler.step` returns fp32 (its dt is an fp32 tensor) and every later step feeds a
  `.to(bf16)` copy.  `alias_first_step_only=True` reproduces that; False is what an all-fp32 run of the reference
  does (aliasing on every step);
* the skip-layer mask zeroes columns `ptb_index::num_conds` (:199-201), which addresses the perturbed copies only for
  batch size 1;
* `rescaling_scale == 1` means "no rescale" (:1093), not "full rescale";
* the Euler step uses the timestep row of the first batch entry (`current_timestep[:1]`, :1262) for every sample,
  while the conditioning select uses each sample's own mask."""
from typing import List, Optional, Sequence, Union

import torch

import ref_block as rb

Tensor = torch.Tensor


def skip_layer_mask(num_layers: int, batch_size: int, num_conds: int, ptb_index: int,
                    skip_block_list: Optional[Sequence[int]], dtype=torch.float32) -> Optional[Tensor]:
    if skip_block_list is None or len(skip_block_list) == 0:
        return None
    mask = torch.ones((num_layers, batch_size * num_conds), dtype=dtype)
    for block_idx in skip_block_list:
        mask[block_idx, ptb_index::num_conds] = 0
    return mask


def _per_step(value, n: int) -> list:
    return list(value) if isinstance(value, (list, tuple)) else [value] * n


def guidance_combine(noise_pred: Tensor, batch_size: int, num_conds: int, do_cfg: bool, do_stg: bool, guidance_scale: float,
                     stg_scale: float, rescaling_scale: float, cfg_star_rescale: bool) -> Tensor:
    """pipeline :1217-1260 on the stacked model output [num_conds * B, N, C]."""
    do_rescaling = rescaling_scale != 1.0
    if do_stg:
        text, perturb = noise_pred.chunk(num_conds)[-2:]
    if do_cfg:
        uncond, text = noise_pred.chunk(num_conds)[:2]
        if cfg_star_rescale:
            pos = text.reshape(batch_size, -1)
            neg = uncond.reshape(batch_size, -1)
            dot = torch.sum(pos * neg, dim=1, keepdim=True)
            sq = torch.sum(neg ** 2, dim=1, keepdim=True) + 1e-8
            alpha = dot / sq
            uncond = alpha.view(batch_size, 1, 1) * uncond
        noise_pred = uncond + guidance_scale * (text - uncond)
    elif do_stg:
        noise_pred = text
    if do_stg:
        noise_pred = noise_pred + stg_scale * (text - perturb)
        if do_rescaling and stg_scale > 0.0:
            text_std = text.reshape(batch_size, -1).std(dim=1, keepdim=True)
            pred_std = noise_pred.reshape(batch_size, -1).std(dim=1, keepdim=True)
            factor = text_std / pred_std
            factor = rescaling_scale * factor + (1 - rescaling_scale)
            noise_pred = noise_pred * factor.view(batch_size, 1, 1)
    return noise_pred


def denoising_step(timesteps: Tensor, latents: Tensor, noise_pred: Tensor, cur_t: Tensor,
                   conditioning_mask: Optional[Tensor], t: Tensor, t_eps: float = 1e-6) -> Tensor:
    """pipeline :1346-1379: Euler step of every token, then keep the tokens whose conditioning level says they are
    not being denoised yet (hard-conditioned tokens, mask 1.0, never move)."""
    denoised = rb.rf_step(timesteps, noise_pred, cur_t, latents)
    if conditioning_mask is None:
        return denoised
    move = (t - t_eps < (1.0 - conditioning_mask)).unsqueeze(-1)
    return torch.where(move, denoised, latents)


def denoise_loop(P, cfg: dict, latents: Tensor, fractional_coords: Tensor, ref: Tensor, pose: Tensor,
                 prompt_embeds: Tensor, prompt_mask: Tensor, timesteps: Tensor,
                 negative_prompt_embeds: Optional[Tensor] = None, negative_prompt_mask: Optional[Tensor] = None,
                 guidance_scale: Union[float, List[float]] = 1.0, stg_scale: Union[float, List[float]] = 0.0,
                 rescaling_scale: Union[float, List[float]] = 1.0, cfg_star_rescale: bool = False,
                 skip_block_list=None, skip_layer_strategy: Optional[str] = None,
                 conditioning_mask: Optional[Tensor] = None, alias_first_step_only: bool = True) -> Tensor:
    """latents [B, N, C] fp32 (modified in place where the reference would); prompt tensors [B, L, *];
    timesteps = scheduler.timesteps (descending, fp32).  Returns the final latents."""
    batch_size = latents.shape[0]
    n_steps = len(timesteps)
    gs_l, stg_l, rs_l = (_per_step(v, n_steps) for v in (guidance_scale, stg_scale, rescaling_scale))
    if skip_block_list is not None and (len(skip_block_list) == 0 or not isinstance(skip_block_list[0], (list, tuple))):
        skip_block_list = [skip_block_list] * n_steps
    if negative_prompt_embeds is None:
        negative_prompt_embeds = torch.zeros_like(prompt_embeds)
    if negative_prompt_mask is None:
        negative_prompt_mask = torch.zeros_like(prompt_mask)
    embeds = torch.cat([negative_prompt_embeds, prompt_embeds, prompt_embeds], dim=0)
    masks = torch.cat([negative_prompt_mask, prompt_mask, prompt_mask], dim=0)
    with torch.no_grad():
        for i, t in enumerate(timesteps):
            do_cfg = gs_l[i] > 1.0
            do_stg = stg_l[i] > 0
            num_conds = 1 + int(do_cfg) + int(do_stg)
            if do_cfg and do_stg:
                sel = slice(0, batch_size * 3)
            elif do_cfg:
                sel = slice(0, batch_size * 2)
            elif do_stg:
                sel = slice(batch_size, batch_size * 3)
            else:
                sel = slice(batch_size, batch_size * 2)
            slm = None
            if do_stg and skip_block_list is not None:
                slm = skip_layer_mask(cfg["num_layers"], batch_size, num_conds, num_conds - 1, skip_block_list[i])
            cm = conditioning_mask
            if cm is not None:
                cm = torch.cat([cm] * num_conds)
            aliased = num_conds == 1 and (i == 0 or not alias_first_step_only)
            model_in = latents if aliased else torch.cat([latents] * num_conds)
            cur_t = t[None].expand(model_in.shape[0]).unsqueeze(-1)
            if cm is not None:
                cur_t = torch.min(cur_t, 1.0 - cm)
            noise_pred = rb.transformer_forward(
                P, cfg, model_in, torch.cat([fractional_coords] * num_conds), torch.cat([ref] * num_conds),
                torch.cat([pose] * num_conds), embeds[sel], cur_t, masks[sel], skip_layer_mask=slm,
                skip_layer_strategy=skip_layer_strategy)
            noise_pred = guidance_combine(noise_pred, batch_size, num_conds, do_cfg, do_stg, gs_l[i], stg_l[i], rs_l[i],
                                          cfg_star_rescale)
            latents = denoising_step(timesteps, latents, noise_pred, cur_t[:1], conditioning_mask, t)
    return latents
