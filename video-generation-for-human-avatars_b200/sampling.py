"""Rectified-flow sampling loop around the B200 transformer forward (BASELINE config 4).

Mirrors the denoising loop of the reference pipeline for the avatar setting (num_conds = 1, no CFG / STG:
pipelines/pipeline_ltx_video.py:1166-1288 with inference-avatars.yaml): per step
    v   = transformer(latents, indices_grid, ref, pose, prompt, timestep = t_i (per sample), mask)
    x  <- scheduler.step(v, t_i, x)            # Euler: x - (t_i - t_{i+1}) v        (rf.py:305-374)
The transformer forward applies the ref/pose conditioning lerp to its input IN PLACE on every call, as the
reference does (transformer3d.py:447-466) -- so the loop must hand it the live latents, not a copy.
Everything runs under no_grad on the forward-only kernels; the Euler update is one element-wise op per step."""
from typing import Callable, Optional

import torch

from .lib import B200Error


@torch.no_grad()
def denoise(model, latents: torch.Tensor, indices_grid: torch.Tensor, ref_image_latents: torch.Tensor,
            pose_latents: torch.Tensor, prompt_embeds: torch.Tensor, prompt_attention_mask: torch.Tensor, scheduler,
            num_inference_steps: int = 40, callback: Optional[Callable[[int, torch.Tensor], None]] = None):
    """latents: [B, N, C] bf16 noise tokens (modified in place and returned); indices_grid: [B, 3, N] (fractional
    pixel coordinates as the pipeline builds them); scheduler: b200_ltx RectifiedFlowScheduler."""
    if latents.dtype != torch.bfloat16 or not latents.is_cuda:
        raise B200Error("denoise: latents must be CUDA bfloat16 tokens [B, N, C]")
    B = latents.shape[0]
    scheduler.set_timesteps(num_inference_steps, samples_shape=latents.shape, device=latents.device)
    enc = prompt_embeds.expand(B, -1, -1) if prompt_embeds.shape[0] != B else prompt_embeds
    msk = prompt_attention_mask.expand(B, -1) if prompt_attention_mask.shape[0] != B else prompt_attention_mask
    was_training = model.training
    model.eval()
    try:
        for i, t in enumerate(scheduler.timesteps):
            tb = t.expand(B).to(torch.float32)
            v = model(hidden_states=latents, indices_grid=indices_grid, ref_image_hidden_states=ref_image_latents,
                      pose_hidden_states=pose_latents, encoder_hidden_states=enc, timestep=tb,
                      encoder_attention_mask=msk, return_dict=False)[0]
            latents.copy_(scheduler.step(v, t, latents, return_dict=False)[0])
            if callback is not None:
                callback(i, latents)
    finally:
        model.train(was_training)
    return latents
