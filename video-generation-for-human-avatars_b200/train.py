"""train_step mirror (ltx_video/training.py:94-166): same arguments and return values; the noising,
velocity target, transformer forward, loss and loss gradient all run on b200 kernels."""
from typing import Optional

import torch

from . import ops


class _RFLossFn(torch.autograd.Function):
    """mean((out - target)^2): the kernel produces the loss and dLoss/dOut in one pass."""

    @staticmethod
    def forward(ctx, out, target):
        loss, dout = ops.rf_loss(out.contiguous(), target.contiguous(), 1.0, True)
        ctx.save_for_backward(dout)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g.to(dout.dtype), None


def rf_mse_loss(out, target):
    return _RFLossFn.apply(out, target)


def sample_timesteps(config, scheduler, samples_shape, B, device, generator: Optional[torch.Generator] = None):
    """LogNormal -> t/(1+t) -> quantile clamp -> resolution shift (training.py:124-136)."""
    mu = torch.tensor(config.rf_log_normal_mu, device=device)
    sigma = torch.tensor(config.rf_log_normal_sigma, device=device)
    raw = torch.exp(mu + sigma * torch.randn(B, device=device, generator=generator))
    t_raw = raw / (1 + raw)
    t_low = torch.quantile(t_raw, config.rf_quantile_min)
    t_high = torch.quantile(t_raw, config.rf_quantile_max)
    t = torch.maximum(torch.minimum(t_raw, t_high), t_low)  # no host sync, unlike float(t_low)
    return scheduler.shift_timesteps(samples_shape, t)


def train_step(model, batch: dict, scheduler, patchifier, config, prompt_embeds, prompt_attention_mask,
               device=None, t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
    """Returns (loss, rel_mse, nrmse, loss_dict) like the reference.  `t` / `noise` may be injected for
    deterministic parity runs; otherwise they are drawn as in training.py:124-138."""
    model_dtype = next(model.parameters()).dtype
    latents = batch["latents"].to(device=device, dtype=model_dtype)
    ref = batch["ref_image_latents"].to(device=device, dtype=model_dtype)
    pose = batch["pose_latents"].to(device=device, dtype=model_dtype)
    B = latents.shape[0]
    enc = prompt_embeds.expand(B, -1, -1).to(device=device, dtype=model_dtype)
    enc_mask = prompt_attention_mask.expand(B, -1).to(device)
    tokens, coords = patchifier.patchify(latents)
    tokens = tokens.contiguous()
    if t is None:
        t = sample_timesteps(config, scheduler, tokens.shape, B, tokens.device)
    if noise is None:
        noise = torch.randn_like(tokens)
    root = getattr(getattr(model, "base_model", None), "model", None) or model
    sp = root.__dict__.get("_b200_sp")
    if sp is not None:
        # sequence-sharded step: every rank of the group sees the same clip and timestep and keeps its
        # contiguous token shard; the loss is the per-shard mean (average gradients over the group)
        torch.distributed.broadcast(t, torch.distributed.get_global_rank(sp.group, 0) if sp.group is not None else 0,
                                    group=sp.group)
        off, n = sp.shard(tokens.shape[1])
        tokens = tokens[:, off:off + n].contiguous()
        coords = coords[:, :, off:off + n].contiguous()
        noise = noise[:, off:off + n].contiguous()
    noisy, v_target = scheduler.noise_and_target(tokens, noise.to(model_dtype), t)
    out = model(hidden_states=noisy, indices_grid=coords, ref_image_hidden_states=ref, pose_hidden_states=pose,
                encoder_hidden_states=enc, timestep=t, attention_mask=None, encoder_attention_mask=enc_mask,
                return_dict=True)
    mse = rf_mse_loss(out.sample, v_target)
    loss = float(getattr(config, "transformer_loss_weight", 1.0)) * mse
    std_target = v_target.float().std()
    rel_mse = loss / (std_target ** 2 + 1e-12)
    nrmse = torch.sqrt(loss) / (std_target + 1e-12)
    return loss, rel_mse, nrmse, {"transformer_mse": mse.detach()}
