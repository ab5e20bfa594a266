"""RectifiedFlowScheduler mirror (ltx_video/schedulers/rf.py:176-426) for the pieces on the hot path:
add_noise / build_velocity_target run the fused b200 kernel on bf16 CUDA tokens; the timestep
schedules, resolution shifts and the Euler step are tiny host / element-wise logic kept in torch."""
import math
from typing import Optional

import torch

from . import ops
from .lib import B200Error


def linear_quadratic_schedule(num_steps, threshold_noise=0.025, linear_steps=None):
    if num_steps == 1:
        return torch.tensor([1.0])
    linear_steps = num_steps // 2 if linear_steps is None else linear_steps
    head = [i * threshold_noise / linear_steps for i in range(linear_steps)]
    diff = linear_steps - threshold_noise * num_steps
    quad = num_steps - linear_steps
    qc = diff / (linear_steps * quad ** 2)
    lc = threshold_noise / linear_steps - 2 * diff / (quad ** 2)
    const = qc * (linear_steps ** 2)
    tail = [qc * (i ** 2) + lc * i + const for i in range(linear_steps, num_steps)]
    return torch.tensor([1.0 - s for s in head + tail + [1.0]][:-1])


def time_shift(mu, sigma, t):
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def _num_tokens(shape):
    if len(shape) == 3:
        return shape[1]
    if len(shape) in (4, 5):
        return math.prod(shape[2:])
    raise ValueError("Samples must have shape (b, t, c), (b, c, h, w) or (b, c, f, h, w)")


def sd3_resolution_dependent_timestep_shift(samples_shape, timesteps, target_shift_terminal=None):
    m = _num_tokens(samples_shape)
    slope = (2.05 - 0.95) / (4096 - 1024)
    shift = slope * m + (0.95 - slope * 1024)
    ts = time_shift(shift, 1, timesteps)
    if target_shift_terminal is not None:
        one_minus = 1 - ts
        ts = 1 - one_minus / (one_minus[-1] / (1 - target_shift_terminal))
    return ts


def simple_diffusion_resolution_dependent_timestep_shift(samples_shape, timesteps, n=32 * 32):
    m = _num_tokens(samples_shape)
    snr = (timesteps / (1 - timesteps)) ** 2
    return torch.sigmoid(0.5 * (torch.log(snr) + 2 * math.log(m / n)))


class RectifiedFlowScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shifting: Optional[str] = None, base_resolution: int = 32 ** 2,
                 target_shift_terminal: Optional[float] = None, sampler: Optional[str] = "Uniform",
                 shift: Optional[float] = None):
        self.config = type("Cfg", (), dict(num_train_timesteps=num_train_timesteps, shifting=shifting,
                                            base_resolution=base_resolution, sampler=sampler, shift=shift,
                                            target_shift_terminal=target_shift_terminal))()
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.sampler, self.shifting, self.base_resolution = sampler, shifting, base_resolution
        self.target_shift_terminal, self.shift = target_shift_terminal, shift
        self.timesteps = self.sigmas = self.get_initial_timesteps(num_train_timesteps, shift=shift)

    def get_initial_timesteps(self, num_timesteps, shift=None):
        if self.sampler == "Uniform":
            return torch.linspace(1, 1 / num_timesteps, num_timesteps)
        if self.sampler == "LinearQuadratic":
            return linear_quadratic_schedule(num_timesteps)
        if self.sampler == "Constant":
            assert shift is not None, "Shift must be provided for constant time shift sampler."
            return time_shift(shift, 1, torch.linspace(1, 1 / num_timesteps, num_timesteps))
        raise ValueError(f"unknown sampler {self.sampler}")

    def shift_timesteps(self, samples_shape, timesteps):
        if self.shifting == "SD3":
            return sd3_resolution_dependent_timestep_shift(samples_shape, timesteps, self.target_shift_terminal)
        if self.shifting == "SimpleDiffusion":
            return simple_diffusion_resolution_dependent_timestep_shift(samples_shape, timesteps, self.base_resolution)
        return timesteps

    def set_timesteps(self, num_inference_steps=None, samples_shape=None, timesteps=None, device=None):
        if timesteps is not None and num_inference_steps is not None:
            raise ValueError("You cannot provide both `timesteps` and `num_inference_steps`.")
        if timesteps is None:
            num_inference_steps = min(self.config.num_train_timesteps, num_inference_steps)
            timesteps = self.get_initial_timesteps(num_inference_steps, shift=self.shift).to(device)
            timesteps = self.shift_timesteps(samples_shape, timesteps)
        else:
            timesteps = torch.Tensor(timesteps).to(device)
            num_inference_steps = len(timesteps)
        self.timesteps = self.sigmas = timesteps
        self.num_inference_steps = num_inference_steps

    def scale_model_input(self, sample, timestep=None):
        return sample

    # ---- hot-path pieces: fused kernel (rf.py:376-386, 400-426) ----
    def add_noise(self, original_samples, noise, timesteps):
        """x_t = (1 - t) x0 + t eps, returned in the tokens' dtype (bf16) as training.py:140-143 uses it."""
        if timesteps.ndim != 1:
            raise B200Error("add_noise: one timestep per sample is built (training.py:124-141)")
        xt, _ = ops.rf_noise(original_samples.contiguous(), noise.contiguous(), timesteps, True, False)
        return xt

    def build_velocity_target(self, tokens, noise, t):
        _, v = ops.rf_noise(tokens.contiguous(), noise.contiguous(), t, False, True)
        return v

    def noise_and_target(self, tokens, noise, t):
        """Both outputs from ONE pass over (x0, eps)."""
        return ops.rf_noise(tokens.contiguous(), noise.contiguous(), t, True, True)

    # ---- Euler step (rf.py:305-374): element-wise, sampling only ----
    def step(self, model_output, timestep, sample, return_dict=True, stochastic_sampling=False, **kwargs):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        eps = 1e-6
        grid = torch.cat([self.timesteps, torch.zeros(1, device=self.timesteps.device)])
        if timestep.ndim == 0:
            lower = grid[grid < timestep - eps][0]
            dt = timestep - lower
        else:
            assert timestep.ndim == 2
            below = grid[:, None, None] < timestep[None] - eps
            lower, _ = (below * grid[:, None, None]).max(dim=0)
            dt = (timestep - lower)[..., None]
        if stochastic_sampling:
            # rf.py:362-365: re-noise the x0 estimate to the next level (off in both shipped configs,
            # inference-avatars.yaml:14; element-wise torch ops, drawn from the global generator as the reference does)
            x0 = sample - timestep[..., None] * model_output
            nxt = timestep[..., None] - dt
            nxt = nxt.reshape(nxt.shape + (1,) * (sample.ndim - nxt.ndim))
            prev = (1 - nxt) * x0 + nxt * torch.randn_like(sample)
        else:
            prev = sample - dt * model_output
        return (prev,) if not return_dict else type("RectifiedFlowSchedulerOutput", (), {"prev_sample": prev})()
