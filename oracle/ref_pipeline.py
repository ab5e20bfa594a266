"""Tier-one oracle for the sampling loop: the reference's OWN, unmodified `LTXVideoPipeline.__call__`
(pipelines/pipeline_ltx_video.py:722-1344) and `denoising_step` (:1346-1379), driven on CPU.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  The pipeline module imports the VAE stack, the
diffusers pipeline base classes and imageio-era helpers, none of which are installable here and all of which are OUT
OF SCOPE for the hot path.  They are replaced in `sys.modules` by the stand-ins below before the
module is imported; the denoising loop itself -- prompt batching [negative, positive, positive], the per-step condition
slice, skip-layer masks, the CFG / CFG* / STG combine, std rescale, `current_timestep[:1]`, `denoising_step` -- runs
as written by the reference, over the reference's own Transformer3DModel and RectifiedFlowScheduler.

What the stand-ins restate (and therefore what stays unpinned): the "VAE" is the identity on latents (the caller hands
latents where the pipeline expects pixels) with the LTXV scale factors (temporal 8, spatial 32); `latent_to_pixel_coords`
follows vae_encode.py:190-225; `randn_tensor` is `torch.randn(shape, generator=...)` as in diffusers 0.35.1."""
import contextlib
import sys
import types

import torch

import ref_import


class _FakeVAE:
    """Stands where the pipeline expects a CausalVideoAutoencoder: scale factors + dtype/device only."""
    spatial_downscale_factor = 32
    temporal_downscale_factor = 8
    dtype = torch.float32
    device = torch.device("cpu")


def _install_stand_ins():
    if "ltx_video.pipelines.pipeline_ltx_video" in sys.modules:
        return
    ref_import._prepare()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class DiffusionPipeline:
        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        @property
        def _execution_device(self):
            return torch.device("cpu")

        @contextlib.contextmanager
        def progress_bar(self, total=None):
            class _Bar:
                def update(self, *a):
                    pass
            yield _Bar()

        def maybe_free_model_hooks(self):
            pass

    class ImagePipelineOutput:
        def __init__(self, images):
            self.images = images

    class _Empty:
        def __init__(self, *a, **k):
            pass

    import diffusers
    mod("diffusers.image_processor", VaeImageProcessor=_Empty)
    diffusers.models.AutoencoderKL = _Empty
    mod("diffusers.pipelines")
    mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=DiffusionPipeline, ImagePipelineOutput=ImagePipelineOutput)
    import diffusers.schedulers as ds
    ds.DPMSolverMultistepScheduler = _Empty
    import diffusers.utils.torch_utils as tu
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(
        shape, generator=generator, device=device, dtype=dtype)

    def get_vae_size_scale_factor(vae):
        return (vae.temporal_downscale_factor, vae.spatial_downscale_factor, vae.spatial_downscale_factor)

    def latent_to_pixel_coords(latent_coords, vae, causal_fix=False):   # vae_encode.py:190-225
        sf = get_vae_size_scale_factor(vae)
        pc = latent_coords * torch.tensor(sf, device=latent_coords.device)[None, :, None]
        if causal_fix:
            pc[:, 0] = (pc[:, 0] + 1 - sf[0]).clamp(min=0)
        return pc

    identity = lambda x, *a, **k: x   # noqa: E731  (the caller passes latents where the pipeline expects pixels)
    mod("ltx_video.models.autoencoders.causal_video_autoencoder", CausalVideoAutoencoder=_FakeVAE)
    mod("ltx_video.models.autoencoders.vae_encode", get_vae_size_scale_factor=get_vae_size_scale_factor,
        latent_to_pixel_coords=latent_to_pixel_coords, vae_decode=identity, vae_encode=identity,
        un_normalize_latents=identity, normalize_latents=identity)
    mod("ltx_video.models.autoencoders.latent_upsampler", LatentUpsampler=_Empty)


def load():
    _install_stand_ins()
    import ltx_video.pipelines.pipeline_ltx_video as pl
    return pl


def make_pipeline(transformer, scheduler, patchifier):
    """An LTXVideoPipeline around the given reference transformer (no tokenizer / text encoder: embeddings are
    passed in), built through the reference's own __init__."""
    pl = load()
    return pl.LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=_FakeVAE(), transformer=transformer,
                               scheduler=scheduler, patchifier=patchifier)


def run_pipeline(pipe, latent_fhw, frame_rate, seed, prompt_embeds, prompt_mask, ref_latents, pose_latents,
                 num_inference_steps, **kw):
    """Calls the reference's `__call__` with `output_type="latent"`; returns (final latents [B,C,F,H,W], the initial
    noise tokens [B,N,C] the pipeline drew from `seed`, the fractional coordinates it fed the transformer)."""
    f, h, w = latent_fhw
    # without a tokenizer / text encoder the negative prompt must come as embeddings too (encode_prompt :406-430)
    kw.setdefault("negative_prompt_embeds", torch.zeros_like(prompt_embeds))
    kw.setdefault("negative_prompt_attention_mask", torch.zeros_like(prompt_mask))
    seen = {}
    tr = pipe.transformer
    orig_forward = tr.forward

    def spy(*a, **k):
        if "coords" not in seen:
            seen["coords"] = k["indices_grid"].clone()
            seen["tokens"] = a[0].clone()
        return orig_forward(*a, **k)
    tr.forward = spy
    try:
        out = pipe(height=h * 32, width=w * 32, num_frames=(f - 1) * 8 + 1, frame_rate=frame_rate, prompt=None,
                   negative_prompt=None, num_inference_steps=num_inference_steps,
                   generator=torch.Generator().manual_seed(seed), prompt_embeds=prompt_embeds,
                   prompt_attention_mask=prompt_mask, output_type="latent", return_dict=False, is_video=True,
                   vae_per_channel_normalize=True, ref_image=ref_latents, pose_frames=pose_latents, **kw)[0]
    finally:
        tr.forward = orig_forward
    return out, seen["tokens"], seen["coords"]
