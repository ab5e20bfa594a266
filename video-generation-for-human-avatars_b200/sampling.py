"""Rectified-flow sampling loop around the B200 transformer forward (BASELINE config 4).

Mirrors the denoising loop of the reference pipeline (pipelines/pipeline_ltx_video.py:1089-1288): per step
    conds = 1 (+1 classifier-free guidance if guidance_scale > 1) (+1 spatio-temporal guidance if stg_scale > 0)
    v     = transformer(cat([x] * conds), ..., prompt batch slice of [negative, positive, positive],
                        timestep = t_i  (or min(t_i, 1 - conditioning_mask) per token), skip-layer mask for STG)
    v     = guidance combine (CFG / CFG* / STG / std rescale)                       (:1217-1260)
    x    <- where(token still to denoise, x - dt v, x)                              (:1346-1379, rf.py:305-374)
Everything after the transformer call is ONE fused launch sequence (ops.guidance_step_: at most three small kernels)
that also writes the next step's bf16 model input, one copy per condition.

Reference behaviours kept (they change results):
* the running latents are fp32 from the first Euler step on (dt is an fp32 tensor), the model sees their bf16
  rounding;
* with one condition the transformer receives the latents tensor itself and lerps the ref / pose conditioning into
  it IN PLACE (transformer3d.py:447-466).  That aliasing only exists while the latents have the model's
  dtype, i.e. on the first step: the caller's bf16 `latents` come back conditioned, later steps condition a copy;
* skip-layer mask columns `ptb_index::num_conds`, `rescaling_scale == 1` meaning "off", and the first batch entry's
  timestep row driving every sample's Euler step -- see oracle/ref_sampling.py for the line references.

`Denoiser(model, scheduler, graph=True)` captures a step (transformer forward + guidance tail) once per guidance
configuration and replays it on this and every later call of the same geometry; the per-step values (timestep tensor,
dt, guidance scalars) and the per-call inputs live in static device buffers refreshed by small copies, so nothing in
the loop synchronises with the host."""
from typing import Callable, List, Optional, Sequence, Union

import torch

from . import ops
from .lib import B200Error

Scale = Union[float, Sequence[float]]


def _per_step(value: Scale, n: int, name: str) -> List[float]:
    if isinstance(value, (list, tuple)):
        if len(value) != n:
            raise B200Error(f"denoise: {name} has {len(value)} entries for {n} steps")
        return [float(v) for v in value]
    return [float(value)] * n


def _step_tables(timesteps: torch.Tensor, conditioning_mask: Optional[torch.Tensor]):
    """Per-step timestep rows and Euler steps, all steps at once and without host round trips.
    Returns (t_rows [S, B, T] fp32 with T = 1 or N, dt [S, T] fp32): the model's timestep input is min(t_i, 1 - mask)
    (pipeline :1143-1171); dt is that row of the FIRST batch entry minus the next lower grid value (rf.py:343-360
    on `current_timestep[:1]`, pipeline :1262)."""
    S = timesteps.numel()
    ts = timesteps.to(torch.float32)
    if conditioning_mask is None:
        rows = ts.view(S, 1, 1)
    else:
        rows = torch.minimum(ts.view(S, 1, 1), (1.0 - conditioning_mask.to(torch.float32))[None])
    grid = torch.cat([ts, torch.zeros(1, device=ts.device)])
    first = rows[:, 0]                                            # [S, T]
    below = grid[:, None, None] < first[None] - 1e-6              # [S+1, S, T]
    lower, _ = (below * grid[:, None, None]).max(dim=0)
    return rows.contiguous(), (first - lower).contiguous()


class Denoiser:
    """The sampling loop as an object: static device buffers and, with graph=True, the captured step of every guidance
    configuration survive between calls, so a serving process pays the capture once per (shape, configuration)."""

    def __init__(self, model, scheduler, graph: bool = False):
        self.model, self.scheduler, self.graph = model, scheduler, graph
        self._sig = None
        self._cfgs = {}
        self._st = {}

    def _static(self, name, src):
        """Refresh (or create) the static tensor `name` with the contents of `src`."""
        cur = self._st.get(name)
        if cur is None or cur.shape != src.shape or cur.dtype != src.dtype:
            self._st[name] = cur = torch.empty_like(src, memory_format=torch.contiguous_format)
        cur.copy_(src)
        return cur

    @torch.no_grad()
    def __call__(self, latents: torch.Tensor, indices_grid: torch.Tensor, ref_image_latents: torch.Tensor,
                 pose_latents: torch.Tensor, prompt_embeds: torch.Tensor, prompt_attention_mask: torch.Tensor,
                 num_inference_steps: int = 40, callback: Optional[Callable[[int, torch.Tensor], None]] = None, *,
                 negative_prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_attention_mask: Optional[torch.Tensor] = None, guidance_scale: Scale = 1.0,
                 stg_scale: Scale = 0.0, rescaling_scale: Scale = 1.0, cfg_star_rescale: bool = False,
                 skip_block_list=None, skip_layer_strategy=None,
                 conditioning_mask: Optional[torch.Tensor] = None, stochastic_sampling: bool = False) -> torch.Tensor:
        model, scheduler = self.model, self.scheduler
        if stochastic_sampling:
            # the fused tail is the deterministic Euler step both shipped configs use (inference-avatars.yaml:14);
            # the re-noising variant exists as RectifiedFlowScheduler.step(..., stochastic_sampling=True)
            raise B200Error("denoise: stochastic_sampling is not part of the fused loop; drive the model with "
                            "RectifiedFlowScheduler.step(..., stochastic_sampling=True) instead")
        if latents.dtype != torch.bfloat16 or not latents.is_cuda or latents.dim() != 3 or not latents.is_contiguous():
            raise B200Error("denoise: latents must be contiguous CUDA bfloat16 tokens [B, N, C]")
        dev = latents.device
        B, N, C = latents.shape
        scheduler.set_timesteps(num_inference_steps, samples_shape=latents.shape, device=dev)
        timesteps = scheduler.timesteps
        S = int(timesteps.numel())
        gs_l = _per_step(guidance_scale, S, "guidance_scale")
        stg_l = _per_step(stg_scale, S, "stg_scale")
        rs_l = _per_step(rescaling_scale, S, "rescaling_scale")
        if skip_block_list is not None and (len(skip_block_list) == 0
                                            or not isinstance(skip_block_list[0], (list, tuple))):
            skip_block_list = [list(skip_block_list)] * S
        if skip_block_list is not None and len(skip_block_list) != S:
            raise B200Error(f"denoise: skip_block_list has {len(skip_block_list)} entries for {S} steps")
        if conditioning_mask is not None and tuple(conditioning_mask.shape) != (B, N):
            raise B200Error("denoise: conditioning_mask must be [B, N]")

        sig = (B, N, C, S, tuple(prompt_embeds.shape[1:]), conditioning_mask is not None, tuple(indices_grid.shape),
               tuple(ref_image_latents.shape), tuple(pose_latents.shape), str(dev), bool(cfg_star_rescale),
               str(skip_layer_strategy))
        if sig != self._sig:        # new geometry: every static buffer and captured step is stale
            self._sig, self._cfgs, self._st = sig, {}, {}

        def _rep(t, lead):
            return t.expand(lead, *t.shape[1:]) if t.shape[0] != lead else t
        pos_e, pos_m = _rep(prompt_embeds, B), _rep(prompt_attention_mask, B)
        neg_e = torch.zeros_like(pos_e) if negative_prompt_embeds is None else _rep(negative_prompt_embeds, B)
        neg_m = (torch.zeros_like(pos_m) if negative_prompt_attention_mask is None
                 else _rep(negative_prompt_attention_mask, B))

        flags = [(gs_l[i] > 1.0, stg_l[i] > 0.0) for i in range(S)]
        scal = torch.tensor([[gs_l[i], stg_l[i], rs_l[i], 0.0] for i in range(S)], device=dev)
        scal[:, 3] = timesteps.to(torch.float32)      # device copy of t: no host read of the schedule
        rows, dts = _step_tables(timesteps, conditioning_mask)
        scal_table, t_rows, dt_table = self._static("scal", scal), self._static("rows", rows), self._static("dts", dts)
        noise_level = None
        if conditioning_mask is not None:
            noise_level = self._static("noise_level", 1.0 - conditioning_mask.to(torch.float32))
        if "ws" not in self._st:
            self._st["ws"] = torch.empty(max(8, ops._L().b200_guidance_step_workspace_bytes(B)), device=dev,
                                         dtype=torch.uint8)
        workspace = self._st["ws"]
        x32 = self._static("x32", latents.float())
        T = t_rows.shape[2]
        touched = set()

        # per guidance configuration: prompt slice, replicated conditioning inputs, model-input and per-step buffers
        def config(i):
            do_cfg, do_stg = flags[i]
            skip = tuple(skip_block_list[i]) if (skip_block_list is not None and do_stg) else None
            rescale = bool(do_stg and rs_l[i] != 1.0)                 # pipeline :1093, :1246
            key = (do_cfg, do_stg, skip, rescale)
            cf = self._cfgs.get(key)
            if cf is None:
                n = 1 + int(do_cfg) + int(do_stg)
                cf = self._cfgs[key] = dict(
                    n=n, do_cfg=do_cfg, do_stg=do_stg, rescale=rescale, runs=0, graph=None, launches=0,
                    slm=model.create_skip_layer_mask(B, n, n - 1, list(skip)) if skip else None,
                    xin=torch.empty((n * B, N, C), device=dev, dtype=torch.bfloat16),
                    t_in=torch.empty((n * B, T), device=dev, dtype=torch.float32),
                    dt=torch.empty((T,), device=dev, dtype=torch.float32),
                    scal=torch.empty((4,), device=dev, dtype=torch.float32))
            if key not in touched:   # this call's conditioning inputs go into the configuration's static tensors
                touched.add(key)
                n = cf["n"]
                parts_e = ([neg_e] if do_cfg else []) + [pos_e] + ([pos_e] if do_stg else [])
                parts_m = ([neg_m] if do_cfg else []) + [pos_m] + ([pos_m] if do_stg else [])
                for name, src in (("enc", torch.cat(parts_e)), ("msk", torch.cat(parts_m)),
                                  ("grid", torch.cat([indices_grid] * n)), ("ref", torch.cat([ref_image_latents] * n)),
                                  ("pose", torch.cat([pose_latents] * n))):
                    if name in cf and cf[name].shape == src.shape and cf[name].dtype == src.dtype:
                        cf[name].copy_(src)
                    else:
                        cf[name] = src.contiguous()
            return cf

        def run_model(cf, x_model):
            return model(hidden_states=x_model, indices_grid=cf["grid"], ref_image_hidden_states=cf["ref"],
                         pose_hidden_states=cf["pose"], encoder_hidden_states=cf["enc"], timestep=cf["t_in"],
                         encoder_attention_mask=cf["msk"], skip_layer_mask=cf["slm"],
                         skip_layer_strategy=skip_layer_strategy, return_dict=False)[0].contiguous()

        def run_tail(cf, v, x_next):
            ops.guidance_step_(v, x32, x_next, cf["dt"], noise_level, cf["scal"], cf["do_cfg"], cf["do_stg"],
                               cfg_star=cfg_star_rescale, rescale=cf["rescale"], workspace=workspace)

        was_training = model.training
        model.eval()
        try:
            for i in range(S):
                cf = config(i)
                n = cf["n"]
                nxt = config(i + 1) if i + 1 < S else None
                cf["t_in"].copy_(t_rows[i].repeat(n, 1) if n > 1 else t_rows[i])
                cf["dt"].copy_(dt_table[i])
                cf["scal"].copy_(scal_table[i])
                if i == 0 and n == 1:
                    # the latents still have the model's dtype: a single condition hands them to the transformer as
                    # they are and gets them back conditioned; the Euler step then starts from the conditioned values
                    v = run_model(cf, latents)
                    x32.copy_(latents)
                    run_tail(cf, v, nxt["xin"] if nxt is not None else None)
                else:
                    if i == 0:
                        cf["xin"].copy_(latents.repeat(n, 1, 1))
                    stays = nxt is None or nxt is cf   # the step may rewrite its own model input for the next one
                    if self.graph and stays and cf["runs"] >= 1:
                        if cf["graph"] is None:
                            torch.cuda.synchronize()
                            cf["graph"] = torch.cuda.CUDAGraph()
                            l0 = ops.launch_count
                            with torch.cuda.graph(cf["graph"]):
                                run_tail(cf, run_model(cf, cf["xin"]), cf["xin"])
                            cf["launches"] = ops.launch_count - l0
                            ops._count(-cf["launches"])  # recorded, not executed
                        cf["graph"].replay()
                        ops._count(cf["launches"])
                    else:
                        run_tail(cf, run_model(cf, cf["xin"]), nxt["xin"] if nxt is not None else None)
                cf["runs"] += 1
                if callback is not None:
                    callback(i, x32)
        finally:
            model.train(was_training)
        return x32.clone()


def denoise(model, latents: torch.Tensor, indices_grid: torch.Tensor, ref_image_latents: torch.Tensor,
            pose_latents: torch.Tensor, prompt_embeds: torch.Tensor, prompt_attention_mask: torch.Tensor, scheduler,
            num_inference_steps: int = 40, callback: Optional[Callable[[int, torch.Tensor], None]] = None,
            graph: bool = False, **guidance) -> torch.Tensor:
    """One sampling run.  latents: [B, N, C] CUDA bf16 noise tokens; indices_grid: [B, 3, N] fractional pixel
    coordinates as the pipeline builds them; prompt tensors [B or 1, L, *].  Keyword arguments as the pipeline's:
    negative_prompt_embeds / negative_prompt_attention_mask, guidance_scale, stg_scale, rescaling_scale (one value or
    one per step), cfg_star_rescale, skip_block_list, skip_layer_strategy, conditioning_mask ([B, N] in [0, 1]; 1 =
    hard conditioning, never denoised).  Returns the final latents as a new fp32 [B, N, C] tensor (the reference's
    latents are fp32 after the first step); `latents` itself is conditioned in place on the first step when that step
    runs a single condition.  A one-shot call gains nothing from graph=True (the capture costs more than it saves
    within one run): keep a Denoiser for repeated runs of one geometry."""
    return Denoiser(model, scheduler, graph=graph)(latents, indices_grid, ref_image_latents, pose_latents,
                                                   prompt_embeds, prompt_attention_mask, num_inference_steps, callback,
                                                   **guidance)
