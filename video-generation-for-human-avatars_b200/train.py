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
    # python scalars, not device tensors: no host-to-device copy, so the step can be captured in a CUDA graph
    mu, sigma = float(config.rf_log_normal_mu), float(config.rf_log_normal_sigma)
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
    from .modules import side
    sp = side(root).get("sp")
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


class GraphedTrainStep:
    """One optimiser micro-step -- zero grads, `train_step`, backward, the bucketed gradient all-reduce (when a
    `dp.GradBucketer` is given) and the optimizer update -- captured ONCE in a CUDA graph and replayed.

    The step is ~1500 kernel launches, a third of them 4-15 us long (LoRA GEMMs, AdaLN / fill / cast glue): issued
    one by one from Python the GPU idles ~8 % of the step waiting for the host (torch.profiler: 87.5 ms of kernels
    in a 95.5 ms step).  Replaying a graph removes the host from the step entirely.  Inputs live in static device
    buffers (`load()` copies a batch into them, from pinned host memory without a sync); the timestep / noise
    draws inside the graph use the graph-safe philox generator, so every replay draws fresh values.
    The optimizer must be capturable (`torch.optim.AdamW(..., fused=True, capturable=True)`).
    Not for the sequence-parallel ring: its NCCL send/recv hops hang under stream capture (tried with both
    capture error modes), so that path stays eager."""

    def __init__(self, model, optimizer, scheduler, patchifier, config, prompt_embeds, prompt_attention_mask,
                 example_batch: dict, bucketer=None, warmup: int = 3, device=None,
                 capture_error_mode: str = "global", side_work: bool = True, dp_mode: str = "one_graph"):
        self.model, self.opt, self.bucketer = model, optimizer, bucketer
        self.mode = "single GPU: whole micro-step replayed as one CUDA graph"
        device = device or next(model.parameters()).device
        self.static = {k: v.to(device).clone() for k, v in example_batch.items()}
        args = (scheduler, patchifier, config, prompt_embeds, prompt_attention_mask)

        def fwd_bwd():
            if bucketer is not None:
                bucketer.zero_grad()
            else:
                optimizer.zero_grad(set_to_none=True)
            ops.side_stream_used = False
            loss, rel_mse, nrmse, _ = train_step(model, self.static, *args, device=device)
            loss.backward()
            # the LoRA weight gradients queued off the critical path: join them (a stream nothing was forked onto --
            # train_mode="full" has no adapters -- must not be waited on inside a capture)
            if ops.side_stream is not None and ops.side_stream_used:
                torch.cuda.current_stream(device).wait_stream(ops.side_stream)
            return loss.detach(), rel_mse.detach(), nrmse.detach()

        def one_step():
            out = fwd_bwd()
            if bucketer is not None:
                bucketer.finish()
            optimizer.step()
            return out

        if ops.timer is not None:
            raise RuntimeError("GraphedTrainStep: kernel timing events cannot be captured (ops.timer must be None)")
        # AccumulateGrad nodes of an earlier eager step that are still alive (a retained loss tensor) would run
        # on the stream they were created on -- the default stream -- and invalidate the capture
        import gc
        gc.collect()
        # Single GPU: the step is captured on a high-priority stream and the work only the optimizer consumes (LoRA
        # weight gradients) on a default-priority one, so it fills SMs the main kernels leave idle.  With a bucketer
        # the gradients are accumulated into bucket views by autograd on the main stream: everything stays there.
        main = torch.cuda.Stream(device=device, priority=-1)
        if bucketer is None and side_work:
            ops.side_stream = torch.cuda.Stream(device=device, priority=0)
        try:
            main.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(main):
                for _ in range(max(warmup, 1)):  # allocator / lazy-init warm-up on a side stream, as torch requires
                    one_step()
            torch.cuda.current_stream(device).wait_stream(main)
            torch.cuda.synchronize(device)
            l0 = ops.launch_count
            self.graph = torch.cuda.CUDAGraph()
            self.graph_opt = None
            if bucketer is None:
                optimizer.zero_grad(set_to_none=True)  # .grad is then allocated from the graph's private pool
                with torch.cuda.graph(self.graph, stream=main, capture_error_mode=capture_error_mode):
                    self.loss, self.rel_mse, self.nrmse = one_step()
            elif dp_mode == "one_graph":
                # data parallel, as specified: ONE graph.  The post-accumulate-grad hooks run during the capture, so
                # each bucket's NCCL all-reduce is recorded where its last gradient becomes ready -- on NCCL's own
                # stream, forked from and joined back into the capture by events -- and overlaps the rest of the
                # backward on every replay; finish() records the join and the 1/world scaling, then the optimizer.
                # (thread_local: NCCL's watchdog thread may query events of earlier eager collectives meanwhile.)
                bucketer.overlap = True
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    self.loss, self.rel_mse, self.nrmse = one_step()
                self.mode = (f"dp{bucketer.world}: whole micro-step as one CUDA graph, the {len(bucketer.buckets)} bucket "
                             "all-reduces captured at their gradient-ready points (overlapped with the backward)")
            else:
                # fallback: graph 1 = zero grads + forward + backward (the bucket hooks only count), the gradient
                # buckets all-reduced eagerly over NCCL, graph 2 = the optimizer update
                bucketer.overlap = False
                with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
                    self.loss, self.rel_mse, self.nrmse = fwd_bwd()
                self.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_opt, pool=self.graph.pool(), capture_error_mode=capture_error_mode):
                    optimizer.step()
                self.mode = (f"dp{bucketer.world}: two CUDA graphs (zero+fwd+bwd | optimizer) around the eager NCCL "
                             "bucket all-reduce")
        finally:
            ops.side_stream = None
        self.launches = ops.launch_count - l0  # b200 kernel launches captured per step

    def describe(self) -> str:
        return self.mode

    def load(self, batch: dict, non_blocking: bool = True):
        for k, dst in self.static.items():
            dst.copy_(batch[k], non_blocking=non_blocking)

    def __call__(self, batch: dict = None):
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        if self.graph_opt is not None:
            self.bucketer.reduce_now()
            self.graph_opt.replay()
        return self.loss
