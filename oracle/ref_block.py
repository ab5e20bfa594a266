"""Tier-two oracle: portable plain-torch restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file; the product path (b200_ltx) never does.

It restates, as pure functions over a flat parameter dict that uses the reference's own
state-dict names, the arithmetic of

  Transformer3DModel.forward          /root/reference/ltx_video/models/transformers/transformer3d.py:361-565
  precompute_freqs_cis                transformer3d.py:209-277
  BasicTransformerBlock.forward       /root/reference/ltx_video/models/transformers/attention.py:198-321
  AttnProcessor2_0.__call__           attention.py:935-1114   (apply_rotary_emb :917-932)
  FeedForward.forward                 attention.py:1257-1264
  SymmetricPatchifier                 models/transformers/symmetric_patchifier.py:33-84
  RectifiedFlowScheduler              schedulers/rf.py:305-426 (step, add_noise, build_velocity_target)
  train_step (deterministic part)     ltx_video/training.py:119-164
plus the diffusers 0.35.1 / peft 0.17.1 leaves those files call.

Pinning: oracle/make_golden.py runs the reference's unmodified modules (tier one,
oracle/ref_import.py) and this file on the same weights and inputs; tests/test_oracle.py asserts
bit-equality on CPU fp32 (output, loss and every LoRA / caption-projection gradient) both against
the live reference when /root/reference is mounted and against the committed tests/golden/*.pt
vectors when it is not.  The third-party leaves (diffusers/peft) themselves are "parity
unpinned": the real packages cannot be installed offline.

Every function is dtype-transparent: run it on fp32 tensors for the fp32 oracle, or on bf16
tensors to reproduce the reference's own bf16 dtype flow (used for the E_ref side of the
two-sided tolerance in tests/).
"""
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# parameter access
# --------------------------------------------------------------------------------------------
def strip_peft_names(state: Params) -> Params:
    """Map peft-wrapped names (base_model.model.X.base_layer.weight) onto plain reference names."""
    out = {}
    for k, v in state.items():
        k = k.replace("base_model.model.", "").replace(".base_layer.", ".")
        out[k] = v
    return out


def linear(P: Params, name: str, x: Tensor, lora_scaling: float = 1.0) -> Tensor:
    """nn.Linear, plus the peft lora.Linear branch when `name.lora_A.default.weight` exists."""
    y = F.linear(x, P[name + ".weight"], P.get(name + ".bias"))
    a = P.get(name + ".lora_A.default.weight")
    if a is not None:
        b = P[name + ".lora_B.default.weight"]
        out_dtype = y.dtype
        xx = x.to(a.dtype)
        y = y + F.linear(F.linear(xx, a), b) * lora_scaling
        y = y.to(out_dtype)
    return y


# --------------------------------------------------------------------------------------------
# leaves (diffusers semantics)
# --------------------------------------------------------------------------------------------
def rms_norm(x: Tensor, weight: Optional[Tensor], eps: float) -> Tensor:
    in_dtype = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if weight is not None:
        if weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(weight.dtype)
        x = x * weight
    else:
        x = x.to(in_dtype)
    return x


def timestep_sinusoid(t: Tensor, dim: int = 256) -> Tensor:
    half = dim // 2
    expo = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device)
    expo = expo / (half - 0)
    ang = t[:, None].float() * torch.exp(expo)[None, :]
    ang = 1 * ang
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip_sin_to_cos


def adaln_single(P: Params, t_flat: Tensor, hidden_dtype: torch.dtype) -> Tuple[Tensor, Tensor]:
    """diffusers AdaLayerNormSingle: returns (linear(silu(e)) [n,6D], e [n,D])."""
    proj = timestep_sinusoid(t_flat).to(dtype=hidden_dtype)
    e = linear(P, "adaln_single.emb.timestep_embedder.linear_1", proj)
    e = linear(P, "adaln_single.emb.timestep_embedder.linear_2", F.silu(e))
    return linear(P, "adaln_single.linear", F.silu(e)), e


def caption_projection(P: Params, enc: Tensor) -> Tensor:
    h = linear(P, "caption_projection.linear_1", enc)
    h = F.gelu(h, approximate="tanh")
    return linear(P, "caption_projection.linear_2", h)


# --------------------------------------------------------------------------------------------
# token layout (symmetric_patchifier.py:33-84, patch size 1)
# --------------------------------------------------------------------------------------------
def latent_coords(f: int, h: int, w: int, batch: int, device=None) -> Tensor:
    g = torch.meshgrid(torch.arange(f, device=device), torch.arange(h, device=device),
                       torch.arange(w, device=device), indexing="ij")
    return torch.stack(g, dim=0).reshape(3, -1).unsqueeze(0).repeat(batch, 1, 1)


def patchify(latents: Tensor) -> Tuple[Tensor, Tensor]:
    b, c, f, h, w = latents.shape
    # einops expresses "b c f h w -> b (f h w) c" as a transposed VIEW of a contiguous input; keep
    # that (randn_like(tokens) in training.py:138 inherits the strides, so the RNG stream depends on it).
    tokens = latents.flatten(2).transpose(1, 2)
    return tokens, latent_coords(f, h, w, b, latents.device)


def unpatchify_view(tokens: Tensor, h: int, w: int) -> Tensor:
    """[B, F*h*w, C] -> [B, C, F, h, w] as a VIEW of `tokens` (the reference relies on that)."""
    b, n, c = tokens.shape
    return tokens.view(b, n // (h * w), h, w, c).permute(0, 4, 1, 2, 3)


# --------------------------------------------------------------------------------------------
# RoPE (transformer3d.py:209-277, attention.py:917-932)
# --------------------------------------------------------------------------------------------
def rope_table(indices_grid: Tensor, dim: int, theta: float, max_pos, out_dtype: torch.dtype):
    frac = torch.stack([indices_grid[:, i] / max_pos[i] for i in range(3)], dim=-1)  # [B,N,3]
    omega = theta ** torch.linspace(math.log(1, theta), math.log(theta, theta), dim // 6,
                                    device=frac.device, dtype=torch.float32)
    omega = omega.to(torch.float32) * math.pi / 2
    ang = (omega * (frac.unsqueeze(-1) * 2 - 1)).transpose(-1, -2).flatten(2)  # [B,N,3*(dim//6)]
    cos = ang.cos().repeat_interleave(2, dim=-1)
    sin = ang.sin().repeat_interleave(2, dim=-1)
    pad = dim % 6
    if pad:
        cos = torch.cat([torch.ones_like(cos[:, :, :pad]), cos], dim=-1)
        sin = torch.cat([torch.zeros_like(cos[:, :, :pad]), sin], dim=-1)
    return cos.to(out_dtype), sin.to(out_dtype)


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    pairs = x.unflatten(-1, (-1, 2))
    x_even, x_odd = pairs.unbind(dim=-1)
    rot = torch.stack((-x_odd, x_even), dim=-1).flatten(-2)
    return x * cos + rot * sin


# --------------------------------------------------------------------------------------------
# attention + block
# --------------------------------------------------------------------------------------------
STG_ATTENTION_SKIP = "attention_skip"
STG_ATTENTION_VALUES = "attention_values"
STG_RESIDUAL = "residual"
STG_TRANSFORMER_BLOCK = "transformer_block"


def attention(P: Params, prefix: str, x: Tensor, heads: int, freqs=None, ctx: Optional[Tensor] = None,
              mask_bias: Optional[Tensor] = None, lora_scaling: float = 1.0,
              skip_layer_mask: Optional[Tensor] = None, skip_layer_strategy: Optional[str] = None) -> Tensor:
    """AttnProcessor2_0 (attention.py:935-1114) for one Attention module named `prefix`."""
    src = x if ctx is None else ctx
    b, lk, _ = src.shape
    if skip_layer_mask is not None:
        skip_layer_mask = skip_layer_mask.reshape(b, 1, 1)
    if mask_bias is not None:  # [B,1,L] additive -> [B,H,1,L]
        mask_bias = mask_bias.unsqueeze(1).repeat_interleave(heads, dim=1)
        mask_bias = mask_bias.view(b, heads, -1, mask_bias.shape[-1])
    q = rms_norm(linear(P, prefix + ".to_q", x, lora_scaling), P.get(prefix + ".q_norm.weight"), 1e-5)
    k = rms_norm(linear(P, prefix + ".to_k", src, lora_scaling), P.get(prefix + ".k_norm.weight"), 1e-5)
    if ctx is None and freqs is not None:
        k = apply_rope(k, *freqs)
        q = apply_rope(q, *freqs)
    v = linear(P, prefix + ".to_v", src, lora_scaling)
    v_for_stg = v
    dh = k.shape[-1] // heads
    qh = q.view(b, -1, heads, dh).transpose(1, 2)
    kh = k.view(b, -1, heads, dh).transpose(1, 2)
    vh = v.view(b, -1, heads, dh).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask_bias, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(b, -1, heads * dh).to(q.dtype)
    if skip_layer_mask is not None and skip_layer_strategy == STG_ATTENTION_SKIP:
        o = o * skip_layer_mask + x * (1.0 - skip_layer_mask)
    elif skip_layer_mask is not None and skip_layer_strategy == STG_ATTENTION_VALUES:
        o = o * skip_layer_mask + v_for_stg * (1.0 - skip_layer_mask)
    return linear(P, prefix + ".to_out.0", o, lora_scaling)


def feed_forward(P: Params, prefix: str, x: Tensor) -> Tensor:
    h = F.gelu(linear(P, prefix + ".net.0.proj", x), approximate="tanh")
    return linear(P, prefix + ".net.2", h)


def block_forward(P: Params, i: int, x: Tensor, freqs, ctx: Tensor, enc_bias: Optional[Tensor],
                  timestep: Tensor, heads: int, lora_scaling: float = 1.0,
                  skip_layer_mask: Optional[Tensor] = None,
                  skip_layer_strategy: Optional[str] = None) -> Tensor:
    """BasicTransformerBlock.forward (attention.py:198-321), adaptive_norm='single_scale_shift'."""
    pre = f"transformer_blocks.{i}"
    b = x.shape[0]
    x_in = x
    table = P[pre + ".scale_shift_table"]
    ada = table[None, None] + timestep.reshape(b, timestep.shape[1], table.shape[0], -1)
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = ada.unbind(dim=2)
    h = rms_norm(x, None, 1e-6) * (1 + scale_msa) + shift_msa
    a = attention(P, pre + ".attn1", h, heads, freqs=freqs, skip_layer_mask=skip_layer_mask,
                  skip_layer_strategy=skip_layer_strategy)
    x = gate_msa * a + x
    c = attention(P, pre + ".attn2", x, heads, freqs=freqs, ctx=ctx, mask_bias=enc_bias,
                  lora_scaling=lora_scaling)
    x = c + x
    h = rms_norm(x, None, 1e-6) * (1 + scale_mlp) + shift_mlp
    x = gate_mlp * feed_forward(P, pre + ".ff", h) + x
    if skip_layer_mask is not None and skip_layer_strategy == STG_TRANSFORMER_BLOCK:
        m = skip_layer_mask.view(-1, 1, 1)
        x = x * m + x_in * (1.0 - m)
    return x


# --------------------------------------------------------------------------------------------
# whole transformer (transformer3d.py:361-565)
# --------------------------------------------------------------------------------------------
LTXV_2B = dict(num_layers=28, num_attention_heads=32, attention_head_dim=64, in_channels=128,
               out_channels=128, caption_channels=4096, cross_attention_dim=2048,
               positional_embedding_theta=10000.0, positional_embedding_max_pos=[20, 2048, 2048],
               timestep_scale_multiplier=1000)


def mask_to_bias(mask: Optional[Tensor], dtype: torch.dtype) -> Optional[Tensor]:
    if mask is not None and mask.ndim == 2:
        mask = ((1 - mask.to(dtype)) * -10000.0).unsqueeze(1)
    return mask


def condition_tokens_(tokens: Tensor, ref: Tensor, pose: Tensor) -> None:
    """In-place ref/pose lerp on the caller's token tensor (transformer3d.py:447-466)."""
    v = unpatchify_view(tokens, ref.shape[3], ref.shape[4])
    v[:, :, 0:1] = torch.lerp(v[:, :, 0:1], ref, 0.85)
    v[:, :, 1:] = torch.lerp(v[:, :, 1:], pose[:, :, 1:], 0.5)


def transformer_forward(P: Params, cfg: dict, hidden_states: Tensor, indices_grid: Tensor,
                        ref_image_hidden_states: Tensor, pose_hidden_states: Tensor,
                        encoder_hidden_states: Tensor, timestep: Tensor,
                        encoder_attention_mask: Optional[Tensor] = None, lora_scaling: float = 1.0,
                        skip_layer_mask: Optional[Tensor] = None,
                        skip_layer_strategy: Optional[str] = None) -> Tensor:
    heads = cfg["num_attention_heads"]
    dim = heads * cfg["attention_head_dim"]
    model_dtype = P["patchify_proj.weight"].dtype
    enc_bias = mask_to_bias(encoder_attention_mask, hidden_states.dtype)
    condition_tokens_(hidden_states, ref_image_hidden_states, pose_hidden_states)
    x = linear(P, "patchify_proj", hidden_states)
    if cfg.get("timestep_scale_multiplier"):
        timestep = cfg["timestep_scale_multiplier"] * timestep
    freqs = rope_table(indices_grid, dim, cfg["positional_embedding_theta"],
                       cfg["positional_embedding_max_pos"], model_dtype)
    b = x.shape[0]
    t6, emb = adaln_single(P, timestep.flatten(), x.dtype)
    t6 = t6.view(b, -1, t6.shape[-1])
    emb = emb.view(b, -1, emb.shape[-1])
    ctx = caption_projection(P, encoder_hidden_states).view(b, -1, x.shape[-1])
    for i in range(cfg["num_layers"]):
        x = block_forward(P, i, x, freqs, ctx, enc_bias, t6, heads, lora_scaling,
                          skip_layer_mask[i] if skip_layer_mask is not None else None,
                          skip_layer_strategy)
    ss = P["scale_shift_table"][None, None] + emb[:, :, None]
    shift, scale = ss[:, :, 0], ss[:, :, 1]
    x = F.layer_norm(x, (dim,), None, None, 1e-6)
    x = x * (1 + scale) + shift
    return linear(P, "proj_out", x)


# --------------------------------------------------------------------------------------------
# rectified flow (schedulers/rf.py:305-426) and the loss (training.py:138-164)
# --------------------------------------------------------------------------------------------
def _append_dims(t: Tensor, ndim: int) -> Tensor:
    return t[(...,) + (None,) * (ndim - t.ndim)]


def rf_add_noise(x0: Tensor, noise: Tensor, t: Tensor) -> Tensor:
    sig = _append_dims(t, x0.ndim)
    return (1 - sig) * x0 + sig * noise


def rf_velocity_target(x0: Tensor, noise: Tensor, t: Tensor) -> Tensor:
    a_dot = _append_dims(torch.full_like(t, -1.0), x0.ndim)
    s_dot = _append_dims(torch.full_like(t, 1.0), x0.ndim)
    return a_dot * x0 + s_dot * noise


def rf_step(timesteps_grid: Tensor, model_output: Tensor, timestep: Tensor, sample: Tensor,
            stochastic_sampling: bool = False) -> Tensor:
    """Euler step (rf.py:305-374); `timesteps_grid` is scheduler.timesteps (descending).  stochastic_sampling: the x0
    estimate re-noised to the next level with a draw from the global generator (rf.py:362-365)."""
    eps = 1e-6
    grid = torch.cat([timesteps_grid, torch.zeros(1, device=timesteps_grid.device)])
    if timestep.ndim == 0:
        lower = grid[grid < timestep - eps][0]
        dt = timestep - lower
    else:
        assert timestep.ndim == 2
        below = grid[:, None, None] < timestep[None] - eps
        lower, _ = (below * grid[:, None, None]).max(dim=0)
        dt = (timestep - lower)[..., None]
    if stochastic_sampling:
        x0 = sample - timestep[..., None] * model_output
        nxt = timestep[..., None] - dt
        nxt = nxt.reshape(nxt.shape + (1,) * (sample.ndim - nxt.ndim))     # append_dims, torch_utils.py:16
        return (1 - nxt) * x0 + nxt * torch.randn_like(sample)
    return sample - dt * model_output


def uniform_timesteps(n: int) -> Tensor:
    return torch.linspace(1, 1 / n, n)


def linear_quadratic_timesteps(n: int, threshold_noise: float = 0.025) -> Tensor:
    if n == 1:
        return torch.tensor([1.0])
    lin = n // 2
    head = [i * threshold_noise / lin for i in range(lin)]
    diff = lin - threshold_noise * n
    quad = n - lin
    qc = diff / (lin * quad ** 2)
    lc = threshold_noise / lin - 2 * diff / (quad ** 2)
    const = qc * (lin ** 2)
    tail = [qc * (i ** 2) + lc * i + const for i in range(lin, n)]
    sched = [1.0 - s for s in head + tail + [1.0]]
    return torch.tensor(sched[:-1])


def train_step_loss(P: Params, cfg: dict, latents: Tensor, ref_latents: Tensor, pose_latents: Tensor,
                    prompt_embeds: Tensor, prompt_mask: Tensor, t: Tensor, noise: Tensor,
                    lora_scaling: float = 1.0) -> Tuple[Tensor, Tensor]:
    """train_step (training.py:94-166) with the random draws (t, noise) passed in.

    Returns (loss, model output).  Dtypes follow the reference: x_t and the target are computed
    from model-dtype tokens with fp32 t, then cast to the model dtype; the MSE runs in model dtype.
    """
    model_dtype = P["patchify_proj.weight"].dtype
    latents = latents.to(model_dtype)
    b = latents.shape[0]
    enc = prompt_embeds.expand(b, -1, -1).to(model_dtype)
    mask = prompt_mask.expand(b, -1)
    tokens, coords = patchify(latents)
    noisy = rf_add_noise(tokens, noise, t).to(model_dtype)
    target = rf_velocity_target(tokens, noise, t).to(model_dtype)
    out = transformer_forward(P, cfg, noisy, coords, ref_latents.to(model_dtype),
                              pose_latents.to(model_dtype), enc, t, mask, lora_scaling)
    return F.mse_loss(out, target, reduction="mean"), out


# --------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------------
def param_shapes(cfg: dict, lora_rank: int = 0) -> Dict[str, Tuple[int, ...]]:
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    cin, cout, cc = cfg["in_channels"], cfg["out_channels"], cfg["caption_channels"]
    xd = cfg.get("cross_attention_dim", d)
    s = {"patchify_proj.weight": (d, cin), "patchify_proj.bias": (d,),
         "adaln_single.emb.timestep_embedder.linear_1.weight": (d, 256),
         "adaln_single.emb.timestep_embedder.linear_1.bias": (d,),
         "adaln_single.emb.timestep_embedder.linear_2.weight": (d, d),
         "adaln_single.emb.timestep_embedder.linear_2.bias": (d,),
         "adaln_single.linear.weight": (6 * d, d), "adaln_single.linear.bias": (6 * d,),
         "caption_projection.linear_1.weight": (d, cc), "caption_projection.linear_1.bias": (d,),
         "caption_projection.linear_2.weight": (d, d), "caption_projection.linear_2.bias": (d,),
         "scale_shift_table": (2, d), "proj_out.weight": (cout, d), "proj_out.bias": (cout,)}
    for i in range(cfg["num_layers"]):
        p = f"transformer_blocks.{i}"
        s[p + ".scale_shift_table"] = (6, d)
        for a in ("attn1", "attn2"):
            kv_in = xd if a == "attn2" else d
            for n, fin in (("to_q", d), ("to_k", kv_in), ("to_v", kv_in), ("to_out.0", d)):
                s[f"{p}.{a}.{n}.weight"] = (d, fin)
                s[f"{p}.{a}.{n}.bias"] = (d,)
                if lora_rank and a == "attn2":
                    s[f"{p}.{a}.{n}.lora_A.default.weight"] = (lora_rank, fin)
                    s[f"{p}.{a}.{n}.lora_B.default.weight"] = (d, lora_rank)
            s[f"{p}.{a}.q_norm.weight"] = (d,)
            s[f"{p}.{a}.k_norm.weight"] = (d,)
        s[p + ".ff.net.0.proj.weight"] = (4 * d, d)
        s[p + ".ff.net.0.proj.bias"] = (4 * d,)
        s[p + ".ff.net.2.weight"] = (d, 4 * d)
        s[p + ".ff.net.2.bias"] = (d,)
    return s


def is_trainable(name: str, train_mode: str = "lora_audio") -> bool:
    """lora_audio strategy (training.py:69-73); any other mode: the key list of training.py:75-91."""
    if train_mode == "lora_audio":
        return ("lora_" in name) or ("caption_projection" in name)
    return any(k in name for k in ("proj_out", "scale_shift_table", "adaln_single", "caption_projection", "attn",
                                   "attn2"))


def init_params(cfg: dict, lora_rank: int = 32, seed: int = 0, dtype=torch.float32,
                device="cpu", lora_b_std: float = 0.02) -> Params:
    """Random-init weights with the reference's initialiser *distributions* (Linear
    kaiming-uniform(a=sqrt5) == U(+-1/sqrt(fan_in)), scale_shift_table N(0,1)/sqrt(D), RMSNorm
    weight 1, LoRA A kaiming-uniform, LoRA B N(0, lora_b_std) so dA != 0 -- SURVEY.md 8d).
    Generated per tensor from a name-keyed CPU generator so any subset is reproducible.
    LoRA adapters stay fp32 (peft autocast_adapter_dtype); everything else is cast to `dtype`."""
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    P = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg, lora_rank).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        if name.endswith("norm.weight"):
            w = torch.ones(shape)
        elif name.endswith("scale_shift_table"):
            w = torch.randn(shape, generator=g) / d ** 0.5
        elif "lora_B" in name:
            w = torch.randn(shape, generator=g) * lora_b_std
        else:
            fan_in = shape[-1] if len(shape) > 1 else None
            if fan_in is None:  # bias: fan_in of the matching weight
                fan_in = P[name[:-4] + "weight"].shape[-1]
            bound = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(shape, generator=g) * 2 - 1) * bound
        keep32 = "lora_" in name
        P[name] = w.to(device=device, dtype=torch.float32 if keep32 else dtype)
    return P


def synthetic_batch(cfg: dict, b: int, f: int, h: int, w: int, n_ctx: int = 256, seed: int = 1234,
                    valid_ctx: Optional[int] = None, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    c, cc = cfg["in_channels"], cfg["caption_channels"]
    batch = dict(latents=torch.randn(b, c, f, h, w, generator=g),
                 pose_latents=torch.randn(b, c, f, h, w, generator=g),
                 ref_image_latents=torch.randn(b, c, 1, h, w, generator=g),
                 prompt_embeds=torch.randn(1, n_ctx, cc, generator=g),
                 noise=torch.randn(b, f * h * w, c, generator=g))
    mask = torch.ones(1, n_ctx, dtype=torch.long)
    if valid_ctx is not None:
        mask[:, valid_ctx:] = 0
    batch["prompt_mask"] = mask
    return {k: v.to(device) for k, v in batch.items()}
