"""Host-side mirror of the reference's operator surface for the LTXV block path.

Same class names, constructor arguments, parameter / state-dict names and forward signatures as
  ltx_video/models/transformers/transformer3d.py:49-565   (Transformer3DModel)
  ltx_video/models/transformers/attention.py:38-321       (BasicTransformerBlock)
  ltx_video/models/transformers/attention.py:325-718      (Attention + processor protocol)
  ltx_video/models/transformers/attention.py:1204-1264    (FeedForward)
  ltx_video/models/transformers/symmetric_patchifier.py   (SymmetricPatchifier)
so that ltx_video/training.py:94-166 (train_step) and pipelines/pipeline_ltx_video.py:1202-1215 can
call it unchanged — but every tensor op on the path is a kernel from libb200ltx.so (ops.py).

The functions `attention_forward`, `block_forward`, `transformer_forward` only read attributes by
the reference's names, so `install()` (api.py) can also bind them onto an instance of the
reference's own classes (with or without peft LoRA wrappers)."""
import math
import weakref
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .lib import B200Error

BF16 = torch.bfloat16


class SkipLayerStrategy:
    """Mirror of ltx_video/utils/skip_layer_strategy.py (enum values compared by name)."""
    AttentionSkip = "AttentionSkip"
    AttentionValues = "AttentionValues"
    Residual = "Residual"
    TransformerBlock = "TransformerBlock"


def _strategy_name(s) -> Optional[str]:
    if s is None:
        return None
    return getattr(s, "name", s)


# ---------------------------------------------------------------------------------------------
# derived per-module state (weight concatenations, RoPE table, sequence-parallel handle, per-forward K/V views)
# ---------------------------------------------------------------------------------------------
# Kept OUTSIDE the modules, keyed weakly by the module object: `copy.deepcopy(model)`, `state_dict()`, pickling and
# `merge_and_unload()` (torch_utils.py:66-102 deep-copies the whole model for every checkpoint export) never see it --
# as module attributes the cached concatenations alone would add 1.2 GB to each copy.
_side_tables = weakref.WeakKeyDictionary()


def side(mod) -> dict:
    d = _side_tables.get(mod)
    if d is None:
        d = _side_tables[mod] = {}
    return d


def drop_weight_caches(model) -> None:
    """Forget every cached weight concatenation / table of `model` and its sub-modules (they are re-built on the
    next forward).  Called by install / uninstall, LoraLinear.merge and merge_and_unload: those change weights
    through raw pointers or `.data`, which does not move the tensors' version counters."""
    for m in model.modules():
        d = _side_tables.get(m)
        if d:
            for k in ("wqkv", "wkv_all", "rope"):
                d.pop(k, None)


# ---------------------------------------------------------------------------------------------
# parameter access that works for nn.Linear and for a peft lora.Linear wrapper
# ---------------------------------------------------------------------------------------------
def linear_parts(mod):
    """-> (weight, bias, (A, B, scaling) | None).  A wrapper whose adapter has been merged into the base weight
    (peft `merge_adapter()` keeps the wrapper and sets `merged`) contributes no separate LoRA term."""
    if hasattr(mod, "base_layer"):
        base = mod.base_layer
        if getattr(mod, "merged", False) or getattr(mod, "disable_adapters", False):
            return base.weight, base.bias, None
        a = mod.lora_A["default"].weight
        b = mod.lora_B["default"].weight
        return base.weight, base.bias, (a, b, float(mod.scaling["default"]))
    return mod.weight, mod.bias, None


def _weights_key(parts):
    """Identity of a set of (weight, bias, lora) triples for the concatenation caches: storage address and version
    counter of every tensor plus whether an adapter rides beside it (so unloading / merging adapters re-keys)."""
    return tuple((w.data_ptr(), w._version, None if b is None else (b.data_ptr(), b._version), lo is not None)
                 for w, b, lo in parts)


def _capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def _require_bf16(t: torch.Tensor, what: str):
    if t.dtype != BF16 or not t.is_cuda:
        raise B200Error(f"{what} must be a CUDA bfloat16 tensor (got {t.dtype} on {t.device}); the b200_ltx "
                        "path has no fp32 / CPU fallback")


def apply_linear(mod, x2d, gate=None, rows_per_gate=0, res=None, join=None, join_role=None):
    W, b, lora = linear_parts(mod)
    return ops.linear(x2d, W, b, lora, gate, rows_per_gate, res, join, join_role)


def _frozen_plain(mod) -> bool:
    """No adapter and no gradient wanted for weight / bias: frozen parameters, or any parameters under no_grad (a
    freshly loaded sampling model has requires_grad = True everywhere and runs under torch.no_grad())."""
    W, b, lora = linear_parts(mod)
    return lora is None and not ops.wants_grad(W, b)


def _cached_wqkv(attn):
    """[3D, D] concatenation of the frozen q/k/v weights (+ [3D] biases): one forward GEMM, and the
    [K,N] operand of the fused dgrad.  Re-built if the parameters are replaced or modified in place."""
    parts = [linear_parts(m) for m in (attn.to_q, attn.to_k, attn.to_v)]
    key = _weights_key(parts)
    tab = side(attn)
    cache = tab.get("wqkv")
    if cache is None or cache[0] != key:
        W = torch.cat([w.detach() for w, _, _ in parts], dim=0).contiguous()
        bias = None
        if parts[0][1] is not None:
            bias = torch.cat([b.detach() for _, b, _ in parts], dim=0).contiguous()
        cache = (key, W, bias)
        tab["wqkv"] = cache
    return cache[1], cache[2]


def _batched_ctx_kv(model, ctx):
    """Project the caption tokens to the attn2 keys / values of EVERY block in one strided-batched GEMM
    (ops.CtxKVFn) and hand each block its column views.  Returns the list of attentions that were given a
    `side(attn)["kv"]` entry (to be cleared after the forward), or [] when the layout does not allow it (then every
    block projects for itself, exactly as the reference does: attention.py:999-1005)."""
    blocks = list(model.transformer_blocks)
    attns = [blk.attn2 for blk in blocks if blk.attn2 is not None]
    if len(attns) != len(blocks) or len(attns) < 2 or ctx is None:
        return []
    B, L, Dc = ctx.shape
    mods = [m for a in attns for m in (a.to_k, a.to_v)]
    parts = [linear_parts(m) for m in mods]
    W0 = parts[0][0]
    D = W0.shape[0]
    if (B * L) % 128 or D % 256 or Dc % 64:
        return []
    has_bias = parts[0][1] is not None
    has_lora = parts[0][2] is not None
    tab = side(model)
    for W, b, lo in parts:
        if W.shape != W0.shape or ops.wants_grad(W, b) or (b is not None) != has_bias \
                or (lo is not None) != has_lora or W.dtype != BF16:
            tab.pop("wkv_all", None)   # trainable weights change through raw pointers: never keep a stale copy
            return []
    r, scaling = 0, 1.0
    adapters = ()
    if has_lora:
        r, scaling = parts[0][2][0].shape[0], parts[0][2][2]
        if any(lo[0].shape[0] != r or lo[2] != scaling for _, _, lo in parts):
            return []
        adapters = tuple(lo[0] for _, _, lo in parts) + tuple(lo[1] for _, _, lo in parts)
    key = _weights_key(parts)
    cache = tab.get("wkv_all")
    if cache is None or cache[0] != key:
        Wkv = torch.cat([W.detach() for W, _, _ in parts], dim=0).contiguous()
        bkv = torch.cat([b.detach() for _, b, _ in parts], dim=0).contiguous() if has_bias else None
        cache = (key, Wkv, bkv)
        tab["wkv_all"] = cache
    G = len(mods)
    outs = ops.CtxKVFn.apply(ctx.reshape(B * L, Dc), cache[1], cache[2], G, float(scaling), int(r), *adapters)
    for i, a in enumerate(attns):
        side(a)["kv"] = (ctx, outs[2 * i], outs[2 * i + 1])
    return attns


# ---------------------------------------------------------------------------------------------
# functional forward
# ---------------------------------------------------------------------------------------------
def _key_bias(mask_bias: Optional[torch.Tensor], B: int, Nk: int) -> Optional[torch.Tensor]:
    """[B,1,Nk] (or [B,Nk]) additive bias -> contiguous fp32 [B,Nk]."""
    if mask_bias is None:
        return None
    if mask_bias.numel() != B * Nk:
        raise B200Error(f"attention mask of shape {tuple(mask_bias.shape)} is not a per-key bias [B,1,{Nk}]")
    return mask_bias.reshape(B, Nk).to(torch.float32).contiguous()


def attention_forward(attn, hidden_states, freqs_cis=None, encoder_hidden_states=None, attention_mask=None,
                      skip_layer_mask=None, skip_layer_strategy=None, gate=None, rows_per_gate=0, res=None):
    """AttnProcessor2_0 semantics (attention.py:935-1114) on [B,N,D] bf16 tensors.  `gate`/`res`
    (2-D) optionally fuse the block's `gate * out + residual` into the output projection."""
    _require_bf16(hidden_states, "hidden_states")
    B, Nq, Din = hidden_states.shape
    H = attn.heads
    D = H * 64
    if linear_parts(attn.to_q)[0].shape[0] != D:
        raise B200Error("only head_dim 64 is built")
    x2d = hidden_states.reshape(B * Nq, Din)
    is_self = encoder_hidden_states is None
    strat = _strategy_name(skip_layer_strategy)
    if is_self:
        src2d, Nk = x2d, Nq
    else:
        _require_bf16(encoder_hidden_states, "encoder_hidden_states")
        Nk = encoder_hidden_states.shape[1]
        src2d = encoder_hidden_states.reshape(B * Nk, encoder_hidden_states.shape[-1])
    kb = _key_bias(attention_mask, B, Nk)
    use_rope = is_self and getattr(attn, "use_rope", False) and freqs_cis is not None
    cos = sin = None
    if use_rope:
        cos = freqs_cis[0].reshape(B * Nq, D)
        sin = freqs_cis[1].reshape(B * Nq, D)
    wqn, wkn = attn.q_norm.weight, attn.k_norm.weight
    scale = float(attn.scale)

    sp = side(attn).get("sp") if is_self else None
    # STG skips of whole batch entries ("attention values" / "attention skip", attention.py:1071-1086) ride inside the
    # attention launch when the mask is known to be 0 / 1 (built by our create_skip_layer_mask) and no gradient is
    # recorded: the skipped entries' CTAs copy their pass-through rows and do no attention work
    batch_keep, pass_input = None, False
    if (is_self and skip_layer_mask is not None and getattr(skip_layer_mask, "_b200_binary", False)
            and strat in (SkipLayerStrategy.AttentionSkip, SkipLayerStrategy.AttentionValues)
            and not torch.is_grad_enabled() and sp is None):
        batch_keep = skip_layer_mask.reshape(B).to(torch.float32).contiguous()
        pass_input = strat == SkipLayerStrategy.AttentionSkip
    qkv_parts = [linear_parts(m) for m in (attn.to_q, attn.to_k, attn.to_v)]
    no_lora = all(lo is None for _, _, lo in qkv_parts) and linear_parts(attn.to_out[0])[2] is None
    frozen = all(_frozen_plain(m) for m in (attn.to_q, attn.to_k, attn.to_v))
    fast = is_self and (skip_layer_mask is None or batch_keep is not None) and no_lora
    if fast and frozen:
        Wqkv, bqkv = _cached_wqkv(attn)   # [3D, D] concatenation of the frozen weights, built once
    elif fast and sp is None:
        # train_mode="full" (training.py:75-91): the projections train.  Same node, fed with an autograd-visible
        # concatenation (25 MB per block and step; its backward hands the [3D, D] weight gradient out as three views)
        side(attn).pop("wqkv", None)
        Wqkv = torch.cat([w for w, _, _ in qkv_parts], dim=0)
        bqkv = torch.cat([b for _, b, _ in qkv_parts], dim=0) if qkv_parts[0][1] is not None else None
    else:
        fast = False
    if fast:
        Wo, bo, _ = linear_parts(attn.to_out[0])
        y = ops.SelfAttnFn.apply(x2d, Wqkv, bqkv, wqn, wkn, cos, sin, Wo, bo, gate, rows_per_gate, res, kb, B, H,
                                 Nq, scale, sp, batch_keep, pass_input)
        return y.view(B, Nq, -1)
    if is_self:
        side(attn).pop("wqkv", None)   # trainable / adapted projections: a cached concatenation would go stale
    if sp is not None:
        raise B200Error("sequence-sharded attn1 needs frozen, adapter-free attn1 projections and no skip-layer mask")

    # attn2 of a block is y = attn(x) + x with x entering at to_q and at the residual of to_out: let the two gradients
    # of x meet in to_q's dgrad epilogue (ops.GradJoin) instead of in an autograd add
    join = None
    if (not is_self and res is not None and gate is None and torch.is_grad_enabled() and x2d.requires_grad
            and res.data_ptr() == x2d.data_ptr() and res.shape == x2d.shape and res.stride() == x2d.stride()):
        join = ops.GradJoin()
    q_pre = apply_linear(attn.to_q, x2d, join=join, join_role="recv")
    pre = side(attn).get("kv") if not is_self else None
    if pre is not None and pre[0] is encoder_hidden_states:
        k_pre, v = pre[1], pre[2]  # projected once for all blocks by transformer_forward (ops.CtxKVFn)
    else:
        k_pre = apply_linear(attn.to_k, src2d)
        v = apply_linear(attn.to_v, src2d)
    # (trainable q_norm / k_norm weights included: the backward kernel emits the products their gradients sum)
    o = ops.AttnCoreFn.apply(q_pre, k_pre, v, wqn, wkn, cos, sin, kb, B, H, Nq, Nk, scale)
    if skip_layer_mask is not None and strat in (SkipLayerStrategy.AttentionSkip, SkipLayerStrategy.AttentionValues):
        m = skip_layer_mask.reshape(B, 1, 1).to(o.dtype)
        other = hidden_states if strat == SkipLayerStrategy.AttentionSkip else v.view(B, Nk, D)
        o = (o.view(B, Nq, D) * m + other * (1.0 - m)).reshape(B * Nq, D)
    y = apply_linear(attn.to_out[0], o, gate, rows_per_gate, res, join=join, join_role="send")
    return y.view(B, Nq, -1)


def block_forward(block, hidden_states, freqs_cis=None, attention_mask=None, encoder_hidden_states=None,
                  encoder_attention_mask=None, timestep=None, cross_attention_kwargs=None, class_labels=None,
                  skip_layer_mask=None, skip_layer_strategy=None):
    """BasicTransformerBlock.forward (attention.py:198-321), adaptive_norm == 'single_scale_shift'."""
    if getattr(block, "adaptive_norm", "single_scale_shift") != "single_scale_shift":
        raise B200Error("only adaptive_norm='single_scale_shift' (the LTXV-2B configuration) is built")
    _require_bf16(hidden_states, "hidden_states")
    B, N, D = hidden_states.shape
    if timestep is None or timestep.ndim != 3:
        raise B200Error("timestep must be [batch, 1 or num_tokens, 6*dim]")
    T = timestep.shape[1]
    if T not in (1, N):
        raise B200Error("timestep must have 1 or num_tokens entries per sample")
    rpm = N // T
    ada = (block.scale_shift_table[None, None] + timestep.reshape(B, T, 6, -1)).reshape(B * T, 6 * D)
    shift_msa, scale_msa, gate_msa = ada[:, 0:D], ada[:, D:2 * D], ada[:, 2 * D:3 * D]
    shift_mlp, scale_mlp, gate_mlp = ada[:, 3 * D:4 * D], ada[:, 4 * D:5 * D], ada[:, 5 * D:]
    x2d = hidden_states.reshape(B * N, D)
    strat = _strategy_name(skip_layer_strategy)

    # the residual input leaves the norm node as its own output: its gradient is added inside norm_mod_bwd
    h, x_res = ops.norm_mod(x2d, scale_msa, shift_msa, rpm, 1e-6, False, with_res=True)
    x1 = attention_forward(block.attn1, h.view(B, N, D), freqs_cis=freqs_cis,
                           encoder_hidden_states=encoder_hidden_states if block.only_cross_attention else None,
                           attention_mask=attention_mask, skip_layer_mask=skip_layer_mask,
                           skip_layer_strategy=skip_layer_strategy, gate=gate_msa, rows_per_gate=rpm, res=x_res)
    x1_2d = x1.reshape(B * N, D)
    if block.attn2 is not None:
        x2 = attention_forward(block.attn2, x1, freqs_cis=freqs_cis, encoder_hidden_states=encoder_hidden_states,
                               attention_mask=encoder_attention_mask, res=x1_2d)
        x2_2d = x2.reshape(B * N, D)
    else:
        x2_2d = x1_2d
    h2, x2_res = ops.norm_mod(x2_2d, scale_mlp, shift_mlp, rpm, 1e-6, False, with_res=True)
    W1, b1, l1 = linear_parts(block.ff.net[0].proj)
    W2, b2, l2 = linear_parts(block.ff.net[2])
    if l1 is not None or l2 is not None:
        raise B200Error("LoRA on the feed-forward is not built (reference targets attn2 only, training.py:51-60)")
    out = ops.FeedForwardFn.apply(h2, W1, b1, W2, b2, gate_mlp, rpm, x2_res).view(B, N, D)
    if skip_layer_mask is not None and strat == SkipLayerStrategy.TransformerBlock:
        m = skip_layer_mask.view(-1, 1, 1).to(out.dtype)
        out = out * m + hidden_states * (1.0 - m)
    return out


def timestep_sinusoid(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def adaln_single_forward(adaln, t_flat: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """diffusers AdaLayerNormSingle (transformer3d.py:160,481-486): (linear(silu(e)), e)."""
    emb = adaln.emb.timestep_embedder
    p = timestep_sinusoid(t_flat).to(BF16)
    e = apply_linear(emb.linear_1, p)
    e = apply_linear(emb.linear_2, torch.nn.functional.silu(e))
    return apply_linear(adaln.linear, torch.nn.functional.silu(e)), e


def rope_table(indices_grid: torch.Tensor, dim: int, theta: float, max_pos: List[int]):
    """precompute_freqs_cis (transformer3d.py:209-277): fp32 angles, bf16 cos/sin [B,N,dim]; same fp32
    op order as the reference because angles reach ~1.5e4 rad."""
    nfreq = dim // 6
    frac = torch.stack([indices_grid[:, i] / max_pos[i] for i in range(3)], dim=-1)
    omega = theta ** torch.linspace(math.log(1, theta), math.log(theta, theta), nfreq, device=frac.device,
                                    dtype=torch.float32)
    omega = omega * math.pi / 2
    ang = (omega * (frac.unsqueeze(-1) * 2 - 1)).transpose(-1, -2).flatten(2)
    cos = ang.cos().repeat_interleave(2, dim=-1)
    sin = ang.sin().repeat_interleave(2, dim=-1)
    pad = dim % 6
    if pad:
        cos = torch.cat([torch.ones_like(cos[:, :, :pad]), cos], dim=-1)
        sin = torch.cat([torch.zeros_like(sin[:, :, :pad]), sin], dim=-1)
    return cos.to(BF16).contiguous(), sin.to(BF16).contiguous()


@dataclass
class Transformer3DModelOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i] if not isinstance(i, str) else getattr(self, i)


def transformer_forward(model, hidden_states, indices_grid, ref_image_hidden_states=None, pose_hidden_states=None,
                        encoder_hidden_states=None, timestep=None, class_labels=None, cross_attention_kwargs=None,
                        attention_mask=None, encoder_attention_mask=None, skip_layer_mask=None,
                        skip_layer_strategy=None, return_dict=True):
    """Transformer3DModel.forward (transformer3d.py:361-565), including the in-place conditioning lerp
    on the caller's token tensor."""
    _require_bf16(hidden_states, "hidden_states")
    if ref_image_hidden_states is None or pose_hidden_states is None:
        raise B200Error("ref_image_hidden_states and pose_hidden_states are required (transformer3d.py:447-465)")
    B, N, C = hidden_states.shape
    dt = hidden_states.dtype
    if attention_mask is not None and attention_mask.ndim == 2:
        attention_mask = ((1 - attention_mask.to(dt)) * -10000.0).unsqueeze(1)
    if encoder_attention_mask is not None and encoder_attention_mask.ndim == 2:
        encoder_attention_mask = ((1 - encoder_attention_mask.to(dt)) * -10000.0).unsqueeze(1)

    with torch.no_grad():
        adapters = []
        for blk in model.transformer_blocks:
            for attn in (blk.attn1, blk.attn2):
                if attn is None:
                    continue
                for m in (attn.to_q, attn.to_k, attn.to_v, attn.to_out[0]):
                    lora = linear_parts(m)[2]
                    if lora is not None:
                        adapters.append(lora)
        ops.prestage_lora(adapters)  # all fp32 adapters -> padded bf16 GEMM operands in a few launches
        tok = hidden_states if hidden_states.is_contiguous() else hidden_states.contiguous()
        # sequence-sharded call: `hidden_states` / `indices_grid` are this rank's contiguous token shard,
        # the conditioning latents are the whole clip
        sp = side(model).get("sp")
        tok_off = sp.rank * N if sp is not None else 0
        ops.lerp_condition_(tok, ref_image_hidden_states.to(dt).contiguous(), pose_hidden_states.to(dt).contiguous(),
                            token_offset=tok_off)
        if tok is not hidden_states:
            hidden_states.copy_(tok)  # keep the reference's side effect on the caller's tensor
    x = apply_linear(model.patchify_proj, tok.view(B * N, C))
    D = x.shape[1]

    mult = getattr(model, "timestep_scale_multiplier", None)
    if mult:
        timestep = mult * timestep
    cfg_theta = model.positional_embedding_theta
    # the table depends only on the coordinates: a caller that passes the SAME tensor again (the sampler's 40 steps)
    # gets the cached bf16 cos / sin instead of ~10 element-wise launches over [B, N, D] fp32
    rkey = (indices_grid.data_ptr(), indices_grid._version, tuple(indices_grid.shape), indices_grid.dtype, D)
    rcache = side(model).get("rope")
    if rcache is not None and rcache[0] == rkey and not _capturing():
        freqs = rcache[1]
    else:
        freqs = rope_table(indices_grid, D, cfg_theta, model.positional_embedding_max_pos)
        if not _capturing():
            # (indices_grid itself is kept alive so that its address cannot be reused by another tensor)
            side(model)["rope"] = (rkey, freqs, indices_grid)
    t6, emb = adaln_single_forward(model.adaln_single, timestep.flatten())
    t6 = t6.view(B, -1, t6.shape[-1])
    emb = emb.view(B, -1, emb.shape[-1])

    ctx = encoder_hidden_states
    if model.caption_projection is not None:
        _require_bf16(encoder_hidden_states, "encoder_hidden_states")
        L = encoder_hidden_states.shape[1]
        cp = model.caption_projection
        W1, b1, _ = linear_parts(cp.linear_1)
        W2, b2, _ = linear_parts(cp.linear_2)
        e2d = encoder_hidden_states.reshape(B * L, encoder_hidden_states.shape[-1])
        ctx = ops.FeedForwardFn.apply(e2d, W1, b1, W2, b2, None, 0, None).view(B, L, D)

    h = x.view(B, N, D)
    checkpointing = model.training and model.gradient_checkpointing
    # everything on the attn2 key/value side depends only on the caption tokens: one batched projection
    # for all blocks (not under activation checkpointing, whose recomputation runs block by block)
    shared_kv = [] if checkpointing else _batched_ctx_kv(model, ctx)
    try:
        for i, block in enumerate(model.transformer_blocks):
            slm = skip_layer_mask[i] if skip_layer_mask is not None else None
            if slm is not None and i not in getattr(skip_layer_mask, "_b200_skip_blocks", (i,)):
                slm = None  # an all-ones row (see create_skip_layer_mask): nothing is skipped in this block
            elif slm is not None and hasattr(skip_layer_mask, "_b200_skip_blocks"):
                slm._b200_binary = True   # our own builder: entries are exactly 0 or 1
            if checkpointing:
                h = torch.utils.checkpoint.checkpoint(block, h, freqs, attention_mask, ctx, encoder_attention_mask,
                                                      t6, cross_attention_kwargs, class_labels, slm,
                                                      skip_layer_strategy, use_reentrant=False)
            else:
                h = block(h, freqs_cis=freqs, attention_mask=attention_mask, encoder_hidden_states=ctx,
                          encoder_attention_mask=encoder_attention_mask, timestep=t6,
                          cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels,
                          skip_layer_mask=slm, skip_layer_strategy=skip_layer_strategy)
    finally:
        for a in shared_kv:
            side(a).pop("kv", None)
        ops.clear_lora_stage()   # the staged bf16 adapter copies are valid for this forward only

    T = emb.shape[1]
    ss = (model.scale_shift_table[None, None] + emb[:, :, None]).reshape(B * T, 2 * D)
    hn, _ = ops.norm_mod(h.reshape(B * N, D), ss[:, D:], ss[:, :D], N // T, 1e-6, True, with_res=False)
    out = apply_linear(model.proj_out, hn).view(B, N, -1)
    if not return_dict:
        return (out,)
    return Transformer3DModelOutput(sample=out)


# ---------------------------------------------------------------------------------------------
# module mirrors
# ---------------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    """diffusers RMSNorm parameter container; the arithmetic runs inside the fused kernels."""

    def __init__(self, dim, eps, elementwise_affine=True):
        super().__init__()
        self.eps = eps
        self.dim = dim
        self.weight = nn.Parameter(torch.ones(dim)) if elementwise_affine else None

    def forward(self, x):
        shp = x.shape
        x2d = x.reshape(-1, shp[-1])
        if self.weight is None:
            return ops.norm_mod_fwd(x2d, None, None, max(x2d.shape[0], 1), self.eps, False).view(shp)
        out = torch.empty_like(x2d)
        ops.qknorm_rope_fwd(x2d, None, self.weight, None, None, None, out, None, self.eps)
        return out.view(shp)


class B200AttnProcessor:
    """Drop-in for AttnProcessor2_0 (attention.py:935-955): install with Attention.set_processor."""

    def __call__(self, attn, hidden_states, freqs_cis=None, encoder_hidden_states=None, attention_mask=None,
                 temb=None, skip_layer_mask=None, skip_layer_strategy=None, *args, **kwargs):
        return attention_forward(attn, hidden_states, freqs_cis=freqs_cis,
                                 encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                                 skip_layer_mask=skip_layer_mask, skip_layer_strategy=skip_layer_strategy)


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 out_bias=True, qk_norm=None, use_rope=False, processor=None, **unused):
        super().__init__()
        if dim_head != 64:
            raise B200Error("only dim_head 64 is built (LTXV-2B: 32 heads x 64)")
        if qk_norm != "rms_norm":
            raise B200Error("only qk_norm='rms_norm' is built (LTXV-2B)")
        self.inner_dim = dim_head * heads
        self.query_dim = query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.use_rope = use_rope
        self.use_tpu_flash_attention = False
        self.q_norm = RMSNorm(self.inner_dim, eps=1e-5)
        self.k_norm = RMSNorm(self.inner_dim, eps=1e-5)
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else B200AttnProcessor())

    def set_processor(self, processor) -> None:
        self.processor = processor

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, freqs_cis=None, encoder_hidden_states=None, attention_mask=None,
                skip_layer_mask=None, skip_layer_strategy=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, freqs_cis=freqs_cis, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, skip_layer_mask=skip_layer_mask,
                              skip_layer_strategy=skip_layer_strategy)


class GELU(nn.Module):
    """diffusers GELU(approximate='tanh') parameter container (`ff.net.0.proj.*`)."""

    def __init__(self, dim_in, dim_out, approximate="tanh", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="gelu-approximate", inner_dim=None,
                 bias=True, **unused):
        super().__init__()
        if activation_fn != "gelu-approximate":
            raise B200Error("only activation_fn='gelu-approximate' is built (LTXV-2B)")
        inner_dim = inner_dim or int(dim * mult)
        self.net = nn.ModuleList([GELU(dim, inner_dim, bias=bias), nn.Dropout(dropout),
                                  nn.Linear(inner_dim, dim_out or dim, bias=bias)])

    def forward(self, hidden_states):
        shp = hidden_states.shape
        W1, b1, _ = linear_parts(self.net[0].proj)
        W2, b2, _ = linear_parts(self.net[2])
        y = ops.FeedForwardFn.apply(hidden_states.reshape(-1, shp[-1]), W1, b1, W2, b2, None, 0, None)
        return y.view(*shp[:-1], -1)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="gelu-approximate", attention_bias=False, only_cross_attention=False,
                 double_self_attention=False, norm_elementwise_affine=False, adaptive_norm="single_scale_shift",
                 standardization_norm="rms_norm", norm_eps=1e-6, qk_norm="rms_norm", use_rope=True, **unused):
        super().__init__()
        if standardization_norm != "rms_norm" or norm_elementwise_affine or adaptive_norm != "single_scale_shift":
            raise B200Error("only the LTXV-2B block flavour (rms_norm, no affine, single_scale_shift) is built")
        self.only_cross_attention = only_cross_attention
        self.adaptive_norm = adaptive_norm
        self.norm1 = RMSNorm(dim, eps=norm_eps, elementwise_affine=False)
        self.attn1 = Attention(dim, cross_attention_dim if only_cross_attention else None, num_attention_heads,
                               attention_head_dim, dropout, attention_bias, True, qk_norm, use_rope)
        self.attn2 = None
        if cross_attention_dim is not None or double_self_attention:
            self.attn2 = Attention(dim, cross_attention_dim if not double_self_attention else None,
                                   num_attention_heads, attention_head_dim, dropout, attention_bias, True, qk_norm,
                                   use_rope)
        self.norm2 = RMSNorm(dim, eps=norm_eps, elementwise_affine=False)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.scale_shift_table = nn.Parameter(torch.randn(6, dim) / dim ** 0.5)

    forward = block_forward


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class _CombinedTimestepEmbeddings(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim)


class AdaLayerNormSingle(nn.Module):
    def __init__(self, dim, use_additional_conditions=False):
        super().__init__()
        self.emb = _CombinedTimestepEmbeddings(dim)
        self.linear = nn.Linear(dim, 6 * dim, bias=True)

    def forward(self, timestep, added_cond_kwargs=None, batch_size=None, hidden_dtype=None):
        return adaln_single_forward(self, timestep)


class PixArtAlphaTextProjection(nn.Module):
    def __init__(self, in_features, hidden_size, out_features=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size, bias=True)
        self.linear_2 = nn.Linear(hidden_size, out_features or hidden_size, bias=True)

    def forward(self, caption):
        shp = caption.shape
        y = ops.FeedForwardFn.apply(caption.reshape(-1, shp[-1]), self.linear_1.weight, self.linear_1.bias,
                                    self.linear_2.weight, self.linear_2.bias, None, 0, None)
        return y.view(*shp[:-1], -1)


class SymmetricPatchifier:
    """symmetric_patchifier.py:33-84 (patch size 1): pure layout, no arithmetic."""

    def __init__(self, patch_size: int = 1):
        if patch_size != 1:
            raise B200Error("only patch_size 1 is built (LTXV)")
        self._patch_size = (1, patch_size, patch_size)

    @property
    def patch_size(self):
        return self._patch_size

    def get_latent_coords(self, latent_num_frames, latent_height, latent_width, batch_size, device):
        g = torch.meshgrid(torch.arange(latent_num_frames, device=device), torch.arange(latent_height, device=device),
                           torch.arange(latent_width, device=device), indexing="ij")
        return torch.stack(g, dim=0).reshape(3, -1).unsqueeze(0).repeat(batch_size, 1, 1)

    def patchify(self, latents):
        b, c, f, h, w = latents.shape
        return latents.flatten(2).transpose(1, 2), self.get_latent_coords(f, h, w, b, latents.device)

    def unpatchify(self, latents, output_height, output_width, out_channels):
        b, n, c = latents.shape
        f = n // (output_height * output_width)
        return latents.reshape(b, f, output_height, output_width, c).permute(0, 4, 1, 2, 3)


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class Transformer3DModel(nn.Module):
    """Same constructor keywords / parameter names as transformer3d.py:49-180 for the LTXV-2B family."""
    _supports_gradient_checkpointing = True

    def __init__(self, num_attention_heads=32, attention_head_dim=64, in_channels=128, out_channels=None,
                 num_layers=28, dropout=0.0, cross_attention_dim=2048, attention_bias=True,
                 activation_fn="gelu-approximate", adaptive_norm="single_scale_shift",
                 standardization_norm="rms_norm", norm_elementwise_affine=False, norm_eps=1e-6,
                 caption_channels=4096, qk_norm="rms_norm", positional_embedding_type="rope",
                 positional_embedding_theta=10000.0, positional_embedding_max_pos=None,
                 timestep_scale_multiplier=1000, causal_temporal_positioning=False, patchifier=None,
                 only_cross_attention=False, double_self_attention=False, **unused):
        super().__init__()
        if positional_embedding_type != "rope":
            raise ValueError("Absolute positional embedding is no longer supported")
        if positional_embedding_theta is None or positional_embedding_max_pos is None:
            raise ValueError("rope needs positional_embedding_theta and positional_embedding_max_pos")
        self.config = _Config(num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                              in_channels=in_channels, out_channels=out_channels or in_channels,
                              num_layers=num_layers, cross_attention_dim=cross_attention_dim,
                              caption_channels=caption_channels,
                              positional_embedding_theta=positional_embedding_theta,
                              positional_embedding_max_pos=positional_embedding_max_pos,
                              timestep_scale_multiplier=timestep_scale_multiplier,
                              causal_temporal_positioning=causal_temporal_positioning)
        self.use_tpu_flash_attention = False
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        self.inner_dim = inner_dim = num_attention_heads * attention_head_dim
        self.patchify_proj = nn.Linear(in_channels, inner_dim, bias=True)
        self.positional_embedding_type = positional_embedding_type
        self.positional_embedding_theta = positional_embedding_theta
        self.positional_embedding_max_pos = positional_embedding_max_pos
        self.use_rope = True
        self.timestep_scale_multiplier = timestep_scale_multiplier
        self.patchifier = patchifier
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  attention_bias=attention_bias, only_cross_attention=only_cross_attention,
                                  double_self_attention=double_self_attention,
                                  norm_elementwise_affine=norm_elementwise_affine, adaptive_norm=adaptive_norm,
                                  standardization_norm=standardization_norm, norm_eps=norm_eps, qk_norm=qk_norm,
                                  use_rope=True) for _ in range(num_layers)])
        self.out_channels = out_channels or in_channels
        self.norm_out = nn.LayerNorm(inner_dim, elementwise_affine=False, eps=1e-6)
        self.scale_shift_table = nn.Parameter(torch.randn(2, inner_dim) / inner_dim ** 0.5)
        self.proj_out = nn.Linear(inner_dim, self.out_channels)
        self.adaln_single = AdaLayerNormSingle(inner_dim)
        self.caption_projection = None
        if caption_channels is not None:
            self.caption_projection = PixArtAlphaTextProjection(caption_channels, inner_dim)
        self.gradient_checkpointing = False

    @classmethod
    def from_config(cls, config: Dict[str, Any], **kwargs):
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        cfg.update(kwargs)
        return cls(**cfg)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def create_skip_layer_mask(self, batch_size, num_conds, ptb_index, skip_block_list=None):
        if skip_block_list is None or len(skip_block_list) == 0:
            return None
        mask = torch.ones((len(self.transformer_blocks), batch_size * num_conds), device=self.device, dtype=self.dtype)
        for block_idx in skip_block_list:
            mask[block_idx, ptb_index::num_conds] = 0
        # host-side note of which rows differ from all-ones: the other blocks keep the fused attn1 path
        # (o * 1 + other * 0 == o exactly) without reading the mask back from the device
        mask._b200_skip_blocks = frozenset(int(b) % len(self.transformer_blocks) for b in skip_block_list)
        return mask

    def precompute_freqs_cis(self, indices_grid, spacing="exp"):
        if spacing != "exp":
            raise B200Error("only spacing='exp' is built")
        return rope_table(indices_grid, self.inner_dim, self.positional_embedding_theta,
                          self.positional_embedding_max_pos)

    def load_state_dict(self, state_dict, *args, **kwargs):
        if any(k.startswith("model.diffusion_model.") for k in state_dict):
            state_dict = {k.replace("model.diffusion_model.", ""): v for k, v in state_dict.items()
                          if k.startswith("model.diffusion_model.")}
        return super().load_state_dict(state_dict, *args, **kwargs)

    forward = transformer_forward
