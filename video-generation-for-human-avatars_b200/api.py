"""Public entry points: build the mirror model, or install the b200 kernels behind an existing
instance of the reference's own Transformer3DModel (with or without peft LoRA wrappers)."""
import types

import torch

from . import lib, modules
from .lora import apply_training_strategy  # noqa: F401
from .modules import (B200AttnProcessor, BasicTransformerBlock, SymmetricPatchifier,  # noqa: F401
                      Transformer3DModel)
from .scheduler import RectifiedFlowScheduler  # noqa: F401
from .train import train_step  # noqa: F401
from .sampling import Denoiser, denoise  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .io import DeviceFeeder, LatentTripleDataset, collate_latent_triples, shard_indices  # noqa: F401

LTXV_2B_CONFIG = dict(num_attention_heads=32, attention_head_dim=64, in_channels=128, out_channels=128,
                      num_layers=28, cross_attention_dim=2048, attention_bias=True,
                      activation_fn="gelu-approximate", caption_channels=4096, qk_norm="rms_norm",
                      standardization_norm="rms_norm", norm_elementwise_affine=False, norm_eps=1e-6,
                      positional_embedding_type="rope", positional_embedding_theta=10000.0,
                      positional_embedding_max_pos=[20, 2048, 2048], timestep_scale_multiplier=1000)


def install(model):
    """Make a reference `Transformer3DModel` (or a peft-wrapped one) run on the b200 kernels.

    (a) every `Attention` gets `B200AttnProcessor` through the reference's own `set_processor`
        (attention.py:532-552); (b) `BasicTransformerBlock.forward` and `Transformer3DModel.forward`
        are rebound, per instance, to the fused versions with identical signatures.  No parameter
        or module is renamed, so peft targeting, state_dict, deepcopy and merge_and_unload keep working."""
    lib.require_device()
    root = model
    inner = getattr(getattr(model, "base_model", None), "model", None)
    if inner is not None:
        root = inner
    for blk in root.transformer_blocks:
        for attn in (blk.attn1, blk.attn2):
            if attn is not None:
                attn.set_processor(B200AttnProcessor())
        blk.forward = types.MethodType(modules.block_forward, blk)
    root.forward = types.MethodType(modules.transformer_forward, root)
    modules.drop_weight_caches(root)
    return model


def uninstall(model):
    root = getattr(getattr(model, "base_model", None), "model", None) or model
    for blk in root.transformer_blocks:
        blk.__dict__.pop("forward", None)
    root.__dict__.pop("forward", None)
    modules.drop_weight_caches(root)
    return model


def enable_sequence_parallel(model, group=None, impl=None, mode="ring", disable=False):
    """Shard the latent tokens of every forward over the ranks of `group` (BASELINE config 5): the model
    then takes this rank's contiguous shard of `hidden_states` / `indices_grid` and attn1 runs as a
    K/V ring (ring.py; `mode="gather"`: one all-gather of K/V and a single attention launch per layer instead).  Trainable gradients come out as per-shard partial means: average them over the
    group (dp.GradBucketer does exactly that).  `group=None` with no default group, or `disable=True`, switches it off again."""
    from .ring import SequenceParallel
    root = getattr(getattr(model, "base_model", None), "model", None) or model
    sp = None
    if not disable and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        sp = SequenceParallel(group, impl, mode)
    modules.side(root)["sp"] = sp
    for blk in root.transformer_blocks:
        modules.side(blk.attn1)["sp"] = sp
    return sp


def build_model(config=None, device="cuda", dtype=torch.bfloat16, patchifier=None):
    cfg = dict(LTXV_2B_CONFIG)
    cfg.update(config or {})
    with torch.device(device):
        m = Transformer3DModel.from_config(cfg)
    m = m.to(dtype)
    m.patchifier = patchifier or SymmetricPatchifier(1)
    return m
