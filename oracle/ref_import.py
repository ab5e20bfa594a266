"""Tier-one oracle: import the reference's OWN, unmodified hot-path modules.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference is mounted (the build container);
the GPU box uses the portable restatement oracle/ref_block.py plus the golden vectors that
oracle/make_golden.py generated from this tier.  Imports, through the stand-ins in
oracle/diffusers_shim and oracle/peft_shim:
  ltx_video/models/transformers/transformer3d.py   (Transformer3DModel)
  ltx_video/models/transformers/attention.py       (BasicTransformerBlock, Attention, AttnProcessor2_0)
  ltx_video/models/transformers/symmetric_patchifier.py
  ltx_video/schedulers/rf.py                       (RectifiedFlowScheduler)
  ltx_video/training.py                            (train_step, apply_training_strategy)
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("B200LTX_REFERENCE_ROOT", "/root/reference")
SHIMMED = {}   # package -> True when the stand-in under oracle/ is what the reference imported


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ltx_video"))


def _prepare():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # A stand-in is used only for a package that is NOT installed: wherever the real diffusers / peft exist, the
    # reference imports them and every test of this tier pins the third-party leaves too (B200LTX_FORCE_SHIMS=1
    # keeps the stand-ins regardless).
    import importlib.util
    for pkg in ("peft", "diffusers"):
        p = os.path.join(HERE, pkg + "_shim")
        if p in sys.path:
            continue
        real = os.environ.get("B200LTX_FORCE_SHIMS") != "1" and importlib.util.find_spec(pkg) is not None
        SHIMMED[pkg] = not real
        if not real:
            sys.path.insert(0, p)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    os.environ.setdefault("WANDB_MODE", "disabled")
    # ltx_video.validation pulls imageio/lpips and the full pipeline; train_step never calls it.
    if "ltx_video.validation" not in sys.modules:
        stub = types.ModuleType("ltx_video.validation")
        stub.validate_epoch = lambda *a, **k: None
        stub.validate_video = lambda *a, **k: None
        sys.modules["ltx_video.validation"] = stub


def load():
    """Returns a namespace with the reference classes/functions on the hot path."""
    _prepare()
    from ltx_video.models.transformers import attention as ref_attention
    from ltx_video.models.transformers import transformer3d as ref_t3d
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.utils.diffusers_config_mapping import OURS_TRANSFORMER_CONFIG
    from ltx_video.utils.skip_layer_strategy import SkipLayerStrategy

    ns = types.SimpleNamespace(
        attention=ref_attention, transformer3d=ref_t3d,
        Transformer3DModel=ref_t3d.Transformer3DModel,
        BasicTransformerBlock=ref_attention.BasicTransformerBlock,
        Attention=ref_attention.Attention, AttnProcessor2_0=ref_attention.AttnProcessor2_0,
        SymmetricPatchifier=SymmetricPatchifier, RectifiedFlowScheduler=RectifiedFlowScheduler,
        OURS_TRANSFORMER_CONFIG=dict(OURS_TRANSFORMER_CONFIG), SkipLayerStrategy=SkipLayerStrategy)
    return ns


def load_training():
    """ltx_video.training (needs transformers' T5 classes importable; wandb disabled)."""
    _prepare()
    import ltx_video.training as tr
    return tr
