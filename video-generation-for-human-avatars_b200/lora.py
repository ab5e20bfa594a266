"""peft-compatible LoRA injection for the attn2 projections.

peft 0.17.1 is not installable in this image, so this module creates the SAME attribute layout as
peft's lora.Linear (`base_layer`, `lora_A["default"]`, `lora_B["default"]`, `scaling["default"]`,
fp32 adapters, A kaiming-uniform(a=sqrt 5), B zeros) and the same wrapper shape as get_peft_model
(`model.base_model.model` is the original module), so parameter names match what
ltx_video/training.py:42-74 (apply_training_strategy) produces and `"lora_" in name` selects them.
When the real peft is present, b200_ltx reads its layers through the same attributes."""
import math
from dataclasses import dataclass, field
from typing import List

import torch
from torch import nn


@dataclass
class LoraConfig:
    r: int = 32
    lora_alpha: int = 32
    target_modules: List[str] = field(default_factory=list)
    lora_dropout: float = 0.0
    bias: str = "none"


class LoraLinear(nn.Module):
    """Parameter container with peft's lora.Linear layout; the math runs in ops.LinearFn."""

    def __init__(self, base_layer: nn.Linear, r: int, lora_alpha: float):
        super().__init__()
        self.base_layer = base_layer
        self.in_features, self.out_features = base_layer.in_features, base_layer.out_features
        self.r = {"default": r}
        self.lora_alpha = {"default": lora_alpha}
        self.scaling = {"default": lora_alpha / r}
        dev = base_layer.weight.device
        self.lora_dropout = nn.ModuleDict({"default": nn.Identity()})
        self.lora_A = nn.ModuleDict({"default": nn.Linear(self.in_features, r, bias=False, device=dev,
                                                          dtype=torch.float32)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, self.out_features, bias=False, device=dev,
                                                          dtype=torch.float32)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)
        self.merged = False

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x):
        from . import modules
        shp = x.shape
        return modules.apply_linear(self, x.reshape(-1, shp[-1])).view(*shp[:-1], -1)

    def merge(self):
        """W += scaling * B A  (peft merge_and_unload; torch_utils.py:66-102 exports merged weights).
        On the device this is one rank-r GEMM accumulating into the bf16 weight in place (residual epilogue):
        a = scaling * B [out, r], b = A [r, in] stored [K, N]; on a CPU copy of the model (checkpoint export of an
        off-loaded model) it is the plain matmul."""
        w = self.base_layer.weight
        A, B, s = self.lora_A["default"].weight, self.lora_B["default"].weight, self.scaling["default"]
        if w.is_cuda and w.dtype == torch.bfloat16:
            from . import ops
            a = (B.detach().float() * s).to(torch.bfloat16)
            b = A.detach().to(torch.bfloat16)
            if a.shape[1] % 8:   # the reduction extent of a GEMM is a multiple of 8: zero-pad the rank
                pad = 8 - a.shape[1] % 8
                a = torch.nn.functional.pad(a, (0, pad))
                b = torch.nn.functional.pad(b, (0, 0, 0, pad))
            ops.gemm(a.contiguous(), b.contiguous(), b_rows_are_k=True, res=w.data, out=w.data)
            # the kernel wrote through the raw pointer: move the version counter so that every cache keyed on it
            # (modules._cached_wqkv / _batched_ctx_kv) sees a new weight
            torch.autograd.graph.increment_version(w)
        else:
            with torch.no_grad():
                w.add_(((B.float() @ A.float()) * s).to(w.dtype))
        self.merged = True

    def unmerge(self):
        """Inverse of merge() (peft `unmerge_adapter`): W -= scaling * B A."""
        if not getattr(self, "merged", False):
            return
        w = self.base_layer.weight
        A, B, s = self.lora_A["default"].weight, self.lora_B["default"].weight, self.scaling["default"]
        with torch.no_grad():
            w.sub_(((B.float() @ A.float()) * s).to(w.dtype))
        self.merged = False


class _LoraModel(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, *a, **k):
        return self.model(*a, **k)


class PeftModel(nn.Module):
    def __init__(self, model, config: LoraConfig):
        super().__init__()
        for name in config.target_modules:
            parent_name, _, child = name.rpartition(".")
            parent = model.get_submodule(parent_name) if parent_name else model
            base = parent[int(child)] if child.isdigit() else getattr(parent, child)
            wrapped = LoraLinear(base, config.r, config.lora_alpha)
            if child.isdigit():
                parent[int(child)] = wrapped
            else:
                setattr(parent, child, wrapped)
        self.base_model = _LoraModel(model)
        self.peft_config = {"default": config}

    def forward(self, *a, **k):
        return self.base_model(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model.model, name)

    def merge_and_unload(self):
        from . import modules
        model = self.base_model.model
        modules.drop_weight_caches(model)
        for parent in model.modules():
            for cname, child in list(parent.named_children()):
                if isinstance(child, LoraLinear):
                    child.merge()
                    if cname.isdigit():
                        parent[int(cname)] = child.base_layer
                    else:
                        setattr(parent, cname, child.base_layer)
        return model


def get_peft_model(model, config: LoraConfig, **kwargs):
    return PeftModel(model, config)


def apply_training_strategy(model, lora_rank: int, lora_alpha: int, train_mode: str = "lora_audio"):
    """training.py:42-91.  "lora_audio": LoRA on attn2.{to_q,to_k,to_v,to_out.0} of every block; LoRA +
    caption_projection trainable, everything else frozen.  Any other mode (the reference's "full"): no adapters;
    proj_out, every scale_shift_table, adaln_single, caption_projection and all attention parameters train (the
    feed-forwards and patchify_proj stay frozen).  In that mode the AdaLN modulate, the gates and the qk-norm affine
    run un-fused so autograd sees them (ops.norm_mod / gate_residual, modules.attention_forward)."""
    if train_mode != "lora_audio":
        # training.py:75-91: no adapters; every parameter whose name contains one of these keys trains
        keys = ("proj_out", "scale_shift_table", "adaln_single", "caption_projection", "attn", "attn2")
        for n, p in model.named_parameters():
            p.requires_grad = any(k in n for k in keys)
        return model
    targets = []
    for i in range(len(model.transformer_blocks)):
        targets += [f"transformer_blocks.{i}.attn2.to_q", f"transformer_blocks.{i}.attn2.to_k",
                    f"transformer_blocks.{i}.attn2.to_v", f"transformer_blocks.{i}.attn2.to_out.0"]
    model = get_peft_model(model, LoraConfig(r=lora_rank, lora_alpha=lora_alpha, target_modules=targets))
    for n, p in model.named_parameters():
        p.requires_grad = ("lora_" in n) or ("caption_projection" in n)
    return model
