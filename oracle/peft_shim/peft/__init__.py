"""Stand-in for peft 0.17.1 LoraConfig / get_peft_model (reference call sites:
ltx_video/training.py:9,61-68).  TEST INFRASTRUCTURE ONLY (oracle/).  Restates the published
lora.Linear contract: base_layer + lora_A/lora_B ModuleDicts keyed "default", fp32 adapters
(autocast_adapter_dtype), A kaiming-uniform(a=sqrt 5), B zeros, y = base(x) + B(A(x.float()))*s
cast back to the base dtype.  **Parity unpinned** against the real package (not installable)."""
import math
from dataclasses import dataclass, field
from typing import List

import torch
from torch import nn


@dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: int = 8
    target_modules: List[str] = field(default_factory=list)
    lora_dropout: float = 0.0
    bias: str = "none"


class LoraLinear(nn.Module):
    def __init__(self, base_layer: nn.Linear, r: int, lora_alpha: float):
        super().__init__()
        self.base_layer = base_layer
        self.r = {"default": r}
        self.lora_alpha = {"default": lora_alpha}
        self.scaling = {"default": lora_alpha / r}
        self.lora_dropout = nn.ModuleDict({"default": nn.Identity()})
        dev = base_layer.weight.device
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base_layer.in_features, r, bias=False,
                                                          device=dev, dtype=torch.float32)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base_layer.out_features, bias=False,
                                                          device=dev, dtype=torch.float32)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)
        self.in_features, self.out_features = base_layer.in_features, base_layer.out_features

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x):
        result = self.base_layer(x)
        torch_result_dtype = result.dtype
        a, b = self.lora_A["default"], self.lora_B["default"]
        xx = x.to(a.weight.dtype)
        result = result + b(a(self.lora_dropout["default"](xx))) * self.scaling["default"]
        return result.to(torch_result_dtype)

    def merge(self):
        w = self.base_layer.weight
        delta = self.lora_B["default"].weight @ self.lora_A["default"].weight
        w.data += (delta * self.scaling["default"]).to(w.dtype)


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
            base = getattr(parent, child) if not child.isdigit() else parent[int(child)]
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
        model = self.base_model.model
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
