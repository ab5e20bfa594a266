"""ModelMixin stand-in: nn.Module + dtype/device + config-attribute fallback."""
import torch
from torch import nn


class ModelMixin(nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.float32

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            cfg = self.__dict__.get("_internal_dict", None)
            if cfg is not None and name in cfg:
                return cfg[name]
            raise
