"""diffusers.utils stand-ins: BaseOutput, logging, deprecate, is_torch_version."""
from collections import OrderedDict
from dataclasses import fields

from . import logging  # noqa: F401


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return tuple(self.values())[k]

    def to_tuple(self):
        return tuple(self.values())


def is_torch_version(op, version):
    return True


def deprecate(*args, **kwargs):
    return None
