"""ConfigMixin / register_to_config stand-ins (diffusers/configuration_utils.py)."""
import functools
import inspect


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def __init__(self, *args, **kwargs):
        super().__init__()

    @property
    def config(self):
        return self.__dict__.get("_internal_dict", _AttrDict())

    def register_to_config(self, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        d = _AttrDict(self.__dict__.get("_internal_dict", {}))
        d.update(kwargs)
        object.__setattr__(self, "_internal_dict", d)

    @classmethod
    def from_config(cls, config, **kwargs):
        params = inspect.signature(cls.__init__).parameters
        cfg = {k: v for k, v in dict(config).items() if k in params and k != "self"}
        cfg.update({k: v for k, v in kwargs.items() if k in params})
        return cls(**cfg)


def register_to_config(init):
    """Decorator: record every __init__ argument (defaults included) in `.config`."""

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        names = [n for n in sig.parameters if n != "self"]
        rec = {n: p.default for n, p in sig.parameters.items()
               if n != "self" and p.default is not inspect.Parameter.empty}
        for n, a in zip(names, args):
            rec[n] = a
        rec.update(kwargs)
        init(self, *args, **{k: v for k, v in kwargs.items() if not k.startswith("_")})
        ConfigMixin.register_to_config(self, **rec)

    return wrapper
