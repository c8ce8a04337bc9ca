"""ConfigMixin / register_to_config: record ctor kwargs (with defaults) on `self.config`."""
import functools
import inspect


class FrozenConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kw):
        cfg = FrozenConfig(getattr(self, "_internal_dict", {}))
        cfg.update(kw)
        object.__setattr__(self, "_internal_dict", cfg)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__)
        kw = {k: v for k, v in dict(config).items() if k in sig.parameters}
        kw.update({k: v for k, v in kwargs.items() if k in sig.parameters})
        return cls(**kw)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = {n: p.default for n, p in list(sig.parameters.items())[1:] if p.default is not inspect._empty}
        names = list(sig.parameters.keys())[1:]
        for n, a in zip(names, args):
            bound[n] = a
        bound.update(kwargs)
        ConfigMixin.register_to_config(self, **bound)
        init(self, *args, **kwargs)

    return inner
