"""Minimal `_target_` resolver for the reference's configs (what `hydra.utils.instantiate(cfg, _recursive_=False)` does
for the classes on this path: DNeRF.py:21-28, train.py:27-28).  Hydra / OmegaConf themselves (defaults lists, `${}`
interpolation, run directories) are control plane and out of scope; a caller passes plain dicts (e.g. `yaml.safe_load`)."""
from __future__ import annotations

import importlib


class Cfg(dict):
    """dict with attribute access and `.get`, enough of DictConfig for the mirror classes (`opt.center`, `opt.get(k)`)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @staticmethod
    def wrap(v):
        if isinstance(v, dict) and not isinstance(v, Cfg):
            return Cfg({k: Cfg.wrap(x) for k, x in v.items()})
        if isinstance(v, (list, tuple)):
            return type(v)(Cfg.wrap(x) for x in v)
        return v


def resolve(target: str):
    """'pkg.mod.Class' -> the class object"""
    mod, _, name = target.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(cfg, **overrides):
    """cfg = {"_target_": "a.b.C", **kwargs} -> C(**kwargs, **overrides); nested configs are passed through
    un-instantiated (`_recursive_=False`, as every call site of the reference does)."""
    cfg = Cfg.wrap(dict(cfg))
    cls = resolve(cfg["_target_"])
    kwargs = {k: v for k, v in cfg.items() if not k.startswith("_")}
    kwargs.update(overrides)
    return cls(**kwargs)
