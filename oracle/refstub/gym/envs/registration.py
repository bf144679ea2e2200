import importlib

_REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, base_kwargs = _REGISTRY[id]
    mod_name, cls_name = entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(base_kwargs.get("kwargs", {}))
    kw.update(kwargs)
    return cls(**kw)
