"""Environment registry: `make("fishing-v1", **ctor_kwargs)` like gym.make on the reference ids
(gym_fishing/envs/__init__.py:17-35).  No wrappers are applied (the reference registers no max_episode_steps).
When a real gym/gymnasium is importable the ids are also registered there under the same names."""
from __future__ import annotations

import importlib

_REGISTRY = {}


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = (entry_point, dict(kwargs))
    for modname in ("gymnasium", "gym"):
        try:  # pragma: no cover - not installed in the build image
            mod = importlib.import_module(modname + ".envs.registration")
            if id not in getattr(mod, "registry", {}):
                mod.register(id=id, entry_point=entry_point, **kwargs)
        except Exception:  # noqa: BLE001
            pass


def spec_ids():
    return sorted(_REGISTRY)


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError(f"unknown environment id {id!r}; available: {spec_ids()} "
                       "(the reference has no fishing-v3)")
    entry_point, base = _REGISTRY[id]
    if callable(entry_point):
        cls = entry_point
    else:
        mod_name, cls_name = entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(base.get("kwargs", {}))
    kw.update(kwargs)
    return cls(**kw)
