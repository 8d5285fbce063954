import importlib

_REGISTRY = {}


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, kw = _REGISTRY[id]
    mod_name, cls_name = entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    return cls(**{**kw, **kwargs})
