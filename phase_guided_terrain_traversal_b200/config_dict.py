"""Minimal stand-in for `ml_collections.config_dict` (not installed in this image).

The reference builds its env config with `config_dict.create(...)` and reads it with attribute
access, item access, `.to_dict()` and in-place mutation (`go2/configs.py:6-79`,
`training/train.py:119-129,178-179`). This class supports exactly that surface.
"""
from __future__ import annotations

import copy


class ConfigDict:
    def __init__(self, initial=None, **kwargs):
        object.__setattr__(self, "_fields", {})
        for k, v in dict(initial or {}, **kwargs).items():
            self[k] = v

    # attribute + item access -------------------------------------------------------------
    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_fields")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __getitem__(self, name):
        return self._fields[name]

    def __setitem__(self, name, value):
        if isinstance(value, dict):
            value = ConfigDict(value)
        self._fields[name] = value

    def __contains__(self, name):
        return name in self._fields

    def __iter__(self):
        return iter(self._fields)

    def keys(self):
        return self._fields.keys()

    def items(self):
        return self._fields.items()

    def values(self):
        return self._fields.values()

    def get(self, name, default=None):
        return self._fields.get(name, default)

    def update(self, other=None, **kwargs):
        for k, v in dict(other or {}, **kwargs).items():
            if isinstance(v, (dict, ConfigDict)) and isinstance(self._fields.get(k), ConfigDict):
                self._fields[k].update(v if isinstance(v, dict) else v.to_dict())
            else:
                self[k] = v

    def update_from_flattened_dict(self, flat):
        for k, v in flat.items():
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else copy.deepcopy(v)) for k, v in self._fields.items()}

    def copy_and_resolve_references(self):
        return ConfigDict(self.to_dict())

    def __deepcopy__(self, memo):
        return ConfigDict(self.to_dict())

    def __repr__(self):
        return f"ConfigDict({self.to_dict()!r})"

    def __eq__(self, other):
        return isinstance(other, ConfigDict) and self.to_dict() == other.to_dict()


def create(**kwargs) -> ConfigDict:
    return ConfigDict(kwargs)
