"""Small helpers of the registry (counterparts of nbkode/util.py:14-86)."""

from __future__ import annotations

from collections.abc import MutableMapping


class classproperty:
    """Read-only property evaluated on the class (nbkode/util.py:14-19)."""

    def __init__(self, fget):
        self.fget = fget

    def __get__(self, obj, owner):
        return self.fget(owner)


class CaseInsensitiveDict(MutableMapping):
    """String-keyed mapping whose lookups ignore case while iteration yields the keys as
    they were last written (the behaviour nbkode/util.py:22-86 relies on for
    ``get_solver("forwardeuler")`` and ``get_solvers("euler")``)."""

    def __init__(self, data=None, **kwargs):
        self._items = {}  # folded key -> (original key, value)
        self.update(data or {}, **kwargs)

    @staticmethod
    def _fold(key):
        return key.lower()

    def __setitem__(self, key, value):
        self._items[self._fold(key)] = (key, value)

    def __getitem__(self, key):
        return self._items[self._fold(key)][1]

    def __delitem__(self, key):
        del self._items[self._fold(key)]

    def __iter__(self):
        return (orig for orig, _ in self._items.values())

    def __len__(self):
        return len(self._items)

    def lower_items(self):
        return ((k, v[1]) for k, v in self._items.items())

    def __eq__(self, other):
        if not isinstance(other, MutableMapping) and not isinstance(other, dict):
            return NotImplemented
        return dict(self.lower_items()) == dict(CaseInsensitiveDict(other).lower_items())

    def copy(self):
        return CaseInsensitiveDict(dict(self.items()))

    def __repr__(self):
        return repr(dict(self.items()))
