"""Small helpers of the registry (the roles of nbkode/util.py:14-86)."""

from __future__ import annotations

from collections.abc import Mapping, MutableMapping


class classproperty:
    """Read-only property evaluated on the class (nbkode/util.py:14-19)."""

    def __init__(self, fget):
        self.fget = fget

    def __get__(self, obj, owner):
        return self.fget(owner)


class CaseInsensitiveDict(MutableMapping):
    """String-keyed mapping for the solver registry: ``get_solver("forwardeuler")`` and ``get_solvers("euler")``
    must find ``ForwardEuler`` / the ``Euler`` group, while listing the registry shows the names as registered.
    Two plain dicts share the case-folded key: one holds the values, one the spelling last used to store them."""

    __slots__ = ("_values", "_spelling")

    def __init__(self, data=None, **kwargs):
        self._values, self._spelling = {}, {}
        if data is not None:
            self.update(data)
        if kwargs:
            self.update(kwargs)

    def __setitem__(self, key, value):
        folded = key.casefold()
        self._values[folded] = value
        self._spelling[folded] = key

    def __getitem__(self, key):
        return self._values[key.casefold()]

    def __delitem__(self, key):
        folded = key.casefold()
        del self._values[folded]
        del self._spelling[folded]

    def __contains__(self, key):
        return isinstance(key, str) and key.casefold() in self._values

    def __iter__(self):
        return iter(self._spelling.values())

    def __len__(self):
        return len(self._values)

    def lower_items(self):
        """(case-folded key, value) pairs."""
        return self._values.items()

    def __eq__(self, other):
        if not isinstance(other, Mapping):
            return NotImplemented
        if not isinstance(other, CaseInsensitiveDict):
            other = CaseInsensitiveDict(other)
        return self._values == other._values

    def copy(self):
        return CaseInsensitiveDict(self)

    def __repr__(self):
        return "{" + ", ".join(f"{k!r}: {self[k]!r}" for k in self) + "}"
