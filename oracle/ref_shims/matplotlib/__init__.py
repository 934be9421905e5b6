"""Inert stand-in for matplotlib (plotting only). Test infrastructure."""
import sys
import types


class _Anything(types.ModuleType):

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return object


for _n in ("pyplot", "cm", "colors", "figure"):
    _m = _Anything("matplotlib." + _n)
    sys.modules["matplotlib." + _n] = _m
    globals()[_n] = _m


def use(*a, **k):
    pass
