"""Inert stand-in for bokeh (plotting only). Test infrastructure."""
import sys
import types


class _Anything(types.ModuleType):

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return object


for _n in ("io", "models", "palettes", "plotting", "resources", "layouts", "transform"):
    _m = _Anything("bokeh." + _n)
    sys.modules["bokeh." + _n] = _m
    globals()[_n] = _m
