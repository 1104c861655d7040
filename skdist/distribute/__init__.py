import sys as _sys

from skdist_b200.distribute import search  # noqa: F401

_sys.modules[__name__ + ".search"] = search
for _name in ("multiclass", "ensemble", "predict", "eliminate"):
    try:
        _mod = __import__("skdist_b200.distribute." + _name, fromlist=[_name])
        _sys.modules[__name__ + "." + _name] = _mod
        globals()[_name] = _mod
    except ImportError:  # module not built yet
        pass
__all__ = ["search", "multiclass", "ensemble", "predict", "eliminate"]
