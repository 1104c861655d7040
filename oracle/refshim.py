"""Import the UNMODIFIED reference (``/root/reference/skdist``) under sklearn 1.9.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Works only where /root/reference
exists (the build container); used by tests/golden/make_golden.py to produce
the committed fixtures and by tests that are skipped when it is absent.

The reference targets sklearn<0.23.2.  Three in-memory patches make its
hot-path modules importable without touching any file:
  * ``np.int`` alias (search.py:510, multiclass.py:159),
  * ``sklearn.utils.metaestimators.if_delegate_has_method`` (search.py:23),
  * ``sklearn.ensemble.forest`` module alias (ensemble.py:11).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("SKDIST_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "skdist", "distribute"))


def load():
    """Return the reference's (search, multiclass, ensemble) modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    import sklearn.utils.metaestimators as _m
    from sklearn.utils.metaestimators import available_if

    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(_m, "if_delegate_has_method"):
        def if_delegate_has_method(delegate):
            names = tuple(delegate) if isinstance(delegate, (list, tuple)) else (delegate,)
            return available_if(lambda self: any(hasattr(self, n) for n in names))
        _m.if_delegate_has_method = if_delegate_has_method
    if "sklearn.ensemble.forest" not in sys.modules:
        import sklearn.ensemble._forest as _f
        sys.modules["sklearn.ensemble.forest"] = _f
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own drop-in alias package is also called ``skdist``: make sure the
    # reference's copy is the one imported here, under a private name.
    import importlib.util
    saved = {k: v for k, v in sys.modules.items() if k == "skdist" or k.startswith("skdist.")}
    for k in saved:
        del sys.modules[k]
    try:
        spec = importlib.util.spec_from_file_location(
            "skdist", os.path.join(REFERENCE_ROOT, "skdist", "__init__.py"),
            submodule_search_locations=[os.path.join(REFERENCE_ROOT, "skdist")])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["skdist"] = pkg
        spec.loader.exec_module(pkg)
        import importlib
        search = importlib.import_module("skdist.distribute.search")
        multiclass = importlib.import_module("skdist.distribute.multiclass")
        ensemble = importlib.import_module("skdist.distribute.ensemble")
    finally:
        ref_mods = {k: v for k, v in sys.modules.items() if k == "skdist" or k.startswith("skdist.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    return search, multiclass, ensemble
