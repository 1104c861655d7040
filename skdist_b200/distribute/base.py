"""Helpers shared by the distributed meta-estimators.

Mirrors /root/reference/skdist/distribute/base.py: ``_clone`` (base.py:8-50, a
sklearn ``clone`` that keeps a ``sc`` attribute by reference), ``_parse_partitions``
(base.py:53-64) and ``_get_value`` (base.py:67-72).  ``sc`` (a SparkContext in
the reference) is accepted everywhere for drop-in compatibility and ignored:
the "cluster" here is the set of B200s reached through ``skdist_b200.engine``.
"""
import copy


def _clone(estimator, safe=True):
    """Construct a new unfitted estimator with the same parameters; ``sc`` is
    carried over by reference instead of being deep-copied (ref base.py:8-50)."""
    found_sc = hasattr(estimator, "sc")
    estimator_type = type(estimator)
    if estimator_type in (list, tuple, set, frozenset):
        return estimator_type([_clone(e, safe=safe) for e in estimator])
    elif not hasattr(estimator, "get_params") or isinstance(estimator, type):
        if not safe:
            return copy.deepcopy(estimator)
        raise TypeError(
            "Cannot clone object '%s' (type %s): it does not seem to be a scikit-learn "
            "estimator as it does not implement a 'get_params' methods."
            % (repr(estimator), type(estimator)))
    klass = estimator.__class__
    params = estimator.get_params(deep=False)
    for name, param in params.items():
        if name != "sc":
            params[name] = _clone(param, safe=False)
    new_object = klass(**params)
    params_set = new_object.get_params(deep=False)
    for name in params:
        if params[name] is not params_set[name]:
            raise RuntimeError(
                "Cannot clone object %s, as the constructor either does not set or "
                "modifies parameter %s" % (estimator, name))
    if found_sc:
        new_object.sc = estimator.sc
    return new_object


class _Cloner:
    """Repeated `_clone` of one template (a thousand label estimators in one-vs-rest): the template is
    cloned and validated once; further clones rebuild from the captured constructor arguments, which
    skips the two `get_params` signature inspections per clone (0.3 ms each)."""

    def __init__(self, estimator):
        first = _clone(estimator)              # validates like `_clone` does (raises the same errors)
        self._klass = first.__class__
        self._params = first.get_params(deep=False)
        self._sc = (estimator.sc,) if hasattr(estimator, "sc") else None
        self._first = first

    def __call__(self):
        if self._first is not None:
            new, self._first = self._first, None
            return new
        params = {k: (v if k == "sc" else _clone(v, safe=False)) for k, v in self._params.items()}
        new = self._klass(**params)
        if self._sc is not None:
            new.sc = self._sc[0]
        return new


def _parse_partitions(partitions, auto_n):
    """ref base.py:53-64.  Kept for signature compatibility; partitions do not
    affect the device engine (columns are sharded round-robin over GPUs)."""
    if partitions is None:
        return None
    if partitions == "auto":
        return auto_n
    try:
        return int(partitions)
    except Exception:
        return None


def _get_value(obj):
    """ref base.py:67-72: unwrap a broadcast variable."""
    return obj.value if hasattr(obj, "value") else obj


class _ScParamMixin:
    """The reference deletes ``self.sc`` at the end of ``fit`` (search.py:568,
    multiclass.py:283, ensemble.py:335) so that the fitted object pickles.  Under
    sklearn >= 1.0 ``get_params`` would then raise; report ``sc=None`` instead."""

    def get_params(self, deep=True):
        missing = not hasattr(self, "sc")
        if missing:
            self.sc = None
        try:
            return super().get_params(deep=deep)
        finally:
            if missing:
                del self.sc


def _merged_params(estimator, candidate_params):
    """Per-candidate parameter dicts (base params overridden by the candidate) without cloning the
    estimator once per candidate; unknown names raise like ``set_params`` does."""
    base = estimator.get_params(deep=False)
    out = []
    for p in candidate_params:
        for k in p:
            if k not in base:
                raise ValueError("Invalid parameter %r for estimator %s. Valid parameters are: %r."
                                 % (k, estimator, sorted(base)))
        q = dict(base)
        q.update(p)
        out.append(q)
    return out

