"""Helpers shared by the distributed meta-estimators.

Counterparts of /root/reference/skdist/distribute/base.py: ``_clone`` (base.py:8-50: sklearn's
``clone`` keeping a ``sc`` attribute by reference) and ``_parse_partitions`` (base.py:53-64).  ``sc`` (a SparkContext in
the reference) is accepted everywhere for drop-in compatibility and ignored:
the "cluster" here is the set of B200s reached through ``skdist_b200.engine``.
"""
from sklearn.base import clone as _sk_clone


def _clone(estimator, safe=True):
    """scikit-learn's ``clone`` plus the one thing the reference adds to it (ref base.py:8-50): a
    ``sc`` attribute is carried over by reference instead of being deep-copied.  ``sc`` is an
    ignored placeholder here, so it is simply re-attached after the clone."""
    if hasattr(estimator, "get_params") and not isinstance(estimator, type) and hasattr(estimator, "sc"):
        sc = estimator.sc
        estimator.sc = None              # never deep-copy a context object
        try:
            new = _sk_clone(estimator, safe=safe)
        finally:
            estimator.sc = sc
        new.sc = sc
        return new
    return _sk_clone(estimator, safe=safe)


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
                self.__dict__.pop("sc", None)


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

