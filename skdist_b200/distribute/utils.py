"""Scoring helpers of the distribute module (host side).

The one helper of /root/reference/skdist/distribute/utils.py the host path needs:
``_check_multimetric_scoring`` (utils.py:75-143).  ``_safe_split`` (utils.py:171-209) has no
counterpart: train/test subsets are never materialised, rows are masked by fold id on the
device.
"""
from sklearn.metrics import check_scoring


def _check_multimetric_scoring(estimator, scoring=None):
    """Return ({name: scorer}, is_multimetric) (ref utils.py:75-143)."""
    if callable(scoring) or scoring is None or isinstance(scoring, str):
        return {"score": check_scoring(estimator, scoring=scoring)}, False
    if isinstance(scoring, (list, tuple, set)):
        try:
            keys = set(scoring)
        except TypeError:
            raise ValueError("The list/tuple elements must be unique strings of predefined scorers.")
        if len(keys) != len(scoring) or len(keys) == 0 or not all(isinstance(k, str) for k in keys):
            raise ValueError("scoring must be a non-empty list of unique scorer names; got %r" % (scoring,))
        return {s: check_scoring(estimator, scoring=s) for s in scoring}, True
    if isinstance(scoring, dict):
        if len(scoring) == 0 or not all(isinstance(k, str) for k in scoring):
            raise ValueError("scoring dict must be non-empty with string keys; got %r" % (scoring,))
        return {k: check_scoring(estimator, scoring=v) for k, v in scoring.items()}, True
    raise ValueError(
        "scoring should either be a single string or callable for single metric evaluation or a "
        "list/tuple of strings or a dict of scorer name mapped to the callable for multiple "
        "metric evaluation. Got %s of type %s" % (repr(scoring), type(scoring)))
