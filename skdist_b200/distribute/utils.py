"""Scoring helpers of the distribute module (host side).

Mirrors the parts of /root/reference/skdist/distribute/utils.py that the hot path uses:
``_check_multimetric_scoring`` (utils.py:75-143), ``_num_samples`` (utils.py:146-168),
``_aggregate_score_dicts`` (utils.py:13-15).  ``_safe_split`` (utils.py:171-209) has no
counterpart: train/test subsets are never materialised, rows are masked by fold id on the
device.
"""
import numbers

import numpy as np
from sklearn.metrics import check_scoring


def _aggregate_score_dicts(scores):
    """Aggregate a list of dicts into a dict of arrays (ref utils.py:13-15)."""
    return {key: np.asarray([score[key] for score in scores]) for key in scores[0]}


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


def _num_samples(x):
    """Number of samples in array-like x (ref utils.py:146-168)."""
    if hasattr(x, "fit") and callable(x.fit):
        raise TypeError("Expected sequence or array-like, got estimator %s" % x)
    if not hasattr(x, "__len__") and not hasattr(x, "shape"):
        if hasattr(x, "__array__"):
            x = np.asarray(x)
        else:
            raise TypeError("Expected sequence or array-like, got %s" % type(x))
    if hasattr(x, "shape"):
        if len(x.shape) == 0:
            raise TypeError("Singleton array %r cannot be considered a valid collection." % x)
        if isinstance(x.shape[0], numbers.Integral):
            return x.shape[0]
        return len(x)
    return len(x)
