"""Distributed hyper-parameter search meta-estimators on B200s.

Drop-in for /root/reference/skdist/distribute/search.py (class names, constructor
signatures incl. positional order, fitted attributes).  The reference fans
``_fit_and_score`` (search.py:180-288) out over Spark executors, one task per
(candidate, fold) (search.py:378-437); here every (candidate, fold) pair is one *column*
of a batched solve on the GPU(s):

  reference                                    this module
  ---------                                    -----------
  cv.split -> (train_idx, test_idx) per task   one int8 fold id per row      (_fold_ids)
  _safe_split copies X[train], X[test]         none: rows are masked by fold id in-kernel
  estimator.fit per task                       Engine.logreg_fit_batch / ridge_cv (all columns)
  scorer(estimator, X_test, y_test)            Engine.linear_score_batch (accuracy counts) / r2
  sc.parallelize(...).map(...).collect()       round-robin column shards over ranks + all_gather
  cv_results_ assembly, best_*, refit          same arithmetic as search.py:461-550 (host)

Base estimators with a device path: ``LogisticRegression`` (binary y, penalty l2, solver
lbfgs) and ``Ridge`` (dense, single target).  Anything else raises NotImplementedError: by
design there is no CPU fallback (the reference's joblib branch, search.py:388-409, is what
the CPU baseline in bench.py times).
"""
import numbers
import time
from collections import defaultdict
from functools import partial

import numpy as np
from numpy.ma import MaskedArray
from scipy.stats import rankdata
from sklearn.base import BaseEstimator, is_classifier
from sklearn.linear_model import LogisticRegression, Ridge
from sklearn.model_selection import (GridSearchCV, ParameterGrid, ParameterSampler,
                                     RandomizedSearchCV, check_cv)
from sklearn.utils.validation import indexable

from .. import parallel
from ..engine import get_engine
from .base import _clone, _merged_params, _parse_partitions, _ScParamMixin
from .utils import _check_multimetric_scoring, _num_samples
from .validation import _check_estimator, _check_n_iter, _validate_models

__all__ = ["DistGridSearchCV", "DistRandomizedSearchCV", "DistMultiModelSearch"]


# ----------------------------------------------------------------------------------------
# cross-validation layout
# ----------------------------------------------------------------------------------------
def _fold_ids(cv_splitted, n_samples):
    """Turn the list of (train, test) index arrays (ref search.py:379) into one fold id per
    row.  Requires what KFold / StratifiedKFold / GroupKFold / LeaveOneGroupOut produce: test
    sets partition the rows and each train set is the complement of its test set."""
    if len(cv_splitted) > 127:
        raise NotImplementedError("more than 127 cv splits are not supported on the device path")
    fold = np.full(n_samples, -1, dtype=np.int8)
    for k, (train, test) in enumerate(cv_splitted):
        test = np.asarray(test)
        if np.any(fold[test] != -1):
            raise NotImplementedError(
                "cv splits with overlapping test sets (e.g. ShuffleSplit, RepeatedKFold) are not "
                "supported on the device path")
        fold[test] = k
        if len(train) + len(test) != n_samples:
            raise NotImplementedError(
                "cv splits whose train set is not the complement of the test set are not "
                "supported on the device path")
    if np.any(fold < 0):
        raise NotImplementedError("cv splits must cover every row exactly once on the device path")
    for k, (train, test) in enumerate(cv_splitted):
        if np.any(fold[np.asarray(train)] == k):
            raise NotImplementedError("cv train/test sets overlap")
    return fold


class _TargetCodes:
    """One hash pass over a 1-d integer / bool target: `codes` numbers the classes by order of first
    appearance (what StratifiedKFold's `_make_test_folds` works on), `classes` are the sorted labels and
    `y_class` the index of every row's label in them (what `np.unique` + `searchsorted` give, without
    the sort over the rows)."""

    def __init__(self, y):
        import pandas as pd
        self.y = y
        self.codes, uniques = pd.factorize(y)
        uniques = np.asarray(uniques, dtype=y.dtype)
        order = np.argsort(uniques, kind="stable")
        self.classes = uniques[order]
        rank = np.empty(len(order), dtype=np.int32)
        rank[order] = np.arange(len(order), dtype=np.int32)
        self.y_class = rank[self.codes]


def _encode_target(y):
    """_TargetCodes for targets the fast path covers (1-d integer / bool arrays), else None."""
    y1 = np.asarray(y) if y is not None else None
    if y1 is None or y1.ndim != 1 or y1.dtype.kind not in "biu" or len(y1) == 0:
        return None
    return _TargetCodes(y1)


def _classes_and_ids(y, enc=None):
    """(sorted class labels, int32 class id per row)."""
    if enc is not None and enc.y is y:
        return enc.classes, enc.y_class
    classes = np.unique(y)
    return classes, np.searchsorted(classes, y).astype(np.int32)


def _cv_fold_ids(cv, X, y, groups, n_samples, enc=None):
    """(fold id per row, n_splits) of a cross-validator.  The generic route materialises every
    (train, test) index pair like the reference does (search.py:379) and converts them; the two
    splitters `check_cv` produces for an integer `cv` -- unshuffled `StratifiedKFold` / `KFold` -- are
    restated directly (no per-split index arrays, no sorts over the rows), fold for fold what
    SK/model_selection/_split.py:774-841 (`_make_test_folds`) and :531-547 (`_iter_test_indices`) give.
    `enc` (a _TargetCodes of y) saves the hash pass when the caller already has it."""
    from sklearn.model_selection import KFold, StratifiedKFold
    if type(cv) is KFold and not cv.shuffle and groups is None:
        k = cv.n_splits
        if k > n_samples:
            return _fold_ids(list(cv.split(X, y, groups)), n_samples), k     # let scikit-learn raise its error
        sizes = np.full(k, n_samples // k, dtype=np.int64)
        sizes[: n_samples % k] += 1
        return np.repeat(np.arange(k, dtype=np.int8), sizes), k
    if type(cv) is StratifiedKFold and not cv.shuffle and groups is None and y is not None:
        k = cv.n_splits
        if enc is None or enc.y is not y:
            enc = _encode_target(y)
        if enc is not None and k <= 127:
            y_encoded = enc.codes                             # classes numbered by order of appearance
            counts = np.bincount(y_encoded)
            if counts.min() >= k:                             # otherwise: scikit-learn's own warnings / errors
                # y_order = sorted codes: class c occupies positions [start_c, start_c + counts[c]); fold i
                # takes the positions congruent to i modulo k, class by class in original row order:
                # alloc[i] = #{p in [start, start + count): p mod k == i}, in closed form
                starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
                fold = np.empty(n_samples, dtype=np.int8)
                folds = np.arange(k, dtype=np.int64)
                for c in range(len(counts)):
                    s0, cnt = int(starts[c]), int(counts[c])
                    alloc = (s0 + cnt - 1 - folds) // k - (s0 - 1 - folds) // k
                    fold[np.flatnonzero(y_encoded == c)] = np.repeat(np.arange(k, dtype=np.int8), alloc)
                return fold, k
    cv_splitted = list(cv.split(X, y, groups))
    return _fold_ids(cv_splitted, n_samples), len(cv_splitted)


# ----------------------------------------------------------------------------------------
# estimator families with a device path
# ----------------------------------------------------------------------------------------
_LOGREG_SEARCHABLE = {"C", "tol", "max_iter", "fit_intercept"}


def _resolve(estimator, params):
    est = _clone(estimator)
    if params:
        est.set_params(**params)
    return est


def _check_logreg(est):
    """Raise unless `est` is a configuration the batched lbfgs kernel path reproduces
    (SK/linear_model/_logistic.py:1355-1593)."""
    p = est if isinstance(est, dict) else est.get_params(deep=False)
    bad = []
    if p.get("solver", "lbfgs") != "lbfgs":
        bad.append("solver=%r (only 'lbfgs')" % p["solver"])
    pen = p.get("penalty", "deprecated")
    if pen not in ("l2", "deprecated"):
        bad.append("penalty=%r (only 'l2')" % (pen,))
    if p.get("l1_ratio", 0.0) not in (None, 0, 0.0):
        bad.append("l1_ratio=%r" % (p["l1_ratio"],))
    if p.get("class_weight", None) is not None:
        bad.append("class_weight")
    if p.get("dual", False):
        bad.append("dual=True")
    if p.get("warm_start", False):
        bad.append("warm_start=True")
    if bad:
        raise NotImplementedError(
            "LogisticRegression configuration without a device path: " + ", ".join(bad))
    return p


_COUNT_METRICS = {"accuracy_score": "accuracy", "f1_score": "f1", "precision_score": "precision",
                  "recall_score": "recall", "balanced_accuracy_score": "balanced_accuracy"}


def _count_metric(scorer):
    """(kind, average) of the count-based metric a scikit-learn scorer computes on predict(), or None.
    average is None for accuracy / balanced accuracy, else "binary" / "micro" / "macro" / "weighted".
    All of them are functions of the confusion counts the scoring kernels deliver."""
    if type(scorer).__name__ == "_PassthroughScorer":       # estimator.score == accuracy (ref utils.py:75-143)
        return "accuracy", None
    f = getattr(scorer, "_score_func", None)
    kind = _COUNT_METRICS.get(getattr(f, "__name__", ""))
    kwargs = dict(getattr(scorer, "_kwargs", {}) or {})
    if getattr(f, "__name__", "") == "log_loss" and not kwargs and getattr(scorer, "_sign", 1) == -1:
        # scoring="neg_log_loss": -log_loss(y, predict_proba(X)) -- summed on the device (csrc/logreg_multi.cu)
        return "neg_log_loss", None
    if getattr(f, "__name__", "") == "roc_auc_score" and not kwargs and getattr(scorer, "_sign", 1) == 1:
        # scoring="roc_auc": roc_auc_score(y, decision_function(X)) -- exact pair counts on the device (csrc/auc.cu)
        return "roc_auc", None
    if kind is None or getattr(scorer, "_sign", 1) != 1:
        return None
    if kind in ("accuracy", "balanced_accuracy"):
        return None if kwargs else (kind, None)
    average = kwargs.pop("average", "binary")
    pos_label = kwargs.pop("pos_label", 1)      # the named averaged scorers ("f1_weighted", ...) carry pos_label=None
    if kwargs or average not in ("binary", "micro", "macro", "weighted"):
        return None
    if pos_label != 1 and not (average != "binary" and pos_label is None):
        return None
    return kind, average


def _metric_from_confusion(kind, average, conf):
    """scikit-learn's formulas on confusion matrices conf[..., true, predicted]
    (SK/metrics/_classification.py: accuracy_score, balanced_accuracy_score,
    precision_recall_fscore_support with zero_division -> 0.0; labels = classes present in y_true or
    y_pred, as unique_labels gives them)."""
    conf = np.asarray(conf, dtype=np.float64)
    tp = np.diagonal(conf, axis1=-2, axis2=-1)
    support = conf.sum(axis=-1)          # rows per true class
    pred = conf.sum(axis=-2)             # rows per predicted class
    total = support.sum(axis=-1)

    def div(a, b):
        return np.divide(a, b, out=np.zeros(np.broadcast(a, b).shape), where=b != 0)
    if kind == "accuracy" or average == "micro":
        return div(tp.sum(axis=-1), total)
    if kind == "balanced_accuracy":      # mean recall over the classes that occur in y_true
        has = support > 0
        return div((div(tp, support) * has).sum(axis=-1), has.sum(axis=-1).astype(np.float64))
    if kind == "precision":
        per_class = div(tp, pred)
    elif kind == "recall":
        per_class = div(tp, support)
    elif kind == "f1":
        per_class = div(2.0 * tp, support + pred)
    else:
        raise ValueError(kind)
    if average == "macro":
        present = (support + pred) > 0
        return div((per_class * present).sum(axis=-1), present.sum(axis=-1).astype(np.float64))
    if average == "weighted":
        return div((per_class * support).sum(axis=-1), total)
    raise ValueError(average)


def _metric_from_counts(kind, correct, count, pred_pos, actual_pos):
    """scikit-learn's formulas on confusion counts (SK/metrics/_classification.py: accuracy_score,
    precision_recall_fscore_support with zero_division -> 0.0, balanced_accuracy_score)."""
    correct = np.asarray(correct, dtype=np.float64)
    count = np.asarray(count, dtype=np.float64)
    if kind == "accuracy":
        return correct / np.maximum(count, 1)
    pred_pos = np.asarray(pred_pos, dtype=np.float64)
    actual_pos = np.asarray(actual_pos, dtype=np.float64)
    tp = (pred_pos + actual_pos + correct - count) / 2.0
    fp, fn = pred_pos - tp, actual_pos - tp
    tn = count - tp - fp - fn

    def div(a, b):
        return np.divide(a, b, out=np.zeros_like(a), where=b != 0)
    if kind == "precision":
        return div(tp, pred_pos)
    if kind == "recall":
        return div(tp, actual_pos)
    if kind == "f1":
        return div(2.0 * tp, actual_pos + pred_pos)
    if kind == "balanced_accuracy":
        return (div(tp, tp + fn) + div(tn, tn + fp)) / 2.0
    raise ValueError(kind)


class _LogRegFamily:
    """(candidate x fold) columns of binary L2 logistic regression."""

    name = "logreg"

    def __init__(self, estimator, candidate_params, X, y, scorers, enc=None):
        self.estimator = estimator
        self.cands = [_check_logreg(q) for q in _merged_params(estimator, candidate_params)]
        for p in candidate_params:
            extra = set(p) - _LOGREG_SEARCHABLE
            if extra:
                raise NotImplementedError(
                    "searching LogisticRegression over %s has no device path (searchable: %s)"
                    % (sorted(extra), sorted(_LOGREG_SEARCHABLE)))
        self.classes_, self.y_class = _classes_and_ids(y, enc)
        if len(self.classes_) != 2:
            raise NotImplementedError(
                "this family is binary (got %d classes)" % len(self.classes_))
        # every scorer must be a count-based metric (accuracy / precision / recall / f1 / balanced
        # accuracy on predict); scoring=None -> _PassthroughScorer -> estimator.score == accuracy
        self.metrics = {}
        for name, scorer in scorers.items():
            m = _count_metric(scorer)
            if m is None:
                raise NotImplementedError(
                    "scorer %r has no device path for classifiers (supported: accuracy, balanced_accuracy, "
                    "precision / recall / f1 with average binary, micro, macro or weighted, roc_auc, "
                    "neg_log_loss)" % (scorer,))
            # binary averaging keeps the plain name; averaged variants carry (kind, average)
            self.metrics[name] = m[0] if m[1] in (None, "binary") else m
        self.needs_pred_pos = any(k not in ("accuracy", "roc_auc", "neg_log_loss") for k in self.metrics.values())

    def stage(self, eng, X, fold, n_splits, x_staged=False):
        if not x_staged:
            parallel.stage_x_replicated(eng, X)
        eng.stage_labels(self.y_class)
        eng.stage_folds(fold, n_splits)
        if self.needs_pred_pos:     # positives per fold: only the precision / recall / f1 formulas use them
            self.pos_in_fold = np.bincount(np.asarray(fold)[self.y_class == 1], minlength=n_splits).astype(np.int64)
            self.total_pos = int(self.pos_in_fold.sum())
        else:
            self.pos_in_fold = np.zeros(n_splits, dtype=np.int64)
            self.total_pos = 0

    def _scores(self, eng, coef, codes, pos, actual_pos):
        """{scorer name: per-column value} on the rows selected by the scoring codes."""
        correct, count = eng.linear_score_batch(coef, codes, pos)
        pred_pos = None
        if self.needs_pred_pos:
            # a positive class id that matches no row makes "correct" count the predicted negatives
            neg_correct, _ = eng.linear_score_batch(coef, codes, np.full(len(pos), -7, dtype=np.int32))
            pred_pos = count - neg_correct
        out = {}
        for name, kind in self.metrics.items():
            if kind == "roc_auc":
                out[name], _ = eng.linear_auc_batch(coef, codes, pos)
            elif kind == "neg_log_loss":
                out[name] = -eng.linear_logloss_batch(coef, codes, pos)[0]
            elif isinstance(kind, tuple):      # micro / macro / weighted: 2 x 2 confusion [true, predicted]
                tp = (pred_pos + actual_pos + correct - count) / 2.0
                fp, fn = pred_pos - tp, actual_pos - tp
                conf = np.stack([np.stack([count - tp - fp - fn, fp], -1), np.stack([fn, tp], -1)], -2)
                out[name] = _metric_from_confusion(kind[0], kind[1], conf)
            else:
                out[name] = _metric_from_counts(kind, correct, count, pred_pos, actual_pos)
        return out, count

    def run_columns(self, eng, cols, n_splits, return_train_score):
        """Fit + score the given global column ids (col = cand * n_splits + fold).
        Returns dict of per-column arrays aligned with `cols`."""
        cols = np.asarray(cols, dtype=np.int64)
        out = {
            "n_test": np.zeros(len(cols), dtype=np.int64),
            "fit_time": np.zeros(len(cols)), "score_time": np.zeros(len(cols)),
            "n_iter": np.zeros(len(cols), dtype=np.int32), "status": np.zeros(len(cols), dtype=np.int32),
        }
        for name in self.metrics:           # one array per scorer: "test_<name>" (+ "train_<name>")
            out["test_%s" % name] = np.zeros(len(cols))
            if return_train_score:
                out["train_%s" % name] = np.zeros(len(cols))
        cand = cols // n_splits
        fold = (cols % n_splits).astype(np.int32)
        groups = defaultdict(list)
        for i, c in enumerate(cand):
            p = self.cands[c]
            groups[(bool(p["fit_intercept"]), float(p["tol"]), int(p["max_iter"]))].append(i)
        for (fi, tol, mi), idx in groups.items():
            idx = np.asarray(idx)
            C = np.array([self.cands[c]["C"] for c in cand[idx]], dtype=np.float64)
            pos = np.ones(len(idx), dtype=np.int32)
            t0 = time.time()
            res = eng.logreg_fit_batch(C, fold[idx], pos, fit_intercept=fi, tol=tol, max_iter=mi)
            t1 = time.time()
            vals, count = self._scores(eng, res["coef"], fold[idx], pos, self.pos_in_fold[fold[idx]])
            t2 = time.time()
            for name, v in vals.items():
                out["test_%s" % name][idx] = v
            out["n_test"][idx] = count
            out["fit_time"][idx] = (t1 - t0) / len(idx)
            out["score_time"][idx] = (t2 - t1) / len(idx)
            out["n_iter"][idx] = res["n_iter"]
            out["status"][idx] = res["status"]
            if return_train_score:
                vals, _ = self._scores(eng, res["coef"], (-3 - fold[idx]).astype(np.int32), pos,
                                       self.total_pos - self.pos_in_fold[fold[idx]])
                for name, v in vals.items():
                    out["train_%s" % name][idx] = v
        return out

    def refit(self, eng, params, X_dtype, n_features):
        p = _check_logreg(_resolve(self.estimator, params))
        res = eng.logreg_fit_batch(np.array([p["C"]]), np.array([-1], dtype=np.int32),
                                   np.array([1], dtype=np.int32), fit_intercept=p["fit_intercept"],
                                   tol=p["tol"], max_iter=p["max_iter"])
        return self.make_estimator(params, res["coef"][0], res["n_iter"][0], X_dtype, n_features)

    def make_estimator(self, params, coef_row, n_iter, X_dtype, n_features):
        """A genuine fitted sklearn LogisticRegression (attributes as set by
        SK/linear_model/_logistic.py:1561-1593) so inherited predict* work."""
        est = _resolve(self.estimator, params)
        dt = np.float64 if X_dtype == np.float64 else np.float32
        est.coef_ = coef_row[None, :n_features].astype(dt)
        if est.fit_intercept:
            est.intercept_ = coef_row[n_features:n_features + 1].astype(dt)
        else:
            est.intercept_ = np.zeros(1, dtype=dt)
        est.classes_ = self.classes_
        est.n_iter_ = np.array([n_iter], dtype=np.int32)
        est.n_features_in_ = n_features
        return est

    def fold_proba(self, eng, params, fold, n_splits):
        """preds_ support (ref search.py:551-560): per-fold refit of the best params,
        predict_proba on the held-out rows, stacked in fold order."""
        p = _check_logreg(_resolve(self.estimator, params))
        f = np.arange(n_splits, dtype=np.int32)
        res = eng.logreg_fit_batch(np.full(n_splits, p["C"]), f, np.ones(n_splits, dtype=np.int32),
                                   fit_intercept=p["fit_intercept"], tol=p["tol"], max_iter=p["max_iter"])
        dec = eng.linear_decision(res["coef"])
        preds = []
        for k in range(n_splits):
            z = dec[fold == k, k].astype(np.float64)
            p1 = 1.0 / (1.0 + np.exp(-z))
            preds.append(np.column_stack([1.0 - p1, p1]))
        return np.vstack(preds)


class _MultinomialFamily(_LogRegFamily):
    """(candidate x fold) problems of multinomial L2 logistic regression: what LogisticRegression(lbfgs)
    fits when the target has more than two classes (SK/linear_model/_logistic.py:523-547).  One
    device optimiser problem per (candidate, fold) with n_classes x (d + 1) variables."""

    name = "logreg_multinomial"

    def __init__(self, estimator, candidate_params, X, y, scorers, enc=None):
        self.estimator = estimator
        self.cands = [_check_logreg(q) for q in _merged_params(estimator, candidate_params)]
        for p in candidate_params:
            extra = set(p) - _LOGREG_SEARCHABLE
            if extra:
                raise NotImplementedError(
                    "searching LogisticRegression over %s has no device path (searchable: %s)"
                    % (sorted(extra), sorted(_LOGREG_SEARCHABLE)))
        self.classes_, self.y_class = _classes_and_ids(y, enc)
        self.n_classes = len(self.classes_)
        self.metrics = {}
        for name, scorer in scorers.items():
            m = _count_metric(scorer)
            if m is None or m[1] == "binary" or m[0] == "roc_auc":    # scikit-learn itself rejects these on a multiclass target
                raise NotImplementedError(
                    "scorer %r has no device path for a multiclass target (supported: accuracy, "
                    "balanced_accuracy, precision / recall / f1 with average micro, macro or weighted)" % (scorer,))
            self.metrics[name] = m
        self.needs_pred_pos = False

    def stage(self, eng, X, fold, n_splits, x_staged=False):
        if not x_staged:
            parallel.stage_x_replicated(eng, X)
        eng.stage_labels(self.y_class)
        eng.stage_folds(fold, n_splits)

    def run_columns(self, eng, cols, n_splits, return_train_score):
        cols = np.asarray(cols, dtype=np.int64)
        out = {
            "n_test": np.zeros(len(cols), dtype=np.int64),
            "fit_time": np.zeros(len(cols)), "score_time": np.zeros(len(cols)),
            "n_iter": np.zeros(len(cols), dtype=np.int32), "status": np.zeros(len(cols), dtype=np.int32),
        }
        for name in self.metrics:
            out["test_%s" % name] = np.zeros(len(cols))
            if return_train_score:
                out["train_%s" % name] = np.zeros(len(cols))
        cand = cols // n_splits
        fold = (cols % n_splits).astype(np.int32)
        groups = defaultdict(list)
        for i, c in enumerate(cand):
            p = self.cands[c]
            groups[(bool(p["fit_intercept"]), float(p["tol"]), int(p["max_iter"]))].append(i)
        for (fi, tol, mi), idx in groups.items():
            idx = np.asarray(idx)
            C = np.array([self.cands[c]["C"] for c in cand[idx]], dtype=np.float64)
            t0 = time.time()
            res = eng.logreg_multinomial_fit_batch(C, fold[idx], self.n_classes, fit_intercept=fi, tol=tol,
                                                   max_iter=mi)
            t1 = time.time()
            conf = eng.multinomial_confusion_batch(res["coef"], fold[idx])
            t2 = time.time()
            for name, (kind, average) in self.metrics.items():
                out["test_%s" % name][idx] = self._metric(eng, kind, average, conf, res["coef"], fold[idx])
            out["n_test"][idx] = conf.sum(axis=(1, 2))
            out["fit_time"][idx] = (t1 - t0) / len(idx)
            out["score_time"][idx] = (t2 - t1) / len(idx)
            out["n_iter"][idx] = res["n_iter"]
            out["status"][idx] = res["status"]
            if return_train_score:
                conf = eng.multinomial_confusion_batch(res["coef"], (-3 - fold[idx]).astype(np.int32))
                for name, (kind, average) in self.metrics.items():
                    out["train_%s" % name][idx] = self._metric(eng, kind, average, conf, res["coef"],
                                                               (-3 - fold[idx]).astype(np.int32))
        return out

    @staticmethod
    def _metric(eng, kind, average, conf, coef, codes):
        if kind == "neg_log_loss":
            return -eng.linear_logloss_batch(coef, codes)[0]
        return _metric_from_confusion(kind, average, conf)

    def refit(self, eng, params, X_dtype, n_features):
        p = _check_logreg(_resolve(self.estimator, params))
        res = eng.logreg_multinomial_fit_batch(np.array([p["C"]]), np.array([-1], dtype=np.int32), self.n_classes,
                                               fit_intercept=p["fit_intercept"], tol=p["tol"],
                                               max_iter=p["max_iter"])
        return self.make_estimator(params, res["coef"][0], res["n_iter"][0], X_dtype, n_features)

    def make_estimator(self, params, coef_rows, n_iter, X_dtype, n_features):
        """Fitted sklearn LogisticRegression with the multiclass attribute shapes
        (SK/linear_model/_logistic.py:1561-1593): coef_ (K, d), intercept_ (K,), n_iter_ (1,)."""
        est = _resolve(self.estimator, params)
        dt = np.float64 if X_dtype == np.float64 else np.float32
        est.coef_ = coef_rows[:, :n_features].astype(dt)
        if est.fit_intercept:
            est.intercept_ = coef_rows[:, n_features].astype(dt)
        else:
            est.intercept_ = np.zeros(self.n_classes, dtype=dt)
        est.classes_ = self.classes_
        est.n_iter_ = np.array([n_iter], dtype=np.int32)
        est.n_features_in_ = n_features
        return est

    def fold_proba(self, eng, params, fold, n_splits):
        from sklearn.utils.extmath import softmax
        p = _check_logreg(_resolve(self.estimator, params))
        f = np.arange(n_splits, dtype=np.int32)
        res = eng.logreg_multinomial_fit_batch(np.full(n_splits, p["C"]), f, self.n_classes,
                                               fit_intercept=p["fit_intercept"], tol=p["tol"],
                                               max_iter=p["max_iter"])
        K = self.n_classes
        dec = eng.linear_decision(res["coef"].reshape(n_splits * K, -1))
        return np.vstack([softmax(dec[fold == k, k * K:(k + 1) * K].astype(np.float64)) for k in range(n_splits)])


def _pick_family(estimator, candidate_params, X, y, scorers, enc=None):
    if type(estimator) is LogisticRegression:
        n_classes = len(enc.classes) if enc is not None and enc.y is y else len(np.unique(y))
        if n_classes > 2:
            return _MultinomialFamily(estimator, candidate_params, X, y, scorers, enc)
        return _LogRegFamily(estimator, candidate_params, X, y, scorers, enc)
    if type(estimator) is Ridge:
        from .ridge_family import _RidgeFamily
        return _RidgeFamily(estimator, candidate_params, X, y, scorers)
    raise NotImplementedError(
        "%s has no device path; supported base estimators: LogisticRegression(solver='lbfgs'), "
        "Ridge.  (No CPU fallback by design.)" % type(estimator).__name__)


# ----------------------------------------------------------------------------------------
# the meta-estimators
# ----------------------------------------------------------------------------------------
class DistBaseSearchCV(_ScParamMixin):
    """Same role as the reference's DistBaseSearchCV (search.py:291-581)."""

    def fit(self, X, y=None, groups=None, **fit_params):
        """Run fit with all sets of parameters (ref search.py:315-571)."""
        if fit_params:
            raise NotImplementedError("fit_params are not supported on the device path")
        _check_estimator(self, verbose=self.verbose)
        estimator = self.estimator
        cv = check_cv(self.cv, y, classifier=is_classifier(estimator))
        scorers, self.multimetric_ = _check_multimetric_scoring(self.estimator, scoring=self.scoring)
        if self.multimetric_:                                   # ref search.py:340-358
            if self.refit is not False and (not isinstance(self.refit, str) or self.refit not in scorers):
                raise ValueError(
                    "For multi-metric scoring, the parameter refit must be set to a scorer key to refit an "
                    "estimator with the best parameter setting on the whole data and make the best_* "
                    "attributes available for that metric. If this is not needed, refit should be set to "
                    "False explicitly. %r was passed." % self.refit)
            refit_metric = self.refit
        else:
            refit_metric = "score"

        X, y, groups = indexable(X, y, groups)
        n_splits = cv.get_n_splits(X, y, groups)

        X_arr = np.asarray(X)
        if X_arr.ndim != 2:
            raise ValueError("X must be a 2-d array")
        y_arr = np.asarray(y)
        n_samples, n_features = X_arr.shape
        # the host-to-device copy of X (the C-ABI call releases the GIL) runs while the host draws the
        # candidates (ParameterSampler: one scipy rvs call per candidate), computes the cv splits and
        # validates the candidates
        eng = get_engine()
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as pool:
            staged = pool.submit(parallel.stage_x_replicated, eng, X_arr)
            time.sleep(0)       # hand the GIL to the worker so that the copy starts before the host work below
            try:
                candidate_params = list(self._get_param_iterator())
                n_candidates = len(candidate_params)
                if self.verbose > 0:
                    print("Fitting {0} folds for each of {1} candidates, totalling {2} fits".format(
                        n_splits, n_candidates, n_candidates * n_splits))
                _parse_partitions(self.partitions, n_candidates * n_splits)
                enc = _encode_target(y_arr) if is_classifier(estimator) else None   # one hash pass over y for both
                fold, _ = _cv_fold_ids(cv, X, y_arr, groups, n_samples, enc)
                family = _pick_family(estimator, candidate_params, X_arr, y_arr, scorers, enc)
                if hasattr(family, "prepare"):      # host-only statistics of the folds (no engine calls)
                    family.prepare(fold, n_splits)
            finally:
                staged.result()

        rank, world, _ = parallel.dist_info()
        family.stage(eng, X_arr, fold, n_splits, x_staged=True)

        # task order: candidate-major, fold-minor (ref search.py:378-383), column = cand * n_splits + fold.
        # Ranks are dealt blocks of 128 consecutive candidates of ONE fold (fold-major order).
        n_cols = n_candidates * n_splits
        deal_order = (np.arange(n_candidates)[None, :] * n_splits + np.arange(n_splits)[:, None]).ravel()
        my_cols = parallel.shard_blocks(n_cols, rank, world, deal_order)
        loc = family.run_columns(eng, my_cols, n_splits, bool(self.return_train_score))
        metric_names = list(family.metrics)
        keys = ["n_test", "fit_time", "score_time"] + ["test_%s" % m for m in metric_names]
        if self.return_train_score:
            keys += ["train_%s" % m for m in metric_names]
        # one collective for all per-column results (counts are exact in float64)
        stacked = np.stack([np.asarray(loc[k], dtype=np.float64) for k in keys], axis=1)
        gathered = parallel.all_gather_blocks(stacked, n_cols, rank, world, deal_order)
        res = {k: gathered[:, i] for i, k in enumerate(keys)}
        res["n_test"] = np.rint(res["n_test"]).astype(np.int64)

        error_score = self.error_score
        for m in metric_names:
            bad = ~np.isfinite(res["test_%s" % m])
            if np.any(bad):
                # ref search.py:226-259 semantics for a failed fit
                if isinstance(error_score, numbers.Number):
                    res["test_%s" % m][bad] = error_score
                else:
                    raise ValueError("a fit produced a non-finite score and error_score=%r" % (error_score,))

        results = {}

        def _store(key_name, array, weights=None, splits=False, rank=False):
            """ref search.py:463-484"""
            array = np.array(array, dtype=np.float64).reshape(n_candidates, n_splits)
            if splits:
                for split_i in range(n_splits):
                    results["split%d_%s" % (split_i, key_name)] = array[:, split_i]
            array_means = np.average(array, axis=1, weights=weights)
            results["mean_%s" % key_name] = array_means
            array_stds = np.sqrt(np.average((array - array_means[:, np.newaxis]) ** 2, axis=1,
                                            weights=weights))
            results["std_%s" % key_name] = array_stds
            if rank:
                results["rank_%s" % key_name] = np.asarray(rankdata(-array_means, method="min"),
                                                           dtype=np.int32)

        _store("fit_time", res["fit_time"])
        _store("score_time", res["score_time"])
        param_results = defaultdict(partial(MaskedArray, np.empty(n_candidates,), mask=True, dtype=object))
        for cand_i, params in enumerate(candidate_params):
            for name, value in params.items():
                param_results["param_%s" % name][cand_i] = value
        results.update(param_results)
        results["params"] = candidate_params

        # ref search.py:510-519: weights = test-fold sizes when `iid` is truthy ("warn" default)
        test_sample_counts = np.array(res["n_test"][:n_splits], dtype=int)
        for m in metric_names:                                  # ref search.py:512-533
            _store("test_%s" % m, res["test_%s" % m], splits=True, rank=True,
                   weights=test_sample_counts if self.iid else None)
            if self.return_train_score:
                _store("train_%s" % m, res["train_%s" % m], splits=True)

        # ref search.py:535-541: best_* only when a refit metric is defined
        if self.refit or not self.multimetric_:
            self.best_index_ = results["rank_test_%s" % refit_metric].argmin()
            self.best_params_ = candidate_params[self.best_index_]
            self.best_score_ = results["mean_test_%s" % refit_metric][self.best_index_]

        if self.refit:
            # ref search.py:543-550 (one more full-data fit, here on the device)
            t0 = time.time()
            self.best_estimator_ = family.refit(eng, self.best_params_, X_arr.dtype, n_features)
            self.refit_time_ = time.time() - t0
            if self.preds:
                self.preds_ = family.fold_proba(eng, self.best_params_, fold, n_splits)

        self.scorer_ = scorers if self.multimetric_ else scorers["score"]   # ref search.py:563
        self.cv_results_ = results
        self.n_splits_ = n_splits

        # ref search.py:568-570
        del self.sc
        if hasattr(self.estimator, "sc"):
            del self.estimator.sc
        return self

    def get_preds(self):
        """Get CV predictions (ref search.py:573-576)."""
        if hasattr(self, "preds_"):
            return self.preds_

    def drop_preds(self):
        """Remove preds_ attribute (ref search.py:578-581)."""
        if hasattr(self, "preds_"):
            del self.preds_


class DistGridSearchCV(DistBaseSearchCV, GridSearchCV):
    """Same as sklearn `GridSearchCV` but with the fits batched on B200s.
    Constructor mirrors ref search.py:608-641 (``sc`` is the 3rd positional argument)."""

    def __init__(self, estimator, param_grid, sc=None, partitions="auto", preds=False,
                 scoring=None, n_jobs=None, iid="warn", refit=True, cv=5, verbose=0,
                 pre_dispatch="2*n_jobs", error_score="raise-deprecating",
                 return_train_score=False):
        self.estimator = estimator
        self.param_grid = param_grid
        self.sc = sc
        self.partitions = partitions
        self.preds = preds
        self.scoring = scoring
        self.n_jobs = n_jobs
        self.iid = iid
        self.refit = refit
        self.cv = cv
        self.verbose = verbose
        self.pre_dispatch = pre_dispatch
        self.error_score = error_score
        self.return_train_score = return_train_score

    def _get_param_iterator(self):
        """ref search.py:643-645"""
        return ParameterGrid(self.param_grid)


class DistRandomizedSearchCV(DistBaseSearchCV, RandomizedSearchCV):
    """Same as sklearn `RandomizedSearchCV` but with the fits batched on B200s.
    Constructor mirrors ref search.py:671-708."""

    def __init__(self, estimator, param_distributions, sc=None, partitions="auto", preds=False,
                 n_iter=10, scoring=None, n_jobs=None, iid="warn", refit=True, cv=5, verbose=0,
                 pre_dispatch="2*n_jobs", random_state=None, error_score="raise-deprecating",
                 return_train_score=False):
        self.estimator = estimator
        self.param_distributions = param_distributions
        self.sc = sc
        self.partitions = partitions
        self.preds = preds
        self.n_iter = n_iter
        self.scoring = scoring
        self.n_jobs = n_jobs
        self.iid = iid
        self.refit = refit
        self.cv = cv
        self.verbose = verbose
        self.pre_dispatch = pre_dispatch
        self.random_state = random_state
        self.error_score = error_score
        self.return_train_score = return_train_score

    def _get_param_iterator(self):
        """ref search.py:710-714"""
        return ParameterSampler(self.param_distributions, self.n_iter, random_state=self.random_state)


# ----------------------------------------------------------------------------------------
# multi-model randomized search (ref search.py:60-177, 717-908)
# ----------------------------------------------------------------------------------------
def _raw_sampler(models, n, random_state=None):
    """`n` sampled parameter sets per model, every model sampled with the same random_state
    (ref search.py:60-90)."""
    param_sets = []
    for index, (_, _, dist) in enumerate(models):
        sampler = list(ParameterSampler(dist, n_iter=_check_n_iter(n, dist), random_state=random_state))
        for sample_index, params in enumerate(sampler):
            param_sets.append({"model_index": index, "params_index": sample_index, "param_set": params})
    return param_sets


class DistMultiModelSearch(_ScParamMixin, BaseEstimator):
    """Randomized search over several (name, estimator, param_set) models at once; the
    (model, sampled params, fold) fits of every model run as columns of that model's batched
    device solve.  Constructor and fitted attributes mirror ref search.py:717-908."""

    def __init__(self, models, sc=None, partitions="auto", n=5, cv=5, scoring=None, random_state=None,
                 verbose=0, refit=True, n_jobs=None, pre_dispatch="2*n_jobs"):
        self.models = models
        self.sc = sc
        self.partitions = partitions
        self.n = n
        self.cv = cv
        self.scoring = scoring
        self.random_state = random_state
        self.verbose = verbose
        self.refit = refit
        self.n_jobs = n_jobs
        self.pre_dispatch = pre_dispatch

    def fit(self, X, y=None, groups=None, **fit_params):
        """ref search.py:800-866"""
        if fit_params:
            raise NotImplementedError("fit_params are not supported on the device path")
        _check_estimator(self, verbose=self.verbose)
        models = _validate_models(self.models, self)
        cv = check_cv(self.cv, y, classifier=is_classifier(models[0][1]))
        X, y, groups = indexable(X, y, groups)
        X_arr, y_arr = np.asarray(X), np.asarray(y)
        n_samples, n_features = X_arr.shape
        fold, n_splits = _cv_fold_ids(cv, X, y, groups, n_samples)
        param_sets = _raw_sampler(models, self.n, random_state=self.random_state)
        _parse_partitions(self.partitions, len(param_sets) * n_splits)

        rank, world, _ = parallel.dist_info()
        eng = get_engine()
        scores = np.zeros(len(param_sets))
        families = {}
        for index, (_, estimator, _) in enumerate(models):
            rows = [i for i, ps in enumerate(param_sets) if ps["model_index"] == index]
            cands = [param_sets[i]["param_set"] for i in rows]
            scorers, _ = _check_multimetric_scoring(estimator, scoring=self.scoring)
            family = _pick_family(estimator, cands, X_arr, y_arr, scorers)
            families[index] = family
            family.stage(eng, X_arr, fold, n_splits)
            n_cols = len(cands) * n_splits
            deal_order = (np.arange(len(cands))[None, :] * n_splits + np.arange(n_splits)[:, None]).ravel()
            my_cols = parallel.shard_blocks(n_cols, rank, world, deal_order)
            loc = family.run_columns(eng, my_cols, n_splits, False)
            test = parallel.all_gather_blocks(loc["test_score"], n_cols, rank, world, deal_order)
            # plain mean over folds (ref search.py:166-176: groupby(...).agg({"score": "mean"}))
            scores[rows] = np.asarray(test, dtype=np.float64).reshape(len(cands), n_splits).mean(axis=1)
        if self.verbose:
            for index, (name, _, _) in enumerate(models):
                best = max(scores[i] for i, ps in enumerate(param_sets) if ps["model_index"] == index)
                print("model %d (%s): best score %.6f" % (index, name, best))

        best_index = int(np.argmax(scores))                              # ref :838
        self.best_model_index_ = param_sets[best_index]["model_index"]
        self.best_model_name_ = models[self.best_model_index_][0]
        self.best_params_ = param_sets[best_index]["param_set"]
        self.best_score_ = scores[best_index]
        self.worst_score_ = scores[best_index]                           # sic (ref :843)
        self.cv_results_ = {                                             # ref :844-859
            "model_index": [ps["model_index"] for ps in param_sets],
            "model_name": [models[ps["model_index"]][0] for ps in param_sets],
            "params": [ps["param_set"] for ps in param_sets],
            "rank_test_score": list(np.asarray(rankdata(-scores), dtype=np.int32)),
            "mean_test_score": list(scores),
        }
        if self.refit:                                                   # ref :861-864
            family = families[self.best_model_index_]
            family.stage(eng, X_arr, fold, n_splits)
            self.best_estimator_ = family.refit(eng, self.best_params_, X_arr.dtype, n_features)
        del self.sc
        return self

    def _check_is_fitted(self):
        """ref search.py:868-880"""
        from sklearn.exceptions import NotFittedError
        if not self.refit:
            raise NotFittedError(
                "This %s instance was initialized with refit=False. The method is available only after "
                "refitting on the best parameters. You can refit an estimator manually using the "
                "``best_params_`` attribute" % (type(self).__name__))
        from sklearn.utils.validation import check_is_fitted
        check_is_fitted(self, "best_estimator_")

    def _delegate(self, name, X):
        self._check_is_fitted()
        if not hasattr(self.best_estimator_, name):
            raise AttributeError("%s has no %s" % (type(self.best_estimator_).__name__, name))
        return getattr(self.best_estimator_, name)(X)

    def predict(self, X):
        return self._delegate("predict", X)

    def predict_proba(self, X):
        return self._delegate("predict_proba", X)

    def predict_log_proba(self, X):
        return self._delegate("predict_log_proba", X)

    def decision_function(self, X):
        return self._delegate("decision_function", X)

    @property
    def classes_(self):
        self._check_is_fitted()
        return self.best_estimator_.classes_
