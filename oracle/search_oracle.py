"""CPU restatement of the reference's search fan-out.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/skdist/distribute/search.py:
  * fit_and_score()  <- _fit_and_score            search.py:180-288
  * search_cv()      <- DistBaseSearchCV.fit      search.py:315-571
      task order (candidate-major, fold-minor)    search.py:378-383
      _store / weighted mean+std / rank           search.py:463-484
      iid weighting by test-fold size             search.py:509-519
      best_index_ = rank.argmin() (first best)    search.py:538-541
      refit on full (X, y)                        search.py:543-550
The arithmetic inside ``estimator.fit`` / the scorer is the installed
scikit-learn (the reference's third-party dependency), exactly as in the
reference.  ``DistGridSearchCV.__init__`` itself cannot be instantiated under
sklearn 1.9 (it passes ``iid=``, search.py:625-638), which is why the driver
loop is restated instead of called.
"""
import time
from collections import defaultdict
from functools import partial
from itertools import product

import numpy as np
from numpy.ma import MaskedArray
from scipy.stats import rankdata
from sklearn.base import clone, is_classifier
from sklearn.metrics import check_scoring
from sklearn.model_selection import check_cv


def fit_and_score(estimator, X, y, scorer, train, test, parameters,
                  return_train_score=False):
    """search.py:180-288 with error_score='raise', fit_params={}."""
    est = clone(estimator)                                   # :197
    if parameters is not None:
        est.set_params(**parameters)                         # :216
    t0 = time.time()
    X_train, y_train = X[train], y[train]                    # :220  (_safe_split -> fancy-index copy)
    X_test, y_test = X[test], y[test]                        # :221
    est.fit(X_train, y_train)                                # :230
    fit_time = time.time() - t0
    test_score = scorer(est, X_test, y_test)                 # :264
    if hasattr(test_score, "item"):
        test_score = test_score.item()
    score_time = time.time() - t0 - fit_time
    ret = []
    if return_train_score:
        tr = scorer(est, X_train, y_train)
        ret.append({"score": tr.item() if hasattr(tr, "item") else tr})
    ret += [{"score": test_score}, len(test), fit_time, score_time]
    return ret


def search_cv(estimator, candidate_params, X, y, cv=5, scoring=None, iid=True,
              refit=True, return_train_score=False, task_fn=None):
    """Restatement of DistBaseSearchCV.fit (search.py:315-571), sc=None branch.

    ``task_fn(estimator, X, y, scorer, train, test, params)`` defaults to
    :func:`fit_and_score`; tests pass the reference's own ``_fit_and_score``
    (through oracle.refshim) to pin this loop against it.
    Returns a dict with cv_results_, best_index_, best_params_, best_score_,
    best_estimator_ (if refit), n_splits_.
    """
    cv_obj = check_cv(cv, y, classifier=is_classifier(estimator))      # :334
    scorer = check_scoring(estimator, scoring=scoring)                 # :336 (single metric)
    candidate_params = list(candidate_params)                          # :365
    n_candidates = len(candidate_params)
    cv_splitted = list(cv_obj.split(X, y, None))                       # :379
    n_splits = len(cv_splitted)
    if task_fn is None:
        task_fn = partial(fit_and_score, return_train_score=return_train_score)
    out = []
    for params, (train, test) in product(candidate_params, cv_splitted):   # :381
        out.append(task_fn(estimator, X, y, scorer, train, test, params))
    if return_train_score:
        train_d, test_d, counts, fit_time, score_time = zip(*out)
    else:
        test_d, counts, fit_time, score_time = zip(*out)               # :451
    test_scores = np.asarray([d["score"] for d in test_d])             # :457

    results = {}

    def _store(key_name, array, weights=None, splits=False, rank=False):   # :463-484
        array = np.array(array, dtype=np.float64).reshape(n_candidates, n_splits)
        if splits:
            for i in range(n_splits):
                results["split%d_%s" % (i, key_name)] = array[:, i]
        means = np.average(array, axis=1, weights=weights)
        results["mean_%s" % key_name] = means
        results["std_%s" % key_name] = np.sqrt(
            np.average((array - means[:, np.newaxis]) ** 2, axis=1, weights=weights))
        if rank:
            results["rank_%s" % key_name] = np.asarray(
                rankdata(-means, method="min"), dtype=np.int32)

    _store("fit_time", fit_time)
    _store("score_time", score_time)
    param_results = defaultdict(partial(MaskedArray, np.empty(n_candidates,),
                                        mask=True, dtype=object))     # :488-497
    for i, params in enumerate(candidate_params):
        for name, value in params.items():
            param_results["param_%s" % name][i] = value
    results.update(param_results)
    results["params"] = candidate_params
    counts = np.array(counts[:n_splits], dtype=int)                    # :510
    _store("test_score", test_scores, splits=True, rank=True,
           weights=counts if iid else None)                            # :513-519
    if return_train_score:
        _store("train_score", [d["score"] for d in train_d], splits=True)

    res = {"cv_results_": results, "n_splits_": n_splits}
    res["best_index_"] = int(results["rank_test_score"].argmin())      # :539
    res["best_params_"] = candidate_params[res["best_index_"]]
    res["best_score_"] = results["mean_test_score"][res["best_index_"]]
    if refit:
        best = clone(estimator).set_params(**res["best_params_"])      # :544
        best.fit(X, y)
        res["best_estimator_"] = best
    return res
