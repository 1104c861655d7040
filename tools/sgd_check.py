"""Column-batched exact-order SGD vs scikit-learn's SGDClassifier: coefficients, intercepts, n_iter_
and t_ must be bit-identical (hinge and log_loss, several shapes)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from sklearn.linear_model import SGDClassifier
from skdist_b200.datasets import make_multiclass
from skdist_b200.engine import Engine
eng = Engine(0)
X, y = make_multiclass(3000, 40, 7, seed=12)
eng.stage_x(X); eng.stage_labels(y.astype(np.int32)); eng.stage_folds(None, 0)
for loss in ("hinge", "log_loss"):
    for shuffle in (True, False):
        for mi in (1, 3, 1000):
            est = SGDClassifier(loss=loss, random_state=1, shuffle=shuffle, max_iter=mi, tol=None if mi < 1000 else 1e-3)
            res = eng.sgd_fit_batch(est, np.array([0, 3], np.int32))
            for j, k in enumerate([0, 3]):
                m = SGDClassifier(loss=loss, random_state=1, shuffle=shuffle, max_iter=mi, tol=None if mi < 1000 else 1e-3).fit(X, (y == k).astype(int))
                d = np.abs(res["coef32"][j] - m.coef_[0]).max() / np.abs(m.coef_[0]).max()
                print("%-8s shuffle=%-5s max_iter=%-4d class %d: n_iter %d/%d rel coef diff %.2e intercept diff %.2e"
                      % (loss, shuffle, mi, k, res["n_iter"][j], m.n_iter_, d, abs(res["intercept"][j] - m.intercept_[0])))
