"""numpy restatement of the dense Ridge fit the reference's per-task function runs
(ref search.py:230 -> sklearn Ridge.fit).  TEST INFRASTRUCTURE ONLY.

The algorithm lives in scikit-learn 1.9.0 (the reference's third-party dependency):
  * centring            SK/linear_model/_base.py:113-220  (_preprocess_data: X -= mean, y -= mean,
                        both in X.dtype; mean via np.average)
  * solver auto->cholesky for dense X        SK/linear_model/_ridge.py (resolve_solver)
  * _solve_cholesky     SK/linear_model/_ridge.py:215-234 (A = X.T X; Xy = X.T y;
                        A.flat[::d+1] += alpha; scipy.linalg.solve(assume_a="pos") -> LAPACK posv)
  * intercept           SK/linear_model/_base.py (_set_intercept: y_offset - X_offset @ coef)
  * score               RegressorMixin.score -> r2_score
tests/test_oracle.py checks fit_ridge() against sklearn.linear_model.Ridge bit for bit.
"""
import numpy as np
from scipy import linalg


def fit_ridge(X, y, alpha=1.0, fit_intercept=True):
    """Returns (coef[d] in X.dtype, intercept)."""
    X = np.array(X, copy=True)
    y = np.array(y, dtype=X.dtype, copy=True)
    n, d = X.shape
    if fit_intercept:
        X_offset = np.average(X, axis=0).astype(X.dtype, copy=False)
        X -= X_offset
        y_offset = np.average(y, axis=0)
        y -= y_offset
    else:
        X_offset = np.zeros(d, X.dtype)
        y_offset = X.dtype.type(0)
    A = X.T @ X
    Xy = X.T @ y.reshape(-1, 1)
    A.flat[:: d + 1] += alpha
    coef = linalg.solve(A, Xy, assume_a="pos", overwrite_a=True).T[0]
    intercept = y_offset - X_offset @ coef if fit_intercept else 0.0
    return coef, intercept


def r2(X, y, coef, intercept):
    """RegressorMixin.score (r2_score, uniform weights)."""
    from sklearn.metrics import r2_score
    pred = X @ coef + intercept
    return float(r2_score(y, pred))
