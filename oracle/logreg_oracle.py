"""numpy restatement of the binary L2 LogisticRegression(lbfgs) fit that the
reference's per-task function runs (search.py:230 -> sklearn).  TEST
INFRASTRUCTURE ONLY.

The algorithm lives in the reference's third-party dependency scikit-learn
(1.9.0 here) + scipy 1.18.1 (L-BFGS-B 3.0, C translation); citations are to
site-packages/sklearn (``SK/``) and site-packages/scipy (``SP/``):
  * objective/gradient   SK/linear_model/_linear_loss.py:291-379 (loss_gradient)
                         SK/linear_model/_linear_loss.py:186-223 (weight_intercept_raw)
  * pointwise loss/grad  SK/_loss/_loss.pyx.tp:728-751 (closs_grad_half_binomial)
  * driver + options     SK/linear_model/_logistic.py:580-604
                         (l2_reg_strength = 1/(C*n); maxiter, maxls=50, gtol=tol,
                          ftol=64*eps), SP/optimize/_lbfgsb_py.py:272-437
  * n_iter_              SK/utils/optimize.py (_check_optimize_result: min(nit, max_iter))
  * scoring              accuracy of (X @ coef.T + intercept > 0)
tests/test_oracle.py checks fit_binary_lbfgs() is bit-identical to
sklearn.linear_model.LogisticRegression().fit on the same inputs.
"""
import numpy as np
from scipy import optimize


def loss_grad_pointwise(y, raw):
    """closs_grad_half_binomial in double, outputs stored in raw's dtype
    (SK/_loss/_loss.pyx.tp:728-751, 1057-1084)."""
    y = y.astype(np.float64)
    z = raw.astype(np.float64)
    loss = np.empty_like(z)
    grad = np.empty_like(z)
    a = z <= -37
    b = (~a) & (z <= -2)
    c = (~a) & (~b) & (z <= 18)
    d = z > 18
    with np.errstate(over="ignore"):
        e = np.exp(z[a]); loss[a] = e - y[a] * z[a]; grad[a] = e - y[a]
        e = np.exp(z[b]); loss[b] = np.log1p(e) - y[b] * z[b]
        grad[b] = ((1 - y[b]) * e - y[b]) / (1 + e)
        e = np.exp(-z[c]); loss[c] = np.log1p(e) + (1 - y[c]) * z[c]
        grad[c] = ((1 - y[c]) - y[c] * e) / (1 + e)
        e = np.exp(-z[d]); loss[d] = e + (1 - y[d]) * z[d]
        grad[d] = ((1 - y[d]) - y[d] * e) / (1 + e)
    return loss.astype(raw.dtype), grad.astype(raw.dtype)


def loss_gradient(coef, X, y, l2_reg_strength, fit_intercept=True):
    """LinearModelLoss.loss_gradient, binary, sample_weight=None
    (SK/linear_model/_linear_loss.py:291-379)."""
    n, d = X.shape
    if fit_intercept:
        weights, intercept = coef[:-1], coef[-1]
    else:
        weights, intercept = coef, 0.0
    w_x = np.asarray(weights, dtype=X.dtype)                   # :216
    b_x = np.asarray(intercept, dtype=X.dtype)                 # :217
    raw = X @ w_x + b_x                                        # :219
    loss_i, g_i = loss_grad_pointwise(y, raw)
    loss = float(np.sum(loss_i) / n)                           # :349
    loss += float(0.5 * l2_reg_strength * (weights @ weights))  # :350, l2_penalty
    g_i /= n                                                   # :352
    grad = np.empty_like(coef, dtype=weights.dtype)
    grad[:d] = X.T @ g_i + l2_reg_strength * weights           # :356-359
    if fit_intercept:
        grad[-1] = np.sum(g_i)                                 # :361
    return loss, grad


def fit_binary_lbfgs(X, y01, C=1.0, tol=1e-4, max_iter=100, fit_intercept=True,
                     return_trace=False):
    """_logistic_regression_path, solver='lbfgs', binary
    (SK/linear_model/_logistic.py:443-457, 580-604, 681-685).
    y01: array of 0/1 in X.dtype.  Returns (coef[d] X.dtype, intercept, n_iter)."""
    n, d = X.shape
    w0 = np.zeros(d + int(fit_intercept), dtype=X.dtype)       # :443
    l2 = 1.0 / (C * n)                                         # :580
    trace = []

    def func(w):
        f, g = loss_gradient(w, X, y01, l2, fit_intercept)
        if return_trace:
            trace.append((w.copy(), f, g.copy()))
        return f, g

    res = optimize.minimize(func, w0, method="L-BFGS-B", jac=True,
                            options={"maxiter": max_iter, "maxls": 50, "gtol": tol,
                                     "ftol": 64 * np.finfo(float).eps})
    n_iter = min(res.nit, max_iter)
    w = np.asarray(res.x, dtype=X.dtype)                       # :681-685
    coef = w[:d] if fit_intercept else w
    intercept = w[-1] if fit_intercept else X.dtype.type(0)
    out = (coef, intercept, n_iter)
    if return_trace:
        out += (trace, res)
    return out


def accuracy(X, y01, coef, intercept):
    """ClassifierMixin.score -> accuracy_score(y, predict(X)); predict is
    decision_function > 0 (SK/linear_model/_base.py)."""
    scores = X @ coef.astype(X.dtype) + X.dtype.type(intercept)
    pred = scores > 0
    return float(np.mean(pred == (y01 > 0.5)))
