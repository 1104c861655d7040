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


# ----------------------------------------------------------------------------------------
# multinomial (more than two classes): SK/linear_model/_logistic.py:523-547, 584-598, 704-712
# ----------------------------------------------------------------------------------------
def multinomial_loss_grad_pointwise(y, raw):
    """CyHalfMultinomialLoss.loss_gradient, sample_weight=None (SK/_loss/_loss.pyx.tp:1293-1327)
    with sum_exp_minus_max (:269-305): exp in double stored in raw's dtype, the sum of those
    accumulated in double class by class, everything else in raw's dtype."""
    dt = raw.dtype
    n, K = raw.shape
    mx = raw.max(axis=1)
    p = np.exp(raw.astype(np.float64) - mx.astype(np.float64)[:, None]).astype(dt)
    s = np.zeros(n, dtype=np.float64)
    for k in range(K):                      # val2 += p[k], k ascending
        s += p[:, k].astype(np.float64)
    sum_exps = s.astype(dt)
    loss = (np.log(sum_exps.astype(np.float64)) + mx.astype(np.float64)).astype(dt)
    yi = y.astype(np.intp)
    loss = (loss - raw[np.arange(n), yi]).astype(dt)
    p = (p / sum_exps[:, None]).astype(dt)
    onehot = np.zeros((n, K), dtype=dt)
    onehot[np.arange(n), yi] = 1
    return loss, (p - onehot).astype(dt)


def multinomial_loss_gradient(coef, X, y, l2_reg_strength, n_classes, fit_intercept=True):
    """LinearModelLoss.loss_gradient, multiclass branch, coef 1-d with the classes of one feature
    contiguous (SK/linear_model/_linear_loss.py:147-223, 291-379)."""
    n, d = X.shape
    W = coef.reshape((n_classes, -1), order="F")               # :173
    if fit_intercept:
        intercept, weights = W[:, -1], W[:, :-1]
    else:
        intercept, weights = 0.0, W
    w_x = np.asarray(weights, dtype=X.dtype)                   # :216
    b_x = np.asarray(intercept, dtype=X.dtype)                 # :217
    raw = X @ w_x.T + b_x                                      # :222
    loss_i, g_i = multinomial_loss_grad_pointwise(y, raw)
    loss = float(np.sum(loss_i) / n)                           # :349
    loss += float(0.5 * l2_reg_strength * np.dot(weights.ravel(order="K"), weights.ravel(order="K")))   # :350 squared_norm
    g_i /= n                                                   # :352
    grad = np.empty((n_classes, d + int(fit_intercept)), dtype=weights.dtype, order="F")   # :363
    grad[:, :d] = g_i.T @ X + l2_reg_strength * weights        # :365-368
    if fit_intercept:
        grad[:, -1] = np.sum(g_i, axis=0)                      # :369-372
    return loss, grad.ravel(order="F")                         # :373-374


def fit_multinomial_lbfgs(X, y_cls, n_classes, C=1.0, tol=1e-4, max_iter=100, fit_intercept=True):
    """_logistic_regression_path, solver='lbfgs', n_classes > 2
    (SK/linear_model/_logistic.py:443-457, 523-547, 580-604, 704-712).
    y_cls: class ids 0..n_classes-1.  Returns (coef [K, d] X.dtype, intercept [K], n_iter)."""
    n, d = X.shape
    w0 = np.zeros((n_classes, d + int(fit_intercept)), dtype=X.dtype, order="F").ravel(order="F")
    y = np.asarray(y_cls, dtype=X.dtype)
    l2 = 1.0 / (C * n)
    res = optimize.minimize(lambda w: multinomial_loss_gradient(w, X, y, l2, n_classes, fit_intercept),
                            w0, method="L-BFGS-B", jac=True,
                            options={"maxiter": max_iter, "maxls": 50, "gtol": tol,
                                     "ftol": 64 * np.finfo(float).eps})
    W = np.asarray(np.reshape(res.x, (n_classes, -1), order="F"), dtype=X.dtype)
    n_iter = min(res.nit, max_iter)
    if fit_intercept:
        return W[:, :d], W[:, d], n_iter
    return W, np.zeros(n_classes, dtype=X.dtype), n_iter
