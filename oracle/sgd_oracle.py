"""Pure-numpy restatement of the binary SGDClassifier fit that the reference's `_fit_binary`
(ref multiclass.py:109-152) runs per label column.  TEST INFRASTRUCTURE ONLY (slow: small cases).

The algorithm lives in scikit-learn 1.9.0 (the reference's third-party dependency):
  * driver        SK/linear_model/_stochastic_gradient.py:387-515 (fit_binary: y in {-1,+1} for
                  hinge, {0,1} for log_loss = CyHalfBinomialLoss,
                  seed = RandomState(random_state).randint(MAX_INT), intercept_decay = 1 dense)
  * inner loop    SK/linear_model/_sgd_fast.pyx.tp:274-640 (_plain_sgd32: float32 weights,
                  float64 scalars, "optimal" schedule eta = 1/(alpha*(optimal_init+t-1)),
                  lazy weight scale, objective-based stopping n_iter_no_change / tol)
  * weight vector SK/utils/_weight_vector.pyx.tp (WeightVector32: dot/add/scale/reset_wscale with
                  their exact float/double mixing)
  * shuffle       SK/utils/_seq_dataset.pyx.tp:137-145 (Fisher-Yates, the SAME seed every epoch,
                  applied to the evolving index array), SK/utils/_random.pxd:20-34 (xorshift32)
tests/test_oracle.py checks fit_binary_sgd() equals sklearn's SGDClassifier bit for bit.
"""
import math

import numpy as np

f32 = np.float32
f64 = np.float64
MAX_INT = np.iinfo(np.int32).max


def rand_r(seed):
    """SK/utils/_random.pxd:20-34; returns (value, new_seed)."""
    if seed == 0:
        seed = 1
    seed ^= (seed << 13) & 0xFFFFFFFF
    seed ^= seed >> 17
    seed ^= (seed << 5) & 0xFFFFFFFF
    return seed % (2 ** 31), seed


def shuffle_inplace(ind, seed):
    n = len(ind)
    for i in range(n - 1):
        r, seed = rand_r(seed)
        j = i + r % (n - i)
        ind[i], ind[j] = ind[j], ind[i]


def _log1pexp(x):
    """SK/_loss/_loss.pyx.tp:256-266."""
    if x <= -37:
        return math.exp(x)
    if x <= -2:
        return math.log1p(math.exp(x))
    if x <= 18:
        return math.log(1.0 + math.exp(x))
    if x <= 33.3:
        return x + math.exp(-x)
    return x


def _loss_grad(loss, y, p):
    """(cy_loss, cy_gradient).  Hinge takes y in {-1,+1} (SK/linear_model/_sgd_fast.pyx.tp:131-146);
    log_loss is CyHalfBinomialLoss on y in {0,1} (SK/_loss/_loss.pyx.tp:686-725,
    SK/linear_model/_stochastic_gradient.py:350-361,450)."""
    if loss == "hinge":
        z = p * y
        return (1.0 - z, -y) if z <= 1.0 else (0.0, 0.0)
    y01 = 1.0 if y > 0 else 0.0
    cur = _log1pexp(p) - y01 * p
    if p > -37:
        e = math.exp(-p)
        return cur, ((1 - y01) - y01 * e) / (1 + e)
    return cur, math.exp(p) - y01


def fit_binary_sgd(X, y_pm1, loss="hinge", alpha=1e-4, fit_intercept=True, max_iter=1000, tol=1e-3,
                   shuffle=True, random_state=0, n_iter_no_change=5):
    """Returns (coef float32[d], intercept float64, n_iter, t_).  learning_rate='optimal', l2."""
    X = np.ascontiguousarray(X, dtype=f32)
    n, d = X.shape
    y = np.asarray(y_pm1, dtype=f32)
    w = np.zeros(d, f32)
    wscale = 1.0
    sq_norm = 0.0
    intercept = 0.0
    rs = np.random.RandomState(random_state)
    rs.randint(1, MAX_INT)          # make_dataset() draws the dataset's own seed first (SK/linear_model/_base.py:make_dataset)
    seed = int(rs.randint(MAX_INT))  # SK/linear_model/_stochastic_gradient.py:473
    typw = np.sqrt(1.0 / np.sqrt(alpha))
    g0 = _loss_grad(loss, 1.0, -typw)[1]
    optimal_init = 1.0 / ((typw / max(1.0, g0)) * alpha)
    t = 1.0
    ind = np.arange(n)
    best_objective = np.inf
    no_improve = 0
    epoch = 0
    with np.errstate(over="ignore"):
        for epoch in range(max_iter):
            objective_sum = 0.0
            if shuffle:
                shuffle_inplace(ind, seed)
            for i in range(n):
                x = X[ind[i]]
                yi = float(y[ind[i]])
                # WeightVector32.dot: float products, double sequential sum, * wscale, return float
                innerprod = 0.0
                prods = (w * x)                      # float32 products
                for v in prods:
                    innerprod += float(v)
                p = float(f32(innerprod * wscale)) + intercept
                eta = 1.0 / (alpha * (optimal_init + t - 1))
                cur_loss, dloss = _loss_grad(loss, yi, p)
                norm_f = f32(np.sqrt(sq_norm))
                objective_sum += cur_loss + alpha * (0.5 * float(f32(norm_f * norm_f)))
                dloss = min(max(dloss, -1e12), 1e12)
                update = -eta * dloss
                # w.scale(c)
                c = f32(max(0.0, 1.0 - eta * alpha))
                wscale *= float(c)
                sq_norm *= float(f32(c * c))
                if wscale < 1e-6:
                    w *= f32(wscale)
                    wscale = 1.0
                if update != 0.0:
                    cf = f32(update)
                    wsf = f32(wscale)
                    q = f32(cf / wsf)
                    w = (w.astype(f64) + x.astype(f64) * float(q)).astype(f32)
                    acc = 0.0
                    for v in (w * w):
                        acc += float(v)
                    sq_norm = acc * float(f32(wsf * wsf))
                if fit_intercept:
                    intercept += update
                t += 1
            if not np.isfinite(intercept) or not np.all(np.isfinite(w)):
                raise ValueError("Floating-point under-/overflow occurred at epoch #%d" % (epoch + 1))
            obj = objective_sum / n
            if tol > -np.inf and obj > best_objective - tol:
                no_improve += 1
            else:
                no_improve = 0
            if obj < best_objective:
                best_objective = obj
            if no_improve >= n_iter_no_change:
                break
    w = w * f32(wscale)
    return w.astype(f32), intercept, epoch + 1, t
