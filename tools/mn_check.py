"""Multinomial device fit vs the oracle restatement on 10-class digits: iteration counts, final
objective, coefficient deviation, at the default and at a tight tolerance."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.datasets import load_digits
from sklearn.model_selection import StratifiedKFold
from oracle import logreg_oracle as lo
from skdist_b200.engine import Engine

dg = load_digits()
X, y = (dg.data / 16).astype(np.float32), dg.target.astype(np.int32)
fold = np.zeros(len(y), np.int8)
for k, (_, te) in enumerate(StratifiedKFold(3).split(X, y)):
    fold[te] = k
eng = Engine(0)
eng.stage_x(X); eng.stage_labels(y); eng.stage_folds(fold, 3)
Cs = np.array([0.01, 0.1, 1.0, 10.0])
C = np.repeat(Cs, 3)
cf = np.tile(np.arange(3, dtype=np.int32), 4)
for tol, mi in ((1e-4, 300), (1e-7, 3000)):
    t0 = time.time()
    res = eng.logreg_multinomial_fit_batch(C, cf, 10, tol=tol, max_iter=mi)
    print("tol %g: device %.3f s (gpu %.3f s)" % (tol, time.time() - t0, res["gpu_seconds"]))
    for j in range(len(C)):
        tr = fold != cf[j]
        W, b, it = lo.fit_multinomial_lbfgs(X[tr], y[tr], 10, C=C[j], tol=tol, max_iter=mi)
        l2 = 1.0 / (C[j] * tr.sum())
        w_or = np.concatenate([W, b[:, None]], 1).astype(np.float64).ravel(order="F")
        w_dev = res["coef"][j].astype(np.float64).ravel(order="F")
        yt = y[tr].astype(np.float32)
        f_or, g_or = lo.multinomial_loss_gradient(w_or, X[tr], yt, l2, 10)
        f_dev, g_dev = lo.multinomial_loss_gradient(w_dev, X[tr], yt, l2, 10)
        print("  C=%-5g fold %d  n_iter %3d / %3d  status %d  f %.9f / %.9f (dev reports %.9f)  |g|inf %.2e / %.2e  coef dev %.2e"
              % (C[j], cf[j], res["n_iter"][j], it, res["status"][j], f_dev, f_or, res["loss"][j],
                 np.abs(g_dev).max(), np.abs(g_or).max(),
                 np.abs(res["coef"][j, :, :64] - W).max() / np.abs(W).max()))
