"""tcgen05 path vs oracle: objective / gradient at random points, several shapes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import logreg_oracle as lo
from skdist_b200.datasets import make_g1_classification
from skdist_b200.engine import Engine
from sklearn.model_selection import StratifiedKFold

eng = Engine(0)
bad = 0
shapes = [(1000, 40, 9, 4), (5000, 100, 20, 5), (20000, 256, 130, 5), (70000, 200, 300, 3)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for (n, d, B, cv) in shapes:
    X, y = make_g1_classification(n, d, seed=n % 97)
    X[:, 0] *= 1000.0            # exercise the per-feature scaling
    X[:, d - 1] *= 1e-3
    fold = np.zeros(n, np.int8)
    for k, (_, te) in enumerate(StratifiedKFold(cv).split(X, y)):
        fold[te] = k
    eng.stage_x(X); eng.stage_labels(y); eng.stage_folds(fold, cv)
    rng = np.random.default_rng(1)
    W = rng.standard_normal((B, d + 1)) * 0.3
    W[:, 0] /= 1000.0; W[:, d - 1] *= 1e3
    W[0] = 0.0
    C = np.logspace(-3, 3, B)
    cf = (np.arange(B) % (cv + 1) - 1).astype(np.int32)
    pos = np.ones(B, np.int32)
    out = {}
    for kern in (1, 2):
        eng.set_kernel(kern)
        t0 = time.time()
        out[kern] = eng.logreg_loss_grad(W, C, cf, pos)
        print("  kernel %d: %.3fs" % (kern, time.time() - t0))
    ef = eg = 0.0
    for j in range(min(B, 12)):
        m = np.ones(n, bool) if cf[j] < 0 else fold != cf[j]
        fo, go = lo.loss_gradient(W[j], X[m], y[m].astype(np.float32), 1.0 / (C[j] * m.sum()))
        for kern in (1, 2):
            f, g = out[kern]
            rf = abs(f[j] - fo) / abs(fo); rg = np.abs(g[j] - go).max() / np.abs(go).max()
            if kern == 2:
                ef = max(ef, rf); eg = max(eg, rg)
            if j < 3:
                print("    col %d kern %d: f %.9g (oracle %.9g) rel %.2e | grad rel %.2e" % (j, kern, f[j], fo, rf, rg))
    dcol = np.abs(out[1][1] - out[2][1]).max(1) / np.abs(out[1][1]).max(1)
    jw = int(dcol.argmax())
    m = np.ones(n, bool) if cf[jw] < 0 else fold != cf[jw]
    X64 = X[m].astype(np.float64); w32 = W[jw].astype(np.float32).astype(np.float64)
    z = X64 @ w32[:d] + w32[d]; yy = y[m].astype(np.float64)
    gp = 1.0 / (1.0 + np.exp(-z)) - yy
    g64 = np.r_[X64.T @ gp / m.sum() + W[jw, :d] / (C[jw] * m.sum()), gp.sum() / m.sum()]
    for kern in (1, 2):
        e = np.abs(out[kern][1][jw] - g64)
        print("    worst col %d (C=%.3g fold=%d |z|max=%.1f): kern %d vs float64: max abs err %.3e at k=%d (|g|max %.3e, g64[k]=%.3e)"
              % (jw, C[jw], cf[jw], np.abs(z).max(), kern, e.max(), e.argmax(), np.abs(g64).max(), g64[e.argmax()]))
    d12f = np.abs(out[1][0] - out[2][0]).max() / np.abs(out[1][0]).max()
    d12g = (np.abs(out[1][1] - out[2][1]).max(1) / np.abs(out[1][1]).max(1)).max()
    print("shape n=%d d=%d B=%d: TC vs oracle max rel f %.2e grad %.2e | TC vs SIMT all cols f %.2e grad %.2e"
          % (n, d, B, ef, eg, d12f, d12g))
    # fp32 accumulation inside the tensor core rounds toward zero: with random W (large, same-sign
    # partial sums) the gradient carries a ~1e-5 relative bias; near an optimum it vanishes
    if not (ef < 5e-6 and eg < 5e-5 and d12g < 5e-4):
        bad += 1
print("TC_CHECK", "FAIL" if bad else "PASS")
sys.exit(1 if bad else 0)
