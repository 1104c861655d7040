"""Phase trace of the headline step (fit 2560 columns + score) on one GPU: SKDIST_B200_TRACE=1."""
import os, sys, time
import numpy as np
os.environ.setdefault("SKDIST_B200_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from skdist_b200.datasets import make_g1_classification
from skdist_b200.engine import Engine
n, d, cand, folds = 1000000, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 512, 5
X, y = make_g1_classification(n, d, seed=0)
fold = bench.fold_ids(y, folds)
Cs = np.logspace(-4, 4, cand)
eng = Engine(0)
eng.profile(1)
eng.stage_x(X); eng.stage_labels(y.astype(np.int32)); eng.stage_folds(fold, folds)
C_cols = np.repeat(Cs, folds); f_cols = np.tile(np.arange(folds, dtype=np.int32), cand); pos = np.ones(len(C_cols), np.int32)
for s in range(int(os.environ.get('STEPS', 3))):
    t0 = time.perf_counter()
    res = eng.logreg_fit_batch(C_cols, f_cols, pos)
    t1 = time.perf_counter()
    correct, count = eng.linear_score_batch(res["coef"], f_cols, pos)
    t2 = time.perf_counter()
    print("step %d: fit %.1f ms (gpu %.1f ms)  score %.1f ms" % (s, 1e3 * (t1 - t0), 1e3 * res.get("gpu_seconds", 0), 1e3 * (t2 - t1)), file=sys.stderr)
