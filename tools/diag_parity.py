"""Print GPU-vs-golden deviations (coefficients, n_iter, flipped predictions) per column."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_parity import _case, _fold_ids, GOLD
from skdist_b200.engine import Engine

eng = Engine(0)
if len(sys.argv) > 1:
    eng.set_kernel(int(sys.argv[1]))
for name in ["search_logreg_g1_4000x16", "search_logreg_digits3", "search_logreg_g1_20000x64"]:
    g = np.load(os.path.join(GOLD, name + ".npz"))
    X, y, cv = _case(name)
    fold = _fold_ids(y, cv)
    eng.stage_x(X); eng.stage_labels(y); eng.stage_folds(fold, cv)
    Cs = g["C"]; C = np.repeat(Cs, cv); cf = np.tile(np.arange(cv, dtype=np.int32), len(Cs))
    t0 = time.time()
    res = eng.logreg_fit_batch(C, cf, np.ones(len(C), np.int32))
    dt = time.time() - t0
    correct, count = eng.linear_score_batch(res["coef"], cf, np.ones(len(C), np.int32))
    gold = np.stack([g["split%d_test_score" % i] for i in range(cv)], 1).ravel()
    flips = correct - np.rint(gold * count)
    gc = g["coef"].reshape(len(C), -1)
    rel = np.abs(res["coef"] - gc).max(1) / np.abs(gc).max(1)
    print("==", name, "fit %.3fs gpu %.3fs" % (dt, res["gpu_seconds"]))
    for j in range(len(C)):
        print("  C=%-8g fold=%d n_iter gpu/gold %3d/%3d status %d evals %3d flips %+d/%d rel_coef %.2e" % (
            C[j], cf[j], res["n_iter"][j], g["n_iter"].ravel()[j], res["status"][j], res["n_evals"][j],
            flips[j], count[j], rel[j]))
