"""Multinomial search throughput: BASELINE config 1 (digits, 4 C x 3 folds, 10 classes) and a larger
synthetic multiclass case; one JSON line with a sampled CPU baseline (scikit-learn's own fit, the
arithmetic the reference's per-task function runs)."""
import argparse, json, os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=200_000)
p.add_argument("--d", type=int, default=128)
p.add_argument("--classes", type=int, default=10)
p.add_argument("--cands", type=int, default=32)
p.add_argument("--folds", type=int, default=5)
p.add_argument("--cpu-sample", type=int, default=2)
a = p.parse_args()

from sklearn.datasets import load_digits
from sklearn.linear_model import LogisticRegression
from sklearn.model_selection import StratifiedKFold
from skdist.distribute.search import DistGridSearchCV
from skdist_b200.datasets import make_multiclass

line = {}
dg = load_digits()
grid = {"C": [0.01, 0.1, 1.0, 10.0]}
for rep in range(2):
    t0 = time.perf_counter()
    gs = DistGridSearchCV(LogisticRegression(max_iter=200), grid, None, cv=3, refit=False).fit(dg.data, dg.target)
    t1 = time.perf_counter() - t0
t0 = time.perf_counter()
from sklearn.model_selection import GridSearchCV
sk = GridSearchCV(LogisticRegression(max_iter=200), grid, cv=3, refit=False).fit(dg.data, dg.target)
t2 = time.perf_counter() - t0
line["config1_digits"] = {"fits": 12, "device_path_seconds": t1, "fits_per_s": 12 / t1,
                          "sklearn_cpu_seconds": t2, "mean_test_score": list(map(float, gs.cv_results_["mean_test_score"])),
                          "sklearn_mean_test_score": list(map(float, sk.cv_results_["mean_test_score"]))}

X, y = make_multiclass(a.n, a.d, a.classes, seed=0)
Cs = list(np.logspace(-3, 3, a.cands))
for rep in range(2):
    t0 = time.perf_counter()
    gs = DistGridSearchCV(LogisticRegression(), {"C": Cs}, None, cv=a.folds, refit=False).fit(X, y)
    t1 = time.perf_counter() - t0
fits = a.cands * a.folds
line["synthetic"] = {"workload": "DistGridSearchCV(LogisticRegression) %d C x %d folds, %d classes, %dx%d fp32"
                     % (a.cands, a.folds, a.classes, a.n, a.d), "fits": fits, "e2e_seconds": t1,
                     "fits_per_s_e2e": fits / t1, "best_C": float(gs.best_params_["C"]),
                     "best_score": float(gs.best_score_)}
if a.cpu_sample:
    sp = list(StratifiedKFold(a.folds).split(X, y))
    t0 = time.time()
    scores = []
    for i in range(a.cpu_sample):
        tr, te = sp[i % a.folds]
        est = LogisticRegression(C=Cs[(i * 7) % a.cands]).fit(X[tr], y[tr])
        scores.append(est.score(X[te], y[te]))
    dt = time.time() - t0
    line["synthetic"]["cpu_baseline"] = {"value": a.cpu_sample / dt, "unit": "fits/s", "cores": os.cpu_count(),
                                         "kind": "port", "sample": "%d fits in %.1f s (threaded BLAS)" % (a.cpu_sample, dt)}
print(json.dumps(line))
