"""Config 5 of BASELINE.json: DistRandomizedSearchCV(Ridge) n_iter=2048 x 5 folds on synthetic
1M x 256 fp32 + batched inference; prints one JSON line (candidate-fits/s, predict rows/s) with a
sampled CPU baseline (the reference's per-task function on the host cores)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=1_000_000)
p.add_argument("--d", type=int, default=256)
p.add_argument("--n-iter", type=int, default=2048)
p.add_argument("--folds", type=int, default=5)
p.add_argument("--predict-rows", type=int, default=2_000_000)
p.add_argument("--cpu-sample", type=int, default=3)
a = p.parse_args()

from scipy.stats import loguniform
from sklearn.linear_model import Ridge
from sklearn.model_selection import KFold, ParameterSampler
from skdist.distribute.search import DistRandomizedSearchCV
from skdist_b200.datasets import make_g1_regression
from skdist_b200.engine import get_engine

X, y = make_g1_regression(a.n, a.d, seed=0)
dist = {"alpha": loguniform(1e-3, 1e3)}
eng = get_engine()
times = []
for rep in range(2):
    t0 = time.perf_counter()
    rs = DistRandomizedSearchCV(Ridge(), dist, None, n_iter=a.n_iter, cv=a.folds, random_state=0, refit=True).fit(X, y)
    times.append(time.perf_counter() - t0)
fits = a.n_iter * a.folds
# device-resident: fit + score calls only
cands = list(ParameterSampler(dist, a.n_iter, random_state=0))
alpha = np.repeat([c["alpha"] for c in cands], a.folds)
cf = np.tile(np.arange(a.folds, dtype=np.int32), a.n_iter)
eng.timer_start()
res = eng.ridge_fit_batch(alpha, cf)
sse, cnt = eng.linear_r2_batch(res["coef"], cf)
dev = eng.timer_stop()
# predict
Xp, _ = make_g1_regression(a.predict_rows, a.d, seed=1)
coef = np.r_[rs.best_estimator_.coef_, rs.best_estimator_.intercept_][None, :].astype(np.float32)
eng.predict_linear(Xp[:100000], coef)
t0 = time.perf_counter(); out = eng.predict_linear(Xp, coef); tp = time.perf_counter() - t0
eng.stage_x(Xp); eng.timer_start(); dec = eng.linear_decision(coef); td = eng.timer_stop()
line = {"workload": "DistRandomizedSearchCV(Ridge) n_iter=%d x %d folds, G1 %dx%d fp32" % (a.n_iter, a.folds, a.n, a.d),
        "fits_per_s_e2e": fits / times[-1], "e2e_seconds": times[-1], "fits_per_s_device": fits / dev,
        "device_seconds": dev, "best_alpha": rs.best_params_["alpha"], "best_r2": float(rs.best_score_),
        "predict_rows_per_s_e2e_host_arrays": a.predict_rows / tp,
        "predict_device_seconds_staged_incl_d2h": td,
        "predict_hbm_GBps_kernel_floor": a.predict_rows * a.d * 4 / td / 1e9}
if a.cpu_sample:
    from sklearn.metrics import check_scoring
    from oracle.search_oracle import fit_and_score
    sp = list(KFold(a.folds).split(X))
    t0 = time.time()
    for i in range(a.cpu_sample):
        fit_and_score(Ridge(), X, y, check_scoring(Ridge()), sp[i % a.folds][0], sp[i % a.folds][1], cands[i])
    dt = time.time() - t0
    line["cpu_baseline"] = {"value": a.cpu_sample / dt, "unit": "fits/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": "%d fits in %.1f s (reference per-task function, threaded BLAS)" % (a.cpu_sample, dt)}
print(json.dumps(line))
