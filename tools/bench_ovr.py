"""Config 3 of BASELINE.json: DistOneVsRestClassifier(SGDClassifier) with K classes on synthetic
n x d fp32; prints one JSON line (label-fits/s) with a sampled CPU baseline."""
import argparse, json, os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
p = argparse.ArgumentParser()
p.add_argument("--n", type=int, default=500_000)
p.add_argument("--d", type=int, default=512)
p.add_argument("--k", type=int, default=1000)
p.add_argument("--cpu-sample", type=int, default=1)
a = p.parse_args()
from sklearn.linear_model import SGDClassifier
from skdist.distribute.multiclass import DistOneVsRestClassifier
from skdist_b200.datasets import make_multiclass
from skdist_b200.engine import get_engine
warnings.simplefilter("ignore")
X, y = make_multiclass(a.n, a.d, a.k, seed=0)
get_engine()          # CUDA context / library load: one-time process start-up, not part of a fit
t0 = time.perf_counter()
ovr = DistOneVsRestClassifier(SGDClassifier(random_state=0), None).fit(X, y)
dt = time.perf_counter() - t0
n_iter = np.array([e.n_iter_ for e in ovr.estimators_])
line = {"workload": "DistOneVsRestClassifier(SGDClassifier(random_state=0)) K=%d, %dx%d fp32" % (a.k, a.n, a.d),
        "label_fits_per_s_e2e": a.k / dt, "seconds": dt, "epochs_min_mean_max": [int(n_iter.min()), float(n_iter.mean()), int(n_iter.max())],
        "class_epochs_per_s": float(n_iter.sum() / dt), "counters": get_engine().counters()}
if a.cpu_sample:
    t0 = time.time()
    for k in range(a.cpu_sample):
        m = SGDClassifier(random_state=0).fit(X, (y == k).astype(int))
        same = np.array_equal(m.coef_, ovr.estimators_[k].coef_)
    dtc = time.time() - t0
    line["cpu_baseline"] = {"value": a.cpu_sample / dtc, "unit": "label-fits/s", "cores": 1, "kind": "port",
                            "sample": "%d label column(s) in %.1f s (SGD is single-threaded per column)" % (a.cpu_sample, dtc),
                            "bit_identical_to_gpu": bool(same)}
print(json.dumps(line))
