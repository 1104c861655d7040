"""Run under torchrun with N >= 2 GPUs: every sharded entry point must give, on every rank, the
result scikit-learn gives in one process (bit-identical where the single-GPU path is)."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world)
from sklearn.ensemble import RandomForestClassifier
from sklearn.linear_model import LogisticRegression, Ridge, SGDClassifier
from sklearn.multiclass import OneVsOneClassifier, OneVsRestClassifier
from skdist.distribute.ensemble import DistRandomForestClassifier
from skdist.distribute.multiclass import DistOneVsOneClassifier, DistOneVsRestClassifier
from skdist.distribute.search import DistGridSearchCV, DistRandomizedSearchCV
from skdist_b200.datasets import make_g1_regression, make_multiclass
X, y = make_multiclass(4000, 24, 5, seed=3)
ovr = DistOneVsRestClassifier(SGDClassifier(random_state=0), None).fit(X, y)
ref = OneVsRestClassifier(SGDClassifier(random_state=0)).fit(X, y)
assert all(np.array_equal(a.coef_, b.coef_) for a, b in zip(ovr.estimators_, ref.estimators_)), "ovr sgd"
ovo = DistOneVsOneClassifier(LogisticRegression(C=0.05), None).fit(X, y)
refo = OneVsOneClassifier(LogisticRegression(C=0.05)).fit(X, y)
assert np.mean(ovo.predict(X) == refo.predict(X)) > 0.999, "ovo"
Xq = np.round(X * 16).astype(np.float32)
rf = DistRandomForestClassifier(n_estimators=7, random_state=4).fit(Xq, y)
rr = RandomForestClassifier(n_estimators=7, random_state=4).fit(Xq, y)
assert all(np.array_equal(a.tree_.threshold, b.tree_.threshold) for a, b in zip(rf.estimators_, rr.estimators_)), "forest"
np.testing.assert_array_equal(rf.predict_proba(Xq[:200]), rr.predict_proba(Xq[:200]))
Xr, yr = make_g1_regression(5000, 16, seed=5)
rs = DistRandomizedSearchCV(Ridge(), {"alpha": list(np.logspace(-2, 2, 30))}, None, n_iter=11, random_state=0, cv=3).fit(Xr, yr)
from sklearn.model_selection import RandomizedSearchCV
rref = RandomizedSearchCV(Ridge(), {"alpha": list(np.logspace(-2, 2, 30))}, n_iter=11, random_state=0, cv=3).fit(Xr, yr)
np.testing.assert_allclose(rs.cv_results_["mean_test_score"], rref.cv_results_["mean_test_score"], atol=2e-4)
assert rs.best_params_ == rref.best_params_, "ridge"
# multiclass target -> multinomial problems dealt over the ranks; X replicated by one NVLink broadcast
from sklearn.datasets import load_digits
from sklearn.model_selection import GridSearchCV
dg = load_digits()
Xd, yd = (dg.data / 16).astype(np.float32), dg.target
gm = DistGridSearchCV(LogisticRegression(max_iter=300), {"C": [0.01, 0.1, 1.0, 10.0]}, None, cv=3).fit(Xd, yd)
gk = GridSearchCV(LogisticRegression(max_iter=300), {"C": [0.01, 0.1, 1.0, 10.0]}, cv=3).fit(Xd, yd)
np.testing.assert_allclose(gm.cv_results_["mean_test_score"], gk.cv_results_["mean_test_score"], atol=4.01 / len(yd))
assert gk.cv_results_["mean_test_score"][gm.best_index_] >= gk.best_score_ - 4.01 / len(yd), "multinomial"
# a matrix rank 0 cannot stage (NaN) must raise on every rank, not hang the broadcast
Xbad = Xd.copy(); Xbad[3, 5] = np.nan
try:
    DistGridSearchCV(LogisticRegression(), {"C": [1.0]}, None, cv=3).fit(Xbad, yd)
    raise SystemExit("NaN input was accepted")
except ValueError:
    pass
# round 2: overlapping test sets (several fold-id layouts), multilabel one-vs-rest (row bit matrices), both
# over ranks; X staged in slices + in-place all-gather
from sklearn.model_selection import ShuffleSplit
from skdist_b200.datasets import make_g1_classification
Xc, yc = make_g1_classification(6000, 20, seed=8)
cvs = ShuffleSplit(n_splits=3, test_size=0.25, random_state=1)
gsh = DistGridSearchCV(LogisticRegression(), {"C": [0.01, 0.1]}, None, cv=cvs).fit(Xc, yc)
gsk = GridSearchCV(LogisticRegression(), {"C": [0.01, 0.1]}, cv=cvs).fit(Xc, yc)
np.testing.assert_allclose(gsh.cv_results_["mean_test_score"], gsk.cv_results_["mean_test_score"], atol=2.01 / 1500)
Yml = np.stack([Xc[:, 0] > 0.3, Xc[:, 1] + Xc[:, 2] > 0.1, yc == 1], axis=1).astype(int)
ml = DistOneVsRestClassifier(LogisticRegression(C=0.05), None).fit(Xc, Yml)
for k, e in enumerate(ml.estimators_):
    r = LogisticRegression(C=0.05).fit(Xc, Yml[:, k])
    assert np.mean(e.predict(Xc) == r.predict(Xc)) > 0.998, "multilabel ovr"
gathered = [None] * world
dist.all_gather_object(gathered, float(rs.cv_results_["mean_test_score"].sum()))
assert len(set(gathered)) == 1, "ranks disagree"
print("rank %d/%d: multi-GPU checks passed" % (rank, world))
dist.destroy_process_group()
