"""Smallest end-to-end check of the search host path on a GPU (a few seconds): binary and multiclass
DistGridSearchCV against scikit-learn's GridSearchCV."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from sklearn.linear_model import LogisticRegression
from sklearn.model_selection import GridSearchCV
from skdist.distribute.search import DistGridSearchCV
from skdist_b200.datasets import make_g1_classification, make_multiclass
X, y = make_g1_classification(4096, 32, seed=0)
grid = {"C": [0.01, 0.1, 1.0]}
for scoring in (None, "roc_auc"):
    gs = DistGridSearchCV(LogisticRegression(), grid, None, cv=3, scoring=scoring).fit(X, y)
    sk = GridSearchCV(LogisticRegression(), grid, cv=3, scoring=scoring).fit(X, y)
    d = np.abs(gs.cv_results_["mean_test_score"] - sk.cv_results_["mean_test_score"]).max()
    assert d <= 2.0 / 1365 and gs.best_params_ == sk.best_params_, (scoring, d)
Xm, ym = make_multiclass(3000, 16, 4, seed=1)
gm = DistGridSearchCV(LogisticRegression(max_iter=200), grid, None, cv=3, scoring="f1_macro").fit(Xm, ym)
sm = GridSearchCV(LogisticRegression(max_iter=200), grid, cv=3, scoring="f1_macro").fit(Xm, ym)
assert np.abs(gm.cv_results_["mean_test_score"] - sm.cv_results_["mean_test_score"]).max() <= 5e-3
print("quick check ok")
