import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from sklearn.linear_model import SGDClassifier
from skdist.distribute.multiclass import DistOneVsRestClassifier
from skdist_b200.datasets import make_multiclass
X, y = make_multiclass(100000, 512, 148, seed=0)
DistOneVsRestClassifier(SGDClassifier(random_state=0, max_iter=3, tol=None), None).fit(X, y)
