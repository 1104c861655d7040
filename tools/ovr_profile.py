"""cProfile of DistOneVsRestClassifier(SGDClassifier).fit at config-3 size: host-side overheads."""
import cProfile, os, pstats, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from sklearn.linear_model import SGDClassifier
from skdist.distribute.multiclass import DistOneVsRestClassifier
from skdist_b200.datasets import make_multiclass
X, y = make_multiclass(500000, 512, 1000, seed=0)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
ovr = DistOneVsRestClassifier(SGDClassifier(random_state=0), None).fit(X, y)
pr.disable()
print("fit wall %.3f s" % (time.perf_counter() - t0), file=sys.stderr)
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(22)
