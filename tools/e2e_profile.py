"""cProfile of one end-to-end DistGridSearchCV.fit on host arrays (where does the non-kernel time go)."""
import cProfile, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SKDIST_B200_TRACE", "1")
from sklearn.linear_model import LogisticRegression
from skdist.distribute.search import DistGridSearchCV
from skdist_b200.datasets import make_g1_classification
X, y = make_g1_classification(1000000, 256, seed=0)
Cs = np.logspace(-4, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 512)
for rep in range(2):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    gs = DistGridSearchCV(LogisticRegression(), {"C": list(Cs)}, None, cv=5, refit=False).fit(X, y)
    pr.disable()
    print("fit wall %.3f s" % (time.perf_counter() - t0), file=sys.stderr)
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(28)
