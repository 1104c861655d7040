"""bench.py's checker pieces on the CPU: the task sample of the CPU leg and the `parity` block that compares its
scores with cv_results_ at the same (candidate, fold) -- on an engine double whose fits are scikit-learn's, so the
block must report no difference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_cpu_tasks_are_whole_candidates_spread_over_the_grid():
    import bench
    tasks = bench.cpu_tasks(512, 5, 40)
    assert len(tasks) == 40
    cands = sorted({c for c, _ in tasks})
    assert cands[0] == 0 and cands[-1] == 511 and len(cands) == 8
    assert all(sorted(f for c, f in tasks if c == ci) == list(range(5)) for ci in cands)
    assert len(bench.cpu_tasks(512, 5, 3)) == 3


def test_parity_block_against_the_cpu_leg(fake_engine):
    import bench
    from sklearn.linear_model import LogisticRegression
    from skdist.distribute.search import DistGridSearchCV
    from skdist_b200.datasets import make_g1_classification
    X, y = make_g1_classification(3000, 12, seed=2)
    Cs = np.logspace(-3, 2, 6)
    fold = bench.fold_ids(y, 3)
    gs = DistGridSearchCV(LogisticRegression(), {"C": list(Cs)}, None, cv=3).fit(X, y)
    tasks = bench.cpu_tasks(len(Cs), 3, 9)
    fps, dt, scores, n_jobs, inner = bench.cpu_fits_per_sec(X, y, fold, Cs, tasks, n_jobs=1)
    assert fps > 0 and len(scores) == len(tasks) == 9
    blk = bench.parity_block(gs.cv_results_, tasks, scores, fold, Cs, 3)
    assert blk["n_compared"] == 9 and blk["max_flips_per_fold"] == 0 and blk["max_abs_dscore"] < 1e-12
    assert blk["best_C_equal_on_subgrid"] and blk["best_C_tied"]
    assert blk["test_rows_per_fold"] == 1000
    # a device result that differs by 3 test rows in one fold is reported as such
    bad = {k: np.array(v, dtype=float).copy() if k.startswith(("split", "mean_test")) else v for k, v in gs.cv_results_.items()}
    ci, f = tasks[4]
    bad["split%d_test_score" % f][ci] -= 3 / 1000
    assert bench.parity_block(bad, tasks, scores, fold, Cs, 3)["max_flips_per_fold"] == 3
