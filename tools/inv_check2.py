"""Batch invariance of the tensor-core objective / gradient evaluation at a fixed point: a column's loss
and gradient inside a 160-group batch vs alone, group by group (diagnostic for multi-group CTAs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sklearn.model_selection import StratifiedKFold
from skdist_b200.datasets import make_g1_classification
from skdist_b200.engine import Engine
e = Engine(0); e.set_kernel(2)
X, y = make_g1_classification(6000, 32, seed=77)
fold = np.zeros(len(y), np.int8)
for k, (_, te) in enumerate(StratifiedKFold(4).split(X, y)):
    fold[te] = k
e.stage_x(X); e.stage_labels(y.astype(np.int32)); e.stage_folds(fold, 4)
rng = np.random.default_rng(0)
B = 160 * 128
w = (rng.standard_normal((B, 33)) * 0.1)
C = np.ones(B); cf = (np.arange(B) % 4).astype(np.int32); pos = np.ones(B, np.int32)
lb, gb = e.logreg_loss_grad(w, C, cf, pos)
bad = []
for j0 in (0, 128 * 37, 128 * 40, 128 * 41, 128 * 100, 128 * 159):
    sl = slice(j0, j0 + 128)
    ls, gs = e.logreg_loss_grad(w[sl], C[sl], cf[sl], pos[sl])
    print("group", j0 // 128, "loss equal", np.array_equal(ls, lb[sl]), "grad equal", np.array_equal(gs, gb[sl]),
          "max grad diff", np.abs(gs - gb[sl]).max(), "rows differing", int((np.abs(gs - gb[sl]).max(1) > 0).sum()))
