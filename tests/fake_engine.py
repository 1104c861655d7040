"""Oracle-backed stand-in for skdist_b200.engine.Engine -- TEST DOUBLE, CPU only.

Lets the `-m "not gpu"` suite exercise the host-side logic (cv_results_ assembly, column
sharding/gather, refit, pickling) without a GPU.  It is never importable from the product
package; tests install it through engine.set_engine_factory()."""
import numpy as np

from oracle import logreg_oracle as lo


class FakeEngine:
    def __init__(self, device=0):
        self.n = self.d = 0
        self.calls = []

    def close(self):
        pass

    def stage_x(self, X):
        self.X = np.ascontiguousarray(X, dtype=np.float32)
        self.n, self.d = self.X.shape

    def stage_labels(self, y):
        self.y = np.asarray(y, dtype=np.int32)

    def stage_targets(self, y):
        self.yr = np.asarray(y, dtype=np.float32)

    def stage_folds(self, fold, n_folds):
        self.fold = None if fold is None else np.asarray(fold, dtype=np.int8)

    def _train_mask(self, f):
        return np.ones(self.n, bool) if f < 0 else self.fold != f

    def logreg_fit_batch(self, C, col_fold, col_pos, fit_intercept=True, tol=1e-4, max_iter=100):
        B = len(C)
        self.calls.append(("fit", B))
        coef = np.zeros((B, self.d + 1), np.float32)
        n_iter = np.zeros(B, np.int32)
        for j in range(B):
            m = self._train_mask(int(col_fold[j]))
            y01 = (self.y[m] == col_pos[j]).astype(np.float32)
            w, b, it = lo.fit_binary_lbfgs(self.X[m], y01, C=float(C[j]), tol=tol, max_iter=max_iter,
                                           fit_intercept=fit_intercept)
            coef[j, :self.d] = w
            coef[j, self.d] = b
            n_iter[j] = it
        return {"coef": coef, "n_iter": n_iter, "status": np.ones(B, np.int32),
                "loss": np.zeros(B), "n_evals": n_iter + 1, "gpu_seconds": 0.0}

    def _rows(self, code):
        if code == -2:
            return np.ones(self.n, bool)
        if code >= 0:
            return self.fold == code
        return self.fold != (-3 - code)

    def linear_score_batch(self, coef, col_fold, col_pos):
        B = coef.shape[0]
        correct = np.zeros(B, np.int64)
        count = np.zeros(B, np.int64)
        for j in range(B):
            m = self._rows(int(col_fold[j]))
            z = self.X[m] @ coef[j, :self.d] + coef[j, self.d]
            correct[j] = np.sum((z > 0) == (self.y[m] == col_pos[j]))
            count[j] = m.sum()
        return correct, count

    def linear_decision(self, coef):
        return (self.X @ coef[:, :self.d].T + coef[:, self.d][None, :]).astype(np.float32)
