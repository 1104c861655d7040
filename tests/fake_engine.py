"""Oracle-backed stand-in for skdist_b200.engine.Engine -- TEST DOUBLE, CPU only.

Lets the `-m "not gpu"` suite exercise the host-side logic (cv_results_ assembly, column
sharding/gather, refit, pickling) without a GPU.  It is never importable from the product
package; tests install it through engine.set_engine_factory()."""
import numpy as np

from oracle import logreg_oracle as lo
from oracle import ridge_oracle as ro
from oracle import sgd_oracle as so


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

    def stage_column_masks(self, mask):
        self._fmask = None if mask is None else np.asarray(mask, dtype=bool)

    def stage_row_bits(self, labels=None, train=None):
        self._ybits = None if labels is None else np.asarray(labels, dtype=bool)
        self._mbits = None if train is None else np.asarray(train, dtype=bool)

    def logreg_fit_batch(self, C, col_fold, col_pos, fit_intercept=True, tol=1e-4, max_iter=100, col_neg=None):
        B = len(C)
        fmask = getattr(self, "_fmask", None)
        self._fmask = None
        ybits, mbits = getattr(self, "_ybits", None), getattr(self, "_mbits", None)
        self._ybits = self._mbits = None
        self.calls.append(("fit", B))
        coef = np.zeros((B, self.d + 1), np.float32)
        n_iter = np.zeros(B, np.int32)
        for j in range(B):
            m = self._train_mask(int(col_fold[j]))
            if col_neg is not None and col_neg[j] >= 0:
                m = m & ((self.y == col_pos[j]) | (self.y == col_neg[j]))
            if mbits is not None:
                m = m & mbits[j]
            y01 = (self.y[m] == col_pos[j]).astype(np.float32) if ybits is None else ybits[j][m].astype(np.float32)
            keep = np.arange(self.d) if fmask is None else np.flatnonzero(fmask[j])
            Xm = self.X[m] if fmask is None else np.ascontiguousarray(self.X[m][:, keep])
            w, b, it = lo.fit_binary_lbfgs(Xm, y01, C=float(C[j]), tol=tol, max_iter=max_iter,
                                           fit_intercept=fit_intercept)
            coef[j, keep] = w
            coef[j, self.d] = b
            n_iter[j] = it
        return {"coef": coef, "n_iter": n_iter, "status": np.ones(B, np.int32),
                "loss": np.zeros(B), "n_evals": n_iter + 1, "gpu_seconds": 0.0}

    def logreg_multinomial_fit_batch(self, C, col_fold, n_classes, fit_intercept=True, tol=1e-4, max_iter=100):
        B = len(C)
        fmask = getattr(self, "_fmask", None)
        self._fmask = None
        self.calls.append(("fit_multinomial", B))
        coef = np.zeros((B, n_classes, self.d + 1), np.float32)
        n_iter = np.zeros(B, np.int32)
        for j in range(B):
            m = self._train_mask(int(col_fold[j]))
            keep = np.arange(self.d) if fmask is None else np.flatnonzero(fmask[j])
            Xm = self.X[m] if fmask is None else np.ascontiguousarray(self.X[m][:, keep])
            W, b, it = lo.fit_multinomial_lbfgs(Xm, self.y[m], n_classes, C=float(C[j]), tol=tol,
                                                max_iter=max_iter, fit_intercept=fit_intercept)
            coef[j][:, keep] = W
            coef[j, :, self.d] = b
            n_iter[j] = it
        return {"coef": coef, "n_iter": n_iter, "status": np.ones(B, np.int32),
                "loss": np.zeros(B), "n_evals": n_iter + 1, "gpu_seconds": 0.0}

    def multinomial_score_batch(self, coef, col_fold):
        B = coef.shape[0]
        correct = np.zeros(B, np.int64)
        count = np.zeros(B, np.int64)
        for j in range(B):
            m = self._rows(int(col_fold[j]))
            z = self.X[m] @ coef[j, :, :self.d].T + coef[j, :, self.d]
            correct[j] = np.sum(z.argmax(axis=1) == self.y[m])
            count[j] = m.sum()
        return correct, count

    def multinomial_confusion_batch(self, coef, col_fold):
        B, K = coef.shape[0], coef.shape[1]
        conf = np.zeros((B, K, K), np.int64)
        for j in range(B):
            m = self._rows(int(col_fold[j]))
            z = self.X[m] @ coef[j, :, :self.d].T + coef[j, :, self.d]
            np.add.at(conf[j], (self.y[m], z.argmax(axis=1)), 1)
        return conf

    def linear_auc_batch(self, coef, col_fold, col_pos):
        from sklearn.metrics import roc_auc_score
        B = coef.shape[0]
        auc = np.full(B, np.nan)
        count = np.zeros(B, np.int64)
        for j in range(B):
            m = self._rows(int(col_fold[j]))
            z = self.X[m] @ coef[j, :self.d] + coef[j, self.d]
            yb = self.y[m] == col_pos[j]
            count[j] = m.sum()
            if 0 < yb.sum() < len(yb):
                auc[j] = roc_auc_score(yb, z)
        return auc, count

    def linear_logloss_batch(self, coef, col_fold, col_pos=None):
        from sklearn.metrics import log_loss
        from sklearn.utils.extmath import softmax
        B = coef.shape[0]
        loss = np.zeros(B)
        count = np.zeros(B, np.int64)
        for j in range(B):
            m = self._rows(int(col_fold[j]))
            if coef.ndim == 2:
                z = self.X[m] @ coef[j, :self.d] + coef[j, self.d]
                from scipy.special import expit
                p1 = expit(z)
                proba = np.c_[1 - p1, p1]
                yt = (self.y[m] == col_pos[j]).astype(int)
                labels = [0, 1]
            else:
                proba = softmax(self.X[m] @ coef[j, :, :self.d].T + coef[j, :, self.d])
                yt = self.y[m]
                labels = list(range(coef.shape[1]))
            loss[j] = log_loss(yt, proba, labels=labels)
            count[j] = m.sum()
        return loss, count

    def forest_fit(self, sample_counts, rand_states, n_classes, max_features, max_depth, min_samples_split,
                   min_samples_leaf, min_weight_leaf, min_impurity_decrease, splitter=0, y_regression=None):
        """Trees by scikit-learn itself.  The device receives the splitter seed (rand_r) of every tree; the
        double maps it back to the tree seed through `seed_of_rand_r` (filled by the test from the seeds the
        forest draws) and checks that the bootstrap multiplicities it was handed are numpy's for that seed."""
        from sklearn.tree import DecisionTreeClassifier, DecisionTreeRegressor, ExtraTreeClassifier, ExtraTreeRegressor
        from sklearn.utils import check_random_state
        self.calls.append(("forest_fit", len(rand_states)))
        self.last_forest_seconds = 0.0
        out = []
        for t, r in enumerate(rand_states):
            seed = self.seed_of_rand_r[int(r)]
            if sample_counts is not None:
                want = np.bincount(check_random_state(seed).randint(0, self.n, self.n), minlength=self.n)
                assert np.array_equal(sample_counts[t], want), "bootstrap multiplicities differ from numpy's"
            reg = y_regression is not None
            cls = ({0: DecisionTreeRegressor, 1: ExtraTreeRegressor} if reg else
                   {0: DecisionTreeClassifier, 1: ExtraTreeClassifier})[splitter]
            est = cls(max_features=max_features, max_depth=None if max_depth >= 2 ** 31 - 1 else max_depth,
                      min_samples_split=min_samples_split, min_samples_leaf=min_samples_leaf,
                      min_impurity_decrease=min_impurity_decrease, random_state=seed)
            sw = None if sample_counts is None else sample_counts[t].astype(np.float64)
            est.fit(self.X, y_regression if reg else self.y, sample_weight=sw)
            tr = est.tree_
            out.append({"left": tr.children_left.astype(np.int32), "right": tr.children_right.astype(np.int32),
                        "feature": tr.feature.astype(np.int32), "threshold": tr.threshold.copy(),
                        "impurity": tr.impurity.copy(), "n_node_samples": tr.n_node_samples.astype(np.int32),
                        "weighted_n_node_samples": tr.weighted_n_node_samples.copy(),
                        "missing_go_to_left": np.zeros(tr.node_count, np.uint8),
                        "value": tr.value[:, 0, :].copy(), "max_depth": tr.max_depth})
        return out

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

    def sgd_fit_batch(self, est, col_pos):
        p = est.get_params(deep=False)
        B = len(col_pos)
        coef = np.zeros((B, self.d + 1))
        n_iter = np.zeros(B, np.int32)
        t = np.zeros(B)
        for j in range(B):
            ypm = np.where(self.y == col_pos[j], 1, -1)
            w, b, it, tt = so.fit_binary_sgd(self.X, ypm, loss=p["loss"], alpha=p["alpha"],
                                             fit_intercept=p["fit_intercept"], max_iter=p["max_iter"],
                                             tol=-np.inf if p["tol"] is None else p["tol"], shuffle=p["shuffle"],
                                             random_state=p["random_state"], n_iter_no_change=p["n_iter_no_change"])
            coef[j, :self.d] = w
            coef[j, self.d] = b
            n_iter[j] = it
            t[j] = tt
        return {"coef": coef, "n_iter": n_iter, "t": t, "status": np.ones(B, np.int32), "gpu_seconds": 0.0}

    def ridge_fit_batch(self, alpha, col_fold, fit_intercept=True):
        B = len(alpha)
        self.calls.append(("ridge", B))
        coef = np.zeros((B, self.d + 1), np.float32)
        for j in range(B):
            m = self._train_mask(int(col_fold[j]))
            w, b = ro.fit_ridge(self.X[m], self.yr[m], float(alpha[j]), fit_intercept)
            coef[j, :self.d] = w
            coef[j, self.d] = b
        return {"coef": coef, "status": np.ones(B, np.int32), "gpu_seconds": 0.0}

    def linear_r2_batch(self, coef, col_fold):
        B = coef.shape[0]
        sse = np.zeros(B)
        count = np.zeros(B, np.int64)
        for j in range(B):
            m = self._rows(int(col_fold[j]))
            pred = self.X[m] @ coef[j, :self.d] + coef[j, self.d]
            sse[j] = np.sum((self.yr[m].astype(np.float64) - pred.astype(np.float64)) ** 2)
            count[j] = m.sum()
        return sse, count

    def linear_decision(self, coef):
        return (self.X @ coef[:, :self.d].T + coef[:, self.d][None, :]).astype(np.float32)

    def predict_linear(self, Xnew, coef):
        Xnew = np.asarray(Xnew, dtype=np.float32)
        coef = np.asarray(coef, dtype=np.float32)
        return (Xnew @ coef[:, :-1].T + coef[:, -1]).astype(np.float32)

    def forest_predict(self, Xnew, tree_offset, left, right, feature, threshold, value):
        """numpy walk of the concatenated tree arrays (test double of skd_forest_predict)."""
        Xnew = np.asarray(Xnew, dtype=np.float32)
        m = Xnew.shape[0]
        out = np.zeros((m, value.shape[1]))
        n_trees = len(tree_offset) - 1
        for t in range(n_trees):
            base = int(tree_offset[t])
            node = np.zeros(m, dtype=np.int64)
            active = np.ones(m, dtype=bool)
            while True:
                k = base + node
                active = left[k] != -1
                if not active.any():
                    break
                go_left = Xnew[np.arange(m), feature[k]].astype(np.float64) <= threshold[k]
                nxt = np.where(go_left, left[k], right[k])
                node = np.where(active, nxt, node)
            out += value[base + node]
        return out / n_trees
