"""Host-side handle on one GPU's staged dataset + batched solvers (ctypes over the C-ABI)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, ptr


class Engine:
    """One context per process / GPU.  Replaces the reference's SparkContext
    (ref skdist/distribute/search.py:309-313, `sc`): data is staged once in
    HBM instead of being shipped with every task."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        check(self._lib.skd_ctx_create(int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self.n = self.d = 0

    # -- lifetime -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.skd_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- staging ----------------------------------------------------------------
    def stage_x(self, X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        n, d = X.shape
        check(self._lib.skd_stage_x(self._h, ptr(X), n, d, d), self._h)
        self.n, self.d = n, d

    def stage_x_device(self, dev_ptr, n, d, ldx=None):
        check(self._lib.skd_stage_x_device(self._h, ctypes.c_void_p(int(dev_ptr)), n, d, ldx or d), self._h)
        self.n, self.d = n, d

    def stage_labels(self, y_class):
        y = np.ascontiguousarray(y_class, dtype=np.int32)
        check(self._lib.skd_stage_labels(self._h, ptr(y), y.shape[0]), self._h)

    def stage_targets(self, y):
        y = np.ascontiguousarray(y, dtype=np.float32)
        check(self._lib.skd_stage_targets(self._h, ptr(y), y.shape[0]), self._h)

    def stage_folds(self, fold_id, n_folds):
        if fold_id is None:
            check(self._lib.skd_stage_folds(self._h, None, 0, 0), self._h)
            return
        f = np.ascontiguousarray(fold_id, dtype=np.int8)
        check(self._lib.skd_stage_folds(self._h, ptr(f), f.shape[0], int(n_folds)), self._h)

    def set_kernel(self, which):
        return self._lib.skd_set_kernel(self._h, int(which))

    def counters(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(self._lib.skd_get_counters(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), self._h)
        return {"launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def profile(self, enable=-1):
        """Read (and with enable in {0,1} reset + switch) the evaluation-kernel timers."""
        ms, fl = ctypes.c_double(), ctypes.c_double()
        nl, nr = ctypes.c_int64(), ctypes.c_int64()
        check(self._lib.skd_profile(self._h, int(enable), ctypes.byref(ms), ctypes.byref(fl),
                                    ctypes.byref(nl), ctypes.byref(nr)), self._h)
        return {"eval_ms": ms.value, "eval_flops": fl.value, "eval_launches": nl.value,
                "rounds": nr.value}

    def timer_start(self):
        check(self._lib.skd_timer_start(self._h), self._h)

    def timer_stop(self):
        """Elapsed device time (seconds) on the library's stream since timer_start()."""
        ms = ctypes.c_double()
        check(self._lib.skd_timer_stop(self._h, ctypes.byref(ms)), self._h)
        return ms.value * 1e-3

    # -- solvers ------------------------------------------------------------------
    def logreg_fit_batch(self, C, col_fold, col_pos, fit_intercept=True, tol=1e-4, max_iter=100):
        C = np.ascontiguousarray(C, dtype=np.float64)
        B = C.shape[0]
        col_fold = np.ascontiguousarray(col_fold, dtype=np.int32)
        col_pos = np.ascontiguousarray(col_pos, dtype=np.int32)
        assert col_fold.shape == (B,) and col_pos.shape == (B,)
        coef = np.empty((B, self.d + 1), dtype=np.float32)
        n_iter = np.empty(B, dtype=np.int32)
        status = np.empty(B, dtype=np.int32)
        loss = np.empty(B, dtype=np.float64)
        n_evals = np.empty(B, dtype=np.int32)
        secs = ctypes.c_double(0.0)
        check(self._lib.skd_logreg_fit_batch(
            self._h, B, ptr(C), ptr(col_fold), ptr(col_pos), int(bool(fit_intercept)), float(tol),
            int(max_iter), ptr(coef), ptr(n_iter), ptr(status), ptr(loss), ptr(n_evals),
            ctypes.byref(secs)), self._h)
        return {"coef": coef, "n_iter": n_iter, "status": status, "loss": loss,
                "n_evals": n_evals, "gpu_seconds": secs.value}

    def logreg_loss_grad(self, w, C, col_fold, col_pos, fit_intercept=True):
        w = np.ascontiguousarray(w, dtype=np.float64)
        B = w.shape[0]
        assert w.shape[1] == self.d + 1
        C = np.ascontiguousarray(C, dtype=np.float64)
        col_fold = np.ascontiguousarray(col_fold, dtype=np.int32)
        col_pos = np.ascontiguousarray(col_pos, dtype=np.int32)
        loss = np.empty(B, dtype=np.float64)
        grad = np.empty((B, self.d + 1), dtype=np.float64)
        check(self._lib.skd_logreg_loss_grad(self._h, B, ptr(w), ptr(C), ptr(col_fold), ptr(col_pos),
                                             int(bool(fit_intercept)), ptr(loss), ptr(grad)), self._h)
        return loss, grad

    def linear_score_batch(self, coef, col_fold, col_pos):
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        B = coef.shape[0]
        assert coef.shape[1] == self.d + 1
        col_fold = np.ascontiguousarray(col_fold, dtype=np.int32)
        col_pos = np.ascontiguousarray(col_pos, dtype=np.int32)
        correct = np.empty(B, dtype=np.int64)
        count = np.empty(B, dtype=np.int64)
        check(self._lib.skd_linear_score_batch(self._h, B, ptr(coef), ptr(col_fold), ptr(col_pos),
                                               ptr(correct), ptr(count)), self._h)
        return correct, count

    def linear_decision(self, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        B = coef.shape[0]
        out = np.empty((self.n, B), dtype=np.float32)
        check(self._lib.skd_linear_decision(self._h, B, ptr(coef), ptr(out)), self._h)
        return out


# -- per-process singleton -------------------------------------------------------------------
_ENGINE = None
_FACTORY = None


def set_engine_factory(factory):
    """Test hook: replace the object returned by get_engine() (e.g. with an oracle-backed
    stand-in so host logic can be exercised without a GPU).  Pass None to restore."""
    global _FACTORY, _ENGINE
    _FACTORY = factory
    _ENGINE = None


def get_engine():
    """The process-wide Engine, bound to cuda:LOCAL_RANK (one process per GPU)."""
    global _ENGINE
    if _ENGINE is None:
        if _FACTORY is not None:
            _ENGINE = _FACTORY()
        else:
            import os
            _ENGINE = Engine(int(os.environ.get("LOCAL_RANK", 0)))
    return _ENGINE


def reset_engine():
    global _ENGINE
    if _ENGINE is not None and hasattr(_ENGINE, "close"):
        _ENGINE.close()
    _ENGINE = None
