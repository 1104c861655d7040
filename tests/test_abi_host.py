"""C-ABI: the shared library loads, exports every symbol include/skdist_b200.h declares,
fails loudly without a GPU, and its host-side L-BFGS-B core tracks scipy's setulb."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy import optimize

from oracle import logreg_oracle as lo
from skdist_b200 import _lib
from skdist_b200.datasets import make_g1_classification

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "skdist_b200.h")).read()
    declared = set(re.findall(r"\b(skd_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"skd_ctx", "skd_lbfgs"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.SYMBOLS) == declared
    assert lib.skd_version() >= 100


def test_no_gpu_fails_loudly():
    lib = _lib.load()
    if lib.skd_device_count() > 0:
        pytest.skip("GPU present")
    from skdist_b200.engine import Engine
    with pytest.raises(_lib.SkdError, match="no CUDA device"):
        Engine(0)


def _mine(func, n, maxiter=100, tol=1e-4):
    lib = _lib.load()
    h = lib.skd_lbfgs_create(n, 10, maxiter, 50, tol, 64 * np.finfo(float).eps)
    x = np.ctypeslib.as_array(lib.skd_lbfgs_x(h), (n,))
    g = np.ctypeslib.as_array(lib.skd_lbfgs_g(h), (n,))
    xs = []
    while True:
        f, gg = func(x.copy())
        g[:] = gg
        xs.append(x.copy())
        st = lib.skd_lbfgs_advance(h, f)
        if st != 0:
            break
    out = x.copy(), lib.skd_lbfgs_nit(h), lib.skd_lbfgs_nfev(h), st, xs
    lib.skd_lbfgs_free(h)
    return out


@pytest.mark.parametrize("C", [1e-4, 1e-2, 1.0])
def test_lbfgs_core_tracks_scipy_setulb(C):
    """Same f,g callback driven through scipy's L-BFGS-B and through csrc/lbfgs_core.h:
    identical iteration / evaluation counts and iterates equal to ~1e-12."""
    X, y = make_g1_classification(5000, 20, seed=11)
    yf = y.astype(np.float32)
    n, d = X.shape
    l2 = 1.0 / (C * n)
    ref_x = []

    def func(w):
        ref_x.append(w.copy())
        return lo.loss_gradient(w, X, yf, l2, True)

    res = optimize.minimize(func, np.zeros(d + 1), method="L-BFGS-B", jac=True,
                            options={"maxiter": 100, "maxls": 50, "gtol": 1e-4,
                                     "ftol": 64 * np.finfo(float).eps})
    xm, nit, nfev, st, xs = _mine(lambda w: lo.loss_gradient(w, X, yf, l2, True), d + 1)
    assert (nit, nfev) == (res.nit, res.nfev)
    assert st in (1, 2)
    assert len(xs) == len(ref_x)
    dev = max(np.abs(a - b).max() for a, b in zip(xs, ref_x))
    assert dev < 1e-9
    np.testing.assert_array_equal(xm.astype(np.float32), res.x.astype(np.float32))


def test_lbfgs_core_maxiter_and_quadratic():
    A = np.diag(np.linspace(1, 50, 30))
    b = np.arange(30.0)

    def func(x):
        return 0.5 * x @ A @ x - b @ x, A @ x - b

    res = optimize.minimize(func, np.zeros(30), method="L-BFGS-B", jac=True,
                            options={"maxiter": 7, "maxls": 50, "gtol": 1e-10, "ftol": 1e-16})
    xm, nit, nfev, st, _ = _mine(func, 30, maxiter=7, tol=1e-10)
    assert st == 3 and nit == res.nit == 7
    np.testing.assert_allclose(xm, res.x, rtol=0, atol=1e-10)
