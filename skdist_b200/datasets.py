"""Seeded synthetic workloads (BASELINE.json configs).

The generator is part of the benchmark contract: it sets the L-BFGS iteration counts
(isotropic N(0,1) features converge in 3-4 iterations and make a degenerate benchmark)."""
import numpy as np


def make_g1_classification(n, d, seed=0, dtype=np.float32):
    """G1 'standardised-correlated' binary classification (config 2).
    Returns X [n, d] float32 (C order), y [n] int64 in {0, 1}."""
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((n, d), dtype=np.float32)
    L = (rng.standard_normal((d, 8)) @ rng.standard_normal((8, d)) / np.sqrt(8)).astype(np.float32)
    X = Z + 0.5 * (Z @ L)
    X /= X.std(axis=0)
    w = rng.standard_normal(d) * 3.0 / np.sqrt(d)
    p = 1.0 / (1.0 + np.exp(-(X @ w.astype(np.float32)).astype(np.float64)))
    y = (rng.random(n) < p).astype(np.int64)
    return np.ascontiguousarray(X, dtype=dtype), y


def make_g1_regression(n, d, seed=0, dtype=np.float32):
    """G1 features with y = X w* + N(0, 1) (config 5)."""
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((n, d), dtype=np.float32)
    L = (rng.standard_normal((d, 8)) @ rng.standard_normal((8, d)) / np.sqrt(8)).astype(np.float32)
    X = Z + 0.5 * (Z @ L)
    X /= X.std(axis=0)
    w = rng.standard_normal(d) * 3.0 / np.sqrt(d)
    y = (X @ w.astype(np.float32)).astype(np.float64) + rng.standard_normal(n)
    return np.ascontiguousarray(X, dtype=dtype), y.astype(dtype)


def make_multiclass(n, d, k, seed=0, dtype=np.float32):
    """G1-style features, y = argmax(X W* + Gumbel) over k classes (config 3)."""
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((n, d), dtype=np.float32)
    L = (rng.standard_normal((d, 8)) @ rng.standard_normal((8, d)) / np.sqrt(8)).astype(np.float32)
    X = Z + 0.5 * (Z @ L)
    X /= X.std(axis=0)
    W = (rng.standard_normal((d, k)) / np.sqrt(d)).astype(np.float32)
    y = np.empty(n, dtype=np.int64)
    step = 65536
    for s in range(0, n, step):
        sc = X[s:s + step] @ W + rng.gumbel(size=(min(step, n - s), k)).astype(np.float32)
        y[s:s + step] = sc.argmax(1)
    return np.ascontiguousarray(X, dtype=dtype), y
