"""Arithmetic of the two screening shortcuts, restated in numpy (no GPU): the bounds the kernels rely on hold
for the emulated tensor-core / float32 arithmetic on random and adversarial inputs.

* csrc/sgd_tc.cu: a sample is skipped as a certain non-violator only if its approximate margin (fp16
  operands, fp32 accumulation) clears 1 by more than  wscale * 2^-9 * |x| * nwb + 2.4e-7 * (|p| + 1),
  nwb >= |w| + sum_j |q_j| |x_j| (the in-block updates).  Here: |approximate dot - exact dot| <= 2^-9 |x| |w|
  for X scaled by one power of two (largest entry in [2^13, 2^14)), W by one per column, as the kernels do.
  Precondition stated in DESIGN.md (K7t): rows whose norm is below 2^-21 of the largest entry of X could
  lose their small entries to fp16 subnormals; their error is absolute (<= 2^-38 * max|X| per entry) and is
  covered by the band's absolute term for margins of ordinary size -- shown below as well.
* csrc/forest_fast.cu: candidate splits are ranked by sq_l / w_l + sq_r / w_r in float32 and scikit-learn's
  float64 proxy is evaluated only within 2^-19 * w_node of the best; here: the float32 value is within
  2^-20 * w_node of the exact one, so a candidate outside the bar cannot win.
"""
import numpy as np
import pytest


def _pow2_scale(m):
    """2^(13 - floor(log2 m)) -- sgd_permute_kernel / sgd_export_row."""
    _, e = np.frexp(np.float32(m))
    return np.float32(np.ldexp(1.0, 13 - (int(e) - 1)))


def _approx_dots(X, W):
    """S = fp16(X * sx) . fp16(W * t)^T with fp32 accumulation, unscaled -- what sgd_gemm_kernel + the scan's
    `s * inv_scale` produce (accumulation order differs on the tensor core; the bound does not depend on it)."""
    sx = _pow2_scale(np.abs(X).max())
    Xh = (X * sx).astype(np.float16).astype(np.float32)
    out = np.empty((X.shape[0], W.shape[0]), np.float64)
    for k in range(W.shape[0]):
        t = _pow2_scale(np.abs(W[k]).max())
        Wh = (W[k] * t).astype(np.float16).astype(np.float32)
        acc = np.zeros(X.shape[0], np.float32)
        for j in range(X.shape[1]):                       # worst-case style sequential fp32 accumulation
            acc = (acc + Xh[:, j] * Wh[j]).astype(np.float32)
        out[:, k] = acc.astype(np.float64) / (float(sx) * float(t))
    return out


@pytest.mark.parametrize("d,seed,kind", [(24, 0, "normal"), (512, 1, "normal"), (100, 2, "heavy"), (64, 3, "sparse")])
def test_sgd_screening_bound_covers_the_fp16_products(d, seed, kind):
    rng = np.random.default_rng(seed)
    n, K = 400, 6
    X = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "heavy":                                   # entries over six orders of magnitude inside every row
        X *= np.float32(10.0) ** rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    if kind == "sparse":
        X *= rng.random((n, d)) < 0.1
        X[5] = 0.0                                        # an all-zero row is exact
    W = (rng.standard_normal((K, d)) * 10.0 ** rng.integers(-4, 3, size=(K, 1))).astype(np.float32)
    exact = X.astype(np.float64) @ W.astype(np.float64).T
    approx = _approx_dots(X, W)
    xn = np.sqrt((X.astype(np.float64) ** 2).sum(1))[:, None]
    wn = np.sqrt((W.astype(np.float64) ** 2).sum(1))[None, :]
    gmax = float(np.abs(X).max())
    rows_ok = (xn[:, 0] == 0) | (xn[:, 0] >= 2.0 ** -21 * gmax)
    assert rows_ok.all()
    bound = 2.0 ** -9 * xn * wn
    assert (np.abs(approx - exact) <= bound).all()
    with np.errstate(invalid="ignore", divide="ignore"):
        used = np.where(bound > 0, np.abs(approx - exact) / bound, 0.0).max()
    assert used <= 0.25          # random data uses about a tenth of the worst-case (Cauchy-Schwarz) bound


def test_sgd_screening_rows_far_below_the_largest_entry():
    """A row eight orders of magnitude below the largest entry of X: its entries are fp16 subnormals after the
    global scaling, the relative bound no longer holds, the error is absolute: at most
    sqrt(d) * 2^-38 * max|X| * |w| -- below the band's absolute term 2.4e-7 whenever
    wscale * |w| * max|X| * sqrt(d) <= 6e4 (margins of ordinary size)."""
    rng = np.random.default_rng(7)
    d = 128
    X = rng.standard_normal((50, d)).astype(np.float32)
    X[3] *= np.float32(1e-8)
    W = rng.standard_normal((3, d)).astype(np.float32)
    exact = X.astype(np.float64) @ W.astype(np.float64).T
    approx = _approx_dots(X, W)
    gmax = float(np.abs(X).max())
    wn = np.sqrt((W.astype(np.float64) ** 2).sum(1))
    err = np.abs(approx - exact)[3]
    assert (err <= np.sqrt(d) * 2.0 ** -38 * gmax * wn).all()
    assert (err <= 2.4e-7).all()


def _rank32(sl, st):
    """ff_rank: float32 fused multiply-adds, two divisions (__fdividef: within 2 ulp, added to the error below)."""
    sl = sl.astype(np.float32); sr = (st - sl).astype(np.float32)
    wl = np.float32(0); sql = np.float32(0); sqr = np.float32(0)
    for c in range(len(st)):
        wl = np.float32(wl + sl[c])
        sql = np.float32(np.float64(sl[c]) * np.float64(sl[c]) + np.float64(sql))      # fmaf: one rounding
        sqr = np.float32(np.float64(sr[c]) * np.float64(sr[c]) + np.float64(sqr))
    wn = np.float32(np.float32(st.sum()))
    return np.float32(sql / wl) + np.float32(sqr / np.float32(wn - wl)), float(sql / wl) + float(sqr / (wn - wl))


@pytest.mark.parametrize("C", [2, 3, 4])
def test_forest_float32_rank_is_within_the_bar(C):
    from fractions import Fraction
    rng = np.random.default_rng(C)
    worst = 0.0
    for trial in range(3000):
        scale = 10 ** rng.integers(0, 7)                      # node weights from a handful to 2^24 (the checked limit n * 255 < 2^32
        st = rng.integers(1, 10 * scale, size=C).astype(np.int64)       # keeps sums exact in uint32; floats round above 2^24)
        if st.sum() >= 2 ** 24:
            st = (st * (2 ** 24 - 1) // st.sum()).clip(1)
        sl = np.array([rng.integers(0, s + 1) for s in st], dtype=np.int64)
        if sl.sum() == 0 or sl.sum() == st.sum():
            continue
        r32, mag = _rank32(sl, st)
        wl, wr, wn = int(sl.sum()), int((st - sl).sum()), int(st.sum())
        exact = sum(Fraction(int(a) ** 2) for a in sl) / wl + sum(Fraction(int(b) ** 2) for b in st - sl) / wr
        err = abs(float(Fraction(float(r32)) - exact)) + 2.0 ** -22 * mag      # + 2 ulp per approximate division
        worst = max(worst, err / wn)
        # proxy_impurity_improvement = rank - w_node exactly: -w_r (1 - sq_r / w_r^2) - w_l (1 - sq_l / w_l^2)
        proxy = -Fraction(wr) * (1 - sum(Fraction(int(b) ** 2) for b in st - sl) / Fraction(wr) ** 2) \
                - Fraction(wl) * (1 - sum(Fraction(int(a) ** 2) for a in sl) / Fraction(wl) ** 2)
        assert proxy == exact - wn
    assert worst <= 2.0 ** -20, worst
