"""DistRandomForestClassifier on the device vs scikit-learn's RandomForestClassifier, which the
reference's `_build_trees` equals tree for tree (live check below)."""
import numpy as np
import pytest
from sklearn.ensemble import RandomForestClassifier

pytestmark = pytest.mark.gpu


def lattice_data(n, d, seed, levels=256, n_classes=2):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((n, d))
    X = np.clip(np.floor((Z + 4.0) / 8.0 * levels), 0, levels - 1).astype(np.float32)
    s = Z[:, 0] + 0.5 * Z[:, 1] * Z[:, 2] - 0.7 * Z[:, 3] + 0.8 * rng.standard_normal(n)
    if n_classes == 2:
        y = (s > 0).astype(np.int64)
    else:
        y = np.digitize(s, np.quantile(s, np.linspace(0, 1, n_classes + 1)[1:-1]))
    return X, y


def assert_same_forest(a, b):
    assert len(a.estimators_) == len(b.estimators_)
    for ta, tb in zip(a.estimators_, b.estimators_):
        x, z = ta.tree_, tb.tree_
        assert x.node_count == z.node_count and x.max_depth == z.max_depth
        np.testing.assert_array_equal(x.children_left, z.children_left)
        np.testing.assert_array_equal(x.children_right, z.children_right)
        np.testing.assert_array_equal(x.feature, z.feature)
        np.testing.assert_array_equal(x.threshold, z.threshold)
        np.testing.assert_array_equal(x.n_node_samples, z.n_node_samples)
        np.testing.assert_array_equal(x.weighted_n_node_samples, z.weighted_n_node_samples)
        np.testing.assert_array_equal(x.impurity, z.impurity)
        np.testing.assert_array_equal(x.value, z.value)


@pytest.mark.parametrize("n,d,levels,k", [(3000, 16, 256, 2), (20000, 64, 256, 2), (5000, 10, 8, 4)])
def test_forest_bit_identical_to_sklearn(n, d, levels, k):
    from skdist.distribute.ensemble import DistRandomForestClassifier
    X, y = lattice_data(n, d, seed=n % 13, levels=levels, n_classes=k)
    ours = DistRandomForestClassifier(n_estimators=6, random_state=5).fit(X, y)
    ref = RandomForestClassifier(n_estimators=6, random_state=5).fit(X, y)
    assert_same_forest(ours, ref)
    np.testing.assert_array_equal(ours.predict(X[:500]), ref.predict(X[:500]))
    np.testing.assert_array_equal(ours.predict_proba(X[:200]), ref.predict_proba(X[:200]))


def test_forest_hyperparameters_and_toy_case():
    from skdist.distribute.ensemble import DistRandomForestClassifier
    X, y = lattice_data(8000, 12, seed=3, levels=32)
    kw = dict(n_estimators=4, random_state=1, max_depth=7, min_samples_leaf=3, min_samples_split=10,
              max_features=5, bootstrap=False)
    assert_same_forest(DistRandomForestClassifier(**kw).fit(X, y), RandomForestClassifier(**kw).fit(X, y))
    # ref skdist/distribute/tests/test_ensemble.py:25-32
    Xt = np.array([[0, 1, 0, 1], [0, 0, 0, 1], [1, 0, 1, 0]])
    yt = np.array([0, 1, 0])
    rfc = DistRandomForestClassifier(n_estimators=10, random_state=5).fit(Xt, yt)
    assert np.allclose(rfc.predict(Xt), np.array([0, 1, 0]))
    assert_same_forest(rfc, RandomForestClassifier(n_estimators=10, random_state=5).fit(Xt, yt))


def test_forest_inference_kernel_matches_sklearn():
    """skd_forest_predict (soft vote over the Tree arrays, float64 sums in tree order) equals
    RandomForest*.predict_proba / predict bit for bit — for sklearn-built forests with continuous
    features too (inference does not need the <= 256-level restriction of the builder)."""
    from sklearn.ensemble import ExtraTreesClassifier, RandomForestClassifier, RandomForestRegressor
    from skdist.distribute.predict import batch_predict, get_prediction_udf
    rng = np.random.default_rng(5)
    X = rng.standard_normal((20000, 12)).astype(np.float32)
    y = (X[:, 0] + X[:, 1] * X[:, 2] + 0.3 * rng.standard_normal(20000) > 0).astype(int) + (X[:, 3] > 1)
    Xt = rng.standard_normal((50001, 12)).astype(np.float32)
    for model in (RandomForestClassifier(n_estimators=12, random_state=2).fit(X, y),
                  ExtraTreesClassifier(n_estimators=6, max_depth=9, random_state=2).fit(X, y)):
        np.testing.assert_array_equal(batch_predict(model, Xt, "predict_proba"), model.predict_proba(Xt))
        np.testing.assert_array_equal(batch_predict(model, Xt, "predict"), model.predict(Xt))
    reg = RandomForestRegressor(n_estimators=9, random_state=3).fit(X, X[:, 0] * 2 + X[:, 4])
    np.testing.assert_allclose(batch_predict(reg, Xt), reg.predict(Xt), rtol=0, atol=1e-12)
    import pandas as pd
    udf = get_prediction_udf(model, method="predict_proba")
    out = udf(*[pd.Series(Xt[:100, j]) for j in range(12)])
    np.testing.assert_array_equal(np.vstack(out.values), model.predict_proba(Xt[:100]))


@pytest.mark.parametrize("n,d,levels,k", [(3000, 16, 256, 2), (20000, 40, 256, 3), (5000, 10, 8, 4)])
def test_extra_trees_bit_identical_to_sklearn(n, d, levels, k):
    """Random splitter (SK/tree/_splitter.pyx:507-736): thresholds are drawn from the same xorshift
    stream as the features, so topology, thresholds and values must all equal scikit-learn's."""
    from sklearn.ensemble import ExtraTreesClassifier
    from skdist.distribute.ensemble import DistExtraTreesClassifier
    X, y = lattice_data(n, d, seed=n % 11, levels=levels, n_classes=k)
    ours = DistExtraTreesClassifier(n_estimators=5, random_state=7).fit(X, y)
    ref = ExtraTreesClassifier(n_estimators=5, random_state=7).fit(X, y)
    assert_same_forest(ours, ref)
    np.testing.assert_array_equal(ours.predict_proba(X[:300]), ref.predict_proba(X[:300]))
    kw = dict(n_estimators=3, random_state=2, max_depth=9, min_samples_leaf=4, max_features=3, bootstrap=True)
    assert_same_forest(DistExtraTreesClassifier(**kw).fit(X, y), ExtraTreesClassifier(**kw).fit(X, y))
    # ref skdist/distribute/tests/test_ensemble.py (ExtraTrees toy case)
    Xt = np.array([[0, 1, 0, 1], [0, 0, 0, 1], [1, 0, 1, 0]])
    yt = np.array([0, 1, 0])
    etc = DistExtraTreesClassifier(n_estimators=10, random_state=5).fit(Xt, yt)
    assert_same_forest(etc, ExtraTreesClassifier(n_estimators=10, random_state=5).fit(Xt, yt))


@pytest.mark.parametrize("n,d,levels", [(4000, 12, 256), (20000, 24, 64)])
def test_regression_forests_bit_identical_on_integer_targets(n, d, levels):
    """MSE trees: with integer-valued targets every float64 sum is exact, so the device trees must
    equal scikit-learn's bit for bit (structure, thresholds, impurities, node means)."""
    from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
    from skdist.distribute.ensemble import DistExtraTreesRegressor, DistRandomForestRegressor
    X, _ = lattice_data(n, d, seed=5, levels=levels)
    rng = np.random.default_rng(1)
    y = np.rint(0.05 * X[:, 0] - 0.03 * X[:, 1] + 0.0004 * X[:, 2] * X[:, 3] + 3 * rng.standard_normal(n))
    kw = dict(n_estimators=4, random_state=9, max_features=5, min_samples_leaf=2)
    assert_same_forest(DistRandomForestRegressor(**kw).fit(X, y), RandomForestRegressor(**kw).fit(X, y))
    assert_same_forest(DistExtraTreesRegressor(**kw).fit(X, y), ExtraTreesRegressor(**kw).fit(X, y))
    ours = DistRandomForestRegressor(n_estimators=3, random_state=2, max_depth=8).fit(X, y)     # max_features="auto" = all
    ref = RandomForestRegressor(n_estimators=3, random_state=2, max_depth=8).fit(X, y)
    assert_same_forest(ours, ref)
    np.testing.assert_array_equal(ours.predict(X[:300]), ref.predict(X[:300]))


def test_regression_forest_real_targets_and_reference_toy_cases():
    """Real-valued targets: sums are formed in a different order than scikit-learn's, so agreement
    is to rounding (predictions), not bitwise.  Toy cases from the reference's own tests
    (ref skdist/distribute/tests/test_ensemble.py:34-58)."""
    from sklearn.ensemble import RandomForestRegressor
    from skdist.distribute.ensemble import DistExtraTreesRegressor, DistRandomForestRegressor
    X, _ = lattice_data(6000, 10, seed=8, levels=64)
    rng = np.random.default_rng(3)
    y = 0.05 * X[:, 0] - 0.03 * X[:, 1] + rng.standard_normal(6000)
    kw = dict(n_estimators=5, random_state=4, max_depth=6, min_samples_leaf=20)
    ours = DistRandomForestRegressor(**kw).fit(X, y)
    ref = RandomForestRegressor(**kw).fit(X, y)
    np.testing.assert_allclose(ours.predict(X), ref.predict(X), rtol=0, atol=1e-9)
    Xt = np.array([[0, 1, 0, 1], [0, 0, 0, 1], [1, 0, 1, 0]])
    yt = np.array([0.1, 0.2, 0.1])
    rfr = DistRandomForestRegressor(n_estimators=10, random_state=5).fit(Xt, yt)
    np.testing.assert_allclose(rfr.predict(Xt), RandomForestRegressor(n_estimators=10, random_state=5).fit(Xt, yt).predict(Xt),
                               rtol=0, atol=1e-12)
    assert np.allclose(rfr.predict(Xt), np.array([0.15, 0.18, 0.12]))     # the reference test's golden values
    etr = DistExtraTreesRegressor(n_estimators=10, random_state=5).fit(Xt, yt)
    assert np.allclose(etr.predict(Xt), np.array([0.1, 0.2, 0.1]))
