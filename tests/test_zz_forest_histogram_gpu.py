"""Opt-in histogram mode of the forests on the device (runs last: it only adds a host transform in front of
the device path the other forest tests cover)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_histogram_mode_on_device(monkeypatch):
    from sklearn.ensemble import RandomForestClassifier
    from skdist.distribute.ensemble import DistRandomForestClassifier
    from skdist_b200.distribute.ensemble import _quantile_codes
    rng = np.random.default_rng(11)
    n, d, bins = 6000, 8, 32
    X = rng.standard_normal((n, d)).astype(np.float32)
    y = ((X[:, 0] + X[:, 1] * X[:, 2]) > 0.3).astype(int) + (X[:, 3] > 1.0)
    with pytest.raises(NotImplementedError, match="SKDIST_B200_FOREST_MAX_BINS"):
        DistRandomForestClassifier(n_estimators=2, random_state=0).fit(X, y)
    monkeypatch.setenv("SKDIST_B200_FOREST_MAX_BINS", str(bins))
    Xc, _ = _quantile_codes(X, bins)
    ours = DistRandomForestClassifier(n_estimators=4, random_state=0).fit(X, y)
    ref = RandomForestClassifier(n_estimators=4, random_state=0).fit(Xc, y)
    for a, b in zip(ours.estimators_, ref.estimators_):
        np.testing.assert_array_equal(a.tree_.feature, b.tree_.feature)
        np.testing.assert_array_equal(a.tree_.children_right, b.tree_.children_right)
        np.testing.assert_array_equal(a.tree_.value, b.tree_.value)
    np.testing.assert_array_equal(ours.predict_proba(X), ref.predict_proba(Xc))


def test_random_trees_embedding_on_device():
    """DistRandomTreesEmbedding: totally random regression trees (one drawn feature per node) on uniform random
    targets -- the leaves every row lands in are scikit-learn's."""
    from sklearn.ensemble import RandomTreesEmbedding
    from skdist.distribute.ensemble import DistRandomTreesEmbedding
    rng = np.random.default_rng(2)
    X = rng.integers(0, 32, size=(5000, 10)).astype(np.float32)
    ours = DistRandomTreesEmbedding(n_estimators=8, random_state=3)
    got = ours.fit_transform(X)
    ref = RandomTreesEmbedding(n_estimators=8, random_state=3)
    want = ref.fit_transform(X)
    assert got.shape == want.shape and (got != want).nnz == 0
