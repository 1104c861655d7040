"""Device runs of the host-side additions made after the last GPU session of round 2 (histogram mode of the
forests, DistRandomTreesEmbedding, one-vs-one with an SGD base): each only adds host logic in front of a device
path the earlier test files cover, and each has a CPU test on the engine double.  Named to run last."""
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


@pytest.mark.filterwarnings("ignore")
def test_ovo_sgd_on_device():
    """DistOneVsOneClassifier(SGDClassifier): one exact-order SGD fit per class pair on the pair's rows, through
    the one-vs-rest SGD engine (warp-per-column kernel for the small pairs, tensor-core screening from 4096 rows)."""
    from sklearn.linear_model import SGDClassifier
    from sklearn.multiclass import OneVsOneClassifier
    from skdist.distribute.multiclass import DistOneVsOneClassifier
    from skdist_b200.datasets import make_multiclass
    for n in (3000, 15000):
        X, y = make_multiclass(n, 20, 3, seed=n % 13)
        ours = DistOneVsOneClassifier(SGDClassifier(random_state=0), None).fit(X, y)
        ref = OneVsOneClassifier(SGDClassifier(random_state=0)).fit(X, y)
        for a, b in zip(ours.estimators_, ref.estimators_):
            np.testing.assert_array_equal(a.coef_, b.coef_)
            np.testing.assert_array_equal(a.intercept_, b.intercept_)
            assert a.n_iter_ == b.n_iter_
