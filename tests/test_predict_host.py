"""Host logic of skdist.distribute.predict (post-processing + UDF plumbing) on the test-double engine."""
import warnings

import numpy as np
import pandas as pd
import pytest
from sklearn.ensemble import RandomForestClassifier, RandomForestRegressor
from sklearn.linear_model import LogisticRegression, Ridge, SGDClassifier
from sklearn.multiclass import OneVsRestClassifier

from skdist.distribute.predict import batch_predict, get_prediction_udf
from skdist_b200.datasets import make_multiclass


@pytest.fixture
def data():
    X, y = make_multiclass(600, 8, 3, seed=4)
    return X, y


@pytest.mark.filterwarnings("ignore")
def test_linear_models_match_sklearn(fake_engine, data):
    X, y = data
    for model in (LogisticRegression().fit(X, y), LogisticRegression().fit(X, y == 1),
                  SGDClassifier(loss="log_loss", random_state=0).fit(X, y),
                  OneVsRestClassifier(LogisticRegression()).fit(X, y)):
        np.testing.assert_array_equal(batch_predict(model, X, "predict"), model.predict(X))
        np.testing.assert_allclose(batch_predict(model, X, "predict_proba"), model.predict_proba(X), atol=2e-6)
    ridge = Ridge().fit(X, y.astype(np.float32))
    np.testing.assert_allclose(batch_predict(ridge, X), ridge.predict(X), atol=1e-5)


def test_forest_models_match_sklearn(fake_engine, data):
    X, y = data
    Xq = np.round(X * 8) / 8
    rf = RandomForestClassifier(n_estimators=7, random_state=1).fit(Xq, y)
    np.testing.assert_array_equal(batch_predict(rf, Xq, "predict_proba"), rf.predict_proba(Xq))
    np.testing.assert_array_equal(batch_predict(rf, Xq, "predict"), rf.predict(Xq))
    rr = RandomForestRegressor(n_estimators=5, random_state=1).fit(Xq, y.astype(float))
    np.testing.assert_allclose(batch_predict(rr, Xq), rr.predict(Xq), rtol=0, atol=1e-12)


def test_prediction_udf_contract(fake_engine, data):
    """Same call shape as the reference's pandas UDF body (ref predict.py:160-179)."""
    X, y = data
    model = LogisticRegression().fit(X, y)
    cols = [pd.Series(X[:, j]) for j in range(X.shape[1])]
    preds = get_prediction_udf(model, method="predict")(*cols)
    probs = get_prediction_udf(model, method="predict_proba")(*cols)
    assert isinstance(preds, pd.Series) and isinstance(probs, pd.Series)
    np.testing.assert_array_equal(preds.values, model.predict(X))
    np.testing.assert_allclose(np.vstack(probs.values), model.predict_proba(X), atol=2e-6)
    with pytest.raises(ValueError):
        get_prediction_udf(model, method="transform")
    with pytest.raises(ValueError):
        get_prediction_udf(model, feature_type="image")
    with pytest.raises(NotImplementedError):
        get_prediction_udf(model, feature_type="text")(pd.Series(["a", "b"]))
    from sklearn.neighbors import KNeighborsClassifier
    with pytest.raises(NotImplementedError):
        get_prediction_udf(KNeighborsClassifier().fit(X, y))
