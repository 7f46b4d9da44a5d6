"""Groundwork for SURVEY.md 8(f) row N3: the TreeSHAP oracle (oracle/treeshap_np.py) is pinned
against the DEFINITION of what it computes - Shapley values of v(S) = E[f(x) | x_S] obtained by
enumerating every feature subset - because the shap package is not installed (parity against it
stays unpinned, as the oracle's header says)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import treeshap_np as T


def _data(seed, n, d):
    rs = np.random.RandomState(seed)
    X = rs.randn(n, d)
    X[:, 1] = rs.randint(0, 3, n)                      # a discrete column: repeated splits on one feature
    y = ((X[:, 0] + 0.5 * X[:, 1] * X[:, 2] + 0.3 * rs.randn(n)) > 0.2).astype(int)
    return X.astype(np.float32).astype(np.float64), y


@pytest.mark.parametrize("seed,d,depth", [(0, 4, 3), (1, 6, 6), (2, 7, None)])
def test_single_tree_equals_enumeration(seed, d, depth):
    from sklearn.tree import DecisionTreeClassifier
    X, y = _data(seed, 300, d)
    est = DecisionTreeClassifier(max_depth=depth, random_state=0).fit(X, y)
    tree = T._Tree(est, 0)
    for r in range(0, 300, 37):
        phi, base = T.tree_shap(tree, X[r], d)
        ref = T.shapley_by_enumeration(tree, X[r], d)
        assert np.allclose(phi, ref, rtol=0, atol=1e-12), (r, np.abs(phi - ref).max())
        # local accuracy: attributions sum to f(x) - E[f]
        fx = est.predict_proba(X[r:r + 1])[0, 0]
        assert abs(phi.sum() - (fx - base)) < 1e-12


@pytest.mark.parametrize("kind", ["rf", "et"])
def test_forest_mean_over_trees_and_local_accuracy(kind):
    """Bootstrap weights (RandomForest) enter through the weighted covers."""
    from sklearn.ensemble import ExtraTreesClassifier, RandomForestClassifier
    X, y = _data(5, 400, 5)
    M = RandomForestClassifier if kind == "rf" else ExtraTreesClassifier
    model = M(n_estimators=7, max_depth=5, random_state=0).fit(X, y)
    rows = X[::57]
    phi = T.forest_shap_values(model, rows, klass=0)
    ref = np.mean([[T.shapley_by_enumeration(T._Tree(e, 0), x, 5) for x in rows] for e in model.estimators_], axis=0)
    assert np.allclose(phi, ref, rtol=0, atol=1e-12)
    base = np.mean([T.expected_value(T._Tree(e, 0), rows[0], frozenset()) for e in model.estimators_])
    assert np.allclose(phi.sum(axis=1), model.predict_proba(rows)[:, 0] - base, atol=1e-12)
    # the two classes' attributions cancel (the class fractions sum to one)
    assert np.allclose(phi + T.forest_shap_values(model, rows, klass=1), 0.0, atol=1e-12)
