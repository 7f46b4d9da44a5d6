"""CPU: the plain-C oracle (oracle/tree_oracle.c) pinned against scikit-learn itself and against
the reference-generated golden trees."""
import os

import numpy as np
import pytest

from util import compare_trees, make_dataset, tree_arrays_sklearn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sk(kind, n_estimators):
    import sklearn.ensemble as E
    import sklearn.tree as T
    return {"DT": lambda: T.DecisionTreeClassifier(random_state=0),
            "RF": lambda: E.RandomForestClassifier(random_state=0, n_estimators=n_estimators),
            "ET": lambda: E.ExtraTreesClassifier(random_state=0, n_estimators=n_estimators)}[kind]()


def test_c_oracle_mt19937_matches_numpy():
    import tree_oracle as TO
    for seed, hi in ((0, 2147483647), (5, 90000), (123, 7)):
        assert list(TO.randint(seed, hi, 50)) == list(np.random.RandomState(seed).randint(0, hi, 50))


@pytest.mark.parametrize("kind", ["DT", "RF", "ET"])
@pytest.mark.parametrize("cfg", [dict(n=1500, fset="Flake16", prep="None"), dict(n=1500, fset="FlakeFlagger", prep="Scaling"),
                                 dict(n=1500, fset="Flake16", prep="PCA", const_col=7)])
def test_c_oracle_trees_equal_sklearn(kind, cfg):
    import tree_oracle as TO
    X, y, _ = make_dataset(**cfg)
    ref = _sk(kind, 4).fit(X[:1300], y[:1300])
    our = TO.OracleForest(kind, 4, 0).fit(X[:1300], y[:1300])
    errs = []
    for t, est in enumerate([ref] if kind == "DT" else ref.estimators_):
        errs += compare_trees(our.tree(t), tree_arrays_sklearn(est), "%s tree %d" % (kind, t))
    assert not errs, "\n".join(errs[:10])
    assert np.array_equal(our.predict(X[1300:]), ref.predict(X[1300:]))


def test_c_oracle_matches_reference_golden_trees():
    """Trees written by tests/golden/make_golden.py (the reference's own code path) - DT and RF,
    no balancing (the ET golden uses SMOTE'd rows, covered by the GPU test)."""
    import tree_oracle as TO
    from flake16_framework_b200 import hostprep as hp, synth
    g = np.load(os.path.join(ROOT, "tests", "golden", "trees_n2000_seed16.npz"))
    parsed = hp.tests_to_arrays(synth.make_tests_dict(2000, 16))
    X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
    tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
    for tag in ("DT", "RF"):
        m = TO.OracleForest(tag, 100, 0).fit(X[tr], y[tr])
        for t in range(1 if tag == "DT" else 3):
            ref = {k: g["%s_%d_%s" % (tag, t, k)] for k in ("children_left", "children_right", "feature", "threshold",
                                                           "n_node_samples", "weighted_n_node_samples", "impurity", "value")}
            assert not compare_trees(m.tree(t), ref, "%s %d" % (tag, t))
        assert np.array_equal(m.predict(X[te]), g["%s_pred" % tag])
