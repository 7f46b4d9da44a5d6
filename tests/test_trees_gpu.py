"""GPU parity: forests grown by the CUDA kernels vs scikit-learn (the reference's dependency),
node for node and bit for bit, through the C ABI."""
import numpy as np
import pytest

from util import compare_trees, make_dataset, tree_arrays_sklearn

pytestmark = pytest.mark.gpu

SK = {"ET": "ExtraTreesClassifier", "RF": "RandomForestClassifier", "DT": "DecisionTreeClassifier"}


def _sk(kind, n_estimators):
    import sklearn.ensemble as E
    import sklearn.tree as T
    if kind == "DT":
        return T.DecisionTreeClassifier(random_state=0)
    return getattr(E, SK[kind])(random_state=0, n_estimators=n_estimators)


def _ours(kind, n_estimators):
    from flake16_framework_b200 import estimators as est
    if kind == "DT":
        return est.DecisionTreeClassifier(random_state=0)
    return getattr(est, SK[kind])(random_state=0, n_estimators=n_estimators)


def _check(kind, X, y, Xte, n_estimators=5):
    ref = _sk(kind, n_estimators).fit(X, y)
    our = _ours(kind, n_estimators).fit(X, y)
    our.forest_.status()
    ref_trees = [ref] if kind == "DT" else ref.estimators_
    counts = our.forest_.node_counts()
    errs = []
    for t, rt in enumerate(ref_trees):
        errs += compare_trees(our.forest_.export_tree(t, int(counts[t])), tree_arrays_sklearn(rt), "%s tree %d" % (kind, t))
    pr, po = ref.predict(Xte), our.predict(Xte)
    if not np.array_equal(pr, po):
        errs.append("%s predict differs on %d/%d rows" % (kind, int((pr != po).sum()), len(pr)))
    assert not errs, "\n".join(errs[:20])


@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
@pytest.mark.parametrize("cfg", [
    dict(n=2000, fset="Flake16", prep="None"),
    dict(n=2000, fset="FlakeFlagger", prep="None"),
    dict(n=3000, fset="Flake16", prep="Scaling"),
    dict(n=3000, fset="Flake16", prep="PCA"),
    dict(n=3000, fset="Flake16", prep="None", const_col=7),
    dict(n=2500, fset="FlakeFlagger", prep="Scaling", flaky="OD"),
])
def test_tree_parity_small(cuda, kind, cfg):
    X, y, _ = make_dataset(**cfg)
    n = len(y)
    cut = int(n * 0.9)
    _check(kind, X[:cut], y[:cut], X[cut:])


@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
def test_tree_parity_20k(cuda, kind):
    X, y, _ = make_dataset(20000, flaky="OD")
    _check(kind, X[:18000], y[:18000], X[18000:], n_estimators=3)


@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
def test_sklearn_toy_kat(cuda, kind):
    """sklearn's own known-answer toy set (sklearn/tree/tests/test_tree.py:154-157,245;
    sklearn/ensemble/tests/test_forest.py:61-64,123)."""
    X = np.array([[-2, -1], [-1, -1], [-1, -2], [1, 1], [1, 2], [2, 1]], dtype=np.float64)
    y = np.array([-1, -1, -1, 1, 1, 1])
    T = np.array([[-1, -1], [2, 2], [3, 2]], dtype=np.float64)
    our = _ours(kind, 10).fit(X, y)
    assert list(our.predict(T)) == [-1, 1, 1]
    _check(kind, X, y, T, n_estimators=10)


@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
def test_duplicates_and_tiny(cuda, kind):
    """Heavy ties, duplicate rows and near-equal values (FEATURE_THRESHOLD = 1e-7 rule)."""
    rs = np.random.RandomState(3)
    X = rs.randint(0, 3, size=(400, 5)).astype(np.float64)
    X[:, 4] = 1.0 + rs.randint(0, 4, size=400) * 4e-8       # chained near-ties
    y = (rs.rand(400) < 0.3)
    _check(kind, X, y, X[:50], n_estimators=7)


def test_full_100_trees_counts(cuda):
    """100-tree forests as the reference uses them: identical predictions -> identical FP/FN/TP."""
    X, y, _ = make_dataset(6000)
    for kind in ("ET", "RF"):
        ref = _sk(kind, 100).fit(X[:5400], y[:5400])
        our = _ours(kind, 100).fit(X[:5400], y[:5400])
        assert np.array_equal(ref.predict(X[5400:]), our.predict(X[5400:])), kind
        assert int(our.forest_.node_counts().sum()) == sum(e.tree_.node_count for e in ref.estimators_), kind


@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
def test_edge_cases(cuda, kind):
    """Degenerate inputs the estimators must survive like scikit-learn does: one row, two rows,
    a single class, all-constant features, empty prediction batch."""
    rs = np.random.RandomState(11)
    # one row / two rows
    for n in (1, 2, 3):
        X = rs.rand(n, 4)
        y = np.arange(n) % 2 == 0
        _check(kind, X, y, X, n_estimators=4)
    # a single class: the root is a leaf
    X = rs.rand(50, 6)
    y = np.zeros(50, dtype=bool)
    our = _ours(kind, 3).fit(X, y)
    assert list(our.forest_.node_counts()) == [1] * (1 if kind == "DT" else 3)
    assert not our.predict(X).any()
    # all features constant: no valid split anywhere
    Xc = np.ones((40, 5))
    yc = rs.rand(40) < 0.5
    _check(kind, Xc, yc, Xc, n_estimators=3)
    # empty prediction batch
    our = _ours(kind, 3).fit(X, rs.rand(50) < 0.5)
    assert our.predict(np.zeros((0, 6))).shape == (0,)


def test_bad_inputs_raise(cuda):
    from flake16_framework_b200 import estimators as est
    from flake16_framework_b200._lib import F16Error
    X = np.random.RandomState(0).rand(30, 4)
    y = np.arange(30) % 2 == 0
    with pytest.raises(ValueError):
        est.ExtraTreesClassifier(random_state=0).fit(X, y[:-1])
    with pytest.raises(ValueError):
        est.RandomForestClassifier(random_state=0).fit(X, np.arange(30) % 3)
    m = est.DecisionTreeClassifier(random_state=0).fit(X, y)
    with pytest.raises(ValueError):
        m.predict(X[:, :3])
    with pytest.raises((ValueError, F16Error)):
        est.RandomForestClassifier(random_state=0).fit(np.zeros((10, 17)), np.arange(10) % 2 == 0)
    with pytest.raises(ValueError):
        est.SMOTE(random_state=0).fit_resample(X, np.arange(30) < 3)       # 3 minority rows < k + 1


def test_global_side_bits_path(cuda):
    """Training sets above 524 288 rows keep the best-splitter's side bits in global memory; force
    that path on a small input (env switch read once per process -> subprocess) and check parity."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from util import make_dataset, compare_trees, tree_arrays_sklearn\n"
        "from sklearn.ensemble import RandomForestClassifier as SK\n"
        "from flake16_framework_b200.estimators import RandomForestClassifier as OURS\n"
        "X, y, _ = make_dataset(3000)\n"
        "ref = SK(random_state=0, n_estimators=4).fit(X, y); our = OURS(random_state=0, n_estimators=4).fit(X, y)\n"
        "errs = []\n"
        "for t, e in enumerate(ref.estimators_): errs += compare_trees(our.forest_.export_tree(t), tree_arrays_sklearn(e), str(t))\n"
        "assert not errs, errs[:3]\n"
        "print('GLOBAL_SIDE_OK')\n" % (root, os.path.join(root, "tests")))
    out = subprocess.check_output([sys.executable, "-c", code], env=dict(os.environ, F16_FORCE_GLOBAL_SIDE="1"))
    assert b"GLOBAL_SIDE_OK" in out
