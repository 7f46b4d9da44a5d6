"""SURVEY.md 8(f) rows N3 / N4 on the GPU: TreeSHAP (f16_forest_shap) against the oracle
(oracle/treeshap_np.py, EXTEND / UNWIND, itself pinned against subset enumeration) and the local-
accuracy property; the Spearman table (f16_spearman) against scipy.stats.spearmanr."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import make_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair(kind, n_estimators):
    import sklearn.ensemble as SE
    import sklearn.tree as ST
    from flake16_framework_b200 import estimators as E
    if kind == "DT":
        return ST.DecisionTreeClassifier(random_state=0), E.DecisionTreeClassifier(random_state=0)
    name = {"ET": "ExtraTreesClassifier", "RF": "RandomForestClassifier"}[kind]
    return (getattr(SE, name)(random_state=0, n_estimators=n_estimators),
            getattr(E, name)(random_state=0, n_estimators=n_estimators))


@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
@pytest.mark.parametrize("cfg", [dict(n=700, fset="Flake16", prep="Scaling"), dict(n=600, fset="FlakeFlagger", prep="None")])
def test_treeshap_equals_oracle(cuda, kind, cfg):
    """Both classes, every feature: relative 1e-10 of the largest attribution (the forests are
    bit-identical to scikit-learn's, so the oracle explains the scikit-learn model)."""
    import treeshap_np as O
    X, y, _ = make_dataset(**cfg)
    ref, our = _pair(kind, 4)
    ref.fit(X, y)
    our.fit(X, y)
    rows = X[:: max(1, len(X) // 40)][:40]
    got = our.shap_values(rows)
    for klass in (0, 1):
        want = O.forest_shap_values(ref, rows, klass)
        scale = np.abs(want).max()
        assert scale > 0
        err = np.abs(got[klass] - want).max() / scale
        assert err <= 1e-10, (kind, klass, err)
    # the two classes' attributions cancel (their outputs sum to one)
    assert np.abs(got[0] + got[1]).max() <= 1e-12


def test_treeshap_local_accuracy_at_scale(cuda):
    """20 000 rows x 30 ExtraTrees: sum of the attributions = f(x) - E[f], with f = scikit-learn's
    predict_proba of the identical forest and E[f] = mean over trees of the root's class fraction."""
    from sklearn.ensemble import ExtraTreesClassifier
    from flake16_framework_b200 import estimators as E
    X, y, _ = make_dataset(20000, prep="Scaling")
    ref = ExtraTreesClassifier(random_state=0, n_estimators=30).fit(X, y)
    our = E.ExtraTreesClassifier(random_state=0, n_estimators=30).fit(X, y)
    phi = our.shap_values(X)[0]
    assert phi.shape == X.shape
    fx = ref.predict_proba(X)[:, 0]
    base = np.mean([e.tree_.value[0, 0, 0] / e.tree_.value[0, 0].sum() for e in ref.estimators_])
    assert np.abs(phi.sum(axis=1) - (fx - base)).max() <= 1e-10
    assert np.isfinite(phi).all() and np.abs(phi).max() > 1e-3


def test_shap_cli_and_reference_configs(cuda, tmp_path):
    """`python experiment.py shap` (experiment.py:520-530): shap.pkl holds the two explained
    configurations as float64 [N, 16] arrays; spot rows equal the oracle on a scikit-learn forest
    fitted on the same resampled set."""
    import samplers_np as S
    import treeshap_np as O
    from sklearn.ensemble import RandomForestClassifier
    from flake16_framework_b200 import hostprep as hp, synth
    synth.make_tests_json(str(tmp_path / "tests.json"), 1200, 16)
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "experiment.py"), "shap"], cwd=str(tmp_path), env=env)
    shap = pickle.load(open(str(tmp_path / "shap.pkl"), "rb"))
    assert isinstance(shap, list) and len(shap) == 2
    assert all(a.shape == (1200, 16) and a.dtype == np.float64 for a in shap)
    parsed = hp.parse_tests(str(tmp_path / "tests.json"))
    X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES["OD"], hp.FEATURE_SETS["Flake16"])
    X = np.ascontiguousarray(hp.preprocess(X, "Scaling"))
    Xb, yb = S.SMOTE(random_state=0).fit_resample(X, y)
    ref = RandomForestClassifier(random_state=0, n_estimators=100).fit(Xb, yb)
    rows = np.arange(0, 1200, 240)
    want = O.forest_shap_values(ref, X[rows], 0)
    assert np.abs(shap[1][rows] - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("n", [5000, 100000])
def test_spearman_equals_scipy(cuda, n):
    """The raw 16-feature table (integer columns full of ties + continuous ones, as the reference
    feeds it) and a variant with duplicated rows, signed zeros and a 7-column subset."""
    from scipy import stats
    from flake16_framework_b200 import ops
    X, _, _ = make_dataset(n)
    got = ops.spearman(torch.from_numpy(X).cuda()).cpu().numpy()
    want = stats.spearmanr(X).correlation
    assert got.shape == (16, 16) and np.abs(got - want).max() <= 1e-12
    assert np.abs(np.diag(got) - 1.0).max() <= 1e-15
    Y = np.ascontiguousarray(X[:, [0, 1, 2, 3, 10, 11, 14]] - 3.0)
    Y[::7] = Y[3]
    Y[5, 1], Y[6, 1] = 0.0, -0.0
    got = ops.spearman(torch.from_numpy(Y).cuda()).cpu().numpy()
    assert np.abs(got - stats.spearmanr(Y).correlation).max() <= 1e-12
