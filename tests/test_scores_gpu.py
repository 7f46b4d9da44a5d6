"""GPU parity of the whole `scores` stage: grid engine + CLI vs the reference-generated golden
fixtures and the oracle; size-independent properties at BASELINE scale."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from util import compare_trees, make_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _gold():
    return pickle.load(open(os.path.join(GOLD, "scores_n1500_seed16.pkl"), "rb"))


def _cmp(scores, gold):
    bad = []
    for cfg, (g_proj, g_total) in gold.items():
        t_train, t_test, per_proj, total = scores[cfg]
        if [int(v) for v in total[:3]] != g_total:
            bad.append((cfg, total[:3], g_total))
            continue
        ours = {str(k): [int(x) for x in v[:3]] for k, v in per_proj.items()}
        if ours != g_proj:
            bad.append((cfg, "per-project", ""))
    return bad


def test_full_grid_matches_reference_golden(cuda):
    """All 216 configs, N=1500: FP/FN/TP per project and overall identical to the reference's
    own write_scores() output (BASELINE.json: integer work bit-exact)."""
    from flake16_framework_b200 import scores as S, hostprep as hp, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(1500, 16))
    scores = S.run_grid(parsed, n_streams=4)
    assert len(scores) == 216
    bad = _cmp(scores, _gold())
    assert not bad, "%d/216 configs differ, first: %r" % (len(bad), bad[:5])
    # P/R/F derived with the reference's expressions from identical counts (tolerance 1e-6)
    from ref_scores import get_prf
    for cfg, (_, _, per_proj, total) in scores.items():
        for row in [*per_proj.values(), total]:
            for a, b in zip(row[3:], get_prf(*row[:3])):
                assert (a is None and b is None) or abs(a - b) <= 1e-6


def test_full_grid_matches_reference_golden_n20000(cuda):
    """All 216 configs at 20 000 tests against the reference's own write_scores() output
    (tests/golden/make_golden_n20000.py).  The SMOTE'd training sets (~35 000 rows) of the Scaling /
    PCA datasets take the tensor-core k-NN filter inside the product flow (n x nq >= 2.5e8), the raw
    ones the early-exit float64 search; the forests hold 100 trees on 18 000 - 35 000 rows."""
    path = os.path.join(GOLD, "scores_n20000_seed16.pkl")
    if not os.path.exists(path):
        pytest.skip("scores_n20000_seed16.pkl not generated")
    from flake16_framework_b200 import scores as S, hostprep as hp, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(20000, 16))
    scores = S.run_grid(parsed, n_streams=4)
    gold = pickle.load(open(path, "rb"))
    assert len(scores) == 216 and len(gold) == 216
    bad = _cmp(scores, gold)
    assert not bad, "%d/216 configs differ, first: %r" % (len(bad), bad[:5])


def test_cli_drop_in(cuda, tmp_path):
    """`python experiment.py scores` reads ./tests.json and writes ./scores.pkl with the
    reference's schema (experiment.py:488-490,498-501)."""
    from flake16_framework_b200 import synth
    synth.make_tests_json(str(tmp_path / "tests.json"), 1500, 16)
    env = dict(os.environ, PYTHONPATH=ROOT, F16_STREAMS="4")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "experiment.py"), "scores"], cwd=str(tmp_path), env=env)
    scores = pickle.load(open(str(tmp_path / "scores.pkl"), "rb"))
    assert len(scores) == 216
    for keys, val in scores.items():
        assert isinstance(keys, tuple) and len(keys) == 5 and all(isinstance(k, str) for k in keys)
        assert isinstance(val, list) and len(val) == 4
        t_train, t_test, per_proj, total = val
        assert isinstance(t_train, float) and isinstance(t_test, float) and t_train > 0
        assert len(per_proj) == 26 and all(len(v) == 6 for v in per_proj.values()) and len(total) == 6
        assert all(isinstance(v, int) for v in total[:3])
    assert not _cmp(scores, _gold())
    with pytest.raises(subprocess.CalledProcessError):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "experiment.py"), "figures"], cwd=str(tmp_path), env=env,
                              stderr=subprocess.DEVNULL)


def test_golden_trees(cuda):
    """BASELINE configs 1-3: trees identical to the reference-generated tree_ arrays."""
    from flake16_framework_b200 import estimators as E, hostprep as hp, synth
    import samplers_np as O
    g = np.load(os.path.join(GOLD, "trees_n2000_seed16.npz"))
    parsed = hp.tests_to_arrays(synth.make_tests_dict(2000, 16))
    for tag, pre, bal, cls in (("DT", "None", None, E.DecisionTreeClassifier), ("RF", "None", None, E.RandomForestClassifier),
                               ("ET", "Scaling", "SMOTE", E.ExtraTreesClassifier)):
        X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
        X = hp.preprocess(X, pre)
        tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
        Xtr, ytr = X[tr], y[tr]
        if bal:
            Xtr, ytr = E.SMOTE(random_state=0).fit_resample(Xtr, ytr)
        m = cls(random_state=0).fit(Xtr, ytr)
        errs = []
        for t in range(1 if tag == "DT" else 3):
            ref = {k: g["%s_%d_%s" % (tag, t, k)] for k in ("children_left", "children_right", "feature", "threshold",
                                                           "n_node_samples", "weighted_n_node_samples", "impurity", "value")}
            errs += compare_trees(m.forest_.export_tree(t), ref, "%s tree %d" % (tag, t))
        assert not errs, "\n".join(errs[:10])
        assert np.array_equal(m.predict(X[te]), g["%s_pred" % tag])


def test_properties_at_baseline_scale(cuda):
    """N = 100k (BASELINE headline size) - properties that need no oracle:
    a fully grown DecisionTree classifies its (duplicate-free) training set perfectly; every
    internal node's children partition its rows and class sums; leaves are pure or unsplittable;
    resampling round trips (SMOTE makes classes equal, Tomek/ENN only remove rows)."""
    from flake16_framework_b200 import estimators as E, ops
    X, y, _ = make_dataset(100000, prep="Scaling")
    m = E.DecisionTreeClassifier(random_state=0).fit(X[:90000], y[:90000])
    assert np.array_equal(m.predict(X[:90000]), y[:90000])
    t = m.forest_.export_tree(0)
    internal = t["children_left"] >= 0
    l, r = t["children_left"][internal], t["children_right"][internal]
    assert np.array_equal(t["n_node_samples"][internal], t["n_node_samples"][l] + t["n_node_samples"][r])
    assert t["n_node_samples"][0] == 90000 and np.all(t["impurity"][~internal] <= 2.3e-16)
    assert np.all(l == np.flatnonzero(internal) + 1)
    for cls in (E.RandomForestClassifier, E.ExtraTreesClassifier):
        f = cls(random_state=0, n_estimators=20).fit(X[:90000], y[:90000])
        p = f.predict(X[90000:])
        assert p.dtype == y.dtype and 0 < p.sum() < 0.1 * len(p)
        c = f.forest_.node_counts()
        assert np.all(c % 2 == 1) and np.all(c > 1000)
    Xs, ys = E.SMOTE(random_state=0).fit_resample(X[:90000], y[:90000])
    assert ys.sum() * 2 == len(ys) and np.array_equal(Xs[:90000], X[:90000])
    Xt, yt = E.TomekLinks().fit_resample(X[:30000], y[:30000])
    assert len(yt) <= 30000 and yt.sum() == y[:30000].sum()
