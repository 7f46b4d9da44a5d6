"""Generates the committed golden fixtures by running THE REFERENCE ITSELF in this container.

/root/reference/experiment.py cannot be imported as is (coverage, shap and imbalanced-learn are
not installed, experiment.py:18-28).  This script imports it with those three modules stubbed:
`coverage` / `shap` are never touched by the `scores` path, and `imblearn.*` is served by
oracle/samplers_np.py (numpy restatement of imbalanced-learn 0.9.0).  Everything else - the
CONFIG_GRID, load_feat_lab_proj, get_scores, write_scores, manage_pool - is the reference's own
code, executed unmodified against the in-image scikit-learn 1.9.0 on a synthetic tests.json.

    python tests/golden/make_golden.py            # ~10 min on 8 cores

Outputs (committed):
    scores_n1500_seed16.pkl    {config_keys: (per-project [fp,fn,tp] dict, total [fp,fn,tp])} - 216 configs
    trees_n2000_seed16.npz     tree_ arrays of the first trees of DT / RF / ET (BASELINE configs 1-3)
/root/reference does not exist on the GPU box; only these fixtures travel.
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
REFERENCE = "/root/reference"


def import_reference():
    import samplers_np
    for name in ("coverage", "shap", "imblearn", "imblearn.over_sampling", "imblearn.combine",
                 "imblearn.under_sampling"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["coverage"].numbits = None
    sys.modules["shap"].TreeExplainer = None
    sys.modules["imblearn.over_sampling"].SMOTE = samplers_np.SMOTE
    sys.modules["imblearn.combine"].SMOTEENN = samplers_np.SMOTEENN
    sys.modules["imblearn.combine"].SMOTETomek = samplers_np.SMOTETomek
    sys.modules["imblearn.under_sampling"].TomekLinks = samplers_np.TomekLinks
    sys.modules["imblearn.under_sampling"].EditedNearestNeighbours = samplers_np.EditedNearestNeighbours
    sys.path.insert(0, REFERENCE)
    import experiment          # the reference file, unmodified
    assert experiment.__file__.startswith(REFERENCE)
    return experiment


def main():
    from flake16_framework_b200 import synth
    experiment = import_reference()
    work = tempfile.mkdtemp(prefix="f16golden")
    os.chdir(work)

    # ---- full-grid scores through the reference's own write_scores()
    synth.make_tests_json("tests.json", 1500, 16)
    experiment.write_scores()
    with open("scores.pkl", "rb") as fd:
        scores = pickle.load(fd)
    gold = {}
    for keys, (t_train, t_test, per_proj, total) in scores.items():
        gold[tuple(keys)] = ({str(p): [int(v) for v in s[:3]] for p, s in per_proj.items()},
                             [int(v) for v in total[:3]])
    assert len(gold) == 216
    with open(os.path.join(HERE, "scores_n1500_seed16.pkl"), "wb") as fd:
        pickle.dump(gold, fd, protocol=4)

    # ---- per-tree arrays for BASELINE configs 1-3 (first fold's training set)
    synth.make_tests_json("tests.json", 2000, 16)
    out = {}
    for tag, keys in (("DT", ("NOD", "Flake16", "None", "None", "Decision Tree")),
                      ("RF", ("NOD", "Flake16", "None", "None", "Random Forest")),
                      ("ET", ("NOD", "Flake16", "Scaling", "SMOTE", "Extra Trees"))):
        vals = [experiment.CONFIG_GRID[i][k] for i, k in enumerate(keys)]
        flaky_label, feature_set, preprocessing, balancing, model = vals
        features, labels, _ = experiment.load_feat_lab_proj(flaky_label, feature_set)
        if preprocessing is not None:
            features = preprocessing.fit_transform(features)
        fold = experiment.StratifiedKFold(n_splits=10, shuffle=True, random_state=0)
        train, test = next(iter(fold.split(features, labels)))
        Xtr, ytr = features[train], labels[train]
        if balancing is not None:
            Xtr, ytr = balancing.fit_resample(Xtr, ytr)
        model.fit(Xtr, ytr)
        trees = [model] if tag == "DT" else model.estimators_[:3]
        for t, est in enumerate(trees):
            tr = est.tree_
            for name in ("children_left", "children_right", "feature", "threshold", "n_node_samples",
                         "weighted_n_node_samples", "impurity"):
                out["%s_%d_%s" % (tag, t, name)] = getattr(tr, name)
            out["%s_%d_value" % (tag, t)] = tr.value[:, 0, :]
        out["%s_pred" % tag] = model.predict(features[test])
    np.savez_compressed(os.path.join(HERE, "trees_n2000_seed16.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
