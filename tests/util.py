"""Shared helpers for the parity tests."""
import numpy as np

from flake16_framework_b200 import hostprep as hp
from flake16_framework_b200 import synth


def make_dataset(n, seed=16, flaky="NOD", fset="Flake16", prep="None", const_col=None):
    tests = synth.make_tests_dict(n, seed)
    parsed = hp.tests_to_arrays(tests)
    X, y, proj = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES[flaky], hp.FEATURE_SETS[fset])
    if const_col is not None:
        X = X.copy()
        X[:, const_col] = 0.0
    X = hp.preprocess(X, prep)
    return np.ascontiguousarray(X), y, proj


def tree_arrays_sklearn(est):
    t = est.tree_
    return {
        "children_left": t.children_left.astype(np.int64), "children_right": t.children_right.astype(np.int64),
        "feature": t.feature.astype(np.int64), "threshold": t.threshold.astype(np.float64),
        "impurity": t.impurity.astype(np.float64), "n_node_samples": t.n_node_samples.astype(np.int64),
        "weighted_n_node_samples": t.weighted_n_node_samples.astype(np.float64),
        "value": t.value[:, 0, :].astype(np.float64),
    }


def compare_trees(ours, ref, label=""):
    """Returns a list of human-readable mismatch descriptions (empty == identical)."""
    errs = []
    if len(ours["feature"]) != len(ref["feature"]):
        errs.append("%s node_count ours=%d ref=%d" % (label, len(ours["feature"]), len(ref["feature"])))
    m = min(len(ours["feature"]), len(ref["feature"]))
    for key in ("feature", "children_left", "children_right", "n_node_samples", "threshold",
                "weighted_n_node_samples", "impurity", "value"):
        a, b = ours[key][:m], ref[key][:m]
        if key == "value" and b.shape[1] < a.shape[1]:
            # single-class training set: sklearn stores one column, ours always two (second is 0)
            if np.any(a[:, b.shape[1]:] != 0):
                errs.append("%s value: extra class column is not zero" % label)
            a = a[:, :b.shape[1]]
        if key in ("threshold", "impurity", "value", "weighted_n_node_samples"):
            bad = a.view(np.int64) != b.view(np.int64)
            if bad.ndim > 1:
                bad = bad.any(axis=1)
        else:
            bad = a != b
        if bad.any():
            i = int(np.flatnonzero(bad)[0])
            errs.append("%s %s first mismatch at node %d: ours=%r ref=%r (%d/%d nodes differ)"
                        % (label, key, i, a[i], b[i], int(bad.sum()), m))
    return errs
