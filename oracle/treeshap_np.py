"""TEST INFRASTRUCTURE - groundwork for SURVEY.md section 8(f) row N3 (`get_shap`, reference
experiment.py:504-530), not part of the product path.

The reference calls ``shap.TreeExplainer(model).shap_values(features)[0]`` on two fitted forests.
``shap`` (pinned 0.40.0 in the reference's requirements.txt) is not installed here and its source
is not under /root/reference, so this file restates the published algorithm it runs in that call:
path-dependent TreeSHAP (Lundberg, Erion, Lee, "Consistent Individualized Feature Attribution for
Tree Ensembles", 2018, Algorithm 2: EXTEND / UNWIND over the root-to-leaf path; the game is
v(S) = E[f(x) | x_S] with the tree's own covers as the conditional distribution, Algorithm 1).

PARITY UNPINNED against the shap package.  What pins it instead (tests/test_treeshap_oracle_cpu.py):
the polynomial algorithm must equal the Shapley values obtained by ENUMERATING all feature subsets
of that game on scikit-learn trees - the definition itself - and satisfy local accuracy
(sum of the attributions = f(x) - E[f]).

Model output, as TreeExplainer takes it for scikit-learn forests: per tree the leaf's class
fraction (``tree_.value`` normalised), covers = ``tree_.weighted_n_node_samples``, the forest's
attribution = mean over trees; ``shap_values(X)[0]`` is class 0.
"""
import itertools
import math

import numpy as np


class _Tree:
    """The arrays of one sklearn tree that the algorithm reads."""

    def __init__(self, est, klass):
        t = est.tree_
        self.left, self.right = t.children_left, t.children_right
        self.feature, self.threshold = t.feature, t.threshold
        self.cover = t.weighted_n_node_samples
        v = t.value[:, 0, :]
        self.value = (v / v.sum(axis=1, keepdims=True))[:, klass]

    def is_leaf(self, j):
        return self.left[j] == -1

    def hot_cold(self, j, x):
        # sklearn's rule: float32(x) <= threshold goes left
        left = np.float32(x[self.feature[j]]) <= self.threshold[j]
        return (self.left[j], self.right[j]) if left else (self.right[j], self.left[j])


# ----------------------------------------------------------------------------- definition
def expected_value(tree, x, subset, j=0):
    """Algorithm 1: E[f(x) | x_S] - follow x on the features in S, average by cover otherwise."""
    if tree.is_leaf(j):
        return tree.value[j]
    if tree.feature[j] in subset:
        hot, _ = tree.hot_cold(j, x)
        return expected_value(tree, x, subset, hot)
    a, b = tree.left[j], tree.right[j]
    return (tree.cover[a] * expected_value(tree, x, subset, a) + tree.cover[b] * expected_value(tree, x, subset, b)) / tree.cover[j]


def shapley_by_enumeration(tree, x, d):
    """phi_i = sum over S of |S|! (d - |S| - 1)! / d! * (v(S + i) - v(S)); exponential in d."""
    phi = np.zeros(d)
    feats = list(range(d))
    for i in feats:
        rest = [f for f in feats if f != i]
        for r in range(d):
            w = math.factorial(r) * math.factorial(d - r - 1) / math.factorial(d)
            for S in itertools.combinations(rest, r):
                S = frozenset(S)
                phi[i] += w * (expected_value(tree, x, S | {i}) - expected_value(tree, x, S))
    return phi


# ----------------------------------------------------------------------------- Algorithm 2
def _extend(path, depth, zero, one, feat):
    """path: list of [feature, zero_fraction, one_fraction, weight]; depth = index to write."""
    path[depth] = [feat, zero, one, 1.0 if depth == 0 else 0.0]
    for i in range(depth - 1, -1, -1):
        path[i + 1][3] += one * path[i][3] * (i + 1) / (depth + 1)
        path[i][3] = zero * path[i][3] * (depth - i) / (depth + 1)


def _unwind(path, depth, k):
    """Removes element k from a path of depth + 1 elements (undoes the EXTEND that added it)."""
    _, zero, one, _ = path[k]
    nxt = path[depth][3]
    for i in range(depth - 1, -1, -1):
        if one != 0.0:
            tmp = path[i][3]
            path[i][3] = nxt * (depth + 1) / ((i + 1) * one)
            nxt = tmp - path[i][3] * zero * (depth - i) / (depth + 1)
        else:
            path[i][3] = path[i][3] * (depth + 1) / (zero * (depth - i))
    for i in range(k, depth):
        path[i][0], path[i][1], path[i][2] = path[i + 1][0], path[i + 1][1], path[i + 1][2]


def _unwound_sum(path, depth, k):
    """Sum of the weights the path would have after _unwind(path, depth, k), without doing it."""
    _, zero, one, _ = path[k]
    nxt = path[depth][3]
    total = 0.0
    for i in range(depth - 1, -1, -1):
        if one != 0.0:
            tmp = nxt * (depth + 1) / ((i + 1) * one)
            total += tmp
            nxt = path[i][3] - tmp * zero * (depth - i) / (depth + 1)
        else:
            total += path[i][3] / zero / ((depth - i) / (depth + 1))
    return total


def _recurse(tree, x, phi, j, depth, parent_path, zero, one, feat):
    path = [list(p) for p in parent_path[:depth]] + [None] * (len(parent_path) - depth)
    _extend(path, depth, zero, one, feat)
    if tree.is_leaf(j):
        for k in range(1, depth + 1):
            w = _unwound_sum(path, depth, k)
            phi[path[k][0]] += w * (path[k][2] - path[k][1]) * tree.value[j]
        return
    hot, cold = tree.hot_cold(j, x)
    f = tree.feature[j]
    in_zero = in_one = 1.0
    k = next((i for i in range(depth + 1) if path[i][0] == f), None)
    if k is not None:                       # the feature was split on higher up: merge the fractions
        in_zero, in_one = path[k][1], path[k][2]
        _unwind(path, depth, k)
        depth -= 1
    _recurse(tree, x, phi, hot, depth + 1, path, in_zero * tree.cover[hot] / tree.cover[j], in_one, f)
    _recurse(tree, x, phi, cold, depth + 1, path, in_zero * tree.cover[cold] / tree.cover[j], 0.0, f)


def tree_shap(tree, x, d):
    """Path-dependent TreeSHAP of one tree for one row: (phi[d], expected value of the tree)."""
    phi = np.zeros(d)
    max_depth = int(_depth(tree, 0)) + 2
    _recurse(tree, x, phi, 0, 0, [None] * max_depth, 1.0, 1.0, -1)
    return phi, expected_value(tree, x, frozenset())


def _depth(tree, j):
    return 0 if tree.is_leaf(j) else 1 + max(_depth(tree, tree.left[j]), _depth(tree, tree.right[j]))


def forest_shap_values(model, X, klass=0):
    """``TreeExplainer(model).shap_values(X)[klass]`` for a fitted sklearn forest or tree:
    float64 [n, d], mean over the trees."""
    ests = getattr(model, "estimators_", [model])
    X = np.asarray(X)
    d = X.shape[1]
    out = np.zeros((X.shape[0], d))
    for est in ests:
        tree = _Tree(est, klass)
        for r in range(X.shape[0]):
            out[r] += tree_shap(tree, X[r], d)[0]
    return out / len(ests)
