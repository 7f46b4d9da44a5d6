"""Host-side (numpy) front of the ``scores`` hot path: rows A1-A3 of SURVEY.md section 8(a).

These three steps are O(N*d), run 2-12 times for the whole 216-config grid and feed the
device path, so they stay on the host (component 9).  They restate,
numpy-op for numpy-op, what the reference gets from scikit-learn, so that the float64
matrices handed to the CUDA kernels are bit-identical to the reference's:

* ``parse_tests`` / ``feat_lab_proj``  <- ``load_feat_lab_proj`` experiment.py:410-427
* ``StandardScaler`` / ``ScalePCA``     <- ``CONFIG_GRID[2]`` experiment.py:82-86, used at :452-453
* ``stratified_kfold_test_folds``      <- ``StratifiedKFold(10, shuffle=True, random_state=0)``
                                          experiment.py:450,458
* ``stratified_group_kfold_test_folds`` <- the 5-fold project-grouped variant of BASELINE.json configs[1]

Provenance: ``StandardScaler``, ``ScalePCA`` and the two fold-map functions are numpy
TRANSLITERATIONS of scikit-learn 1.9.0 internals (``preprocessing/_data.py``, ``decomposition/_pca.py``
covariance_eigh path, ``model_selection/_split.py``) - third-party code of the reference's dependency,
not of the reference itself - kept statement for statement (some of scikit-learn's local names
survive) because bit-identical folds and feature matrices are what the count parity rests on.
"""

import json

import numpy as np

NON_FLAKY, OD_FLAKY, FLAKY = 0, 1, 2

FEATURE_NAMES = (
    "Covered Lines", "Covered Changes", "Source Covered Lines",
    "Execution Time", "Read Count", "Write Count", "Context Switches",
    "Max. Threads", "Max. Memory", "AST Depth", "Assertions",
    "External Modules", "Halstead Volume", "Cyclomatic Complexity",
    "Test Lines of Code", "Maintainability"
)

FLAKY_TYPES = {"NOD": FLAKY, "OD": OD_FLAKY}                       # experiment.py:74-77
FEATURE_SETS = {"Flake16": range(len(FEATURE_NAMES)),               # experiment.py:78-81
                "FlakeFlagger": (0, 1, 2, 3, 10, 11, 14)}


# ----------------------------------------------------------------------------- A1
def parse_tests(path):
    """Parses tests.json ONCE into (all_features list-of-lists, raw labels, projects).

    The reference re-parses the file in every one of the 216 ``get_scores`` calls
    (experiment.py:449 -> :411-412); the parse result does not depend on the config, so
    the grid engine calls this once and derives the 4 (label, feature-set) views below.
    """
    fast = _parse_tests_native(path)
    if fast is not None:
        return fast
    with open(path, "r") as fd:
        tests = json.load(fd)
    return tests_to_arrays(tests)


def _parse_tests_native(path):
    """The same three arrays through the library's one-pass scanner (f16_tests_parse: strtod for
    every number, so the float64 values are the ones ``json.load`` yields); None if the library is
    not built or the file is not in the plain ``write_tests`` format - the caller then uses json."""
    import ctypes
    try:
        from . import _lib
        L = _lib.lib()
    except Exception:
        return None
    h = ctypes.c_void_p()
    if L.f16_tests_parse(str(path).encode(), ctypes.byref(h)) != 0:
        return None
    try:
        n, cols, n_proj = L.f16_tests_rows(h), L.f16_tests_cols(h), L.f16_tests_projects(h)
        if n == 0 or cols < 3:
            return None
        values = np.empty((n, cols), dtype=np.float64)
        proj = np.empty(n, dtype=np.int32)
        names = ctypes.create_string_buffer(max(1, L.f16_tests_names_bytes(h)))
        if L.f16_tests_copy(h, values.ctypes.data, proj.ctypes.data, names) != 0:
            return None
    finally:
        L.f16_tests_free(h)
    labels = values[:, 1]
    if not np.array_equal(labels, np.floor(labels)):
        return None
    name_list = names.raw[:-1].decode("utf-8").split("\0") if n_proj else []
    if len(name_list) != n_proj:
        return None
    # same objects as tests_to_arrays: C-contiguous float64 features, int64 labels, <U project names
    return (np.ascontiguousarray(values[:, 2:]), labels.astype(np.int64), np.array(name_list)[proj])


def tests_to_arrays(tests):
    features, labels, projects = [], [], []
    for proj, tests_proj in tests.items():                         # experiment.py:416
        projects += [proj] * len(tests_proj)                       # :417
        for (_, label_nid, *features_nid) in tests_proj.values():  # :419
            features.append(features_nid)                          # :420
            labels.append(label_nid)                               # :421
    return np.array(features), np.array(labels), np.array(projects)


def feat_lab_proj(parsed, flaky_label, feature_set):
    """experiment.py:423-427.  NB: the fancy column index makes the result
    F-ordered, which decides numpy's summation order in the scaler below; keep the
    expression identical to the reference's."""
    all_features, raw_labels, projects = parsed
    features = all_features[:, feature_set]                        # :423
    labels = raw_labels == flaky_label                             # :424
    return features, labels, projects


# ----------------------------------------------------------------------------- A3
class StandardScaler:
    """``sklearn.preprocessing.StandardScaler().fit_transform`` restated
    (first-call path of ``_incremental_mean_and_var``, sklearn/utils/extmath.py,
    + ``_is_constant_feature`` / ``_handle_zeros_in_scale``, preprocessing/_data.py).
    Dense float64, no NaN, no sample weights: the only regime experiment.py:453 uses."""

    def fit_transform(self, X):
        X = np.asarray(X, dtype=np.float64)
        n = X.shape[0]
        new_sum = np.sum(X, axis=0)
        cnt = np.asarray(n, dtype=np.float64) - np.sum(np.isnan(X).astype(X.dtype), axis=0)
        self.mean_ = (0.0 + new_sum) / cnt
        T = new_sum / cnt
        temp = X - T
        correction = np.sum(temp, axis=0)
        temp **= 2
        unnorm = np.sum(temp, axis=0)
        unnorm -= correction ** 2 / cnt
        self.var_ = unnorm / cnt
        eps = np.finfo(np.float64).eps
        constant = self.var_ <= n * eps * self.var_ + (n * self.mean_ * eps) ** 2
        scale = np.sqrt(self.var_)
        scale[constant] = 1.0
        self.scale_ = scale
        Xt = np.array(X, copy=True)     # keeps X's (F) layout, like check_array(copy=True)
        Xt -= self.mean_
        Xt /= self.scale_
        return Xt


class PCA:
    """``sklearn.decomposition.PCA(random_state=0).fit_transform`` restated for the
    solver scikit-learn 1.9 picks for tall inputs (``covariance_eigh``,
    decomposition/_pca.py ``_fit_full``; ``svd_flip(u_based_decision=False)``;
    ``_transform`` in decomposition/_base.py).  All components are kept
    (``n_components=None``), so the output has d columns."""

    def fit_transform(self, X):
        n, d = X.shape
        if not (d <= 1000 and n >= 10 * d):
            raise ValueError("PCA restatement covers the tall (covariance_eigh) regime only")
        self.mean_ = np.reshape(np.asarray(np.mean(X, axis=0)), (-1,))
        C = X.T @ X
        C -= n * np.reshape(self.mean_, (-1, 1)) * np.reshape(self.mean_, (1, -1))
        C /= n - 1
        eigenvals, eigenvecs = np.linalg.eigh(C)
        eigenvals = np.flip(np.reshape(np.asarray(eigenvals), (-1,)), axis=0)
        eigenvecs = np.flip(np.asarray(eigenvecs), axis=1)
        eigenvals[eigenvals < 0.0] = 0.0
        Vt = eigenvecs.T
        max_abs_v_rows = np.argmax(np.abs(Vt), axis=1)
        shift = np.arange(Vt.shape[0])
        indices = max_abs_v_rows + shift * Vt.shape[1]
        signs = np.sign(np.take(np.reshape(Vt, (-1,)), indices, axis=0))
        Vt *= signs[:, np.newaxis]
        self.components_ = np.array(Vt[:d, :], copy=True)
        self.explained_variance_ = eigenvals
        Xt = X @ self.components_.T
        Xt -= np.reshape(self.mean_, (1, -1)) @ self.components_.T
        return Xt


class ScalePCA:
    """``Pipeline([("s", StandardScaler()), ("p", PCA(random_state=0))])`` experiment.py:85."""

    def fit_transform(self, X):
        return PCA().fit_transform(StandardScaler().fit_transform(X))


PREPROCESSINGS = {"None": None, "Scaling": StandardScaler, "PCA": ScalePCA}


def preprocess(features, name):
    cls = PREPROCESSINGS[name]
    return features if cls is None else cls().fit_transform(features)


# ----------------------------------------------------------------------------- A2
def stratified_kfold_test_folds(labels, n_splits=10, shuffle=True, random_state=0):
    """``StratifiedKFold._make_test_folds`` (sklearn/model_selection/_split.py) restated.
    Returns test_folds int[N]; fold i's test set is ``flatnonzero(test_folds == i)`` and
    its training set ``flatnonzero(test_folds != i)``, both ascending, as sklearn yields."""
    y = np.asarray(labels)
    rng = np.random.RandomState(random_state) if shuffle else None   # legacy MT19937
    _, y_idx, y_inv = np.unique(y, return_index=True, return_inverse=True)
    _, class_perm = np.unique(y_idx, return_inverse=True)            # order of appearance
    y_encoded = class_perm[y_inv]
    n_classes = len(y_idx)
    y_counts = np.bincount(y_encoded)
    if np.all(n_splits > y_counts):
        raise ValueError("n_splits=%d cannot be greater than the number of members in "
                         "each class." % n_splits)
    y_order = np.sort(y_encoded)
    allocation = np.asarray([np.bincount(y_order[i::n_splits], minlength=n_classes)
                             for i in range(n_splits)])
    test_folds = np.empty(len(y), dtype="i")
    for k in range(n_classes):
        folds_for_class = np.arange(n_splits).repeat(allocation[:, k])
        if shuffle:
            rng.shuffle(folds_for_class)
        test_folds[y_encoded == k] = folds_for_class
    return test_folds


def stratified_group_kfold_test_folds(labels, groups, n_splits=5, shuffle=True, random_state=0):
    """``StratifiedGroupKFold._iter_test_indices`` (sklearn/model_selection/_split.py, 1.9.0)
    restated: the 5-fold project-grouped variant BASELINE.json's configs[1] names (the reference
    itself uses the 10-fold splitter above).  Groups (projects) are dealt, in
    order of decreasing spread of their class counts, to the fold that keeps the per-class
    distribution over folds most even; ties go to the emptier fold.  Returns test_folds int[N]."""
    y = np.asarray(labels)
    _, y_inv, y_cnt = np.unique(y, return_inverse=True, return_counts=True)
    if np.all(n_splits > y_cnt):
        raise ValueError("n_splits=%d cannot be greater than the number of members in each class." % n_splits)
    n_classes = len(y_cnt)
    _, groups_inv, groups_cnt = np.unique(groups, return_inverse=True, return_counts=True)
    n_groups = len(groups_cnt)
    if n_splits > n_groups:
        raise ValueError("Cannot have number of splits n_splits=%d greater than the number of groups: %d."
                         % (n_splits, n_groups))
    per_group = np.zeros((n_groups, n_classes))
    np.add.at(per_group, (groups_inv, y_inv), 1)
    if shuffle:
        perm = np.arange(n_groups)
        np.random.RandomState(random_state).shuffle(perm)          # legacy MT19937
        per_group = per_group[perm]
        inv_perm = np.empty_like(perm)
        inv_perm[perm] = np.arange(perm.size)
        groups_inv = inv_perm[groups_inv]
    per_fold = np.zeros((n_splits, n_classes))
    fold_of_group = np.empty(n_groups, dtype="i")
    for g in np.argsort(-np.std(per_group, axis=1), kind="stable"):
        best, min_eval, min_samples = None, np.inf, np.inf
        for i in range(n_splits):
            per_fold[i] += per_group[g]
            fold_eval = np.mean(np.std(per_fold / y_cnt.reshape(1, -1), axis=0))
            per_fold[i] -= per_group[g]
            samples = np.sum(per_fold[i])
            if fold_eval < min_eval or (np.isclose(fold_eval, min_eval) and samples < min_samples):
                best, min_eval, min_samples = i, fold_eval, samples
        per_fold[best] += per_group[g]
        fold_of_group[g] = best
    return fold_of_group[groups_inv]


def kfold_split(test_folds, n_splits=10):
    for i in range(n_splits):
        mask = test_folds == i
        yield np.flatnonzero(~mask), np.flatnonzero(mask)


# ----------------------------------------------------------------------------- A14 (host half)
def div_none(a, b):                                                 # experiment.py:430-431
    return a / b if b else None


def get_prf(fp, fn, tp):                                            # experiment.py:434-443
    p = div_none(tp, tp + fp)
    r = div_none(tp, tp + fn)
    if p is None or r is None:
        f = None
    else:
        f = div_none(2 * p * r, p + r)
    return p, r, f
