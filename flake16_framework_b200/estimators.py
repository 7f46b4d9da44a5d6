"""Drop-in estimator objects for the reference's ``CONFIG_GRID`` (experiment.py:73-100).

The reference only ever calls (experiment.py:453-473, duck typing)

    balancing.fit_resample(X, y) -> (X', y')      model.fit(X, y)      model.predict(X) -> bool[n]

on objects that are re-used across folds and configs and re-seed on every call because
``random_state`` is an int.  These classes keep that protocol (same names, same argument
meaning, same exceptions for bad input) and run every step on the GPU through the C ABI.
Inputs may be numpy arrays (copied host->device, results copied back: the reference-facing
path) or CUDA torch tensors (results stay on the device: what the grid engine uses).
"""

import numpy as np
import torch

from . import ops
from ._lib import F16Error


def _to_dev_f64(X):
    if isinstance(X, torch.Tensor):
        if not X.is_cuda:
            raise F16Error("torch inputs must live on a CUDA device")
        return X.to(torch.float64).contiguous(), True
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim != 2:
        raise ValueError("Expected 2D array, got %dD array instead" % X.ndim)
    return torch.from_numpy(X).cuda(non_blocking=True), False


def _to_dev_y(y):
    """Class labels -> (uint8 index tensor, classes_) as sklearn's np.unique encoding."""
    if isinstance(y, torch.Tensor):
        return y.to(torch.uint8).contiguous(), np.array([False, True])
    y = np.asarray(y)
    classes, yi = np.unique(y, return_inverse=True)
    if len(classes) > 2:
        raise ValueError("only binary targets are supported (got %d classes)" % len(classes))
    return torch.from_numpy(yi.astype(np.uint8)).cuda(non_blocking=True), classes


class _TreeEnsemble:
    _kind = None

    def __init__(self, n_estimators=100, random_state=None):
        self.n_estimators = n_estimators
        self.random_state = random_state
        self._forest = None

    def _seed(self):
        if self.random_state is None or not isinstance(self.random_state, (int, np.integer)):
            raise ValueError("random_state must be an int (the reference uses random_state=0)")
        return int(self.random_state)

    def fit(self, X, y, sorted_idx=None):
        Xd, _ = _to_dev_f64(X)
        yd, self.classes_ = _to_dev_y(y)
        if Xd.shape[0] != yd.shape[0]:
            raise ValueError("Found input variables with inconsistent numbers of samples: [%d, %d]"
                             % (Xd.shape[0], yd.shape[0]))
        self.n_features_in_ = Xd.shape[1]
        Xrow = ops.rows_f32(Xd)
        return self.fit_rows(Xrow, yd, self.n_features_in_, sorted_idx)

    def fit_rows(self, Xrow, y_u8, d, sorted_idx=None):
        """Device fast path: float32 padded rows + uint8 labels already staged."""
        if self._forest is not None:
            self._forest.free()
        self.n_features_in_ = d
        if not hasattr(self, "classes_"):
            self.classes_ = np.array([False, True])
        self._forest = ops.forest_fit(Xrow, y_u8, d, self._kind, self.n_estimators, self._seed(), sorted_idx)
        return self

    def predict_rows(self, Xrow):
        if self._forest is None:
            raise F16Error("This %s instance is not fitted yet." % type(self).__name__)
        return self._forest.predict(Xrow)

    def predict(self, X):
        Xd, on_dev = _to_dev_f64(X)
        if Xd.shape[1] != self.n_features_in_:
            raise ValueError("X has %d features, but %s is expecting %d features as input."
                             % (Xd.shape[1], type(self).__name__, self.n_features_in_))
        pred = self.predict_rows(ops.rows_f32(Xd))
        if on_dev:
            return pred
        self._forest.status()
        return self.classes_.take(pred.cpu().numpy().astype(np.intp))

    def shap_values(self, X):
        """``shap.TreeExplainer(self).shap_values(X)`` (get_shap, experiment.py:517): one float64
        [n, d] array per class, path-dependent TreeSHAP on the device."""
        if self._forest is None:
            raise F16Error("This %s instance is not fitted yet." % type(self).__name__)
        Xd, on_dev = _to_dev_f64(X)
        if Xd.shape[1] != self.n_features_in_:
            raise ValueError("X has %d features, but %s is expecting %d features as input."
                             % (Xd.shape[1], type(self).__name__, self.n_features_in_))
        Xrow = ops.rows_f32(Xd)
        out = [self._forest.shap_values(Xrow, k) for k in (0, 1)]
        return out if on_dev else [o.cpu().numpy() for o in out]

    @property
    def forest_(self):
        return self._forest


class ExtraTreesClassifier(_TreeEnsemble):          # experiment.py:96
    _kind = ops.KIND_ET


class RandomForestClassifier(_TreeEnsemble):        # experiment.py:97
    _kind = ops.KIND_RF


class DecisionTreeClassifier(_TreeEnsemble):        # experiment.py:98
    _kind = ops.KIND_DT

    def __init__(self, random_state=None):
        super().__init__(n_estimators=1, random_state=random_state)


# ----------------------------------------------------------------------------- samplers
def _class_stats(y_u8, counts=None):
    if counts is None:
        c1 = int(y_u8.sum().item())
        counts = (y_u8.shape[0] - c1, c1)
    minority = 0 if counts[0] < counts[1] else 1    # np.argmin tie -> first class, as min() over sorted keys
    if counts[0] == counts[1]:
        minority = 0
    return counts, minority


def _clean_mask(strategy, minority):
    if strategy == "all":
        return 0b11
    if strategy == "auto":                           # "not minority"
        return 0b11 & ~(1 << minority)
    raise ValueError("sampling_strategy must be 'auto' or 'all'")


class _Sampler:
    def fit_resample(self, X, y, counts=None):
        Xd, on_dev = _to_dev_f64(X)
        yd, classes = _to_dev_y(y)
        Xo, yo = self._resample(Xd, yd, counts)
        if on_dev:
            return Xo, yo
        return Xo.cpu().numpy(), classes.take(yo.cpu().numpy().astype(np.intp))


class SMOTE(_Sampler):                               # experiment.py:90
    def __init__(self, random_state=None, k_neighbors=5):
        self.random_state, self.k_neighbors = random_state, k_neighbors

    def _resample(self, X, y, counts=None):
        counts, minority = _class_stats(y, counts)
        n_min, n_maj = min(counts), max(counts)
        if n_min <= self.k_neighbors:
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d"
                             % (self.k_neighbors + 1, n_min))
        return ops.smote(X, y, n_min, n_maj, minority, int(self.random_state), self.k_neighbors)


class TomekLinks(_Sampler):                          # experiment.py:89
    def __init__(self, sampling_strategy="auto"):
        self.sampling_strategy = sampling_strategy

    def _resample(self, X, y, counts=None):
        _, minority = _class_stats(y, counts)
        Xo, yo, self.sample_indices_ = ops.tomek_links(X, y, _clean_mask(self.sampling_strategy, minority))
        return Xo, yo


class EditedNearestNeighbours(_Sampler):             # experiment.py:91
    def __init__(self, sampling_strategy="auto", n_neighbors=3):
        self.sampling_strategy, self.n_neighbors = sampling_strategy, n_neighbors

    def _resample(self, X, y, counts=None):
        _, minority = _class_stats(y, counts)
        Xo, yo, self.sample_indices_ = ops.enn(X, y, _clean_mask(self.sampling_strategy, minority),
                                               self.n_neighbors)
        return Xo, yo


class SMOTEENN(_Sampler):                            # experiment.py:92
    def __init__(self, random_state=None):
        self.random_state = random_state

    def _resample(self, X, y, counts=None):
        Xs, ys = SMOTE(random_state=self.random_state)._resample(X, y, counts)
        return EditedNearestNeighbours(sampling_strategy="all")._resample(Xs, ys, None)


class SMOTETomek(_Sampler):                          # experiment.py:93
    def __init__(self, random_state=None):
        self.random_state = random_state

    def _resample(self, X, y, counts=None):
        Xs, ys = SMOTE(random_state=self.random_state)._resample(X, y, counts)
        return TomekLinks(sampling_strategy="all")._resample(Xs, ys, None)
