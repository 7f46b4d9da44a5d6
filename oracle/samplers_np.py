"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg).

numpy restatement of the five imbalanced-learn samplers the reference puts in
``CONFIG_GRID[3]`` (experiment.py:87-94) and calls at experiment.py:463-466.

The arithmetic lives in the third-party package **imbalanced-learn, pinned ==0.9.0**
(reference ``requirements.txt:9``), whose source is neither under /root/reference nor
installed in this image.  What follows restates its published algorithm:

* ``imblearn/over_sampling/_smote/base.py``      SMOTE._fit_resample / _make_samples /
                                                  _generate_samples
* ``imblearn/under_sampling/_prototype_selection/_tomek_links.py``      TomekLinks
* ``imblearn/under_sampling/_prototype_selection/_edited_nearest_neighbours.py`` ENN
* ``imblearn/combine/_smote_enn.py``, ``_smote_tomek.py``

PARITY UNPINNED for this file: the reference ships no golden vectors for these samplers
and imblearn cannot be imported here to generate any.  The
neighbour search itself is delegated to the in-image scikit-learn ``NearestNeighbors``
exactly as imblearn does (``algorithm="auto"``), so the k-NN part *is* the real
dependency.
"""

import numpy as np
from sklearn.neighbors import NearestNeighbors


def _knn(A, Q, K):
    nn = NearestNeighbors(n_neighbors=K)
    nn.fit(A)
    return nn.kneighbors(Q, return_distance=False)


def _minority(y):
    classes, counts = np.unique(y, return_counts=True)
    return classes[np.argmin(counts)], classes[np.argmax(counts)], counts


class SMOTE:
    """sampling_strategy="auto" (== "not majority"), k_neighbors=5."""

    def __init__(self, random_state=None, k_neighbors=5):
        self.random_state = random_state
        self.k_neighbors = k_neighbors

    def fit_resample(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y)
        classes, counts = np.unique(y, return_counts=True)
        n_maj = counts.max()
        X_res, y_res = [X.copy()], [y.copy()]
        for cls, cnt in zip(classes, counts):            # sorted class order
            n_new = int(n_maj - cnt)
            if n_new == 0:
                continue
            idx = np.flatnonzero(y == cls)
            C = X[idx]
            nns = _knn(C, C, self.k_neighbors + 1)[:, 1:]
            rs = np.random.RandomState(self.random_state)   # check_random_state(int)
            samples_indices = rs.randint(low=0, high=nns.size, size=n_new)
            steps = rs.uniform(size=n_new)[:, np.newaxis]
            rows = np.floor_divide(samples_indices, nns.shape[1])
            cols = np.mod(samples_indices, nns.shape[1])
            diffs = C[nns[rows, cols]] - C[rows]
            X_new = C[rows] + steps * diffs
            X_res.append(X_new.astype(X.dtype))
            y_res.append(np.full(n_new, fill_value=cls, dtype=y.dtype))
        return np.vstack(X_res), np.hstack(y_res)


def _classes_to_clean(y, strategy):
    classes, counts = np.unique(y, return_counts=True)
    if strategy == "all":
        return list(classes)
    # "auto" == "not minority" for cleaning samplers
    cmin = classes[np.argmin(counts)]
    return [c for c in classes if c != cmin]


class TomekLinks:
    def __init__(self, sampling_strategy="auto"):
        self.sampling_strategy = sampling_strategy

    def fit_resample(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y)
        nns = _knn(X, X, 2)[:, 1]
        clean = _classes_to_clean(y, self.sampling_strategy)
        in_clean = np.isin(y, clean)
        links = in_clean & (y[nns] != y) & (nns[nns] == np.arange(len(y)))
        keep = np.flatnonzero(np.logical_not(links))
        self.sample_indices_ = keep
        return X[keep], y[keep]


class EditedNearestNeighbours:
    """n_neighbors=3, kind_sel="all"."""

    def __init__(self, sampling_strategy="auto", n_neighbors=3):
        self.sampling_strategy = sampling_strategy
        self.n_neighbors = n_neighbors

    def fit_resample(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y)
        nn = NearestNeighbors(n_neighbors=self.n_neighbors + 1)
        nn.fit(X)
        clean = _classes_to_clean(y, self.sampling_strategy)
        idx_under = np.empty((0,), dtype=int)
        for cls in np.unique(y):
            I = np.flatnonzero(y == cls)
            if cls in clean:
                nb = nn.kneighbors(X[I], return_distance=False)[:, 1:]
                ok = np.all(y[nb] == cls, axis=1)
                I = I[np.flatnonzero(ok)]
            idx_under = np.concatenate((idx_under, I), axis=0)
        self.sample_indices_ = idx_under
        return X[idx_under], y[idx_under]


class SMOTEENN:
    def __init__(self, random_state=None):
        self.random_state = random_state

    def fit_resample(self, X, y):
        Xs, ys = SMOTE(random_state=self.random_state).fit_resample(X, y)
        return EditedNearestNeighbours(sampling_strategy="all").fit_resample(Xs, ys)


class SMOTETomek:
    def __init__(self, random_state=None):
        self.random_state = random_state

    def fit_resample(self, X, y):
        Xs, ys = SMOTE(random_state=self.random_state).fit_resample(X, y)
        return TomekLinks(sampling_strategy="all").fit_resample(Xs, ys)
