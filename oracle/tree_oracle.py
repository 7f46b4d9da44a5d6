"""ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/tree_oracle.c (plain-C restatement
of scikit-learn's DecisionTree / RandomForest / ExtraTrees fit + predict for the regime the
reference uses: binary target, gini, max_depth=None, min_samples_split=2, sqrt / all features)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtree_oracle.so")
KIND = {"DT": 0, "RF": 1, "ET": 2}


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _lib():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "tree_oracle.c")):
        build()
    L = ctypes.CDLL(_SO)
    L.oracle_forest_fit.restype = ctypes.c_void_p
    L.oracle_forest_fit.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_uint32]
    L.oracle_tree_node_count.restype = ctypes.c_int64
    L.oracle_tree_node_count.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.oracle_tree_export.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 8
    L.oracle_forest_predict.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.oracle_forest_free.argtypes = [ctypes.c_void_p]
    L.oracle_randint.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
    return L


class OracleForest:
    def __init__(self, kind, n_estimators=100, random_state=0):
        self.kind, self.n_estimators, self.seed = KIND[kind], n_estimators, random_state
        self._h, self._L = None, _lib()

    def fit(self, X, y):
        X32 = np.ascontiguousarray(X, dtype=np.float32)
        self.classes_, yi = np.unique(np.asarray(y), return_inverse=True)
        yi = np.ascontiguousarray(yi.astype(np.uint8))
        n, d = X32.shape
        mf = d if self.kind == 0 else max(1, int(np.sqrt(d)))
        self.d = d
        self._h = self._L.oracle_forest_fit(X32.ctypes.data, yi.ctypes.data, n, d, self.kind, self.n_estimators, mf,
                                            self.seed)
        self.n_trees = 1 if self.kind == 0 else self.n_estimators
        return self

    def tree(self, t):
        c = int(self._L.oracle_tree_node_count(self._h, t))
        a = {"children_left": np.zeros(c, np.int64), "children_right": np.zeros(c, np.int64),
             "feature": np.zeros(c, np.int64), "threshold": np.zeros(c, np.float64),
             "impurity": np.zeros(c, np.float64), "n_node_samples": np.zeros(c, np.int64),
             "weighted_n_node_samples": np.zeros(c, np.float64), "value": np.zeros((c, 2), np.float64)}
        self._L.oracle_tree_export(self._h, t, *[a[k].ctypes.data for k in (
            "children_left", "children_right", "feature", "threshold", "impurity", "n_node_samples",
            "weighted_n_node_samples", "value")])
        return a

    def predict(self, X):
        X32 = np.ascontiguousarray(X, dtype=np.float32)
        out = np.zeros(X32.shape[0], dtype=np.uint8)
        self._L.oracle_forest_predict(self._h, X32.ctypes.data, X32.shape[0], out.ctypes.data)
        return self.classes_.take(out.astype(np.intp))

    def __del__(self):
        if self._h:
            self._L.oracle_forest_free(self._h)


def randint(seed, hi, n):
    out = np.zeros(n, dtype=np.uint32)
    _lib().oracle_randint(seed, hi, n, out.ctypes.data)
    return out
