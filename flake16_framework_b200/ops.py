"""Device-level operators: thin Python wrappers over the C ABI (include/f16.h).

Every function takes / returns CUDA torch tensors (used only as device-memory containers)
and launches on torch's current stream.  Nothing here computes on the CPU and nothing falls
back: a missing library or a failing call raises ``F16Error``.
"""

import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib
from ._lib import check

KIND_DT, KIND_RF, KIND_ET = 0, 1, 2
F16_ERR_OVERFLOW = -3


class F16Overflow(_lib.F16Error):
    """A device-side capacity was exceeded (node array of f16_forest_fit_cap, DFS stack, bootstrap weight)."""


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ready(device=None):
    if not torch.cuda.is_available():
        raise _lib.F16Error("a CUDA device is required (no CPU fallback)")
    dev = torch.cuda.current_device() if device is None else device
    _lib.init(dev)
    return _lib.lib()


def padded_dim(d):
    return 8 if d <= 8 else 16


# ----------------------------------------------------------------------------- staging
def rows_f32(X64, idx=None):
    """float64 [N,d] (row-major, device) -> float32 [n,dp] rows, optionally gathered by idx."""
    L = _ready()
    assert X64.dtype == torch.float64 and X64.is_contiguous() and X64.dim() == 2
    d = X64.shape[1]
    n = X64.shape[0] if idx is None else idx.shape[0]
    if idx is not None:
        assert idx.dtype == torch.int64 and idx.is_contiguous()
    out = torch.empty((n, padded_dim(d)), dtype=torch.float32, device=X64.device)
    check(L.f16_gather_rows_f32(_ptr(X64), d, _ptr(idx), n, _ptr(out), _stream()))
    return out


def gather_rows_f64(X64, idx):
    L = _ready()
    assert X64.dtype == torch.float64 and X64.is_contiguous() and idx.dtype == torch.int64
    out = torch.empty((idx.shape[0], X64.shape[1]), dtype=torch.float64, device=X64.device)
    check(L.f16_gather_rows_f64(_ptr(X64), X64.shape[1], _ptr(idx), idx.shape[0], _ptr(out), _stream()))
    return out


def gather_u8(y, idx):
    L = _ready()
    assert y.dtype == torch.uint8 and idx.dtype == torch.int64
    out = torch.empty((idx.shape[0],), dtype=torch.uint8, device=y.device)
    check(L.f16_gather_u8(_ptr(y), _ptr(idx), idx.shape[0], _ptr(out), _stream()))
    return out


def gather_i32(v, idx):
    L = _ready()
    assert v.dtype == torch.int32 and idx.dtype == torch.int64
    out = torch.empty((idx.shape[0],), dtype=torch.int32, device=v.device)
    check(L.f16_gather_i32(_ptr(v), _ptr(idx), idx.shape[0], _ptr(out), _stream()))
    return out


def argsort_columns(Xrow, d):
    L = _ready()
    n = Xrow.shape[0]
    assert Xrow.dtype == torch.float32 and Xrow.shape[1] == padded_dim(d) and Xrow.is_contiguous()
    out = torch.empty((d, n), dtype=torch.int32, device=Xrow.device)
    check(L.f16_argsort_columns(_ptr(Xrow), n, d, _ptr(out), _stream()))
    return out


def zero_bytes(t):
    """Device memset of a contiguous tensor's storage (f16_fill_u8)."""
    L = _ready()
    assert t.is_contiguous()
    check(L.f16_fill_u8(_ptr(t), 0, t.numel() * t.element_size(), _stream()))
    return t


# ----------------------------------------------------------------------------- forests
def tree_seeds(seed, kind, n_trees):
    L = _lib.lib()
    ts = np.zeros(n_trees, dtype=np.uint32)
    rr = np.zeros(n_trees, dtype=np.uint32)
    check(L.f16_tree_seeds(seed, kind, n_trees, ts.ctypes.data, rr.ctypes.data))
    return ts, rr


def bootstrap_counts(tree_seed, n, device="cuda"):
    L = _ready()
    ts = np.ascontiguousarray(tree_seed, dtype=np.uint32)
    stride = (n + 3) // 4 * 4
    w = torch.empty((len(ts), stride), dtype=torch.uint8, device=device)
    check(L.f16_bootstrap_counts(ts.ctypes.data, len(ts), n, _ptr(w), _stream()))
    return w[:, :n]


class Forest:
    """Opaque fitted ensemble living on the device."""

    def __init__(self, handle, kind, d, n_train):
        self._h = handle
        self.kind, self.d, self.n_train = kind, d, n_train

    @property
    def n_trees(self):
        return _lib.lib().f16_forest_n_trees(self._h)

    def predict(self, Xrow):
        L = _lib.lib()
        assert Xrow.dtype == torch.float32 and Xrow.shape[1] == padded_dim(self.d) and Xrow.is_contiguous()
        pred = torch.empty((Xrow.shape[0],), dtype=torch.uint8, device=Xrow.device)
        check(L.f16_forest_predict(self._h, _ptr(Xrow), Xrow.shape[0], _ptr(pred), _stream()))
        return pred

    def shap_values(self, Xrow, klass=0):
        """Path-dependent TreeSHAP of every row: float64 [n, d] (f16_forest_shap)."""
        L = _lib.lib()
        assert Xrow.dtype == torch.float32 and Xrow.shape[1] == padded_dim(self.d) and Xrow.is_contiguous()
        phi = torch.empty((Xrow.shape[0], self.d), dtype=torch.float64, device=Xrow.device)
        check(L.f16_forest_shap(self._h, _ptr(Xrow), Xrow.shape[0], int(klass), _ptr(phi), _stream()))
        return phi

    def status(self):
        """Synchronises the current stream; raises if the device-side fit failed."""
        rc = _lib.lib().f16_forest_status(self._h, _stream())
        if rc == F16_ERR_OVERFLOW:
            raise F16Overflow(_lib.lib().f16_last_error().decode())
        check(rc)

    def max_nodes(self):
        """Largest node count of the forest's trees (valid after ``status``)."""
        return int(_lib.lib().f16_forest_max_nodes(self._h))

    def node_counts(self):
        c = np.zeros(self.n_trees, dtype=np.int32)
        check(_lib.lib().f16_forest_node_counts(self._h, c.ctypes.data, _stream()))
        return c

    def export_tree(self, t, n_nodes=None):
        """sklearn ``tree_``-layout arrays of tree t (parity tests)."""
        if n_nodes is None:
            n_nodes = int(self.node_counts()[t])
        a = {
            "children_left": np.zeros(n_nodes, np.int64), "children_right": np.zeros(n_nodes, np.int64),
            "feature": np.zeros(n_nodes, np.int64), "threshold": np.zeros(n_nodes, np.float64),
            "impurity": np.zeros(n_nodes, np.float64), "n_node_samples": np.zeros(n_nodes, np.int64),
            "weighted_n_node_samples": np.zeros(n_nodes, np.float64),
            "value": np.zeros((n_nodes, 2), np.float64),
        }
        check(_lib.lib().f16_forest_export(
            self._h, t, n_nodes, a["children_left"].ctypes.data, a["children_right"].ctypes.data,
            a["feature"].ctypes.data, a["threshold"].ctypes.data, a["impurity"].ctypes.data,
            a["n_node_samples"].ctypes.data, a["weighted_n_node_samples"].ctypes.data,
            a["value"].ctypes.data, _stream()))
        return a

    def free(self):
        if self._h is not None:
            _lib.lib().f16_forest_free(self._h, _stream())
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def resolve_max_features(kind, d):
    # sklearn: forests "sqrt" -> max(1, int(sqrt(d))); single tree None -> d
    return d if kind == KIND_DT else max(1, int(np.sqrt(d)))


def forest_fit(Xrow, y, d, kind, n_estimators=100, seed=0, sorted_idx=None, max_features=None, node_cap=0):
    """``node_cap``: per-tree node capacity (0 = worst case 2n - 1); see f16_forest_fit_cap."""
    L = _ready()
    n = Xrow.shape[0]
    assert Xrow.dtype == torch.float32 and Xrow.shape[1] == padded_dim(d) and Xrow.is_contiguous()
    assert y.dtype == torch.uint8 and y.shape[0] == n and y.is_contiguous()
    if kind != KIND_ET and sorted_idx is None:
        sorted_idx = argsort_columns(Xrow, d)
    if max_features is None:
        max_features = resolve_max_features(kind, d)
    h = ctypes.c_void_p()
    check(L.f16_forest_fit_cap(_ptr(Xrow), _ptr(y), n, d, _ptr(sorted_idx), kind, n_estimators, max_features,
                               seed, int(node_cap), _stream(), ctypes.byref(h)))
    return Forest(h, kind, d, n)


# ----------------------------------------------------------------------------- k-NN + samplers
_COL_ORDER = threading.local()


class column_order:
    """Context manager: coordinate accumulation order for k-NN calls made inside it (the grid
    engine passes each dataset's columns sorted by descending variance)."""

    def __init__(self, order):
        """``order``: None, a column permutation, or the (permutation, prefix_test) pair that
        ``variance_order`` returns."""
        prefix = 0
        if isinstance(order, tuple):
            order, prefix = order
        self.order = None if order is None else (np.ascontiguousarray(order, dtype=np.int32), int(prefix))

    def __enter__(self):
        self.prev = getattr(_COL_ORDER, "v", None)
        _COL_ORDER.v = self.order

    def __exit__(self, *a):
        _COL_ORDER.v = self.prev


def variance_order(X):
    """(columns of a host matrix by descending variance, filter mode) for ``f16_knn``:
    mode 1 - float64 filter with the early exit on the two leading columns, when they carry most
             of the total variance (raw, unscaled features);
    mode 3 - tensor-core candidate filter + exact float64 selection, for data of moderate range
             (StandardScaler / PCA outputs; every |value| <= 3e4 so the float16 x 3 split holds);
    mode 0 - plain float64 filter otherwise.
    (mode 2, a float32 filter for centred data, exists in the library and is covered by the parity
    tests, but is never selected: measured no faster than mode 0 on B200 - a 3-register FFMA
    issues at the same 2 cycles per warp as a DFMA - and slower on SMOTE'd data, whose dense
    synthetic clusters defeat a 1e-6-relative filter.)
    The modes only differ in speed, never in result."""
    X = np.asarray(X, dtype=np.float64)
    var = np.var(X, axis=0)
    order = np.argsort(-var, kind="stable").astype(np.int32)
    tot = float(var.sum())
    if tot > 0 and float(var[order[:2]].sum()) / tot > 0.8:
        return order, 1
    if X.size and float(np.abs(X).max()) <= 3.0e4:
        return order, 3
    return order, 0


def calibrate_knn(X, col_order, k=4, n_queries=32768):
    """Picks the k-NN strategy for a device matrix by MEASURING it: ``variance_order``'s static
    choice against the plain float64 search on a strided sample of queries (all strategies return
    the same neighbours, so this is a pure speed decision; e.g. uncentred raw columns make the
    tensor-core filter's norm-relative band useless and its candidate lists overflow).
    Synchronises; meant for the one-off preparation of a dataset."""
    order, mode = col_order
    forced = os.environ.get("F16_KNN_STRATEGY")          # e.g. under a profiler, whose per-launch overhead skews the timing
    if forced is not None:
        return order, int(forced)
    if mode == 0 or X.shape[0] < 2 * n_queries:
        return order, mode
    stride = X.shape[0] // n_queries
    Q = X[::stride][:n_queries].contiguous()
    best, best_ms = mode, None
    for m in ((mode, 6, 0) if mode == 1 else (mode, 0)):      # 6 = sorted sweep, only where one column dominates
        # strategy 3 falls back to 0 below its size threshold: force it (4) so the sample tells
        cm = (order, 4 if m == 3 else m)
        knn(X, Q, k, cm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        knn(X, Q, k, cm)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        if best_ms is None or ms < best_ms:
            best, best_ms = m, ms
    return order, best


def knn_tc_probe(A, Q, umma=False):
    """Largest observed relative error of the tensor-core distance estimate (test hook);
    ``umma`` selects the tcgen05 / TMEM implementation instead of the mma.sync one."""
    L = _ready()
    err = ctypes.c_float(0.0)
    check((L.f16_knn_umma_probe if umma else L.f16_knn_tc_probe)(_ptr(A), A.shape[0], _ptr(Q), Q.shape[0], A.shape[1], ctypes.byref(err), _stream()))
    return float(err.value)


def knn(A, Q, k, col_order=None):
    L = _ready()
    assert A.dtype == torch.float64 and Q.dtype == torch.float64 and A.is_contiguous() and Q.is_contiguous()
    out = torch.empty((Q.shape[0], k), dtype=torch.int32, device=A.device)
    if col_order is None:
        col_order = getattr(_COL_ORDER, "v", None)
    co, prefix = None, 0
    if col_order is not None:
        if isinstance(col_order, tuple):
            col_order, prefix = col_order
        co = np.ascontiguousarray(col_order, dtype=np.int32)
        assert co.shape[0] == A.shape[1]
    check(L.f16_knn(_ptr(A), A.shape[0], _ptr(Q), Q.shape[0], A.shape[1], k,
                    ctypes.c_void_p(co.ctypes.data) if co is not None else None, int(prefix), _ptr(out), _stream()))
    return out


def smote(X, y, n_min, n_maj, minority=1, seed=0, k_neighbors=5, idx_min=None):
    """SMOTE(random_state=seed).fit_resample on device.  X float64 [n,d], y uint8 [n].
    ``n_min`` / ``n_maj`` are the class counts (known on the host from the fold map, so no
    device->host read is needed).  The MT19937 draws (randint then uniform, imblearn
    ``_make_samples``) are tiny and made on the host with numpy's legacy RandomState."""
    L = _ready()
    n, d = X.shape
    n_new = int(n_maj - n_min)
    if n_new == 0:
        return X, y
    if idx_min is None:
        idx_min = torch.nonzero(y == minority).squeeze(1)      # original order
    C = gather_rows_f64(X, idx_min)
    nn = knn(C, C, k_neighbors + 1)
    rs = np.random.RandomState(seed)
    sample_idx = rs.randint(low=0, high=n_min * k_neighbors, size=n_new)
    steps = rs.uniform(size=n_new)
    sidx_d = torch.from_numpy(sample_idx.astype(np.int64)).to(X.device, non_blocking=True)
    steps_d = torch.from_numpy(steps).to(X.device, non_blocking=True)
    Xout = torch.empty((n + n_new, d), dtype=torch.float64, device=X.device)
    Xout[:n].copy_(X)
    check(L.f16_smote_generate(_ptr(C), n_min, d, _ptr(nn), k_neighbors, _ptr(sidx_d), _ptr(steps_d), n_new,
                               ctypes.c_void_p(Xout.data_ptr() + n * d * 8), _stream()))
    yout = torch.empty((n + n_new,), dtype=torch.uint8, device=X.device)
    yout[:n].copy_(y)
    check(L.f16_fill_u8(ctypes.c_void_p(yout.data_ptr() + n), int(minority), n_new, _stream()))
    return Xout, yout


def _compact(X, y, keep, grouped):
    L = _lib.lib()
    n, d = X.shape
    Xout = torch.empty_like(X)
    yout = torch.empty_like(y)
    src = torch.empty((n,), dtype=torch.int64, device=X.device)
    n_out = torch.empty((2,), dtype=torch.int64, device=X.device)
    check(L.f16_compact_rows(_ptr(X), _ptr(y), _ptr(keep), n, d, grouped, _ptr(Xout), _ptr(yout), _ptr(src),
                             _ptr(n_out), _stream()))
    m = int(n_out[0].item())         # the one device->host read of a resample: its new size
    return Xout[:m], yout[:m], src[:m]


def tomek_links(X, y, clean_mask):
    """TomekLinks.fit_resample; clean_mask bit c <=> class c is cleaned."""
    L = _ready()
    nn = knn(X, X, 2)
    keep = torch.empty((X.shape[0],), dtype=torch.uint8, device=X.device)
    check(L.f16_tomek_keep(_ptr(nn), 2, _ptr(y), X.shape[0], clean_mask, _ptr(keep), _stream()))
    return _compact(X, y, keep, 0)


def enn(X, y, clean_mask, n_neighbors=3):
    """EditedNearestNeighbours(kind_sel="all").fit_resample (class-grouped output order)."""
    L = _ready()
    nn = knn(X, X, n_neighbors + 1)
    keep = torch.empty((X.shape[0],), dtype=torch.uint8, device=X.device)
    check(L.f16_enn_keep(_ptr(nn), n_neighbors + 1, _ptr(y), X.shape[0], clean_mask, _ptr(keep), _stream()))
    return _compact(X, y, keep, 1)


# ----------------------------------------------------------------------------- figures
def spearman(X64):
    """scipy.stats.spearmanr(X).correlation for a device matrix: float64 [d, d]."""
    L = _ready()
    assert X64.dtype == torch.float64 and X64.is_contiguous() and X64.dim() == 2
    rho = torch.empty((X64.shape[1], X64.shape[1]), dtype=torch.float64, device=X64.device)
    check(L.f16_spearman(_ptr(X64), X64.shape[0], X64.shape[1], _ptr(rho), _stream()))
    return rho


# ----------------------------------------------------------------------------- scoring
def confusion(y, pred, proj, n_proj, counts):
    L = _ready()
    assert counts.dtype == torch.int64 and counts.numel() == (n_proj + 1) * 3
    assert proj.dtype == torch.int32 and y.dtype == torch.uint8 and pred.dtype == torch.uint8
    check(L.f16_confusion(_ptr(y), _ptr(pred), _ptr(proj), y.shape[0], n_proj, _ptr(counts), _stream()))
    return counts
