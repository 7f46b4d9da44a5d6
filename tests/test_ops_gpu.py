"""GPU parity of the supporting operators (RNG, sort, k-NN, samplers, confusion) through the C ABI."""
import numpy as np
import pytest
import torch

from util import make_dataset

pytestmark = pytest.mark.gpu


def test_bootstrap_counts_match_numpy(cuda):
    """sklearn/ensemble/_forest.py:94-112: RandomState(seed).randint(0, n, n) -> bincount."""
    from flake16_framework_b200 import ops
    for n in (1, 2, 5, 624, 1000, 4096, 90001):
        seeds, _ = ops.tree_seeds(0, ops.KIND_RF, 4)
        w = ops.bootstrap_counts(seeds, n).cpu().numpy()
        for t, s in enumerate(seeds):
            ref = np.bincount(np.random.RandomState(int(s)).randint(0, n, n), minlength=n)
            assert np.array_equal(w[t], ref), (n, t)


def test_argsort_columns(cuda):
    from flake16_framework_b200 import ops
    rs = np.random.RandomState(0)
    for n, d in ((1, 3), (7, 16), (5000, 7), (70001, 16)):
        X = rs.randn(n, d)
        X[:, 0] = rs.randint(0, 5, n)            # heavy ties
        if d > 2:
            X[:, 2] = -np.abs(X[:, 2])           # negatives
        Xd = torch.from_numpy(X).cuda()
        rows = ops.rows_f32(Xd)
        idx = ops.argsort_columns(rows, d).cpu().numpy()
        X32 = X.astype(np.float32)
        for f in range(d):
            assert np.array_equal(np.sort(idx[f]), np.arange(n)), (n, d, f)
            v = X32[idx[f], f]
            assert np.all(v[1:] >= v[:-1]), (n, d, f)


def test_rows_f32_gather(cuda):
    from flake16_framework_b200 import ops
    rs = np.random.RandomState(1)
    for d in (7, 16):
        X = rs.randn(100, d) * 1e3
        idx = rs.randint(0, 100, 37)
        r = ops.rows_f32(torch.from_numpy(X).cuda(), torch.from_numpy(idx).cuda()).cpu().numpy()
        assert r.shape[1] == (8 if d == 7 else 16)
        assert np.array_equal(r[:, :d], X[idx].astype(np.float32))
        assert np.all(r[:, d:] == 0)


@pytest.mark.parametrize("cfg", [dict(n=3000, fset="Flake16", prep="None"), dict(n=3000, fset="Flake16", prep="Scaling"),
                                 dict(n=3000, fset="FlakeFlagger", prep="PCA"), dict(n=20000, fset="Flake16", prep="PCA")])
def test_knn_matches_sklearn(cuda, cfg):
    from sklearn.neighbors import NearestNeighbors
    from flake16_framework_b200 import ops
    X, _, _ = make_dataset(**cfg)
    Xd = torch.from_numpy(X).cuda()
    for k in (2, 4, 6):
        ours = ops.knn(Xd, Xd, k).cpu().numpy()
        ref = NearestNeighbors(n_neighbors=k).fit(X).kneighbors(X, return_distance=False)
        frac = (ours == ref).all(axis=1).mean()
        assert frac == 1.0, "k=%d: %.6f of rows have identical neighbour lists" % (k, frac)
        # column order / prefix early exit are pure scheduling: same result
        order, prefix = ops.variance_order(X)
        # (4 / 5 = tensor-core filter whatever the size, tcgen05 / mma.sync implementation; raw
        #  features do not fit float16 and take the exhaustive float64 path)
        # (6 = sweep over the rows sorted by the leading column of the order)
        for co in ((order, prefix), (order, 1), (order, 2), (order, 4), (order, 5), (order, 6), (order[::-1].copy(), 6), (order[::-1].copy(), 0)):
            assert np.array_equal(ops.knn(Xd, Xd, k, co).cpu().numpy(), ours), (k, co[1])


def _smote_like(rs, n_seed, n_new, d, scale=1.0):
    """Seed points + points interpolated on segments between near seeds (dense 1-d clusters),
    exact duplicates and a few far outliers - the geometry that stresses a distance filter."""
    S = rs.randn(n_seed, d) * scale
    a = rs.randint(0, n_seed, n_new)
    b = (a + rs.randint(1, 4, n_new)) % n_seed
    st = rs.rand(n_new, 1)
    N = S[a] + st * (S[b] - S[a])
    X = np.vstack([S, N, S[:50], N[:50]])                  # duplicates => distance ties
    X[rs.randint(0, len(X), 20)] *= 40.0                     # outliers with large norms
    return np.ascontiguousarray(X)


@pytest.mark.parametrize("impl", [4, 5, 6])       # 4: tcgen05 + TMEM (f16_knn_umma.cu), 5: mma.sync (f16_knn_tc.cu), 6: sorted sweep (f16_knn_sweep.cu)
@pytest.mark.parametrize("d", [16, 7, 3])
def test_knn_tensor_filter_equals_float64_search(cuda, d, impl):
    """Both tensor-core filters must return exactly what the float64 kernel returns."""
    from flake16_framework_b200 import ops
    rs = np.random.RandomState(d)
    X = _smote_like(rs, 1500, 9000, d)                       # 10 600 rows: not a multiple of 128
    Xd = torch.from_numpy(X).cuda()
    Q = torch.from_numpy(np.ascontiguousarray(X[::7] + 1e-3 * rs.randn(len(X[::7]), d))).cuda()
    order = np.arange(d, dtype=np.int32)
    for k in (1, 2, 4, 6, 8):
        ref = ops.knn(Xd, Xd, k, (order, 0)).cpu().numpy()
        got = ops.knn(Xd, Xd, k, (order, impl)).cpu().numpy()
        assert np.array_equal(got, ref), "self-join k=%d: %d rows differ" % (k, (got != ref).any(axis=1).sum())
        refq = ops.knn(Xd, Q, k, (order, 0)).cpu().numpy()
        gotq = ops.knn(Xd, Q, k, (order, impl)).cpu().numpy()
        assert np.array_equal(gotq, refq), "A != Q k=%d" % k
    # tiny and huge (but float16-representable) magnitudes
    for scale in (1e-4, 300.0):
        Y = _smote_like(rs, 300, 1500, d, scale)
        Yd = torch.from_numpy(Y).cuda()
        assert np.array_equal(ops.knn(Yd, Yd, 4, (order, impl)).cpu().numpy(), ops.knn(Yd, Yd, 4, (order, 0)).cpu().numpy()), scale


@pytest.mark.parametrize("impl", [4, 5, 6])
def test_knn_tensor_filter_overflow_and_range_fallbacks(cuda, impl):
    from flake16_framework_b200 import ops
    d = 16
    order = np.arange(d, dtype=np.int32)
    # references ordered by DECREASING distance from every query: each one enters the running
    # top-k, the 256-slot candidate lists overflow and the exhaustive path must take over
    n = 3000
    R = np.zeros((n, d)); R[:, 0] = 50.0 - 0.01 * np.arange(n); R[:, 1] = 1e-3 * np.arange(n)
    Qs = np.zeros((300, d)); Qs[:, 2] = 1e-2 * np.arange(300)
    Rd, Qd = torch.from_numpy(R).cuda(), torch.from_numpy(Qs).cuda()
    assert np.array_equal(ops.knn(Rd, Qd, 4, (order, impl)).cpu().numpy(), ops.knn(Rd, Qd, 4, (order, 0)).cpu().numpy())
    # one coordinate outside the float16 range: device-side detection, exhaustive search
    rs = np.random.RandomState(3)
    X = rs.randn(2000, d); X[17, 5] = 1.0e6
    Xd = torch.from_numpy(X).cuda()
    assert np.array_equal(ops.knn(Xd, Xd, 4, (order, impl)).cpu().numpy(), ops.knn(Xd, Xd, 4, (order, 0)).cpu().numpy())


@pytest.mark.parametrize("umma", [True, False])
def test_knn_tensor_filter_error_bound(cuda, umma):
    """The filter assumes |estimate - d^2| <= 6e-5 (|q|^2 + |x|^2) + slack; measured on the device
    it must stay below a quarter of that on every kind of data the grid produces."""
    from flake16_framework_b200 import ops
    rs = np.random.RandomState(11)
    worst = 0.0
    for cfg in (dict(n=1500, fset="Flake16", prep="Scaling"), dict(n=1500, fset="Flake16", prep="PCA"),
                dict(n=1500, fset="FlakeFlagger", prep="Scaling")):
        X, _, _ = make_dataset(**cfg)
        Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
        worst = max(worst, ops.knn_tc_probe(Xd, Xd, umma))
    # (the tcgen05 filter carries the norm seed in three float16: |x|^2 up to 2.6e8)
    for scale in (1e-3, 1.0, 20.0 if umma else 200.0):
        Y = torch.from_numpy(_smote_like(rs, 300, 900, 16, scale)).cuda()
        worst = max(worst, ops.knn_tc_probe(Y, Y, umma))
    assert worst < 1.5e-5, worst


SAMPLERS = ["TomekLinks", "SMOTE", "EditedNearestNeighbours", "SMOTEENN", "SMOTETomek"]


@pytest.mark.parametrize("name", SAMPLERS)
@pytest.mark.parametrize("cfg", [dict(n=4000, fset="Flake16", prep="None"), dict(n=4000, fset="FlakeFlagger", prep="Scaling"),
                                 dict(n=6000, fset="Flake16", prep="PCA", flaky="OD")])
def test_samplers_match_oracle(cuda, name, cfg):
    import samplers_np as O
    from flake16_framework_b200 import estimators as E
    X, y, _ = make_dataset(**cfg)
    kw = {"random_state": 0} if "SMOTE" in name else {}
    Xr, yr = getattr(O, name)(**kw).fit_resample(X, y)
    Xo, yo = getattr(E, name)(**kw).fit_resample(X, y)
    assert Xo.shape == Xr.shape, (Xo.shape, Xr.shape)
    assert np.array_equal(yo, yr)
    assert np.array_equal(Xo.view(np.int64), Xr.view(np.int64)), "resampled rows differ bitwise"


def test_confusion(cuda):
    from flake16_framework_b200 import ops
    rs = np.random.RandomState(5)
    n, n_proj = 100000, 26
    y = (rs.rand(n) < 0.1).astype(np.uint8)
    p = (rs.rand(n) < 0.1).astype(np.uint8)
    proj = rs.randint(0, n_proj, n).astype(np.int32)
    counts = torch.zeros((n_proj + 1, 3), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.confusion(torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(proj).cuda(), n_proj, counts)
    ref = np.zeros((n_proj + 1, 3), dtype=np.int64)
    for j in range(n):                                        # experiment.py:476-483
        k = int(2 * y[j] + p[j]) - 1
        if k == -1:
            continue
        ref[proj[j], k] += 1
        ref[n_proj, k] += 1
    assert np.array_equal(counts.cpu().numpy(), 2 * ref)
