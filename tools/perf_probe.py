"""Per-operator timings at headline scale (run on the GPU box)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from flake16_framework_b200 import ops, _lib, synth, hostprep as hp, scores as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
what = sys.argv[2] if len(sys.argv) > 2 else "all"
t0 = time.time()
tests = synth.make_tests_dict(n, 16)
parsed = hp.tests_to_arrays(tests)
print("synth+parse %.1fs" % (time.time() - t0), flush=True)
L = _lib.lib(); _lib.init(0)
L.f16_set_profiling(1)

def timeit(name, fn, reps=1):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize(); dt = (time.time() - t) / reps
    print("%-40s %9.2f ms" % (name, dt * 1e3), flush=True)
    return r

if what in ("all", "ops"):
    for fs, pre in ((("Flake16", "Scaling"), ("FlakeFlagger", "None")) if os.environ.get("F16_LIB") else
                    (("Flake16", "None"), ("Flake16", "Scaling"), ("FlakeFlagger", "None"))):
        X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES["NOD"], hp.FEATURE_SETS[fs])
        X = np.ascontiguousarray(hp.preprocess(X, pre)); d = X.shape[1]
        tf = hp.stratified_kfold_test_folds(y)
        tr, te = next(iter(hp.kfold_split(tf)))
        Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
        tri = torch.from_numpy(tr).cuda(); tei = torch.from_numpy(te).cuda()
        print("== %s/%s n_train=%d d=%d" % (fs, pre, len(tr), d))
        Xrow = timeit("rows_f32(train)", lambda: ops.rows_f32(Xd, tri))
        ytr = ops.gather_u8(yd, tri); Xte = ops.rows_f32(Xd, tei)
        sidx = timeit("argsort_columns", lambda: ops.argsort_columns(Xrow, d))
        for kind, name in ((ops.KIND_ET, "ET"), (ops.KIND_RF, "RF"), (ops.KIND_DT, "DT")):
            f = timeit("fit %s (100 trees)" % name, lambda: ops.forest_fit(Xrow, ytr, d, kind, 100, 0, sidx))
            print("    build kernel %.2f ms, nodes total %d" % (L.f16_forest_build_ms(f._h), int(f.node_counts().sum())))
            timeit("predict %s" % name, lambda: f.predict(Xte))
            f.free()
        Xtr64 = ops.gather_rows_f64(Xd, tri)
        timeit("knn k=4 n=%d" % len(tr), lambda: ops.knn(Xtr64, Xtr64, 4, ops.variance_order(X)))
        c1 = int(y[tr].sum())
        Xs, ys = timeit("smote", lambda: ops.smote(Xtr64, ytr, c1, len(tr) - c1, 1, 0, 5))
        timeit("knn k=4 n=%d (after SMOTE)" % Xs.shape[0], lambda: ops.knn(Xs, Xs, 4, ops.variance_order(X)))
        Xrs = ops.rows_f32(Xs.contiguous()); sidx2 = ops.argsort_columns(Xrs, d)
        for kind, name in ((ops.KIND_ET, "ET"), (ops.KIND_RF, "RF"), (ops.KIND_DT, "DT")):
            f = timeit("fit %s after SMOTE n=%d" % (name, Xs.shape[0]), lambda: ops.forest_fit(Xrs, ys.contiguous(), d, kind, 100, 0, sidx2))
            print("    build kernel %.2f ms, nodes total %d" % (L.f16_forest_build_ms(f._h), int(f.node_counts().sum())))
            f.free()

if what in ("all", "grid"):
    L.f16_set_profiling(0)
    cfgs = [c for c in S.all_config_keys() if c[0] == "NOD" and c[1] == "Flake16" and c[2] == "Scaling"]
    for ns in ((8, 16) if os.environ.get("F16_LIB") else (1, 4, 8, 16)):
        torch.cuda.synchronize(); t = time.time()
        S.run_grid(parsed, cfgs, n_streams=ns)
        torch.cuda.synchronize(); dt = time.time() - t
        print("grid slice: %d configs (1 dataset x 10 folds), n_streams=%d: %.2f s -> %.2f configs/s" % (len(cfgs), ns, dt, len(cfgs) / dt), flush=True)
