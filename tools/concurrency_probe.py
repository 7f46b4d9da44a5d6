"""How does the GPU scale with N identical forest fits in flight on N streams? (GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flake16_framework_b200 import ops, _lib, synth, hostprep as hp

parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
X = np.ascontiguousarray(hp.preprocess(X, "Scaling")); d = 16
tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
tri = torch.from_numpy(tr).cuda()
Xrow = ops.rows_f32(Xd, tri); ytr = ops.gather_u8(yd, tri)
sidx = ops.argsort_columns(Xrow, d)
Xtr64 = ops.gather_rows_f64(Xd, tri)
streams = [torch.cuda.Stream() for _ in range(16)]
torch.cuda.synchronize()
print(os.path.basename(os.environ.get("F16_LIB", "main")))
for kind, name in ((ops.KIND_ET, "ET"), (ops.KIND_RF, "RF"), (None, "knn")):
    for N in (1, 2, 4, 8, 16):
        for rep in range(2):
            torch.cuda.synchronize(); t = time.time()
            fs = []
            for i in range(N):
                with torch.cuda.stream(streams[i]):
                    if kind is None:
                        fs.append(ops.knn(Xtr64, Xtr64, 4))
                    else:
                        fs.append(ops.forest_fit(Xrow, ytr, d, kind, 100, 0, sidx))
            torch.cuda.synchronize(); dt = time.time() - t
            for f in fs:
                if kind is not None:
                    f.free()
        print("%-3s N=%2d  %7.1f ms  -> %6.1f ms per fit" % (name, N, dt * 1e3, dt * 1e3 / N), flush=True)
