"""Short single-fold target for ncu captures: stage, sort, ET/RF/DT fit + predict, k-NN, at headline scale."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flake16_framework_b200 import ops, _lib, synth, hostprep as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
parsed = hp.tests_to_arrays(synth.make_tests_dict(n, 16))
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
X = np.ascontiguousarray(X); d = 16
tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
tri = torch.from_numpy(tr).cuda(); tei = torch.from_numpy(te).cuda()
Xrow = ops.rows_f32(Xd, tri); ytr = ops.gather_u8(yd, tri); Xte = ops.rows_f32(Xd, tei)
sidx = ops.argsort_columns(Xrow, d)
for kind in (ops.KIND_ET, ops.KIND_RF, ops.KIND_DT):
    f = ops.forest_fit(Xrow, ytr, d, kind, 100, 0, sidx)
    p = f.predict(Xte)
    f.status(); f.free()
Xtr64 = ops.gather_rows_f64(Xd, tri)
nn = ops.knn(Xtr64, Xtr64, 4, ops.variance_order(X))
torch.cuda.synchronize()
print("ncu target done", int(p.sum()), int(nn.sum()))
