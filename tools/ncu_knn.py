"""One tensor-filter k-NN call on the grid's SMOTE'd fold-0 set (target for ncu, GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from flake16_framework_b200 import ops, synth, hostprep as hp, estimators as E
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
X = np.ascontiguousarray(hp.preprocess(X, "Scaling"))
co = ops.variance_order(X)
tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
with ops.column_order(co):
    Xs, ys = E.SMOTE(random_state=0).fit_resample(np.ascontiguousarray(X[tr]), y[tr])
Md = torch.from_numpy(np.ascontiguousarray(Xs)).cuda()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = ops.knn(Md, Md, 4, (co[0], 3))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", out.shape)
