"""ncu target: ONE launch that holds as many trees as 16 forests in flight (n_estimators = 1600 on a
90 000 x 16 fold) - the tree builders as the grid engine runs them, saturated, in a form ncu can
replay.  usage: python tools/ncu_saturated.py [ET|RF] [n_trees]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flake16_framework_b200 import ops, synth, hostprep as hp

which = sys.argv[1] if len(sys.argv) > 1 else "ET"
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
X = np.ascontiguousarray(hp.preprocess(X, "Scaling")); d = 16
tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
tri = torch.from_numpy(tr).cuda()
Xrow = ops.rows_f32(Xd, tri); ytr = ops.gather_u8(yd, tri)
sidx = ops.argsort_columns(Xrow, d)
import time
for kind in [k for k in which.split(",")]:
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        f = ops.forest_fit(Xrow, ytr, d, {"ET": ops.KIND_ET, "RF": ops.KIND_RF}[kind], nt, 0, sidx, node_cap=40000)
        f.status(); dt = time.time() - t
        print(kind, nt, "trees: %.1f ms = %.2f ms per 100 trees, max nodes %d" % (dt * 1e3, dt * 1e5 / nt, f.max_nodes()), flush=True)
        f.free()
