"""ms per 100 trees of one ExtraTrees / RandomForest launch holding 100 .. 4800 trees (90 000 x 16 fold):
the saturated cost of a tree kernel without the grid engine around it.
usage: [F16_LIB=..] [F16_ET_WARP=1] python tools/sat_probe.py ET 100,1600,4800"""
import os, sys, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flake16_framework_b200 import ops, synth, hostprep as hp

kinds = (sys.argv[1] if len(sys.argv) > 1 else "ET").split(",")
nts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "100,1600").split(",")]
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
fsname = sys.argv[3] if len(sys.argv) > 3 else "Flake16"
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS[fsname])
X = np.ascontiguousarray(hp.preprocess(X, "Scaling")); d = X.shape[1]
tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
tri = torch.from_numpy(tr).cuda()
Xrow = ops.rows_f32(Xd, tri); ytr = ops.gather_u8(yd, tri)
sidx = ops.argsort_columns(Xrow, d)
out = [os.path.basename(os.environ.get("F16_LIB", "main")) + " " + fsname]
for kind in kinds:
    for nt in nts:
        best, dig = 1e9, ""
        for rep in range(2):
            torch.cuda.synchronize(); t = time.time()
            f = ops.forest_fit(Xrow, ytr, d, {"ET": ops.KIND_ET, "RF": ops.KIND_RF}[kind], nt, 0, sidx, node_cap=40000)
            f.status(); best = min(best, time.time() - t)
            dig = hashlib.sha1(f.node_counts()[:100].tobytes()).hexdigest()[:8]
            f.free()
        out.append("%s[%s] %d trees: %.2f ms/100" % (kind, dig, nt, best * 1e5 / nt))
print("  ".join(out), flush=True)
