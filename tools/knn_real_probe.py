"""k-NN timing on the grid's real inputs: fold-0 training rows and their SMOTE'd set, for every
preprocessing of both feature sets, float64 strategy vs the strategy the grid selects (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from flake16_framework_b200 import ops, synth, hostprep as hp, estimators as E

n_tests = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
parsed = hp.tests_to_arrays(synth.make_tests_dict(n_tests, 16))
for fs in ("Flake16", "FlakeFlagger"):
    for pre in ("None", "Scaling", "PCA"):
        X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS[fs])
        X = np.ascontiguousarray(hp.preprocess(X, pre))
        co = ops.variance_order(X)
        tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
        Xtr, ytr = np.ascontiguousarray(X[tr]), y[tr]
        with ops.column_order(co):
            Xs, ys = E.SMOTE(random_state=0).fit_resample(Xtr, ytr)
        for M, name in ((Xtr, "train"), (np.ascontiguousarray(Xs), "smote")):
            Md = torch.from_numpy(M).cuda()
            ref = None
            for mode in sorted({co[1], 0} | ({5} if co[1] == 3 else set())):      # 5 = mma.sync filter
                cm = (co[0], mode)
                out = ops.knn(Md, Md, 4, cm); torch.cuda.synchronize()
                if ref is None:
                    ref = out
                same = bool(torch.equal(out, ref))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    ops.knn(Md, Md, 4, cm)
                e1.record(); torch.cuda.synchronize()
                uniq = len(np.unique(M, axis=0)) if mode == 0 else -1
                print("%-12s %-8s %-6s n=%6d mode=%d %8.2f ms  same=%s  unique_rows=%d" % (
                    fs, pre, name, len(M), mode, e0.elapsed_time(e1) / 3, same, uniq), flush=True)
