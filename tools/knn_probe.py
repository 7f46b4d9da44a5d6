"""k-NN kernel timing only (device events), raw / scaled / d=7, 90k and 178k (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from flake16_framework_b200 import ops, synth, hostprep as hp
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
print(os.path.basename(os.environ.get("F16_LIB", "main")))
for fs, pre in (("Flake16", "None"), ("Flake16", "Scaling"), ("FlakeFlagger", "Scaling")):
    X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS[fs])
    X = np.ascontiguousarray(hp.preprocess(X, pre))
    co = ops.variance_order(X)
    A = torch.from_numpy(X[:90000]).cuda()
    B = torch.cat([A, A[:88200] * 0.5 + A[1800:90000] * 0.5]).contiguous()      # 178 200 rows, SMOTE-like
    for M, name in ((A, "90k"), (B, "178k")):
        ref = None
        for mode in sorted({co[1], 0}):
            cm = (co[0], mode)
            out = ops.knn(M, M, 4, cm); torch.cuda.synchronize()
            if ref is None:
                ref = out
            same = bool(torch.equal(out, ref))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.knn(M, M, 4, cm)
            e1.record(); torch.cuda.synchronize()
            print("%-12s %-8s mode=%d %-5s %7.2f ms  same=%s" % (fs, pre, mode, name, e0.elapsed_time(e1) / 3, same), flush=True)
