"""Quick timing of the tensor-filter k-NN on the scaled Flake16 fold-0 sets (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from flake16_framework_b200 import ops, synth, hostprep as hp
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
X = np.ascontiguousarray(hp.preprocess(X, "Scaling"))
co = ops.variance_order(X)
A = torch.from_numpy(X[:90000]).cuda()
B = torch.cat([A, A[:88200] * 0.5 + A[1800:90000] * 0.5]).contiguous()
line = [os.path.basename(os.environ.get("F16_LIB", "main"))]
for M, name in ((A, "90k"), (B, "178k")):
    ref = ops.knn(M, M, 4, (co[0], 5))
    for mode in (3, 5):
        out = ops.knn(M, M, 4, (co[0], mode)); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.knn(M, M, 4, (co[0], mode))
        e1.record(); torch.cuda.synchronize()
        line.append("%s mode%d %.2f ms%s" % (name, mode, e0.elapsed_time(e1) / 3, "" if torch.equal(out, ref) else " MISMATCH"))
print("  ".join(line), flush=True)
