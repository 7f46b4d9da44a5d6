"""Times one forest kind alone and with 16 fits in flight, for the library named by F16_LIB,
and checks the fitted forests against a reference digest (same trees regardless of variant).
usage: F16_LIB=tools/variants/libf16_x.so python tools/variant_probe.py [RF,ET,DT] [N,...]"""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flake16_framework_b200 import ops, synth, hostprep as hp

kinds = (sys.argv[1] if len(sys.argv) > 1 else "RF").split(",")
Ns = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,16").split(",")]
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
X = np.ascontiguousarray(hp.preprocess(X, "Scaling")); d = 16
tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
tri = torch.from_numpy(tr).cuda(); tei = torch.from_numpy(te).cuda()
Xrow = ops.rows_f32(Xd, tri); ytr = ops.gather_u8(yd, tri)
Xte = ops.rows_f32(Xd, tei)
sidx = ops.argsort_columns(Xrow, d)
streams = [torch.cuda.Stream() for _ in range(16)]
torch.cuda.synchronize()
KIND = {"RF": ops.KIND_RF, "ET": ops.KIND_ET, "DT": ops.KIND_DT}
line = [os.path.basename(os.environ.get("F16_LIB", "main"))]
for name in kinds:
    kind = KIND[name]
    nt = 1 if name == "DT" else 100
    f = ops.forest_fit(Xrow, ytr, d, kind, nt, 0, sidx)
    nc = f.node_counts()
    pred = f.predict(Xte).cpu().numpy()
    digest = hashlib.sha1(np.asarray(nc).tobytes() + pred.tobytes()).hexdigest()[:10]
    f.free()
    line.append("%s[%s]" % (name, digest))
    for N in Ns:
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t = time.time()
            fs = []
            for i in range(N):
                with torch.cuda.stream(streams[i]):
                    fs.append(ops.forest_fit(Xrow, ytr, d, kind, nt, 0, sidx))
            torch.cuda.synchronize(); dt = time.time() - t
            for f in fs:
                f.free()
            best = min(best, dt)
        line.append("N=%d %.1f ms/fit" % (N, best * 1e3 / N))
print("  ".join(line), flush=True)
