"""Tiny end-to-end exercise of every kernel (used under compute-sanitizer on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from util import make_dataset
from flake16_framework_b200 import estimators as E, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
X, y, _ = make_dataset(n)
for cls in (E.ExtraTreesClassifier, E.RandomForestClassifier, E.DecisionTreeClassifier):
    kw = {} if cls is E.DecisionTreeClassifier else {"n_estimators": 3}
    t = time.time()
    m = cls(random_state=0, **kw).fit(X, y)
    p = m.predict(X)
    torch.cuda.synchronize()
    print(cls.__name__, "nodes", m.forest_.node_counts().tolist(), "pred+", int(p.sum()), "acc", float((p == y).mean()),
          "%.3fs" % (time.time() - t), flush=True)
for cls in (E.TomekLinks, E.SMOTE, E.EditedNearestNeighbours, E.SMOTEENN, E.SMOTETomek):
    kw = {"random_state": 0} if "SMOTE" in cls.__name__ else {}
    Xr, yr = cls(**kw).fit_resample(X, y)
    print(cls.__name__, Xr.shape, int(yr.sum()), flush=True)
print("SMOKE_OK")
# round 2: TreeSHAP, Spearman, the grid engine with learnt node capacities
m = E.ExtraTreesClassifier(random_state=0, n_estimators=3).fit(X, y)
phi = m.shap_values(X[:200])
print("shap", phi[0].shape, float(np.abs(phi[0]).max()), float(np.abs(phi[0] + phi[1]).max()), flush=True)
rho = ops.spearman(torch.from_numpy(np.ascontiguousarray(X)).cuda()).cpu().numpy()
print("spearman", rho.shape, float(np.abs(np.diag(rho) - 1).max()), flush=True)
if len(sys.argv) > 2:
    from flake16_framework_b200 import scores as S, hostprep as hp, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(n, 16))
    cfgs = [c for c in S.all_config_keys() if c[:3] == ("NOD", "Flake16", "Scaling")]
    prep = S.prepare(parsed, cfgs)
    for rep in range(2):
        out = S.run_grid(parsed, cfgs, prepared=prep, n_streams=2)
    print("grid", len(out), sum(v[3][2] for v in out.values()), flush=True)
print("SMOKE2_OK")
