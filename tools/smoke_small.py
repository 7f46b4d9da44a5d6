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
