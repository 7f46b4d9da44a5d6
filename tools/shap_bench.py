"""`python experiment.py shap` at bench size: the reference's two explained configurations
(experiment.py:520-530) on the synthetic 100 000-test table - seconds for balancing + fit and for
the TreeSHAP of all rows, attribution checks (local accuracy against the forest's own output).
usage: python tools/shap_bench.py [n_tests]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from flake16_framework_b200 import explain, hostprep as hp, synth, ops, estimators as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
parsed = hp.tests_to_arrays(synth.make_tests_dict(n, 16))
for rep in range(2):
    for cfg in explain.SHAP_CONFIGS:
        torch.cuda.synchronize(); t0 = time.time()
        ft, fs, pre, bal, model = cfg
        features, labels, _ = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES[ft], hp.FEATURE_SETS[fs])
        features = np.ascontiguousarray(hp.preprocess(features, pre))
        Xd = torch.from_numpy(features).cuda(); yd = torch.from_numpy(labels.astype(np.uint8)).cuda()
        Xb, yb = explain._BALANCING[bal]().fit_resample(Xd, yd)
        m = explain._MODEL[model](100).fit(Xb.contiguous(), yb.contiguous())
        m.forest_.status()
        torch.cuda.synchronize(); t1 = time.time()
        rows = ops.rows_f32(Xd)
        phi = m.forest_.shap_values(rows, 0)
        torch.cuda.synchronize(); t2 = time.time()
        nodes = int(m.forest_.node_counts().sum())
        print("%s: resample+fit %.2f s (%d rows, %d nodes), TreeSHAP of %d rows x %d trees: %.2f s  -> %.1f M (row, leaf) pairs/s; |phi| max %.4f, sum range [%.4f, %.4f]"
              % ("/".join(cfg), t1 - t0, Xb.shape[0], nodes, n, 100, t2 - t1, n * ((nodes + 100) / 2) / (t2 - t1) / 1e6,
                 float(phi.abs().max()), float(phi.sum(1).min()), float(phi.sum(1).max())), flush=True)
