"""Grid-slice wall time (NOD/Flake16/Scaling: 18 configs x 10 folds, 100 000 tests) under the current
environment (F16_LANES, F16_DT_LANES, F16_STREAMS, F16_LIB).  usage: python tools/grid_env_probe.py [n_datasets]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
nds = int(sys.argv[1]) if len(sys.argv) > 1 else 1
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
keys = [("NOD", "Flake16", "Scaling"), ("OD", "Flake16", "None"), ("NOD", "FlakeFlagger", "PCA")][:nds]
cfgs = [c for c in S.all_config_keys() if c[:3] in keys]
prep = S.prepare(parsed, cfgs)
ns = int(os.environ.get("F16_STREAMS", "4"))
S.run_grid(parsed, cfgs, prepared=prep, n_streams=ns)
best = 1e9
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    S.run_grid(parsed, cfgs, prepared=prep, n_streams=ns)
    torch.cuda.synchronize(); best = min(best, time.time() - t)
print("grid %d dataset(s) streams=%d lanes=%s dt_lanes=%s: %.2f s" % (nds, ns, os.environ.get("F16_LANES", "2"), os.environ.get("F16_DT_LANES", "2"), best), flush=True)
