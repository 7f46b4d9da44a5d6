"""Marginal cost of each model family / of the balancers in a saturated grid slice (GPU box).
usage: python tools/grid_marginals.py [n_datasets=3]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
nds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
keep = [("NOD", "Flake16", "Scaling"), ("OD", "FlakeFlagger", "None"), ("NOD", "Flake16", "None"),
        ("OD", "Flake16", "PCA"), ("NOD", "FlakeFlagger", "Scaling")][:nds]
allc = [c for c in S.all_config_keys() if c[:3] in keep]
models = sorted({c[4] for c in allc})
subsets = {"all": allc}
for m in models:
    subsets["only " + m] = [c for c in allc if c[4] == m]
    subsets["without " + m] = [c for c in allc if c[4] != m]
subsets["balancing None only"] = [c for c in allc if c[3] == "None"]
subsets["SMOTE* only"] = [c for c in allc if "SMOTE" in c[3]]
subsets["no SMOTE*"] = [c for c in allc if "SMOTE" not in c[3]]
prep = S.prepare(parsed, allc)
S.run_grid(parsed, allc, n_streams=8, prepared=prep)
for name, cfgs in subsets.items():
    best = 1e9
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        S.run_grid(parsed, cfgs, n_streams=8, prepared=prep)
        torch.cuda.synchronize(); best = min(best, time.time() - t)
    print("%-28s %3d configs  %6.2f s" % (name, len(cfgs), best), flush=True)
