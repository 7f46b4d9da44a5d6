"""Saturated 3-dataset grid slice with and without the DecisionTree configs, for the library in F16_LIB."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
keep = [("NOD", "Flake16", "Scaling"), ("OD", "FlakeFlagger", "None"), ("NOD", "Flake16", "None")]
allc = [c for c in S.all_config_keys() if c[:3] in keep]
prep = S.prepare(parsed, allc)
S.run_grid(parsed, allc, prepared=prep)
out = [os.path.basename(os.environ.get("F16_LIB", "main"))]
for name, cfgs in (("all", allc), ("without DT", [c for c in allc if c[4] != "Decision Tree"]), ("only DT", [c for c in allc if c[4] == "Decision Tree"])):
    best = 1e9
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        S.run_grid(parsed, cfgs, prepared=prep)
        torch.cuda.synchronize(); best = min(best, time.time() - t)
    out.append("%s %.2f s" % (name, best))
print("   ".join(out), flush=True)
