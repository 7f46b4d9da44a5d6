"""Wall time of the grid engine per dataset (18 configs x 10 folds each, 100 000 tests) and for the
whole 216-config grid: where a full pass spends its time.  usage: python tools/grid_by_dataset.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
allc = S.all_config_keys()
prep = S.prepare(parsed, allc)
S.run_grid(parsed, allc, prepared=prep)
def timed(cfgs):
    best = 1e9
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        S.run_grid(parsed, cfgs, prepared=prep)
        torch.cuda.synchronize(); best = min(best, time.time() - t)
    return best
tot = 0.0
for ds in sorted({c[:3] for c in allc}):
    cfgs = [c for c in allc if c[:3] == ds]
    t = timed(cfgs); tot += t
    per = {m: timed([c for c in cfgs if c[4] == m]) for m in S.MODELS}
    print("%-28s %5.2f s   only ET %.2f  RF %.2f  DT %.2f" % ("/".join(ds), t, per["Extra Trees"], per["Random Forest"], per["Decision Tree"]), flush=True)
print("sum of the 12 datasets %.2f s; full grid in one pass %.2f s" % (tot, timed(allc)), flush=True)
