"""Round-2 variant sweep (GPU box): for every library under tools/variants/ - forest kinds alone and
saturated (tools/variant_probe.py) and one grid slice (1 dataset x 18 configs x 10 folds).
usage: python tools/r2_probe.py [KINDS] [Ns] [grid: 0|1]"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kinds = sys.argv[1] if len(sys.argv) > 1 else "ET,RF"
Ns = sys.argv[2] if len(sys.argv) > 2 else "1,8,16"
grid = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
cfgs = [c for c in S.all_config_keys() if c[:3] == ("NOD", "Flake16", "Scaling")]
prep = S.prepare(parsed, cfgs)
S.run_grid(parsed, cfgs, prepared=prep)
best = 1e9
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    S.run_grid(parsed, cfgs, prepared=prep)
    torch.cuda.synchronize(); best = min(best, time.time() - t)
print("grid slice %%.2f s" %% best, flush=True)
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "libf16_*.so"))):
    env = dict(os.environ, F16_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variant_probe.py"), kinds, Ns], env=env,
                       capture_output=True, text=True, timeout=900)
    print(r.stdout.strip() or ("FAILED " + r.stderr[-1500:]), flush=True)
    if grid:
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
        print("    ", os.path.basename(lib), r.stdout.strip() or ("FAILED " + r.stderr[-1500:]), flush=True)
