"""Grid-slice wall time for (worker streams, side-stream lanes) combinations (GPU box)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
keep = {("NOD", "Flake16", "Scaling"), ("OD", "FlakeFlagger", "None"), ("NOD", "Flake16", "None")}
cfgs = [c for c in S.all_config_keys() if c[:3] in keep]
prep = S.prepare(parsed, cfgs)
S.run_grid(parsed, cfgs, n_streams=8, prepared=prep)
for ns in %s:
    best = 1e9
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        S.run_grid(parsed, cfgs, n_streams=ns, prepared=prep)
        torch.cuda.synchronize(); best = min(best, time.time() - t)
    print("streams=%%d %%.2f s" %% (ns, best), flush=True)
'''
streams = sys.argv[1] if len(sys.argv) > 1 else "(6, 8, 10, 12)"
for lanes in (sys.argv[2] if len(sys.argv) > 2 else "3,4").split(","):
    env = dict(os.environ, F16_LANES=lanes)
    r = subprocess.run([sys.executable, "-c", code % (ROOT, streams)], env=env, capture_output=True, text=True, timeout=1200)
    print("lanes=" + lanes, " | ".join(r.stdout.split("\n")), r.stderr[-400:] if r.returncode else "", flush=True)
