"""Grid-slice throughput for every library variant x (streams, lanes) (GPU box)."""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
from flake16_framework_b200 import synth, hostprep as hp, scores as S
parsed = hp.tests_to_arrays(synth.make_tests_dict(100000, 16))
cfgs = [c for c in S.all_config_keys() if c[:3] == ("NOD", "Flake16", "Scaling")]
prep = S.prepare(parsed, cfgs)
S.run_grid(parsed, cfgs, n_streams=4, prepared=prep)
for ns in (4, 8):
    torch.cuda.synchronize(); t = time.time()
    S.run_grid(parsed, cfgs, n_streams=ns, prepared=prep)
    torch.cuda.synchronize(); print("streams=%%d %%.2f s" %% (ns, time.time() - t), flush=True)
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libf16_*.so"))):
    for lanes in ("3", "6"):
        env = dict(os.environ, F16_LIB=lib, F16_LANES=lanes)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
        print(os.path.basename(lib), "lanes=" + lanes, " | ".join(r.stdout.split("\n")), r.stderr[-300:] if r.returncode else "", flush=True)
