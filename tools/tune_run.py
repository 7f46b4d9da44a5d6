"""Runs the perf probe against every library variant under build_variants/ (GPU box)."""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode = sys.argv[1] if len(sys.argv) > 1 else "grid"
for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "libf16_*.so"))):
    print("=====", os.path.basename(lib), flush=True)
    env = dict(os.environ, F16_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "perf_probe.py"), "100000", mode], env=env,
                       capture_output=True, text=True, timeout=900)
    out = [l for l in r.stdout.splitlines() if ("fit" in l or "grid" in l or "Error" in l)]
    print("\n".join(out[-40:]), flush=True)
    if r.returncode:
        print("FAILED", r.stderr[-2000:], flush=True)
