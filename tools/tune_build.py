"""Builds experimental variants of the library (tree-kernel tunables) under build_variants/."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flake16_framework_b200 import _lib
VARIANTS = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {
    "base": {},
    "et64_s128": {"ET_NT": 64, "ET_MINB": 12, "ET_S16": 128, "ET_S8": 256, "RF_NT": 128, "RF_MINB": 5},
    "et64_s256": {"ET_NT": 64, "ET_MINB": 10, "ET_S16": 256, "ET_S8": 512, "RF_NT": 64, "RF_MINB": 10},
    "et128_s256": {"ET_NT": 128, "ET_MINB": 5, "ET_S16": 256, "ET_S8": 512, "RF_NT": 128, "RF_MINB": 5, "DT_NT": 512},
    "et32_s64": {"ET_NT": 32, "ET_MINB": 24, "ET_S16": 64, "ET_S8": 128, "RF_NT": 32, "RF_MINB": 16},
}
os.makedirs(os.path.join(ROOT, "build_variants"), exist_ok=True)
for name, tune in VARIANTS.items():
    out = os.path.join(ROOT, "build_variants", "libf16_%s.so" % name)
    _lib.build(tune=tune, out=out)
    print("built", out, tune, flush=True)
