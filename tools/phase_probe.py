"""Where do the tree builders spend their cycles? (GPU box)

Builds a variant of the library with -DF16_PHASE_TIMING (thread 0 of every tree CTA accumulates
clock64() deltas per phase; phases are barrier separated, so thread 0's view is the CTA's), fits
one forest of each kind alone and with 8 in flight, and prints the share of every phase.

    python tools/phase_probe.py build      # here (no GPU needed): tools/libf16_phase.so
    F16_LIB=tools/libf16_phase.so python tools/phase_probe.py   # on the GPU box
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "libf16_phase.so")   # *.so: git-ignored, but travels to the GPU box

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from flake16_framework_b200 import _lib
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    _lib.build(tune={"EXTRA": ["-DF16_PHASE_TIMING"]}, out=OUT)
    print("built", OUT)
    sys.exit(0)

import numpy as np
import torch
from flake16_framework_b200 import ops, _lib, synth, hostprep as hp

NPH = 12
ET_NAMES = {0: "root", 1: "relocate", 2: "G minmax+draw", 3: "G count+choose+finish", 4: "G partition",
            5: "S subtree (warp 0)", 6: "G leaf"}
BEST_NAMES = {0: "root", 1: "relocate", 2: "G ends+draw", 3: "G scan+argmax", 4: "G finish+mark+partition",
              5: "S ends+draw", 6: "S scan+argmax", 7: "S finish+mark+partition", 8: "G leaf/no split",
              9: "S leaf/no split"}


def read(sym, reset=True):
    L = _lib.lib()
    fn = getattr(L, sym)
    cyc = (ctypes.c_ulonglong * (2 * NPH))()
    nod = (ctypes.c_ulonglong * (2 * NPH))()
    assert fn(cyc, nod, int(reset)) == 0
    return np.array(cyc).reshape(2, NPH), np.array(nod).reshape(2, NPH)


def report(title, cyc, nod, names):
    tot = cyc.sum()
    print("== %s: %.1f Mcycles per tree" % (title, tot / 100 / 1e6))
    for k in range(NPH):
        if cyc[k]:
            print("   %-28s %5.1f %%   %8d visits/tree  %8.0f cycles/visit" % (
                names.get(k, str(k)), 100.0 * cyc[k] / tot, nod[k] // 100, cyc[k] / max(nod[k], 1)))


n_tests = int(os.environ.get("N_TESTS", "100000"))
parsed = hp.tests_to_arrays(synth.make_tests_dict(n_tests, 16))
for prep in ("None", "Scaling"):
    X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
    X = np.ascontiguousarray(hp.preprocess(X, prep)); d = 16
    tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
    Xd = torch.from_numpy(X).cuda(); yd = torch.from_numpy(y.astype(np.uint8)).cuda()
    tri = torch.from_numpy(tr).cuda()
    Xrow = ops.rows_f32(Xd, tri); ytr = ops.gather_u8(yd, tri)
    sidx = ops.argsort_columns(Xrow, d)
    streams = [torch.cuda.Stream() for _ in range(8)]
    torch.cuda.synchronize()
    for kind, name, sym, kidx, names in ((ops.KIND_ET, "ET", "f16_debug_phases_et", 0, ET_NAMES),
                                         (ops.KIND_RF, "RF", "f16_debug_phases_rf", 1, BEST_NAMES),
                                         (ops.KIND_DT, "DT", "f16_debug_phases_dt", 1, BEST_NAMES)):
        for N in (1, 8):
            if kind == ops.KIND_DT and N > 1:
                continue
            for rep in range(2):
                read(sym)
                fs = []
                for i in range(N):
                    with torch.cuda.stream(streams[i]):
                        fs.append(ops.forest_fit(Xrow, ytr, d, kind, 100 if kind != ops.KIND_DT else 1, 0, sidx))
                torch.cuda.synchronize()
                ms = [f.build_ms() for f in fs] if hasattr(fs[0], "build_ms") else []
                for f in fs:
                    f.free()
            cyc, nod = read(sym)
            scale = N * (1 if kind != ops.KIND_DT else 0.01)
            report("%s prep=%s, %d in flight" % (name, prep, N), cyc[kidx] / scale, (nod[kidx] / scale).astype(np.int64), names)
