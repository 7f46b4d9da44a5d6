"""ctypes binding of libf16_b200.so (C ABI declared in include/f16.h) + the in-tree nvcc build.

There is NO CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  PyTorch tensors are used purely as device-memory containers (``.data_ptr()``).
"""

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("F16_LIB") or os.path.join(_HERE, "libf16_b200.so")

# Tree-kernel tunables (threads per tree CTA, min CTAs/SM, rows of the shared-memory regime).
TUNE = {"ET_NT": 256, "ET_MINB": 4, "ET_S16": 512, "ET_S8": 1024,
        "RF_NT": 256, "RF_MINB": 3, "DT_NT": 512, "DT_MINB": 1}


def _sources(tune):
    """(source, object tag, extra flags).  -fmad=false wherever float64 expressions must
    round like the CPU.  The tree builders are compiled once per variant (ET / RF / DT)."""
    t = dict(TUNE)
    t.update(tune or {})
    x = list(t.get("EXTRA", []))      # e.g. ["-DF16_PHASE_TIMING"] for tools/phase_probe.py
    return (
        ("f16_tree.cu", "", ["-fmad=false"]),
        ("f16_tree_random.cu", "_et", ["-fmad=false", "-DF16_VARIANT=_et", "-DNT=%d" % t["ET_NT"],
                                        "-DF16_MINB=%d" % t["ET_MINB"], "-DF16_S16=%d" % t["ET_S16"],
                                        "-DF16_S8=%d" % t["ET_S8"]] + x),
        ("f16_tree_best.cu", "_rf", ["-fmad=false", "-DF16_VARIANT=_rf", "-DNT=%d" % t["RF_NT"],
                                      "-DF16_MINB=%d" % t["RF_MINB"], "-DF16_WITH_BOOTSTRAP"] + x),
        ("f16_tree_best.cu", "_dt", ["-fmad=false", "-DF16_VARIANT=_dt", "-DNT=%d" % t["DT_NT"],
                                      "-DF16_MINB=%d" % t["DT_MINB"]] + x),
        ("f16_shap.cu", "", []),
        ("f16_stats.cu", "", []),
        ("f16_parse.cu", "", []),
        ("f16_misc.cu", "", ["-fmad=false"]),
        ("f16_sort.cu", "", []),
        ("f16_knn.cu", "", ["-DKQ=%d" % t.get("KNN_KQ", 2), "-DKNN_UNROLL=%d" % t.get("KNN_UNROLL", 1)]),
        ("f16_knn_tc.cu", "", ["-DTC_MT=%d" % t.get("TC_MT", 2)]),
        ("f16_knn_sweep.cu", "", []),
        ("f16_knn_umma.cu", "", ["-DUM_NB=%d" % t.get("UM_NB", 2), "-DUM_STAGES=%d" % t.get("UM_STAGES", 4)]),
    )


_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False, tune=None, out=None):
    """Compiles every CUDA source for sm_100a and links libf16_b200.so in-tree.
    ``tune`` / ``out`` build an experimental variant under another name (tools/tune_*.py)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    objs, rebuilt = [], False
    procs = []
    tag = "" if out is None else "." + os.path.basename(out).replace(".so", "")
    force = force or out is not None
    lib_path = out or os.path.join(_HERE, "libf16_b200.so")
    for src, variant, extra in _sources(tune):
        s = os.path.join(_CSRC, src)
        o = os.path.join(_CSRC, src.replace(".cu", variant + tag + ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in hdrs) or _newer(__file__, o):
            cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", *_ARCH, "-Xcompiler", "-fPIC",
                   "-Xptxas", "-v" if verbose else "-O3", *extra, "-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
            rebuilt = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode()))
        if verbose:
            sys.stderr.write(out.decode())
    if rebuilt or not os.path.exists(lib_path):
        cmd = [nvcc, "-shared", *_ARCH, "-o", lib_path, *objs, "-lcudart"]
        subprocess.check_call(cmd)
    if out is not None:
        for o in objs:
            os.remove(o)
    return lib_path


_lib = None

c_void_p, c_int, c_int32, c_int64, c_uint32 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int32,
                                               ctypes.c_int64, ctypes.c_uint32)

# name -> argtypes ; every symbol include/f16.h declares (tests check the export list)
SIGNATURES = {
    "f16_last_error": ([], ctypes.c_char_p),
    "f16_version": ([], c_int),
    "f16_init": ([c_int], c_int),
    "f16_launch_count": ([c_int], ctypes.c_longlong),
    "f16_set_profiling": ([c_int], None),
    "f16_forest_build_ms": ([c_void_p], ctypes.c_double),
    "f16_gather_rows_f32": ([c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "f16_gather_rows_f64": ([c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "f16_gather_u8": ([c_void_p, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "f16_gather_i32": ([c_void_p, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "f16_argsort_columns": ([c_void_p, c_int64, c_int32, c_void_p, c_void_p], c_int),
    "f16_tree_seeds": ([c_uint32, c_int32, c_int32, c_void_p, c_void_p], c_int),
    "f16_bootstrap_counts": ([c_void_p, c_int32, c_int64, c_void_p, c_void_p], c_int),
    "f16_forest_fit": ([c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32,
                        c_uint32, c_void_p, ctypes.POINTER(c_void_p)], c_int),
    "f16_forest_fit_cap": ([c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32,
                            c_uint32, c_int64, c_void_p, ctypes.POINTER(c_void_p)], c_int),
    "f16_forest_predict": ([c_void_p, c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "f16_forest_status": ([c_void_p, c_void_p], c_int),
    "f16_forest_n_trees": ([c_void_p], c_int),
    "f16_forest_max_nodes": ([c_void_p], c_int),
    "f16_forest_node_counts": ([c_void_p, c_void_p, c_void_p], c_int),
    "f16_forest_export": ([c_void_p, c_int32, c_int64] + [c_void_p] * 8 + [c_void_p], c_int),
    "f16_forest_free": ([c_void_p, c_void_p], None),
    "f16_forest_shap": ([c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p], c_int),
    "f16_spearman": ([c_void_p, c_int64, c_int32, c_void_p, c_void_p], c_int),
    "f16_knn": ([c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_void_p, c_void_p], c_int),
    "f16_knn_tc_probe": ([c_void_p, c_int64, c_void_p, c_int64, c_int32, ctypes.POINTER(ctypes.c_float), c_void_p], c_int),
    "f16_knn_umma_probe": ([c_void_p, c_int64, c_void_p, c_int64, c_int32, ctypes.POINTER(ctypes.c_float), c_void_p], c_int),
    "f16_fill_u8": ([c_void_p, c_int32, c_int64, c_void_p], c_int),
    "f16_smote_generate": ([c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int64,
                            c_void_p, c_void_p], c_int),
    "f16_tomek_keep": ([c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p], c_int),
    "f16_enn_keep": ([c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p], c_int),
    "f16_compact_rows": ([c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p], c_int),
    "f16_tests_parse": ([ctypes.c_char_p, ctypes.POINTER(c_void_p)], c_int),
    "f16_tests_rows": ([c_void_p], c_int64),
    "f16_tests_cols": ([c_void_p], c_int32),
    "f16_tests_projects": ([c_void_p], c_int32),
    "f16_tests_names_bytes": ([c_void_p], c_int64),
    "f16_tests_copy": ([c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "f16_tests_free": ([c_void_p], None),
    "f16_confusion": ([c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p], c_int),
}


class F16Error(RuntimeError):
    pass


def lib():
    """Loads the shared library (never builds silently, never falls back to a CPU path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise F16Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is mandatory; there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise F16Error("libf16_b200 error %d: %s" % (rc, lib().f16_last_error().decode()))


_inited = set()


def init(device=0):
    if device not in _inited:
        check(lib().f16_init(int(device)))
        _inited.add(device)
