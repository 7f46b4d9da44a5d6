#!/usr/bin/env python
"""bench.py - the `scores` hot path benchmark (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the full 216-config grid (2 flaky types x 2 feature sets x 3
preprocessings x 6 balancings x 3 models, 10-fold CV, 100-tree forests) over a synthetic
100 000-test x 16-feature tests.json.  metric = configs/sec (BASELINE.json).
  value : inputs already resident in HBM (12 preprocessed datasets + fold index maps).
  e2e   : `flake16_framework_b200.scores.write_scores("tests.json" -> "scores.pkl")`, i.e. the
          reference CLI path: JSON parse, host preprocessing, H2D of every dataset, the grid,
          D2H of the counts, pickle.
With N > 1 ranks the (dataset, fold, balancing group) work items are sharded over ranks (strong
scaling: the grid is fixed) and the counts are all-reduced once.

The CPU arm (`--impl reference`, and the `cpu_baseline` object of our own line at N = 1) is ONE
fixed sample, independent of --steps: fold 1 of 10 of all 18 (balancing x model) kinds on one
Flake16 dataset and on one FlakeFlagger dataset = 36 single-fold tasks through the oracle
(oracle/ref_scores.py: the reference's get_scores on the in-image scikit-learn), one
single-threaded process per task like the reference's Pool(N_PROC) (experiment.py:47,496),
extrapolated to the 216-config grid with the grid's own multiplicities (cpu_grid_estimate).
"""
import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")    # before any CUDA call (see flake16_framework_b200/__init__.py)
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "configs/sec for full `scores` grid"
UNIT = "configs/s"
BALANCINGS = ("None", "Tomek Links", "SMOTE", "ENN", "SMOTE ENN", "SMOTE Tomek")
MODELS = ("Extra Trees", "Random Forest", "Decision Tree")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-tests", type=int, default=100000)
    ap.add_argument("--n-estimators", type=int, default=100)
    ap.add_argument("--streams", type=int, default=int(os.environ.get("F16_STREAMS", "4")))
    ap.add_argument("--configs", default="grid216",
                    help="grid216 | slice (1 dataset, 18 configs) | config2 (BASELINE configs[1]: RandomForest, Flake16, no "
                         "balancing; use with --cv group --n-splits 5) | config3 (configs[2]: ExtraTrees + SMOTE, Flake16, "
                         "Scaling, NOD + OD) | config5 (configs[4]: the 12 datasets x SMOTE ENN x Extra Trees; use with "
                         "--n-tests 1000000 --n-estimators 500)")
    ap.add_argument("--cv", default="stratified", choices=["stratified", "group"],
                    help="stratified: the reference's StratifiedKFold; group: StratifiedGroupKFold over the projects (configs[1])")
    ap.add_argument("--n-splits", type=int, default=10)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probes", action="store_true", help="skip the single-kernel roofline probes")
    ap.add_argument("--cpu-sample-out", default=None, help=argparse.SUPPRESS)     # internal: run the CPU sample, dump JSON
    ap.add_argument("--cpu-sample-tests", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-budget-s", type=float, default=240.0,
                    help="reference arm: stop repeating the CPU sample when the next repetition would pass this budget")
    return ap.parse_args()


def select_configs(all_keys, which):
    if which == "slice":
        return [c for c in all_keys if c[:3] == ("NOD", "Flake16", "Scaling")]
    if which == "config2":
        return [c for c in all_keys if c == ("NOD", "Flake16", "None", "None", "Random Forest")]
    if which == "config3":
        return [c for c in all_keys if c[1:] == ("Flake16", "Scaling", "SMOTE", "Extra Trees")]
    if which == "config5":
        return [c for c in all_keys if c[3] == "SMOTE ENN" and c[4] == "Extra Trees"]
    return list(all_keys)


def workload_config(args):
    names = {"grid216": "full scores grid: 216 configs x 10-fold CV, %d-tree forests" % args.n_estimators,
             "slice": "grid slice NOD/Flake16/Scaling: 18 configs x 10-fold CV, %d-tree forests" % args.n_estimators,
             "config2": "BASELINE configs[1]: RandomForest n_estimators=%d, Flake16, no preprocessing / balancing, NOD" % args.n_estimators,
             "config3": "BASELINE configs[2]: ExtraTrees n_estimators=%d + SMOTE, Flake16, Scaling, NOD + OD" % args.n_estimators,
             "config5": "BASELINE configs[4]: ExtraTrees n_estimators=%d + SMOTE-ENN on the 12 (flaky type, feature set, "
                        "preprocessing) datasets x 10-fold CV" % args.n_estimators}
    n_cfg = {"grid216": 216, "slice": 18, "config2": 1, "config3": 2, "config5": 12}[args.configs]
    return {"workload": "%s, synthetic tests.json %d tests x 16 features (seed 16)" % (names[args.configs].replace("10-fold", "%d-fold" % args.n_splits), args.n_tests),
            "n_tests": args.n_tests, "n_configs": n_cfg, "n_splits": args.n_splits,
            "cv": "StratifiedKFold(shuffle, random_state=0)" if args.cv == "stratified" else "StratifiedGroupKFold(groups=projects, shuffle, random_state=0)",
            "n_estimators": args.n_estimators, "parallelism": "grid-sharded x%d" % args.gpus,
            "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU reference arm
# The sample is FIXED (it does not depend on --steps / --warmup): fold 1 of all 18 (balancing x
# model) kinds on the NOD / Flake16 / Scaling dataset (d = 16: brute-force k-NN in Tomek / ENN)
# and on the NOD / FlakeFlagger / None dataset (d = 7: KD-tree k-NN) = 36 tasks.
CPU_SAMPLE_DATASETS = (("NOD", "Flake16", "Scaling"), ("NOD", "FlakeFlagger", "None"))


def cpu_sample_tasks():
    return [ds + (bal, model) for ds in CPU_SAMPLE_DATASETS for bal in BALANCINGS for model in MODELS]


def cpu_sample_desc(n_tests, n_estimators, workers):
    return ("sample: fold 1 of 10 of all 18 (balancing x model) kinds on NOD/Flake16/Scaling and on NOD/FlakeFlagger/None "
            "= 36 single-fold tasks of the %d-test tests.json (%d-tree forests), one single-threaded process per task "
            "(%d at a time) through the oracle on scikit-learn; extrapolated to the 216-config grid: every sampled kind "
            "stands for the 6 (flaky type x preprocessing) configs of its feature set, config time = setup + 10 x fold "
            "time, grid time = sum of config times / min(host cores, 216) workers (perfect packing, the best case for "
            "the reference's Pool(N_PROC))" % (n_tests, n_estimators, workers))


def run_cpu_sample(tests_file, n_estimators=100):
    """Runs the 36 tasks; returns (results, wall seconds, workers)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_scores as R
    tasks = cpu_sample_tasks()
    workers = max(1, min(os.cpu_count() or 1, len(tasks)))
    t0 = time.perf_counter()
    res = R.run_configs_timed(tasks, tests_file, workers, n_splits=10, max_folds=1, n_estimators=n_estimators)
    return res, time.perf_counter() - t0, workers


def cpu_sample_subprocess(args, tests_file, out_file):
    """The CPU sample in its own interpreter (no CUDA context, clean fork pool), started while the GPU
    arm prepares and warms up - untimed work on both sides; it is joined BEFORE the timed region."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-sample-out", out_file, "--cpu-sample-tests", tests_file,
           "--n-estimators", str(args.n_estimators)]
    return subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=open(out_file + ".err", "w"))


def cpu_sample_worker_main(args):
    res, wall, workers = run_cpu_sample(args.cpu_sample_tests, args.n_estimators)
    with open(args.cpu_sample_out, "w") as fd:
        json.dump({"results": [[list(k), t, c, tot] for k, t, c, tot in res], "wall": wall, "workers": workers}, fd)


def cpu_grid_estimate(results):
    """216-config grid estimate from the 36-task sample (see cpu_sample_desc)."""
    cores = os.cpu_count() or 1
    pool = min(cores, 216)
    cfg_s = {keys: t["setup_s"] + 10.0 * t["folds_s"] for keys, t, _, _ in results}
    total_cpu_s = 6.0 * sum(cfg_s.values())
    grid_s = total_cpu_s / pool
    by_kind = {}
    for keys, t, _, _ in results:
        by_kind.setdefault("%s/%s" % (keys[3], keys[4]), []).append(round(t["folds_s"], 2))
    return {"value": 216.0 / grid_s, "grid_s_estimate": grid_s, "cpu_core_s_estimate": total_cpu_s, "pool_workers": pool,
            "slowest_config_s_estimate": max(cfg_s.values()),
            "fold_s_by_kind[Flake16,FlakeFlagger]": by_kind}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from flake16_framework_b200 import synth
    tmp = tempfile.mkdtemp(prefix="f16bench_ref")
    tests_file = os.path.join(tmp, "tests.json")
    synth.make_tests_json(tests_file, args.n_tests, 16)
    # W warm-up + K timed repetitions of the fixed sample, as far as the time budget allows:
    # one repetition takes as long as its slowest task (a RandomForest on SMOTE'd rows after a
    # brute-force k-NN: minutes), so the steps ACTUALLY run are printed.
    t_begin = time.perf_counter()
    runs = []
    warm_done = 0
    requested = args.warmup + args.steps
    last = 0.0
    for i in range(requested):
        if runs and (time.perf_counter() - t_begin) + last > args.cpu_budget_s:
            break
        res, wall, workers = run_cpu_sample(tests_file, args.n_estimators)
        last = wall
        runs.append((res, wall, workers))
    if len(runs) > args.steps:          # the first ones count as warm-up
        warm_done = len(runs) - args.steps
    timed = runs[warm_done:]
    ests = [cpu_grid_estimate(r) for r, _, _ in timed]
    value = sum(e["value"] for e in ests) / len(ests)
    workers = timed[0][2]
    wall = sum(w for _, w, _ in timed) / len(timed)
    est = ests[-1]
    cfg = workload_config(args)
    cfg["workload"] = cpu_sample_desc(args.n_tests, args.n_estimators, workers) + " | stands for: " + cfg["workload"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(timed), "warmup": warm_done, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": cpu_sample_desc(args.n_tests, args.n_estimators, workers) + " (%.0f s per repetition)" % wall,
                             "host_cores": os.cpu_count(), "estimate": est},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _traffic(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(key)
    except Exception:
        return None


def tree_roofline_probe(parsed, L, kind_name):
    """One tree-building kernel alone (100 trees on one 90 000 x 16 fold): algorithmic bytes per launch
    = 12 B x sum_{internal nodes} n_node_samples x (F_eval + 1), F_eval =
    max_features = 4, read off the fitted trees; duration = CUDA events around that kernel on its
    launch stream (inside f16_forest_fit)."""
    import numpy as np
    import torch
    from flake16_framework_b200 import hostprep as hp, ops
    kind = {"RF": ops.KIND_RF, "ET": ops.KIND_ET}[kind_name]
    X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
    X = np.ascontiguousarray(X)
    tr, _ = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
    Xd = torch.from_numpy(X).cuda()
    Xrow = ops.rows_f32(Xd, torch.from_numpy(tr).cuda())
    ytr = torch.from_numpy(y[tr].astype(np.uint8)).cuda()
    sidx = ops.argsort_columns(Xrow, 16)
    L.f16_set_profiling(1)
    ms, alg = [], None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for it in range(4):
        flush.fill_(it)
        f = ops.forest_fit(Xrow, ytr, 16, kind, 100, 0, sidx)
        t = L.f16_forest_build_ms(f._h)
        if it > 0:
            ms.append(t)
        if alg is None:
            tot = 0
            counts = f.node_counts()
            for ti in range(100):
                tr_ = f.export_tree(ti, int(counts[ti]))
                tot += int(tr_["n_node_samples"][tr_["children_left"] >= 0].sum())
            alg = 12.0 * tot * (4 + 1)
        f.free()
    L.f16_set_profiling(0)
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_ms = sum(ms) / len(ms)
    achieved = alg / (avg_ms * 1e-3) / 1e9
    kname = {"RF": "k_build_best_rf<16> (RandomForest", "ET": "k_build_random_et<16> (ExtraTrees"}[kind_name]
    out = {"bound": "hbm", "kernel": kname + ", 100 trees, 90000x16 fold, timed alone)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
           "algorithmic_bytes_per_launch": alg, "kernel_ms": avg_ms,
           "traffic": _traffic("k_build_best_bytes_per_launch" if kind_name == "RF" else "k_build_random_dram_bytes_per_launch")}
    if kind_name == "ET":
        out["traffic_l2_to_l1"] = _traffic("k_build_random_l2_to_l1_bytes_per_launch")
        out["note"] = ("the training rows (5.8 MB) stay L2-resident: DRAM traffic is ~1 % of the algorithmic bytes, the "
                       "kernel's memory traffic is L2->L1 gathers; it is bound by per-node latency chains, not by HBM")
    return out


def knn_roofline_probe(parsed):
    """The tensor-core k-NN filter (tcgen05 + TMEM).  Algorithmic flops = 3 d n_query n_ref
    (SURVEY.md 8(d)); duration = CUDA events around the whole f16_knn call (centring, float16 split,
    filter, exact float64 selection) on one standardised 90 000 x 16 fold; peak = measured dense
    bf16 tensor throughput."""
    import numpy as np
    import torch
    from flake16_framework_b200 import hostprep as hp, ops
    X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY, hp.FEATURE_SETS["Flake16"])
    X = np.ascontiguousarray(hp.preprocess(X, "Scaling"))
    tr, _ = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
    A = torch.from_numpy(np.ascontiguousarray(X[tr])).cuda()
    order, _ = ops.variance_order(X)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for it in range(4):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.knn(A, A, 4, (order, 3))
        e1.record()
        e1.synchronize()
        if it > 0:
            ms.append(e0.elapsed_time(e1))
    peaks = _peaks()
    peak = float(peaks.get("bf16_tflops", 1590.0))
    n = A.shape[0]
    flops = 3.0 * 16 * n * n
    avg_ms = sum(ms) / len(ms)
    achieved = flops / (avg_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "f16_knn strategy 3: k_knn_umma_filter (tcgen05) + exact float64 select, %d x %d x 16" % (n, n),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (measured)" if peaks else "fallback 1590 TFLOP/s",
            "algorithmic_flops_per_call": flops, "call_ms": avg_ms, "traffic": None,
            "note": "the filter is bound by its per-element epilogue (one compare per distance), not by the MMA rate"}


def parity_check(cpu_results, S, parsed, dev, n_streams, n_estimators):
    """The CPU sample's fold-1 counts (oracle) against this library's counts for the same 36
    (config, fold) pairs at the bench's own size, per project and in total."""
    tasks = [keys for keys, _, _, _ in cpu_results]
    cfgs, counts, _, gd = S.run_grid(parsed, tasks, n_streams=n_streams, device=dev, n_estimators=n_estimators,
                                     return_counts=True, folds=[0])
    bad = []
    for keys, _, per_proj, total in cpu_results:
        ci = cfgs.index(keys)
        ours_total = [int(v) for v in counts[ci, gd.n_proj]]
        ours_proj = {str(name): [int(v) for v in counts[ci, pid]] for pid, name in enumerate(gd.proj_names)}
        if ours_total != total or ours_proj != per_proj:
            bad.append({"config": list(keys), "cpu_total_fp_fn_tp": total, "gpu_total_fp_fn_tp": ours_total})
    return {"configs": len(tasks), "folds": 1, "n_tests": int(len(parsed[1])), "identical": not bad,
            "what": "FP/FN/TP per project and in total, oracle (scikit-learn + samplers_np) vs GPU, fold 1 of each sampled config",
            "mismatches": bad[:8]}


def ours_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")       # the driver's own setting (e.g. INFO) wins
        dist.init_process_group("nccl", device_id=dev)
    from flake16_framework_b200 import _lib, hostprep as hp, scores as S, synth
    L = _lib.lib()
    _lib.init(local)

    tmp = tempfile.mkdtemp(prefix="f16bench_r%d_" % rank)
    tests_file = os.path.join(tmp, "tests.json")
    synth.make_tests_json(tests_file, args.n_tests, 16)          # every rank: same seeded table
    parsed = hp.parse_tests(tests_file)
    configs = select_configs(S.all_config_keys(), args.configs)
    with_cpu = (world == 1 and not args.no_cpu_baseline and args.configs == "grid216" and args.cv == "stratified"
                and args.n_splits == 10)
    cpu_proc, cpu_out = None, os.path.join(tmp, "cpu_sample.json")
    if with_cpu:
        cpu_proc = cpu_sample_subprocess(args, tests_file, cpu_out)
    prepared = S.prepare(parsed, configs, dev, args.n_splits, args.cv)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize(dev)

    def one_step():
        return S.run_grid(parsed, configs, n_splits=args.n_splits, n_streams=args.streams, device=dev, rank=rank,
                          world=world, n_estimators=args.n_estimators, prepared=prepared, return_counts=True)

    for _ in range(args.warmup):
        one_step()
    # untimed single-kernel probes and the parity pass come BEFORE the timed region: they fill the
    # wait for the CPU sample (which must be over before anything is timed) instead of following it
    probes = {}
    if world == 1 and not args.no_probes:
        probes["roofline"] = tree_roofline_probe(parsed, L, "RF")
        probes["roofline_et"] = tree_roofline_probe(parsed, L, "ET")
        probes["roofline_knn"] = knn_roofline_probe(parsed)
    cpu_sample, cpu_res, parity = None, None, None
    if cpu_proc is not None:
        if cpu_proc.wait() != 0:
            raise RuntimeError("CPU sample failed:\n" + open(cpu_out + ".err").read()[-2000:])
        cpu_sample = json.load(open(cpu_out))
        cpu_res = [(tuple(k), t, c, tot) for k, t, c, tot in cpu_sample["results"]]
        parity = parity_check(cpu_res, S, parsed, dev, args.streams, args.n_estimators)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    L.f16_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    for it in range(args.steps):
        flush.fill_(it & 0xff)
        barrier()
        ev0.record()
        one_step()
        torch.cuda.synchronize(dev)
        ev1.record()
        barrier()
        total_ms += ev0.elapsed_time(ev1)
    launches = int(L.f16_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    peak_mem = torch.cuda.max_memory_allocated(dev)
    try:        # the library's scratch lives in the CUDA stream-ordered pool, not in torch's allocator
        free_b, total_b = torch.cuda.mem_get_info(dev)
        peak_mem = max(peak_mem, total_b - free_b)
    except Exception:
        pass
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    ln = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    n_cfg = len(configs)
    value = n_cfg * args.steps / (total_ms * 1e-3)

    # ---- e2e: tests.json on disk -> scores.pkl through the public write_scores(); a couple of steps
    #      (everything is warm by now; the path differs from `value` by host parsing + H2D + pickle)
    e2e = None
    if not args.no_e2e:
        scores_file = os.path.join(tmp, "scores.pkl")
        kw = dict(n_streams=args.streams, configs=configs, n_estimators=args.n_estimators, n_splits=args.n_splits, cv=args.cv)
        e2e_steps = max(1, min(args.e2e_steps, args.steps))
        if total_ms / args.steps > 15e3:        # long steps: one end-to-end pass is enough (it differs from `value` by < 5 %)
            e2e_steps = 1
        barrier()
        t0 = time.perf_counter()
        stats = None
        for _ in range(e2e_steps):
            _, _, stats = S.write_scores(tests_file, scores_file, return_stats=True, **kw)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_cfg * e2e_steps / float(dt.item()), "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": stats["h2d_bytes"], "d2h_bytes_per_step": stats["d2h_bytes"],
               "includes": "json parse, host preprocessing, H2D, grid, D2H, pickle", "breakdown_s_last_step": stats.get("breakdown_s")}

    if rank != 0:
        return 0
    trees_per_fit = args.n_estimators
    n_forest_cfg = sum(1 for c in configs if c[4] != "Decision Tree")
    n_dt_cfg = n_cfg - n_forest_cfg
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 criterion / f32 features / int counts",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(ln.item()),
            "trees_per_s": (n_forest_cfg * trees_per_fit + n_dt_cfg) * args.n_splits * args.steps / (total_ms * 1e-3),
            "device_memory_in_use_bytes": int(peak_mem)}
    rc = 0
    if world == 1:
        line.update(probes)
        if cpu_sample is not None:
            res = cpu_res
            wall, workers = cpu_sample["wall"], cpu_sample["workers"]
            est = cpu_grid_estimate(res)
            line["cpu_baseline"] = {"value": est["value"], "unit": UNIT, "cores": workers, "kind": "port",
                                    "sample": cpu_sample_desc(args.n_tests, args.n_estimators, workers) +
                                              " (%.0f s; run in a separate process during this arm's untimed preparation "
                                              "and warm-up, finished before the timed region)" % wall,
                                    "host_cores": os.cpu_count(), "estimate": est}
            line["parity_check"] = parity
            if not parity["identical"]:
                rc = 1
    print(json.dumps(line), flush=True)
    return rc


if __name__ == "__main__":
    a = parse_args()
    rc = 0
    if a.cpu_sample_out:
        cpu_sample_worker_main(a)
        sys.exit(0)
    if a.impl == "reference":
        reference_arm(a)
    else:
        rc = ours_arm(a) or 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    sys.exit(rc)
