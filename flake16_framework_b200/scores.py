"""The ``scores`` stage on B200: ``get_scores`` / ``write_scores`` of the reference
(experiment.py:446-501) re-organised for the GPU.

Reference shape: 216 independent ``get_scores(config)`` calls in a ``multiprocessing.Pool``
(experiment.py:493-498), each re-parsing tests.json, re-running the preprocessing and re-doing
the 10-fold loop with its own balancing + fit + predict + Python per-row confusion loop.

Here the grid is de-duplicated: 2 label vectors x 2 feature sets x 3
preprocessings = 12 datasets; the fold map depends only on the labels; a resampled training
set is shared by the 3 models; the k=4 neighbour table of a training set serves both
TomekLinks (column 1) and ENN (columns 1..3).  The unit of work is one (dataset, fold): it
stages the fold on the device once and runs every requested (balancing, model) pair on it.
Units are independent: they are spread over worker threads (one CUDA stream each, so several
forests are in flight per GPU) and, with ``torch.distributed``, over ranks (one process per
GPU); the only exchange step is one all-reduce of the integer confusion counts at the end.

Results are FP/FN/TP per project and in total (int64, accumulated on the device by
``f16_confusion``); precision / recall / F1 are derived on the host with the reference's own
Python-float expressions, so they are bit-identical given identical counts.
"""

import itertools
import os
import pickle
import queue
import threading
import time

import numpy as np
import torch

from . import DEFAULT_LANES, DEFAULT_WORKERS
from . import hostprep as hp
from . import ops

BALANCINGS = ("None", "Tomek Links", "SMOTE", "ENN", "SMOTE ENN", "SMOTE Tomek")     # experiment.py:87-94
MODELS = ("Extra Trees", "Random Forest", "Decision Tree")                           # experiment.py:95-99
MODEL_KIND = {"Extra Trees": ops.KIND_ET, "Random Forest": ops.KIND_RF, "Decision Tree": ops.KIND_DT}
GRID_KEYS = (tuple(hp.FLAKY_TYPES), tuple(hp.FEATURE_SETS), tuple(hp.PREPROCESSINGS), BALANCINGS, MODELS)


def all_config_keys():
    return list(itertools.product(*GRID_KEYS))                                       # experiment.py:494


class GridData:
    """Host-side, config-independent preparation: parse once, 12 datasets, 2 fold maps."""

    def __init__(self, parsed, configs, n_splits=10, cv="stratified"):
        """``cv``: "stratified" - the reference's StratifiedKFold(n_splits, shuffle=True,
        random_state=0) (experiment.py:450); "group" - StratifiedGroupKFold with the projects as
        groups (BASELINE.json configs[1]'s variant)."""
        self.parsed = parsed
        self.n_splits = n_splits
        self.cv = cv
        all_features, raw_labels, projects = parsed
        self.projects = projects
        # project ids in order of first appearance == insertion order of the reference's
        # `scores = {proj: [0] * 6 for proj in projects}` (experiment.py:456)
        uniq, first, inv = np.unique(projects, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        self.proj_names = [projects[first[o]] for o in order]
        self.proj_id = rank[inv].astype(np.int32)
        self.n_proj = len(uniq)
        self.col_order = {}     # (ft, fs, pre) -> columns by descending variance (k-NN early exit)
        self.datasets = {}      # (ft, fs, pre) -> float64 C-contiguous [N, d]
        self.labels = {}        # ft -> bool[N]
        self.folds = {}         # ft -> test_folds int[N]
        matrices = {}           # (fs, pre) -> (matrix, column order): independent of the flaky type
        scaled = {}
        by_fs, by_ft = {}, {}   # the feature view depends on the feature set only, the labels on the flaky type only
        for (ft, fs, pre) in sorted({c[:3] for c in configs}):
            if fs not in by_fs or ft not in by_ft:
                Xv, yv, _ = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES[ft], hp.FEATURE_SETS[fs])
                by_fs.setdefault(fs, Xv)
                by_ft.setdefault(ft, yv)
            X, y = by_fs[fs], by_ft[ft]
            if (fs, pre) not in matrices:
                if pre == "None":
                    M = X
                else:        # the StandardScaler output of a feature set serves "Scaling" and, as PCA's input, "PCA"
                    if fs not in scaled:
                        scaled[fs] = hp.StandardScaler().fit_transform(X)
                    M = scaled[fs] if pre == "Scaling" else hp.PCA().fit_transform(scaled[fs])
                M = np.ascontiguousarray(M)
                matrices[(fs, pre)] = (M, ops.variance_order(M))
            self.datasets[(ft, fs, pre)], self.col_order[(ft, fs, pre)] = matrices[(fs, pre)]
            if ft not in self.labels:
                self.labels[ft] = y
                self.folds[ft] = (hp.stratified_kfold_test_folds(y, n_splits, True, 0) if cv == "stratified" else
                                  hp.stratified_group_kfold_test_folds(y, projects, n_splits, True, 0))


class _DeviceData:
    """Per-process device copies (uploaded once; 12 x <= 12.8 MB at 100 k rows)."""

    def __init__(self, gd, device):
        uploaded = {}           # one device copy per distinct host matrix (6 for the full grid)
        self.X = {}
        for k, v in gd.datasets.items():
            if id(v) not in uploaded:
                uploaded[id(v)] = torch.from_numpy(v).to(device)
            self.X[k] = uploaded[id(v)]
        # k-NN strategy per dataset: the static choice, confirmed by a measurement on the device
        self.col_order, seen = {}, {}
        for key, co in gd.col_order.items():
            if key[1:] not in seen:                      # the matrix depends on (feature set, preprocessing) only
                seen[key[1:]] = ops.calibrate_knn(self.X[key], co)
            self.col_order[key] = seen[key[1:]]
        self.y = {k: torch.from_numpy(v.astype(np.uint8)).to(device) for k, v in gd.labels.items()}
        self.proj = torch.from_numpy(gd.proj_id).to(device)
        self.caps = _NodeCaps()          # learnt node capacities persist across passes over this data
        self.fold_idx = {}
        for ft, tf in gd.folds.items():
            for i, (tr, te) in enumerate(hp.kfold_split(tf, gd.n_splits)):
                ytr = gd.labels[ft][tr]
                c1 = int(ytr.sum())
                self.fold_idx[(ft, i)] = (torch.from_numpy(tr).to(device), torch.from_numpy(te).to(device),
                                          (len(tr) - c1, c1),
                                          torch.from_numpy(np.flatnonzero(ytr)).to(device),
                                          torch.from_numpy(np.flatnonzero(~ytr)).to(device))

    def h2d_bytes(self):
        b = sum(t.numel() * t.element_size() for t in {id(t): t for t in self.X.values()}.values())
        b += sum(t.numel() for t in self.y.values()) + self.proj.numel() * 4
        b += sum((a.numel() + b_.numel()) * 8 for a, b_, *_ in self.fold_idx.values())
        return b


# Balancings that share intermediate results stay in one work item: the k=4 neighbour table of the
# training rows serves TomekLinks and ENN; the SMOTE'd set and its neighbour table serve SMOTE,
# SMOTE ENN and SMOTE Tomek.
BALANCING_GROUPS = (("None", "Tomek Links", "ENN"), ("SMOTE", "SMOTE ENN", "SMOTE Tomek"))


def _unit_cost(gd, unit, wanted):
    (ft, fs, pre), fold = unit[0], unit[1]
    d = gd.datasets[(ft, fs, pre)].shape[1]
    n = gd.datasets[(ft, fs, pre)].shape[0]
    cost = 0.0
    for bal, models in wanted.items():
        m = 2.0 if "SMOTE" in bal else 1.0
        knn = (m * n) ** 2 * d * 1e-9 if bal != "None" and bal != "SMOTE" else 0.0
        cost += knn + m * n * d * len(models) * 1e-3
    return cost


def plan_units(gd, configs, n_splits, world, folds=None):
    """Work items (dataset, fold, balancing group), what each must produce ({balancing: [models]}),
    and their longest-processing-time-first assignment to ``world`` ranks (deterministic, so every
    rank derives the same plan without communicating).  720 (dataset, balancing, fold) resamples
    of the full grid -> 240 items of <= 9 fits.  ``folds`` restricts the plan to those fold indices
    (bench.py's parity check runs fold 0 only)."""
    wanted_by_ds = {}
    for c in configs:
        wanted_by_ds.setdefault(tuple(c[:3]), {}).setdefault(c[3], []).append(c[4])
    wanted, units = {}, []
    for ds, w in wanted_by_ds.items():
        for g, group in enumerate(BALANCING_GROUPS):
            wg = {bal: w[bal] for bal in group if bal in w}
            if not wg:
                continue
            for f in (range(n_splits) if folds is None else folds):
                units.append((ds, f, g))
                wanted[(ds, f, g)] = wg
    cost = {u: _unit_cost(gd, u, wanted[u]) for u in units}
    units.sort(key=lambda u: (-cost[u], u))
    load = [0.0] * world
    shards = [[] for _ in range(world)]
    for u in units:
        r = int(np.argmin(load))
        load[r] += cost[u]
        shards[r].append(u)
    return wanted, shards


def _minority_clean_mask(counts, strategy):
    if strategy == "all":
        return 0b11
    minority = 0 if counts[0] <= counts[1] else 1
    return 0b11 & ~(1 << minority)


_PILOT_BYTES = 8 << 30      # worst-case node bytes of one forest above which its capacity is piloted


class _NodeCaps:
    """Per-tree node capacity for f16_forest_fit_cap, learnt from the fits already settled: the
    worst case 2n - 1 costs 32 B x (2n - 1) per tree (1.1 GB per 100-tree forest at 178 k rows,
    57 GB for 500 trees at 1.8 M), real trees of this path hold 0.03 n - 0.3 n nodes.  Key = the config
    (dataset, balancing, model: its 10 folds are statistically alike); capacity = 1.5 x the largest nodes / n ratio seen + 1024.  The first fit of a
    key uses the worst case; if a later fit overflows its capacity (F16_ERR_OVERFLOW from
    f16_forest_status) run_grid repeats the whole pass with capacities disabled."""

    _seen = {}          # process-wide: ratios learnt by earlier runs seed the next run's capacities
                        # (like the memory pool, warm state of the library; a first run starts empty)

    def __init__(self, enabled=True):
        self.ratio, self.lock, self.enabled = dict(_NodeCaps._seen) if enabled else {}, threading.Lock(), enabled

    def cap(self, key, n):
        with self.lock:
            r = self.ratio.get(key) if self.enabled else None
        return 0 if r is None else min(2 * n - 1, int(1.5 * r * n) + 1024)

    def update(self, key, n, max_nodes):
        with self.lock:
            self.ratio[key] = max(self.ratio.get(key, 0.0), max_nodes / float(n))
            if self.enabled:
                _NodeCaps._seen[key] = self.ratio[key]


def _run_unit(gd, dd, unit, wanted, cfg_index, counts_all, n_estimators, timers, model_streams, caps):
    """One (dataset, fold, balancing group): stage, resample per balancing, fit/predict/count per model."""
    ds_key, fold = unit[0], unit[1]
    ft = ds_key[0]
    X64 = dd.X[ds_key]
    d = X64.shape[1]
    y_all = dd.y[ft]
    tr_idx, te_idx, counts, pos_in_tr, neg_in_tr = dd.fold_idx[(ft, fold)]
    Xte = ops.rows_f32(X64, te_idx)
    yte = ops.gather_u8(y_all, te_idx)
    pte = ops.gather_i32(dd.proj, te_idx)
    ytr = ops.gather_u8(y_all, tr_idx)
    need64 = any(b != "None" for b in wanted)
    Xtr64 = ops.gather_rows_f64(X64, tr_idx) if need64 else None
    n_tr = tr_idx.shape[0]
    minority = 0 if counts[0] <= counts[1] else 1

    cache = {}
    keep = [Xte, yte, pte]     # tensors read by side streams stay referenced until the unit is settled

    def nn4(tag, X):
        if tag not in cache:
            cache[tag] = ops.knn(X, X, 4)
        return cache[tag]

    def smoted():
        if "smote" not in cache:
            idx_min = pos_in_tr if minority == 1 else neg_in_tr
            cache["smote"] = ops.smote(Xtr64, ytr, min(counts), max(counts), minority, 0, 5, idx_min=idx_min)
        return cache["smote"]

    def clean(kind, X, y, tag, strategy, cnts):
        L = ops._ready()
        nn = nn4(tag, X)
        keep = torch.empty((X.shape[0],), dtype=torch.uint8, device=X.device)
        mask = _minority_clean_mask(cnts, strategy)
        fn = L.f16_tomek_keep if kind == "tomek" else L.f16_enn_keep
        ops.check(fn(ops._ptr(nn), 4, ops._ptr(y), X.shape[0], mask, ops._ptr(keep), ops._stream()))
        Xo, yo, _ = ops._compact(X, y, keep, 0 if kind == "tomek" else 1)
        return Xo, yo

    for bal in BALANCINGS:
        if bal not in wanted:
            continue
        if bal == "None":
            Xrow, yb = ops.rows_f32(X64, tr_idx), ytr
        else:
            if bal == "Tomek Links":
                Xb, yb = clean("tomek", Xtr64, ytr, "tr", "auto", counts)
            elif bal == "ENN":
                Xb, yb = clean("enn", Xtr64, ytr, "tr", "auto", counts)
            elif bal == "SMOTE":
                Xb, yb = smoted()
            elif bal == "SMOTE ENN":
                Xs, ys = smoted()
                Xb, yb = clean("enn", Xs, ys, "sm", "all", None)
            elif bal == "SMOTE Tomek":
                Xs, ys = smoted()
                Xb, yb = clean("tomek", Xs, ys, "sm", "all", None)
            Xrow = ops.rows_f32(Xb.contiguous())
            yb = yb.contiguous()
        models = wanted[bal]
        sorted_idx = None
        if any(m != "Extra Trees" for m in models):
            sorted_idx = ops.argsort_columns(Xrow, d)
        # the three models of a resample are independent: fork them onto side streams so the
        # single-CTA DecisionTree does not serialise behind the forests (and the next
        # balancing's k-NN overlaps with these fits)
        keep.append((Xrow, yb, sorted_idx))
        ready = torch.cuda.Event()
        ready.record()
        for model in MODELS:
            if model not in models:
                continue
            ci = cfg_index[ds_key + (bal, model)]
            lanes = model_streams[model]
            side = lanes[BALANCINGS.index(bal) % len(lanes)]
            side.wait_event(ready)
            with torch.cuda.stream(side):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                n_fit = Xrow.shape[0]
                cap_key = ds_key + (bal, model, int(n_fit).bit_length())        # config + size class
                node_cap = caps.cap(cap_key, n_fit)
                if node_cap == 0 and caps.enabled and model != "Decision Tree" and \
                        64 * n_fit * n_estimators > _PILOT_BYTES:
                    # the worst-case node arrays of this forest would not fit comfortably (57 GB for
                    # 500 trees on 1.8 M rows): measure the node count on a few trees first
                    pilot = ops.forest_fit(Xrow, yb, d, MODEL_KIND[model], min(8, n_estimators), 0, sorted_idx)
                    pilot.status()
                    caps.update(cap_key, n_fit, pilot.max_nodes())
                    pilot.free()
                    node_cap = caps.cap(cap_key, n_fit)
                forest = ops.forest_fit(Xrow, yb, d, MODEL_KIND[model], n_estimators, 0, sorted_idx, node_cap=node_cap)
                e1.record()
                pred = forest.predict(Xte)
                e2.record()
                ops.confusion(yte, pred, pte, gd.n_proj, counts_all[ci])
                keep.append(pred)
            timers.append((ci, e0, e1, e2, forest, cap_key, n_fit))
    # completion markers instead of a drain: the worker goes on to its next unit and settles this
    # one (status, timings, frees) later, so its streams never run dry between units
    events = []
    for s in [torch.cuda.current_stream()] + [side for lanes in model_streams.values() for side in lanes]:
        e = torch.cuda.Event()
        e.record(s)
        events.append(e)
    return events, keep


def prepare(parsed, configs=None, device=None, n_splits=10, cv="stratified"):
    """Host preparation + upload: returns (GridData, device copies).  bench.py uses this to
    time the grid with its inputs already resident in HBM."""
    configs = list(configs) if configs is not None else all_config_keys()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    ops._ready(device.index if device.index is not None else torch.cuda.current_device())
    gd = GridData(parsed, configs, n_splits, cv)
    dd = _DeviceData(gd, device)
    torch.cuda.synchronize(device)
    return gd, dd


def run_grid(parsed, configs=None, n_splits=10, n_estimators=100, n_streams=DEFAULT_WORKERS, device=None,
             rank=0, world=1, progress=None, return_counts=False, prepared=None, stats=None, folds=None,
             cv="stratified"):
    """Computes the scores dict for ``configs`` (default: the full 216 grid).

    Returns {config_keys: [t_train / n_splits, t_test / n_splits, scores, scores_total]} - the
    value layout of the reference's scores.pkl (experiment.py:488-490, :498)."""
    configs = list(configs) if configs is not None else all_config_keys()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    gd, dd = prepared if prepared is not None else prepare(parsed, configs, device, n_splits, cv)
    if stats is not None:
        stats["h2d_bytes"] = dd.h2d_bytes()
    cfg_index = {c: i for i, c in enumerate(configs)}
    wanted, shards = plan_units(gd, configs, n_splits, world, folds)
    mine = shards[rank]
    n_lanes = int(os.environ.get("F16_LANES", str(DEFAULT_LANES)))
    n_dt_lanes = int(os.environ.get("F16_DT_LANES", "2"))

    def one_pass(caps):
        """All of this rank's work items once; returns (counts, times, a fit overflowed its node capacity)."""
        counts_all = torch.empty((len(configs), gd.n_proj + 1, 3), dtype=torch.int64, device=device)
        ops.zero_bytes(counts_all)
        # the memset runs on the caller's stream; the workers' non-blocking streams must not
        # start accumulating before it has completed
        torch.cuda.current_stream(device).synchronize()
        times = np.zeros((len(configs), 2), dtype=np.float64)
        q = queue.Queue()
        for u in mine:
            q.put(u)
        errors, overflow = [], [False]
        lock = threading.Lock()
        done = [0]

        def worker():
            torch.cuda.set_device(device)
            stream = torch.cuda.Stream(device=device)
            # several forests of one item in flight: n_lanes side streams per forest model, 2 for the tree
            model_streams = {m: [torch.cuda.Stream(device=device) for _ in range(n_dt_lanes if m == "Decision Tree" else n_lanes)]
                             for m in MODELS}
            util = torch.cuda.Stream(device=device)     # status reads / frees of settled items
            pending = []

            def settle(item):
                events, keep, timers = item
                for e in events:
                    e.synchronize()
                with torch.cuda.stream(util):
                    for ci, e0, e1, e2, forest, cap_key, n_fit in timers:
                        try:
                            forest.status()
                            caps.update(cap_key, n_fit, forest.max_nodes())
                        except ops.F16Overflow:
                            overflow[0] = True
                        with lock:
                            times[ci, 0] += e0.elapsed_time(e1) * 1e-3
                            times[ci, 1] += e1.elapsed_time(e2) * 1e-3
                        forest.free()
                keep.clear()
                with lock:
                    done[0] += 1
                    if progress:
                        progress(done[0], len(mine))

            try:
                with torch.cuda.stream(stream):
                    while True:
                        try:
                            u = q.get_nowait()
                        except queue.Empty:
                            break
                        timers = []
                        with ops.column_order(dd.col_order[u[0]]):
                            events, keep = _run_unit(gd, dd, u, wanted[u], cfg_index, counts_all, n_estimators,
                                                     timers, model_streams, caps)
                        pending.append((events, keep, timers))
                        if len(pending) > 1:            # one item in flight behind the current one
                            settle(pending.pop(0))
                    while pending:
                        settle(pending.pop(0))
            except Exception as ex:  # propagate to the caller
                with lock:
                    errors.append(ex)

        n_thr = max(1, min(n_streams, len(mine)))
        threads = [threading.Thread(target=worker, daemon=True) for _ in range(n_thr)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        torch.cuda.synchronize(device)
        return counts_all, times, overflow[0]

    counts_all, times, overflowed = one_pass(dd.caps)
    if overflowed:          # a tree outgrew its learnt node capacity: redo the pass with the worst case
        counts_all, times, overflowed = one_pass(_NodeCaps(enabled=False))
        if overflowed:
            raise ops.F16Overflow("a tree exceeded the worst-case node capacity 2n - 1 (corrupt input?)")

    if world > 1:
        import torch.distributed as dist
        on_host = dist.get_backend() != "nccl"        # e.g. gloo (tests: two ranks on one device)
        tt = torch.from_numpy(times) if on_host else torch.from_numpy(times).to(device)
        if on_host:
            counts_all = counts_all.cpu()
        dist.all_reduce(counts_all, op=dist.ReduceOp.SUM)          # the one exchange step
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        times = tt.cpu().numpy()
    counts = counts_all.cpu().numpy()
    if stats is not None:
        stats["d2h_bytes"] = counts.nbytes + times.nbytes
    if return_counts:
        return configs, counts, times, gd
    return assemble_scores(configs, counts, times, gd, n_splits)


def assemble_scores(configs, counts, times, gd, n_splits=10):
    """experiment.py:485-490: P/R/F from the integer counts with the reference's expressions."""
    out = {}
    for ci, config_keys in enumerate(configs):
        scores = {}
        for pid, name in enumerate(gd.proj_names):
            fp, fn, tp = (int(v) for v in counts[ci, pid])
            scores[name] = [fp, fn, tp, *hp.get_prf(fp, fn, tp)]
        fp, fn, tp = (int(v) for v in counts[ci, gd.n_proj])
        total = [fp, fn, tp, *hp.get_prf(fp, fn, tp)]
        # the reference divides by the literal 10 (experiment.py:489)
        out[tuple(config_keys)] = [times[ci, 0] / 10, times[ci, 1] / 10, scores, total]
    return out


def write_scores(tests_file="tests.json", scores_file="scores.pkl", return_stats=False, **kw):
    """``python experiment.py scores`` (experiment.py:493-501)."""
    import os
    stats = {}
    kw["stats"] = stats
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
    t0 = time.time()
    parsed = hp.parse_tests(tests_file)
    t1 = time.time()
    if kw.get("prepared") is None:
        kw["prepared"] = prepare(parsed, kw.get("configs"), kw.get("device"), kw.get("n_splits", 10), kw.get("cv", "stratified"))
    t2 = time.time()
    scores = run_grid(parsed, rank=rank, world=world, **kw)
    t3 = time.time()
    if rank == 0:
        with open(scores_file, "wb") as fd:
            pickle.dump(scores, fd)
    stats["breakdown_s"] = {"json_parse": t1 - t0, "host_prep_h2d_knn_calibration": t2 - t1, "grid": t3 - t2,
                            "pickle": time.time() - t3}
    if return_stats:
        return scores, time.time() - t0, stats
    return scores, time.time() - t0
