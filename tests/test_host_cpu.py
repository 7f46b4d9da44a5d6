"""CPU tests (no GPU): host logic, oracle vs golden fixtures, C-ABI surface."""
import ctypes
import json
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest

from util import make_dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _parsed(n):
    from flake16_framework_b200 import hostprep as hp, synth
    return hp.tests_to_arrays(synth.make_tests_dict(n, 16))


@pytest.mark.parametrize("fs", ["Flake16", "FlakeFlagger"])
def test_hostprep_bit_exact_vs_sklearn(fs):
    """A1-A3: load/scale/PCA/fold maps must be bit-identical to what the reference computes."""
    from sklearn.decomposition import PCA
    from sklearn.model_selection import StratifiedKFold
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from flake16_framework_b200 import hostprep as hp
    parsed = _parsed(3000)
    for ft in ("NOD", "OD"):
        X, y, _ = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES[ft], hp.FEATURE_SETS[fs])
        assert np.array_equal(hp.StandardScaler().fit_transform(X), StandardScaler().fit_transform(X))
        ref = Pipeline([("s", StandardScaler()), ("p", PCA(random_state=0))]).fit_transform(X)
        assert np.array_equal(hp.ScalePCA().fit_transform(X), ref)
        tf = hp.stratified_kfold_test_folds(y)
        for (a, b), (c, d) in zip(hp.kfold_split(tf), StratifiedKFold(10, shuffle=True, random_state=0).split(X, y)):
            assert np.array_equal(a, c) and np.array_equal(b, d)


def test_load_feat_lab_proj_matches_reference_expression(tmp_path):
    import json
    import ref_scores as R
    from flake16_framework_b200 import hostprep as hp, synth
    p = str(tmp_path / "tests.json")
    synth.make_tests_json(p, 500, 3)
    parsed = hp.parse_tests(p)
    for ft, lab in hp.FLAKY_TYPES.items():
        for fs, cols in hp.FEATURE_SETS.items():
            a = hp.feat_lab_proj(parsed, lab, cols)
            b = R.load_feat_lab_proj(lab, cols, p)
            for u, v in zip(a, b):
                assert np.array_equal(u, v)
            assert a[0].flags["F_CONTIGUOUS"] == b[0].flags["F_CONTIGUOUS"]


def test_get_prf():
    from flake16_framework_b200 import hostprep as hp
    assert hp.get_prf(0, 0, 0) == (None, None, None)
    assert hp.get_prf(1, 0, 0) == (0.0, None, None)
    assert hp.get_prf(0, 1, 0) == (None, 0.0, None)
    assert hp.get_prf(1, 1, 0) == (0.0, 0.0, None)
    assert hp.get_prf(1, 3, 1) == (0.5, 0.25, 2 * 0.5 * 0.25 / (0.5 + 0.25))


def test_tree_seeds_match_numpy():
    """A4: per-tree seeds as sklearn derives them (host code inside the shared library)."""
    from flake16_framework_b200 import ops
    for seed in (0, 1, 12345):
        ts, rr = ops.tree_seeds(seed, ops.KIND_RF, 100)
        rs = np.random.RandomState(seed)
        assert list(ts) == [rs.randint(np.iinfo(np.int32).max) for _ in range(100)]
        assert list(rr) == [np.random.RandomState(int(s)).randint(0, 2147483647) for s in ts]
        ts, rr = ops.tree_seeds(seed, ops.KIND_DT, 1)
        assert ts[0] == seed and rr[0] == np.random.RandomState(seed).randint(0, 2147483647)
    assert list(ops.tree_seeds(0, ops.KIND_ET, 3)[0]) == [209652396, 398764591, 924231285]


def test_abi_exports_every_declared_symbol():
    from flake16_framework_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "f16.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(f16_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "library does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    out = subprocess.check_output(["nm", "-D", _lib.LIB_PATH]).decode()
    assert "f16_forest_fit" in out


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flake16_framework_b200 import estimators as E, _lib
    X, y, _ = make_dataset(300)
    with pytest.raises(Exception):
        E.ExtraTreesClassifier(random_state=0).fit(X, y)


def test_product_never_imports_oracle_or_sklearn():
    for root, _, files in os.walk(os.path.join(ROOT, "flake16_framework_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import sklearn" not in src and "from sklearn" not in src, f
                assert "ref_scores" not in src and "samplers_np" not in src, f
    src = open(os.path.join(ROOT, "experiment.py")).read()
    assert "sklearn" not in src and "oracle" not in src


GOLD_CFGS = [("NOD", "Flake16", "None", "None", "Decision Tree"),
             ("NOD", "Flake16", "None", "None", "Random Forest"),
             ("OD", "FlakeFlagger", "Scaling", "SMOTE", "Extra Trees"),
             ("NOD", "Flake16", "PCA", "SMOTE ENN", "Extra Trees"),
             ("OD", "Flake16", "None", "Tomek Links", "Decision Tree"),
             ("OD", "FlakeFlagger", "PCA", "ENN", "Random Forest"),
             ("NOD", "FlakeFlagger", "None", "SMOTE Tomek", "Decision Tree")]


def test_oracle_pinned_against_reference_golden(tmp_path):
    """The oracle restatement (oracle/ref_scores.py) reproduces what the reference's own
    get_scores produced here (tests/golden/make_golden.py), count for count."""
    import ref_scores as R
    from flake16_framework_b200 import synth
    gold = pickle.load(open(os.path.join(GOLD, "scores_n1500_seed16.pkl"), "rb"))
    p = str(tmp_path / "tests.json")
    synth.make_tests_json(p, 1500, 16)
    for cfg in GOLD_CFGS:
        _, (keys, _, _, per_proj, total) = R.get_scores(cfg, p, R.make_config_grid())
        g_proj, g_total = gold[cfg]
        assert [int(v) for v in total[:3]] == g_total, cfg
        assert {str(k): [int(x) for x in v[:3]] for k, v in per_proj.items()} == g_proj, cfg


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flake16_framework_b200 import scores as S, hostprep as hp, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(400, 16))
    configs = S.all_config_keys()
    gd = S.GridData(parsed, configs)
    wanted, shards = S.plan_units(gd, configs, 10, world)
    cfg_index = {c: i for i, c in enumerate(configs)}
    counts = torch.zeros((len(configs), gd.n_proj + 1, 3), dtype=torch.int64)
    for u in shards[rank]:                               # fake, deterministic per-(item, config) counts
        ds, fold = u[0], u[1]
        for bal, models in wanted[u].items():
            for m in models:
                ci = cfg_index[ds + (bal, m)]
                counts[ci] += (fold + 1) * (ci + 1)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)        # the path's one exchange step
    if rank == 0:
        q.put((counts.numpy(), [len(s) for s in shards]))
    dist.destroy_process_group()


def test_world_size_2_sharding_and_reduce():
    """N>1 path on CPU (gloo): the unit plan is a partition and the reduced counts equal the
    single-process totals."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, sizes = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert sum(sizes) == 240 and abs(sizes[0] - sizes[1]) <= 60
    expect = np.array([sum(f + 1 for f in range(10)) * (ci + 1) for ci in range(216)])
    assert np.array_equal(counts[:, 0, 0], expect)


@pytest.mark.skipif(not os.path.exists("/root/reference/experiment.py"), reason="reference tree not present")
def test_scores_pickle_is_consumed_by_reference_figures(tmp_path):
    """the reference's own consumers of scores.pkl (get_top_tables,
    get_comparison_table, write_table - experiment.py:559-586, 621-630, 665-684) accept the dict
    this repo assembles (same keys, value layout, None handling, numpy.str_ project names)."""
    import sys
    import types
    import samplers_np
    for name in ("coverage", "shap", "imblearn", "imblearn.over_sampling", "imblearn.combine", "imblearn.under_sampling"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["coverage"].numbits = None
    sys.modules["shap"].TreeExplainer = None
    sys.modules["imblearn.over_sampling"].SMOTE = samplers_np.SMOTE
    sys.modules["imblearn.combine"].SMOTEENN = samplers_np.SMOTEENN
    sys.modules["imblearn.combine"].SMOTETomek = samplers_np.SMOTETomek
    sys.modules["imblearn.under_sampling"].TomekLinks = samplers_np.TomekLinks
    sys.modules["imblearn.under_sampling"].EditedNearestNeighbours = samplers_np.EditedNearestNeighbours
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_experiment", "/root/reference/experiment.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    from flake16_framework_b200 import hostprep as hp, scores as S, synth
    gold = pickle.load(open(os.path.join(GOLD, "scores_n1500_seed16.pkl"), "rb"))
    parsed = hp.tests_to_arrays(synth.make_tests_dict(1500, 16))
    configs = S.all_config_keys()
    gd = S.GridData(parsed, configs)
    counts = np.zeros((len(configs), gd.n_proj + 1, 3), dtype=np.int64)
    for ci, cfg in enumerate(configs):
        g_proj, g_total = gold[cfg]
        for pid, name in enumerate(gd.proj_names):
            counts[ci, pid] = g_proj[str(name)]
        counts[ci, gd.n_proj] = g_total
    scores = S.assemble_scores(configs, counts, np.full((len(configs), 2), 0.5), gd)

    tab_nod, tab_od = ref.get_top_tables(scores)
    assert len(tab_nod[0]) == 10 and len(tab_od[0]) == 10
    comp = ref.get_comparison_table(scores[("NOD", "FlakeFlagger", "None", "Tomek Links", "Extra Trees")],
                                    scores[("NOD", "Flake16", "PCA", "SMOTE", "Extra Trees")])
    assert comp[1][0][0] == "{\\bf Total}" and len(comp[1][0]) == 13
    os.chdir(tmp_path)
    ref.write_table("nod-top.tex", tab_nod)
    ref.write_table("nod-comp.tex", comp)
    assert open("nod-top.tex").read().count("\\\\") == 10


def test_abi_argument_validation_without_gpu():
    """Limits are enforced before any CUDA call (maximum sizes, dimensions, k): safe to check on CPU
    with dummy non-null pointers that are never dereferenced."""
    from flake16_framework_b200 import _lib
    L = _lib.lib()
    fake = ctypes.c_void_p(0x1000)
    h = ctypes.c_void_p()
    # n at / above the 24-bit row-id limit, d above 16, bad kind, bad max_features
    assert L.f16_forest_fit(fake, fake, 1 << 24, 16, fake, 1, 100, 4, 0, None, ctypes.byref(h)) == -1
    assert b"bad arguments" in L.f16_last_error()
    assert L.f16_forest_fit(fake, fake, 1000, 17, fake, 1, 100, 4, 0, None, ctypes.byref(h)) == -1
    assert L.f16_forest_fit(fake, fake, 1000, 16, fake, 7, 100, 4, 0, None, ctypes.byref(h)) == -1
    assert L.f16_forest_fit(fake, fake, 1000, 16, fake, 1, 100, 17, 0, None, ctypes.byref(h)) == -1
    assert L.f16_forest_fit(fake, fake, 1000, 16, None, 1, 100, 4, 0, None, ctypes.byref(h)) == -1   # RF needs sorted_idx
    assert b"sorted_idx" in L.f16_last_error()
    assert L.f16_knn(fake, 100, fake, 100, 16, 9, None, 0, fake, None) == -1                            # k > 8
    assert L.f16_knn(fake, 3, fake, 3, 16, 4, None, 0, fake, None) == -1                                # k > n
    bad_order = (ctypes.c_int32 * 16)(*([0] * 16))
    assert L.f16_knn(fake, 100, fake, 100, 16, 4, bad_order, 0, fake, None) == -1                       # not a permutation
    assert L.f16_confusion(fake, fake, fake, 10, 0, fake, None) == -1
    assert L.f16_tree_seeds(0, 1, 0, fake, fake) == -1


def test_small_modulus_table_in_common_header():
    """f16_rand_int_small (f16_common.cuh) replaces `r % m` (m <= 16) by a multiply-high with
    floor(2^32 / m) and one correction: check the committed table and the identity on the host."""
    import re
    src = open(os.path.join(ROOT, "flake16_framework_b200", "csrc", "f16_common.cuh")).read()
    table = [int(x, 16) for x in re.search(r"f16_magic16\[17\] = \{([^}]*)\}", src).group(1).replace("u", "").split(",")]
    assert len(table) == 17
    rs = np.random.RandomState(0)
    for m in range(1, 17):
        assert table[m] == min(0xffffffff, (1 << 32) // m)
        rr = np.concatenate([np.array([0, 1, m - 1, m, m + 1, 2**31 - 2, 2**31 - 1], dtype=np.uint64),
                             rs.randint(0, 2**31, 20000).astype(np.uint64)])
        q = (rr * np.uint64(table[m])) >> np.uint64(32)
        rem = rr - q * np.uint64(m)
        rem = np.where(rem >= m, rem - np.uint64(m), rem)
        assert np.array_equal(rem, rr % np.uint64(m)), m


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_unit_plan_is_a_balanced_partition(world):
    """(dataset, fold, balancing group) work items are assigned
    longest-processing-time-first; every rank derives the same plan without communicating, every
    item is owned exactly once, together they cover the 720 (dataset, balancing, fold) resamples,
    and the modelled loads of the ranks stay within 5 % of each other up to 8 GPUs."""
    from flake16_framework_b200 import hostprep as hp, scores as S, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(3000, 16))
    cfgs = S.all_config_keys()
    gd = S.GridData(parsed, cfgs)
    wanted, shards = S.plan_units(gd, cfgs, 10, world)
    wanted2, shards2 = S.plan_units(gd, cfgs, 10, world)
    assert shards == shards2                                    # deterministic
    flat = [u for sh in shards for u in sh]
    assert len(flat) == 240 and len(set(flat)) == 240
    assert sum(len(wanted[u]) for u in flat) == 720 and sum(len(m) for u in flat for m in wanted[u].values()) == 2160
    loads = [sum(S._unit_cost(gd, u, wanted[u]) for u in sh) for sh in shards]
    assert max(loads) <= 1.05 * (sum(loads) / world)
    # bench.py's parity check plans a single fold
    _, one = S.plan_units(gd, cfgs, 10, 1, folds=[0])
    assert len(one[0]) == 24 and all(u[1] == 0 for u in one[0])


@pytest.mark.parametrize("flaky", ["NOD", "OD"])
def test_stratified_group_kfold_matches_sklearn(flaky):
    """BASELINE.json configs[1]'s splitter (5-fold StratifiedGroupKFold over the projects): the
    restated fold map equals scikit-learn's, fold by fold."""
    from sklearn.model_selection import StratifiedGroupKFold
    from flake16_framework_b200 import hostprep as hp, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(6000, 16))
    X, y, proj = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES[flaky], hp.FEATURE_SETS["Flake16"])
    tf = hp.stratified_group_kfold_test_folds(y, proj, 5, True, 0)
    ref = StratifiedGroupKFold(n_splits=5, shuffle=True, random_state=0)
    for i, (tr, te) in enumerate(ref.split(X, y, groups=proj)):
        assert np.array_equal(np.flatnonzero(tf == i), te)
        assert np.array_equal(np.flatnonzero(tf != i), tr)
        assert not set(proj[te]) & set(proj[tr])                 # no project on both sides


def test_oracle_pinned_against_reference_golden_n20000(tmp_path):
    """The oracle restatement against the second, larger reference-generated golden (20 000 tests)
    on two configurations that finish in seconds."""
    path = os.path.join(GOLD, "scores_n20000_seed16.pkl")
    if not os.path.exists(path):
        pytest.skip("scores_n20000_seed16.pkl not generated")
    import ref_scores as R
    from flake16_framework_b200 import synth
    gold = pickle.load(open(path, "rb"))
    p = str(tmp_path / "tests.json")
    synth.make_tests_json(p, 20000, 16)
    for cfg in (("NOD", "Flake16", "None", "None", "Decision Tree"), ("OD", "FlakeFlagger", "Scaling", "None", "Extra Trees")):
        _, (keys, _, _, per_proj, total) = R.get_scores(cfg, p, R.make_config_grid())
        g_proj, g_total = gold[cfg]
        assert [int(v) for v in total[:3]] == g_total, cfg
        assert {str(k): [int(x) for x in v[:3]] for k, v in per_proj.items()} == g_proj, cfg


def test_bench_cpu_arm_is_a_fixed_sample_and_extrapolates_with_grid_multiplicities():
    """bench.py's CPU arm: the sample does not depend on --steps, covers all 18
    (balancing x model) kinds on one Flake16 and one FlakeFlagger dataset, and the grid estimate is
    6 x the sum of the sampled configs' (setup + 10 x fold) seconds over min(cores, 216) workers."""
    sys.path.insert(0, ROOT)
    import bench
    tasks = bench.cpu_sample_tasks()
    assert len(tasks) == 36 and len(set(tasks)) == 36
    assert {t[3] for t in tasks} == set(bench.BALANCINGS) and {t[4] for t in tasks} == set(bench.MODELS)
    assert {t[:3] for t in tasks} == set(bench.CPU_SAMPLE_DATASETS)
    assert any("Tomek" in t[3] for t in tasks) and any("ENN" in t[3] for t in tasks)
    res = [(t, {"setup_s": 1.0, "folds_s": 2.0 + i}, {}, [0, 0, 0]) for i, t in enumerate(tasks)]
    est = bench.cpu_grid_estimate(res)
    core_s = 6.0 * sum(1.0 + 10.0 * (2.0 + i) for i in range(36))
    pool = min(os.cpu_count() or 1, 216)
    assert abs(est["cpu_core_s_estimate"] - core_s) < 1e-9
    assert abs(est["value"] - 216.0 / (core_s / pool)) < 1e-12
    assert abs(est["slowest_config_s_estimate"] - (1.0 + 10.0 * 37.0)) < 1e-9


def test_native_tests_json_parser_equals_json_module(tmp_path):
    """hostprep.parse_tests reads tests.json with the library's one-pass scanner (f16_tests_parse);
    the arrays must be the ones json.load + the reference's loop give, bit for bit - on the
    synthetic table and on a hand-made file with the awkward cases (exponents, negative and huge
    numbers, "-0", escapes and unicode in test ids, odd whitespace)."""
    from flake16_framework_b200 import hostprep as hp, synth
    p = str(tmp_path / "tests.json")
    synth.make_tests_json(p, 3000, 16)
    a = hp._parse_tests_native(p)
    assert a is not None, "the library's scanner rejected the synthetic file"
    b = hp.tests_to_arrays(json.load(open(p)))
    assert np.array_equal(a[0].view(np.int64), b[0].view(np.int64)) and a[0].flags["C_CONTIGUOUS"]
    assert np.array_equal(a[1], b[1]) and a[1].dtype == b[1].dtype
    assert np.array_equal(a[2], b[2]) and a[2].dtype == b[2].dtype
    row = [0, 2, 1e-7, -3.5e+10, 123456789012345678901234567890, -0, 0.1, 1E3, 2.2250738585072014e-308,
           17, 5e-324, 1.7976931348623157e308, 9007199254740993, 3, 4, 5.000000000000001, 6, 7.25]
    odd = {"proj\u00e9t": {'t::a["x\\y"]': row, "t\u00e9::b": [1, 0] + [float(i) / 3 for i in range(16)]},
           "p2": {"n": [2, 1] + list(range(16))}}
    text = json.dumps(odd, indent=4, ensure_ascii=False).replace("[\n", "[ \t\n").replace(",\n", " ,\r\n")
    q = str(tmp_path / "odd.json")
    open(q, "w", encoding="utf-8").write(text)
    a = hp._parse_tests_native(q)
    b = hp.tests_to_arrays(json.load(open(q, encoding="utf-8")))
    assert a is not None
    assert np.array_equal(a[0].view(np.int64), np.asarray(b[0], dtype=np.float64).view(np.int64))
    assert np.array_equal(a[1], b[1]) and list(a[2]) == list(b[2])
    # outside the format: the scanner declines and parse_tests falls back to the json module
    r = str(tmp_path / "ragged.json")
    json.dump({"p": {"a": [0, 1, 2.0, 3.0], "b": [0, 1, 2.0]}}, open(r, "w"))
    assert hp._parse_tests_native(r) is None
    e = str(tmp_path / "esc.json")
    json.dump({"p\\q": {"a": [0, 1, 2.0]}}, open(e, "w"))
    assert hp._parse_tests_native(e) is None
    assert list(hp.parse_tests(e)[2]) == ["p\\q"]


def test_node_capacity_bookkeeping_and_bench_config_selectors():
    """The grid engine's learnt node capacities (first fit of a key: worst case; later: 1.5 x the
    largest ratio seen + 1024, never above 2n - 1; disabled instances always ask for the worst case)
    and bench.py's BASELINE config selectors."""
    sys.path.insert(0, ROOT)
    import bench
    from flake16_framework_b200 import scores as S
    S._NodeCaps._seen.clear()
    caps = S._NodeCaps()
    key = ("NOD", "Flake16", "None", "SMOTE", "Extra Trees", 18)
    assert caps.cap(key, 178200) == 0                       # unknown: f16_forest_fit_cap takes 0 as "2n - 1"
    caps.update(key, 178200, 16000)
    caps.update(key, 178200, 15000)                         # the largest ratio is kept
    assert caps.cap(key, 178200) == int(1.5 * (16000 / 178200) * 178200) + 1024
    assert caps.cap(key, 100) == 2 * 100 - 1                # never above the worst case
    assert S._NodeCaps().cap(key, 178200) > 0               # a later run starts from what was learnt
    assert S._NodeCaps(enabled=False).cap(key, 178200) == 0
    S._NodeCaps._seen.clear()
    allc = S.all_config_keys()
    assert len(bench.select_configs(allc, "grid216")) == 216 and len(bench.select_configs(allc, "slice")) == 18
    assert bench.select_configs(allc, "config2") == [("NOD", "Flake16", "None", "None", "Random Forest")]
    assert len(bench.select_configs(allc, "config3")) == 2 and len(bench.select_configs(allc, "config5")) == 12
    assert all(c[3] == "SMOTE ENN" and c[4] == "Extra Trees" for c in bench.select_configs(allc, "config5"))


def test_native_parser_number_round_trip_random(tmp_path):
    """20 000 random doubles (all exponents, subnormals, integers up to 2**63) written the way
    json.dump writes them come back from the library's scanner with the bit patterns json.load gives."""
    from flake16_framework_b200 import hostprep as hp
    rs = np.random.RandomState(7)
    bits = rs.randint(0, 2 ** 63, size=12000, dtype=np.int64).astype(np.uint64) | (rs.randint(0, 2, 12000).astype(np.uint64) << np.uint64(63))
    vals = bits.view(np.float64)
    vals = vals[np.isfinite(vals)]
    ints = [int(v) for v in rs.randint(-2 ** 62, 2 ** 62, size=4000)] + [2 ** 53 + 1, 2 ** 63 - 1, -(2 ** 63) + 1, 10 ** 18, 10 ** 19]
    smalls = [float(v) for v in rs.randn(4000)] + [5e-324, 2.2250738585072009e-308, 1.7976931348623157e308, 0.1, 1e23]
    nums = [float(v) for v in vals] + ints + smalls
    rs.shuffle(nums)
    rows = {}
    width = 18
    for i in range(0, len(nums) - width, width - 2):
        rows["t%d" % i] = [0, 1] + nums[i:i + width - 2]
    p = str(tmp_path / "rand.json")
    json.dump({"p": rows}, open(p, "w"), indent=4)
    a = hp._parse_tests_native(p)
    b = hp.tests_to_arrays(json.load(open(p)))
    assert a is not None
    assert np.array_equal(a[0].view(np.int64), np.asarray(b[0], dtype=np.float64).view(np.int64))
