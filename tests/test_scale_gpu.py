"""GPU parity AT THE SIZE THE BENCH RUNS (100 000 tests, 90 000-row folds, ~178 000 rows after
SMOTE): the paths that only exist at that scale - the DFS stack spilling past its shared-memory
cache, GLOBAL-regime nodes at depth, the calibrated k-NN strategies (tensor-core filter, sorted
sweep), SMOTE'd training sets - compared with scikit-learn / the oracle in the product flow.
Also: the reference's own get_scores loop driving this repo's estimator objects (INTEGRATION.md
route 2), and the N > 1 path of run_grid."""
import os
import sys

import numpy as np
import pytest
import torch

from util import compare_trees, make_dataset, tree_arrays_sklearn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

N_BENCH = 100000


@pytest.fixture(scope="module")
def bench_fold():
    """Fold 1 of the bench's NOD / Flake16 / Scaling dataset and its SMOTE'd training set (oracle)."""
    from flake16_framework_b200 import hostprep as hp
    import samplers_np as O
    X, y, _ = make_dataset(N_BENCH, prep="Scaling")
    tr, te = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
    Xtr, ytr = np.ascontiguousarray(X[tr]), y[tr]
    Xs, ys = O.SMOTE(random_state=0).fit_resample(Xtr, ytr)
    return {"X": X, "y": y, "tr": tr, "te": te, "Xtr": Xtr, "ytr": ytr, "Xs": np.ascontiguousarray(Xs), "ys": ys}


def _sk(kind, n_estimators):
    import sklearn.ensemble as E
    import sklearn.tree as T
    return {"DT": lambda: T.DecisionTreeClassifier(random_state=0),
            "RF": lambda: E.RandomForestClassifier(random_state=0, n_estimators=n_estimators),
            "ET": lambda: E.ExtraTreesClassifier(random_state=0, n_estimators=n_estimators)}[kind]()


def _ours(kind, n_estimators):
    from flake16_framework_b200 import estimators as E
    return {"DT": lambda: E.DecisionTreeClassifier(random_state=0),
            "RF": lambda: E.RandomForestClassifier(random_state=0, n_estimators=n_estimators),
            "ET": lambda: E.ExtraTreesClassifier(random_state=0, n_estimators=n_estimators)}[kind]()


@pytest.mark.parametrize("which", ["fold", "smote"])
@pytest.mark.parametrize("kind", ["ET", "RF", "DT"])
def test_trees_bit_identical_at_bench_size(cuda, bench_fold, kind, which):
    """3 trees of each family on the 90 000-row fold and on its ~178 000-row SMOTE'd set: every
    tree_ array bit-identical to scikit-learn's, and the trees are deep enough to leave the
    shared-memory stack cache (40 records), so the spill path is exercised."""
    Xtr, ytr = (bench_fold["Xtr"], bench_fold["ytr"]) if which == "fold" else (bench_fold["Xs"], bench_fold["ys"])
    Xte = bench_fold["X"][bench_fold["te"]]
    ref = _sk(kind, 3).fit(Xtr, ytr)
    our = _ours(kind, 3).fit(Xtr, ytr)
    our.forest_.status()
    ref_trees = [ref] if kind == "DT" else ref.estimators_
    counts = our.forest_.node_counts()
    errs, depth = [], 0
    for t, rt in enumerate(ref_trees):
        errs += compare_trees(our.forest_.export_tree(t, int(counts[t])), tree_arrays_sklearn(rt), "%s/%s tree %d" % (kind, which, t))
        depth = max(depth, rt.tree_.max_depth)
    assert not errs, "\n".join(errs[:10])
    assert np.array_equal(ref.predict(Xte), our.predict(Xte))
    if which == "smote":     # scikit-learn's depths here: ET 37/43/47, RF 42/37/35, DT 53 - deterministic
        assert depth >= 40, "depth %d: the DFS stack never left its 40-record shared-memory cache" % depth


@pytest.mark.parametrize("cfg", [dict(prep="None"), dict(prep="Scaling"), dict(prep="PCA"),
                                 dict(prep="None", fset="FlakeFlagger")])
def test_knn_at_bench_size_through_calibrated_strategy(cuda, cfg):
    """k = 4 neighbour lists of the 90 000-row fold, searched with the strategy the grid engine's
    own calibration picks for that dataset (tensor-core filter for Scaling / PCA, sorted sweep or
    early exit for raw data), against sklearn.NearestNeighbors."""
    from sklearn.neighbors import NearestNeighbors
    from flake16_framework_b200 import hostprep as hp, ops
    X, y, _ = make_dataset(N_BENCH, **cfg)
    tr, _ = next(iter(hp.kfold_split(hp.stratified_kfold_test_folds(y))))
    A = np.ascontiguousarray(X[tr])
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    strategy = ops.calibrate_knn(Xd, ops.variance_order(X))
    Ad = torch.from_numpy(A).cuda()
    ours = ops.knn(Ad, Ad, 4, strategy).cpu().numpy()
    ref = NearestNeighbors(n_neighbors=4).fit(A).kneighbors(A, return_distance=False)
    same = (ours == ref).all(axis=1)
    assert same.all(), "strategy %d: %d of %d rows differ" % (strategy[1], int((~same).sum()), len(same))


def test_knn_ties_report(cuda):
    """duplicated and integer-only rows.  With exact ties the neighbour ORDER is
    implementation-defined: this kernel's rule is (exact distance, lowest index);
    sklearn's brute path ranks expanded-form distances from a heap, its KD-tree path visits leaves
    in tree order.  What must hold for every row is the neighbour DISTANCE multiset (a tie never
    changes which distances are the k smallest); identical index lists are reported, not required."""
    from sklearn.neighbors import NearestNeighbors
    from flake16_framework_b200 import ops
    rs = np.random.RandomState(5)
    for d in (16, 7):                          # d = 16: brute force in sklearn; d = 7: KD-tree
        X = rs.randint(0, 4, size=(4000, d)).astype(np.float64)        # integer-only, heavy ties
        X[1000:1400] = X[:400]                                          # exact duplicates
        Xd = torch.from_numpy(X).cuda()
        ours = ops.knn(Xd, Xd, 4).cpu().numpy()
        nn = NearestNeighbors(n_neighbors=4).fit(X)
        rd, ri = nn.kneighbors(X)
        od = np.sqrt(((X[:, None, :] - X[ours]) ** 2).sum(-1))
        assert np.allclose(np.sort(od, axis=1), np.sort(rd, axis=1), rtol=0, atol=1e-12), d
        # our tie rule: among equal distances the lower index comes first
        for j in range(3):
            tie = od[:, j] == od[:, j + 1]
            assert np.all(ours[tie, j] < ours[tie, j + 1]), d
        agree = float((np.sort(ours, axis=1) == np.sort(ri, axis=1)).all(axis=1).mean())
        print("d=%d: neighbour-set agreement with sklearn on tie-heavy data: %.3f" % (d, agree))


def test_reference_loop_with_our_estimators(cuda, tmp_path):
    """INTEGRATION.md route 2: the REFERENCE's get_scores loop (oracle/ref_scores.get_scores is
    experiment.py:446-490 line for line) with this repo's estimator objects slotted into
    CONFIG_GRID - numpy in, numpy out - reproduces the reference-generated golden counts."""
    import pickle
    import ref_scores as R
    from flake16_framework_b200 import estimators as E, synth
    gold = pickle.load(open(os.path.join(ROOT, "tests", "golden", "scores_n1500_seed16.pkl"), "rb"))
    p = str(tmp_path / "tests.json")
    synth.make_tests_json(p, 1500, 16)
    ref_grid = R.make_config_grid()
    grid = (ref_grid[0], ref_grid[1], ref_grid[2],
            {"None": None, "Tomek Links": E.TomekLinks(), "SMOTE": E.SMOTE(random_state=0),
             "ENN": E.EditedNearestNeighbours(), "SMOTE ENN": E.SMOTEENN(random_state=0),
             "SMOTE Tomek": E.SMOTETomek(random_state=0)},
            {"Extra Trees": E.ExtraTreesClassifier(random_state=0), "Random Forest": E.RandomForestClassifier(random_state=0),
             "Decision Tree": E.DecisionTreeClassifier(random_state=0)})
    for cfg in (("NOD", "Flake16", "None", "None", "Decision Tree"),
                ("NOD", "Flake16", "Scaling", "SMOTE", "Extra Trees"),
                ("OD", "FlakeFlagger", "PCA", "SMOTE ENN", "Random Forest"),
                ("OD", "Flake16", "None", "Tomek Links", "Random Forest"),
                ("NOD", "FlakeFlagger", "Scaling", "ENN", "Extra Trees"),
                ("NOD", "Flake16", "PCA", "SMOTE Tomek", "Decision Tree")):
        _, (keys, _, _, per_proj, total) = R.get_scores(cfg, p, grid)
        g_proj, g_total = gold[cfg]
        assert [int(v) for v in total[:3]] == g_total, cfg
        assert {str(k): [int(x) for x in v[:3]] for k, v in per_proj.items()} == g_proj, cfg


def _rank_worker(rank, world, port, n_tests, q):
    """One rank of run_grid on cuda:0 (both ranks share the device; the exchange step runs over
    gloo because NCCL refuses two ranks on one GPU - the code path around it is the product's)."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flake16_framework_b200 import scores as S, hostprep as hp, synth
    parsed = hp.tests_to_arrays(synth.make_tests_dict(n_tests, 16))
    cfgs = [c for c in S.all_config_keys() if c[0] == "NOD" and c[2] == "Scaling"]
    out = S.run_grid(parsed, cfgs, rank=rank, world=world, device=torch.device("cuda", 0))
    if rank == 0:
        q.put({k: (v[2], v[3]) for k, v in out.items()})
    dist.destroy_process_group()


def test_run_grid_two_ranks_equals_one(cuda):
    """the multi-rank path of run_grid itself (sharded work items, all-reduce of the
    counts, rank-0 assembly) gives the single-rank result."""
    import torch.multiprocessing as mp
    from flake16_framework_b200 import scores as S, hostprep as hp, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, 1500, q)) for r in range(2)]
    for p in procs:
        p.start()
    two = q.get(timeout=300)
    for p in procs:
        p.join(60)
    parsed = hp.tests_to_arrays(synth.make_tests_dict(1500, 16))
    cfgs = [c for c in S.all_config_keys() if c[0] == "NOD" and c[2] == "Scaling"]
    one = S.run_grid(parsed, cfgs)
    assert set(one) == set(two) and len(one) == 36
    for k in one:
        assert one[k][3][:3] == two[k][1][:3], k
        assert {p: v[:3] for p, v in one[k][2].items()} == {p: v[:3] for p, v in two[k][0].items()}, k


def test_node_capacity_overflow_is_detected_and_retried(cuda):
    """A fit whose trees outgrow an explicit node capacity reports F16_ERR_OVERFLOW through
    f16_forest_status (nothing is written out of bounds, predict stays finite); the grid engine
    redoes the pass with the worst-case capacity when its learnt capacities are too small."""
    from flake16_framework_b200 import ops, scores as S, hostprep as hp, synth
    X, y, _ = make_dataset(3000)
    Xd = torch.from_numpy(X).cuda()
    Xrow = ops.rows_f32(Xd)
    yd = torch.from_numpy(y.astype(np.uint8)).cuda()
    for kind in (ops.KIND_ET, ops.KIND_RF, ops.KIND_DT):
        f = ops.forest_fit(Xrow, yd, 16, kind, 4, 0, node_cap=9)
        f.predict(Xrow)
        with pytest.raises(ops.F16Overflow):
            f.status()
        f.free()
        f = ops.forest_fit(Xrow, yd, 16, kind, 4, 0, node_cap=0)
        f.status()
        assert f.max_nodes() == int(f.node_counts().max()) > 9
        f.free()
    parsed = hp.tests_to_arrays(synth.make_tests_dict(1500, 16))
    cfgs = [c for c in S.all_config_keys() if c[:3] == ("NOD", "Flake16", "None")]
    prepared = S.prepare(parsed, cfgs)
    good = S.run_grid(parsed, cfgs, prepared=prepared)
    assert prepared[1].caps.ratio and all(0 < r < 2 for r in prepared[1].caps.ratio.values())
    prepared[1].caps.cap = lambda key, n: 9             # poison the learnt capacities: every fit overflows
    again = S.run_grid(parsed, cfgs, prepared=prepared)
    assert {k: v[3][:3] for k, v in good.items()} == {k: v[3][:3] for k, v in again.items()}
