"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg).

CPU restatement of the reference's ``scores`` hot path, experiment.py:410-501
(``load_feat_lab_proj``, ``div_none``, ``get_prf``, ``get_scores``, ``write_scores``),
line for line, on top of

* the in-image **scikit-learn 1.9.0** estimators (the reference pins ``scikit-learn==1.0.2``,
  requirements.txt:31 - a third-party dependency whose source is not under /root/reference;
  version skew discussed in SURVEY.md section 7.3 H5) and
* ``oracle/samplers_np.py`` (numpy restatement of imbalanced-learn 0.9.0, not installable here).

``import experiment`` itself fails in this image (``coverage``, ``shap``, ``imblearn`` are
missing, experiment.py:18-28); tests/golden/make_golden.py shows the other route - importing
the reference file with those three modules stubbed - and pins this restatement against it.

Parity status: the reference's own tests hold no golden vector for this path
(SURVEY.md section 4), so the pin is "reference get_scores() run here with sklearn 1.9.0";
the sampler half stays PARITY UNPINNED (see samplers_np.py).
"""

import itertools
import json
import os
import pickle
import sys
import time

import numpy as np
from sklearn.decomposition import PCA
from sklearn.ensemble import ExtraTreesClassifier, RandomForestClassifier
from sklearn.model_selection import StratifiedKFold
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import StandardScaler
from sklearn.tree import DecisionTreeClassifier

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from samplers_np import (SMOTE, SMOTEENN, SMOTETomek, TomekLinks,  # noqa: E402
                         EditedNearestNeighbours)

TESTS_FILE = "tests.json"
SCORES_FILE = "scores.pkl"
NON_FLAKY, OD_FLAKY, FLAKY = 0, 1, 2
N_FEATURES = 16


def make_config_grid(n_estimators=100):
    """experiment.py:73-100 (fresh estimator instances)."""
    return (
        {"NOD": FLAKY, "OD": OD_FLAKY},
        {"Flake16": range(N_FEATURES), "FlakeFlagger": (0, 1, 2, 3, 10, 11, 14)},
        {"None": None, "Scaling": StandardScaler(),
         "PCA": Pipeline([("s", StandardScaler()), ("p", PCA(random_state=0))])},
        {"None": None, "Tomek Links": TomekLinks(), "SMOTE": SMOTE(random_state=0),
         "ENN": EditedNearestNeighbours(), "SMOTE ENN": SMOTEENN(random_state=0),
         "SMOTE Tomek": SMOTETomek(random_state=0)},
        {"Extra Trees": ExtraTreesClassifier(random_state=0, n_estimators=n_estimators),
         "Random Forest": RandomForestClassifier(random_state=0, n_estimators=n_estimators),
         "Decision Tree": DecisionTreeClassifier(random_state=0)},
    )


CONFIG_GRID = make_config_grid()


def load_feat_lab_proj(flaky_label, feature_set, tests_file=None):   # experiment.py:410-427
    with open(tests_file or TESTS_FILE, "r") as fd:
        tests = json.load(fd)
    features, labels, projects = [], [], []
    for proj, tests_proj in tests.items():
        projects += [proj] * len(tests_proj)
        for (_, label_nid, *features_nid) in tests_proj.values():
            features.append(features_nid)
            labels.append(label_nid)
    features = np.array(features)[:, feature_set]
    labels = np.array(labels) == flaky_label
    projects = np.array(projects)
    return features, labels, projects


def div_none(a, b):                                                  # experiment.py:430-431
    return a / b if b else None


def get_prf(fp, fn, tp):                                             # experiment.py:434-443
    p = div_none(tp, tp + fp)
    r = div_none(tp, tp + fn)
    if p is None or r is None:
        f = None
    else:
        f = div_none(2 * p * r, p + r)
    return p, r, f


def get_scores(config_keys, tests_file=None, grid=None, n_splits=10, max_folds=None, timing=None):
    """experiment.py:446-490.  ``max_folds`` (not in the reference) stops after that many
    folds - used only to bound the CPU-baseline sample in bench.py; ``timing`` (a dict, not in
    the reference either) receives the seconds spent before the fold loop (JSON parse +
    preprocessing, ``setup_s``) and inside it (``folds_s``) for that bench leg's extrapolation."""
    t_enter = time.time()
    grid = grid or CONFIG_GRID
    config_vals = [grid[i][k] for i, k in enumerate(config_keys)]
    flaky_label, feature_set, preprocessing, balancing, model = config_vals
    features, labels, projects = load_feat_lab_proj(flaky_label, feature_set, tests_file)
    fold = StratifiedKFold(n_splits=n_splits, shuffle=True, random_state=0)

    if preprocessing is not None:
        features = preprocessing.fit_transform(features)

    t_train = t_test = 0
    scores, scores_total = {proj: [0] * 6 for proj in projects}, [0] * 6
    t_loop = time.time()

    for i, (train, test) in enumerate(fold.split(features, labels)):
        if max_folds is not None and i >= max_folds:
            break
        features_train, labels_train = features[train], labels[train]
        features_test, labels_test = features[test], labels[test]
        projects_test = projects[test]

        if balancing is not None:
            features_train, labels_train = balancing.fit_resample(
                features_train, labels_train
            )

        t_start = time.time()
        model.fit(features_train, labels_train)
        t_train += time.time() - t_start

        t_start = time.time()
        labels_pred = model.predict(features_test)
        t_test += time.time() - t_start

        for j, labels_test_j in enumerate(labels_test):
            k = int(2 * labels_test_j + labels_pred[j]) - 1
            if k == -1:
                continue
            scores[projects_test[j]][k] += 1
            scores_total[k] += 1

    if timing is not None:
        timing["setup_s"] = t_loop - t_enter
        timing["folds_s"] = time.time() - t_loop
    for scores_proj in [*scores.values(), scores_total]:
        scores_proj[3:] = get_prf(*scores_proj[:3])

    return ", ".join(config_keys), (
        config_keys, t_train / 10, t_test / 10, scores, scores_total
    )


def all_config_keys(grid=None):
    grid = grid or CONFIG_GRID
    return list(itertools.product(*[d.keys() for d in grid]))          # experiment.py:494


def _worker(args):
    config_keys, tests_file, n_splits, max_folds, n_estimators = args
    grid = make_config_grid(n_estimators)
    try:        # one single-threaded call chain per worker, like the reference's Pool(N_PROC)
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            return get_scores(config_keys, tests_file, grid, n_splits, max_folds)
    except ImportError:
        return get_scores(config_keys, tests_file, grid, n_splits, max_folds)


def _timed_worker(args):
    """get_scores on ``max_folds`` folds + its setup / fold-loop seconds (bench.py's CPU arm)."""
    config_keys, tests_file, n_splits, max_folds, n_estimators = args
    grid = make_config_grid(n_estimators)
    timing = {}
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            _, (_, _, _, scores, total) = get_scores(config_keys, tests_file, grid, n_splits, max_folds, timing)
    except ImportError:
        _, (_, _, _, scores, total) = get_scores(config_keys, tests_file, grid, n_splits, max_folds, timing)
    counts = {str(p): [int(v) for v in s[:3]] for p, s in scores.items()}
    return tuple(config_keys), timing, counts, [int(v) for v in total[:3]]


def run_configs_timed(configs, tests_file, processes, n_splits=10, max_folds=1, n_estimators=100):
    """One process per task (maxtasksperchild=1: a worker never carries a warm cache from one
    task to the next, like distinct Pool workers in the reference)."""
    from multiprocessing import Pool
    args = [(c, tests_file, n_splits, max_folds, n_estimators) for c in configs]
    with Pool(processes=processes, maxtasksperchild=1) as pool:
        return list(pool.imap_unordered(_timed_worker, args, chunksize=1))


def run_configs(configs, tests_file, processes=None, n_splits=10, max_folds=None,
                n_estimators=100):
    """experiment.py:493-498 (Pool(N_PROC) + imap_unordered), on a config subset."""
    from multiprocessing import Pool
    args = [(c, tests_file, n_splits, max_folds, n_estimators) for c in configs]
    processes = processes or os.cpu_count()
    if processes == 1 or len(args) == 1:
        results = [_worker(a) for a in args]
    else:
        with Pool(processes=processes) as pool:
            results = list(pool.imap_unordered(_worker, args))
    return {config_keys: rest for _, (config_keys, *rest) in results}


def write_scores(tests_file=TESTS_FILE, scores_file=SCORES_FILE, processes=None):
    scores = run_configs(all_config_keys(), tests_file, processes)
    with open(scores_file, "wb") as fd:
        pickle.dump(scores, fd)


if __name__ == "__main__":
    write_scores()
