"""Synthetic ``tests.json`` generator.

The reference's real ``tests.json`` is not in its tree (``.gitignore:84``; it is
shipped through a Google-Drive zip, ``README.rst:46-51``), so every run of this
repo uses a seeded synthetic table with the same wire format the reference's
``write_tests`` emits (``experiment.py:376-407``):

    {project: {test_nid: [req_runs, label, f0, ..., f15]}}

with ``label`` in {0 non-flaky, 1 OD flaky, 2 NOD flaky} (``experiment.py:50``) and the
16 columns in ``FEATURE_NAMES`` order (``experiment.py:65-71``).
"""

import json

import numpy as np

N_PROJECTS = 26
N_FEATURES = 16
NON_FLAKY, OD_FLAKY, FLAKY = 0, 1, 2

# columns emitted as JSON floats (rusage-derived + two continuous static metrics);
# the rest are integers, as in the reference's collated data (experiment.py:305)
_FLOAT_COLS = (3, 4, 5, 6, 7, 8, 12, 15)


def make_feature_table(n_tests, seed=16):
    """Returns (features float64 [n,16], labels int64 [n], project_sizes list)."""
    rng = np.random.RandomState(seed)
    n = int(n_tests)
    cols = [
        rng.poisson(300, n),            # Covered Lines
        rng.poisson(5, n),              # Covered Changes
        rng.poisson(200, n),            # Source Covered Lines
        rng.lognormal(-3, 1.5, n),      # Execution Time (continuous: breaks kNN ties)
        rng.poisson(3, n),              # Read Count
        rng.poisson(2, n),              # Write Count
        rng.poisson(1, n),              # Context Switches
        rng.binomial(3, 0.05, n) + 1,   # Max. Threads
        rng.lognormal(12, 2, n),        # Max. Memory
        rng.randint(1, 8, n),           # AST Depth
        rng.poisson(2, n),              # Assertions
        rng.poisson(3, n),              # External Modules
        rng.lognormal(4, 1, n),         # Halstead Volume
        rng.randint(1, 10, n),          # Cyclomatic Complexity
        rng.randint(1, 60, n),          # Test Lines of Code
        rng.uniform(30, 100, n),        # Maintainability
    ]
    feats = np.stack([np.asarray(c, dtype=np.float64) for c in cols], axis=1)

    # two noisy, partly non-linear scores -> labels
    z = (feats - feats.mean(0)) / (feats.std(0) + 1e-12)
    lz3 = np.log(feats[:, 3]); lz3 = (lz3 - lz3.mean()) / lz3.std()
    lz8 = np.log(feats[:, 8]); lz8 = (lz8 - lz8.mean()) / lz8.std()
    s_nod = (1.2 * lz3 + 0.8 * z[:, 6] + 0.7 * lz8 + 0.5 * z[:, 4]
             + 0.6 * (z[:, 7] > 0) + rng.normal(0, 1.0, n))
    s_od = (0.9 * z[:, 0] + 0.8 * z[:, 11] - 0.6 * z[:, 15] + 0.5 * z[:, 5]
            + 0.5 * z[:, 14] * (z[:, 9] > 0) + rng.normal(0, 1.0, n))
    labels = np.zeros(n, dtype=np.int64)
    n_nod = max(60, int(round(0.01 * n)))
    n_od = max(60, int(round(0.02 * n)))
    nod_idx = np.argsort(-s_nod, kind="stable")[:n_nod]
    labels[nod_idx] = FLAKY
    s_od2 = s_od.copy()
    s_od2[nod_idx] = -np.inf
    od_idx = np.argsort(-s_od2, kind="stable")[:n_od]
    labels[od_idx] = OD_FLAKY

    # 26 projects in contiguous blocks with sizes proportional to (i + 1)
    w = np.arange(1, N_PROJECTS + 1, dtype=np.float64)
    sizes = np.floor(w / w.sum() * n).astype(np.int64)
    sizes[-1] += n - sizes.sum()
    # shuffle rows so that positives are spread over the projects
    perm = rng.permutation(n)
    return feats[perm], labels[perm], sizes.tolist()


def make_tests_dict(n_tests, seed=16):
    feats, labels, sizes = make_feature_table(n_tests, seed)
    tests = {}
    row = 0
    for p, size in enumerate(sizes):
        proj = {}
        for _ in range(size):
            f = feats[row]
            vals = [float(f[j]) if j in _FLOAT_COLS else int(f[j])
                    for j in range(N_FEATURES)]
            proj["t%07d" % row] = [0, int(labels[row])] + vals
            row += 1
        tests["p%02d" % p] = proj
    return tests


def make_tests_json(path, n_tests, seed=16, indent=4):
    """Writes the synthetic table in the reference's wire format
    (``json.dump(tests, fd, indent=4)``, experiment.py:406-407)."""
    tests = make_tests_dict(n_tests, seed)
    with open(path, "w") as fd:
        json.dump(tests, fd, indent=indent)
    return path
