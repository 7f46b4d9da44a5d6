"""the step BEFORE the hot path - raw run data -> ``tests.json``.

Host-side, I/O-bound bookkeeping (no GPU work): the reference's ``python experiment.py tests``
(experiment.py:242-407).  The function names, argument meaning and the in-memory shapes follow the
reference so that its own unit tests (test_experiment.py) read the same against this module:

    collated            {project: [test_data, test_fn_data, test_files, churn]}
    test_data           {nodeid: [runs, coverage, rusage, fn_id]}
    runs                {"baseline" | "shuffle": [n_runs, n_fail, first_fail_run, first_pass_run]}
    coverage            {file relative to the project dir: set(line numbers)}
    tests.json          {project: {nodeid: [req_runs, label, 3 coverage features, 6 rusage values,
                                            7 static features]}}        (the wire format of the hot path)

File naming in ``data/``: ``{project}_{mode}_{run}.{ext}`` with mode in baseline / shuffle (TSV of
``outcome<TAB>nodeid``) and testinspect (``.sqlite3`` coverage.py line data with one context per
test, ``.tsv`` of six resource-usage numbers + nodeid, ``.pkl`` static analysis).
"""

import json
import os
import pickle
import sqlite3

NON_FLAKY, OD_FLAKY, FLAKY = 0, 1, 2                        # experiment.py:48
N_RUNS = {"baseline": 2500, "shuffle": 2500}                # experiment.py:50 (the two run modes)
DATA_DIR = "data"                                           # experiment.py:37
SUBJECTS_DIR = os.path.join("/", "home", "user", "subjects")   # experiment.py:39-40
TESTS_FILE = "tests.json"                                   # experiment.py:32


def numbits_to_nums(numbits):
    """coverage.py's ``numbits`` blob -> sorted line numbers: bit b of byte i stands for the
    number 8 i + b (coverage/numbits.py; the package is not installed here)."""
    return [8 * i + b for i, byte in enumerate(bytes(numbits)) for b in range(8) if (byte >> b) & 1]


def nums_to_numbits(nums):
    """Inverse of :func:`numbits_to_nums` (used by the tests to build coverage databases)."""
    nums = list(nums)
    out = bytearray(max(nums) // 8 + 1 if nums else 0)
    for n in nums:
        out[n // 8] |= 1 << (n % 8)
    return bytes(out)


def iter_data_dir(data_dir=None):
    """experiment.py:242-247: (path, project, mode, run number, extension) per data file."""
    data_dir = DATA_DIR if data_dir is None else data_dir
    for name in os.listdir(data_dir):
        proj, mode, rest = name.split("_", 2)
        run_n, ext = rest.split(".", 1)
        yield os.path.join(data_dir, name), proj, mode, int(run_n), ext


def _fields(lines, n_split):
    for line in lines:
        yield line.strip().split("\t", n_split)


def _entry(collated_proj, nid):
    return collated_proj[0].setdefault(nid, [{}, {}, None, None])


def update_collated_runs(fd, mode, run_n, collated_proj):
    """One run of one mode (experiment.py:260-277): counts runs and failures per test and keeps
    the lowest run number that failed / that passed."""
    for outcome, nid in _fields(fd, 1):
        tally = _entry(collated_proj, nid)[0].setdefault(mode, [0, 0, None, None])
        tally[0] += 1
        failed = "failed" in outcome
        if failed:
            tally[1] += 1
        slot = 2 if failed else 3
        tally[slot] = run_n if tally[slot] is None else min(tally[slot], run_n)


def update_collated_cov(con, proj, collated_proj):
    """Per-test line coverage out of a coverage.py database (experiment.py:280-299): contexts are
    test node ids, file paths are made relative to the project's checkout."""
    cur = con.cursor()
    contexts = dict(cur.execute("SELECT id, context FROM context").fetchall())
    base = os.path.join(SUBJECTS_DIR, proj, proj)
    files = {fid: os.path.relpath(path, start=base) for fid, path in cur.execute("SELECT id, path FROM file").fetchall()}
    for context_id, file_id, blob in cur.execute("SELECT context_id, file_id, numbits FROM line_bits").fetchall():
        _entry(collated_proj, contexts[context_id])[1][files[file_id]] = set(numbits_to_nums(blob))


def update_collated_rusage(fd, collated_proj):
    """experiment.py:302-305: six resource-usage floats, then the node id."""
    for *usage, nid in _fields(fd, 6):
        _entry(collated_proj, nid)[2] = [float(x) for x in usage]


def update_collated_static(fd, collated_proj):
    """experiment.py:308-313: pickle of (nodeid -> function id, per-function static features,
    test files, churn)."""
    fn_ids, *collated_proj[1:] = pickle.load(fd)
    for nid, fid in fn_ids.items():
        _entry(collated_proj, nid)[3] = fid


def get_collated(data_dir=None):
    """experiment.py:316-336."""
    collated = {}
    for path, proj, mode, run_n, ext in iter_data_dir(data_dir):
        collated_proj = collated.setdefault(proj, [{}, None, None, None])
        if mode in ("baseline", "shuffle"):
            with open(path, "r") as fd:
                update_collated_runs(fd, mode, run_n, collated_proj)
        elif mode == "testinspect":
            if ext == "sqlite3":
                with sqlite3.connect(path) as con:
                    update_collated_cov(con, proj, collated_proj)
            elif ext == "tsv":
                with open(path, "r") as fd:
                    update_collated_rusage(fd, collated_proj)
            elif ext == "pkl":
                with open(path, "rb") as fd:
                    update_collated_static(fd, collated_proj)
    return collated


def get_req_runs_label_nid(runs_nid):
    """Label of one test and the number of runs that were needed to see it (experiment.py:339-359):
    incomplete data -> (0, None); never / always failing in both modes -> NON_FLAKY; consistent in
    baseline but not under shuffling -> OD_FLAKY (order dependent); inconsistent in baseline -> FLAKY."""
    base = runs_nid.get("baseline", [0, 0, None, None])
    shuf = runs_nid.get("shuffle", [0, 0, None, None])
    if base[0] != N_RUNS["baseline"] or shuf[0] != N_RUNS["shuffle"]:
        return 0, None
    if base[1] == 0:                                   # always passes in isolation order
        return (0, NON_FLAKY) if shuf[1] == 0 else (shuf[2], OD_FLAKY)
    if base[1] == base[0]:                             # always fails
        return (0, NON_FLAKY) if shuf[1] == shuf[0] else (shuf[3], OD_FLAKY)
    return max(base[2], base[3]), FLAKY


def get_features_nid_cov(cov_nid, test_files, churn):
    """(covered lines, covered changes, covered lines outside test files) - experiment.py:362-373."""
    n_lines = n_changes = n_src_lines = 0
    for name, lines in cov_nid.items():
        n_lines += len(lines)
        per_line = churn.get(name, {})
        n_changes += sum(per_line.get(line, 0) for line in lines)
        if name not in test_files:
            n_src_lines += len(lines)
    return n_lines, n_changes, n_src_lines


def collate_tests(collated):
    """The dict ``write_tests`` dumps (experiment.py:378-404): projects and node ids in
    case-insensitive order; projects / tests with any missing part are dropped."""
    tests = {}
    for proj in sorted(collated, key=str.lower):
        if not all(collated[proj]):
            continue
        test_data, fn_data, test_files, churn = collated[proj]
        rows = {}
        for nid in sorted(test_data, key=str.lower):
            if not all(test_data[nid]):
                continue
            runs, cov, rusage, fid = test_data[nid]
            req_runs, label = get_req_runs_label_nid(runs)
            if label is None:
                continue
            rows[nid] = (req_runs, label, *get_features_nid_cov(cov, test_files, churn), *rusage, *fn_data[fid])
        if rows:
            tests[proj] = rows
    return tests


def write_tests(data_dir=None, tests_file=None):
    """``python experiment.py tests`` (experiment.py:376-407): same file name and JSON layout
    (``indent=4``), so the output is what ``scores`` - the reference's or this repo's - reads."""
    tests = collate_tests(get_collated(data_dir))
    with open(TESTS_FILE if tests_file is None else tests_file, "w") as fd:
        json.dump(tests, fd, indent=4)
    return tests
