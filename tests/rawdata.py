"""Deterministic synthetic ``data/`` directory in the layout ``experiment.py tests`` collates
(reference experiment.py:242-336): per project and mode ``{project}_{mode}_{run}.{ext}``.

Built to exercise every branch of the collation: all three labels by both routes (first failing /
first passing run), tests with incomplete run counts, a test with no resource usage, the reference's
``all([...])`` quirk that drops function id 0, case-insensitive ordering of projects and node ids,
and a project with no static-analysis pickle (dropped as a whole).
"""
import os
import pickle
import sqlite3

from flake16_framework_b200 import collate

N_RUNS_SMALL = 4        # runs per mode in the fixture (the reference's N_RUNS is patched to this)


def _outcomes(pattern, run):
    """pattern: string of 'p'/'f' per run, or None when the test is missing from that run."""
    if pattern is None or run >= len(pattern):
        return None
    return "passed" if pattern[run] == "p" else "failed"


# nodeid -> (baseline outcomes per run, shuffle outcomes per run)
_TESTS = {
    "tests/test_a.py::test_stable":        ("pppp", "pppp"),     # NON_FLAKY, req_runs 0
    "Tests/test_B.py::test_od_fail":       ("pppp", "ppfp"),     # OD_FLAKY via first failing shuffle run (2)
    "tests/test_c.py::test_broken":        ("ffff", "ffff"),     # NON_FLAKY (always fails)
    "tests/test_d.py::test_od_pass":       ("ffff", "fpff"),     # OD_FLAKY via first passing shuffle run (1)
    "tests/test_e.py::Test::test_flaky":   ("pfpp", "pppp"),     # FLAKY, req_runs max(1, 0) = 1
    "tests/test_f.py::test_flaky_late":    ("fffp", "ffpf"),     # FLAKY, req_runs max(0, 3) = 3
    "tests/test_g.py::test_incomplete":    ("ppp", "pppp"),      # only 3 baseline runs -> no label -> dropped
    "tests/test_h.py::test_no_rusage":     ("pppp", "pppp"),     # no resource usage row -> dropped
    "tests/test_i.py::test_fid_zero":      ("pppp", "pppp"),     # function id 0 -> all([...]) is False -> dropped
}


def make_raw_data(data_dir, projects=("ProjB", "proja", "projC")):
    os.makedirs(data_dir, exist_ok=True)
    for pi, proj in enumerate(projects):
        nids = list(_TESTS)
        for mode_i, mode in enumerate(("baseline", "shuffle")):
            for run in range(N_RUNS_SMALL):
                with open(os.path.join(data_dir, "%s_%s_%d.tsv" % (proj, mode, run)), "w") as fd:
                    for nid in (nids if run % 2 == 0 else reversed(nids)):
                        o = _outcomes(_TESTS[nid][mode_i], run)
                        if o is not None:
                            fd.write("%s\t%s\n" % (o, nid))
        # ---- coverage: one context per test, two or three files each
        base = os.path.join(collate.SUBJECTS_DIR, proj, proj)
        files = ["src/core.py", "src/util/io.py", "tests/test_a.py"]
        con = sqlite3.connect(os.path.join(data_dir, "%s_testinspect_0.sqlite3" % proj))
        con.execute("CREATE TABLE context (id integer primary key, context text, unique (context))")
        con.execute("CREATE TABLE file (id integer primary key, path text, unique (path))")
        con.execute("CREATE TABLE line_bits (file_id integer, context_id integer, numbits blob, "
                    "foreign key (file_id) references file (id), foreign key (context_id) references context (id), "
                    "unique (file_id, context_id))")
        for fi, f in enumerate(files):
            con.execute("INSERT INTO file (id, path) VALUES (?, ?)", (fi + 1, os.path.join(base, f)))
        for ti, nid in enumerate(nids):
            con.execute("INSERT INTO context (id, context) VALUES (?, ?)", (ti + 1, nid))
            for fi in range(len(files)):
                if (ti + fi + pi) % 3 == 2:
                    continue
                lines = [1 + ((ti * 7 + fi * 3 + k * (2 + pi)) % 40) for k in range(3 + (ti + fi) % 4)]
                con.execute("INSERT INTO line_bits (file_id, context_id, numbits) VALUES (?, ?, ?)",
                            (fi + 1, ti + 1, collate.nums_to_numbits(lines)))
        con.commit()
        con.close()
        # ---- resource usage: six numbers + node id (one test has none)
        with open(os.path.join(data_dir, "%s_testinspect_0.tsv" % proj), "w") as fd:
            for ti, nid in enumerate(nids):
                if "no_rusage" in nid:
                    continue
                vals = [0.01 * (ti + 1) + pi, 10 * ti + 1, 3 * ti, 2 + ti, 1 + ti % 3, 1000.5 * (ti + 1)]
                fd.write("\t".join(repr(float(v)) for v in vals) + "\t" + nid + "\n")
        # ---- static analysis (the third project has none and is dropped as a whole)
        if proj != projects[-1]:
            fn_ids = {nid: (0 if "fid_zero" in nid else ti + 1) for ti, nid in enumerate(nids)}
            fn_data = {fid: (3 + fid, fid % 5, 2 * fid, 11.5 * fid, 1 + fid % 4, 5 + fid, 80.25 - fid)
                       for fid in set(fn_ids.values())}
            test_files = {"tests/test_a.py"}
            churn = {"src/core.py": {line: 1 + line % 3 for line in range(1, 41, 2)}, "src/util/io.py": {5: 2, 6: 9}}
            with open(os.path.join(data_dir, "%s_testinspect_0.pkl" % proj), "wb") as fd:
                pickle.dump((fn_ids, fn_data, test_files, churn), fd, protocol=4)
    return data_dir
