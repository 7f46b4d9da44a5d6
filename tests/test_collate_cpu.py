"""raw run data -> tests.json (flake16_framework_b200/collate.py).

The first four tests restate the reference's own unit tests for this step with the same inputs and
expected values (/root/reference/test_experiment.py:19-170); the coverage test builds the SQLite
database by hand because coverage.py is not installed.  The last ones compare against the golden
file that the reference's own ``write_tests()`` produced from the same raw directory
(tests/golden/make_golden_collate.py).
"""
import json
import os
import sqlite3
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rawdata
from flake16_framework_b200 import collate
from flake16_framework_b200.collate import FLAKY, NON_FLAKY, OD_FLAKY, N_RUNS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collated_tests.json")


def test_update_collated_runs():
    """test_experiment.py:19-63."""
    cp = [{}, None, None, None]
    collate.update_collated_runs(["passed\ttest1", "passed\ttest2"], "baseline", 0, cp)
    assert cp[0]["test1"][0] == {"baseline": [1, 0, None, 0]}
    assert cp[0]["test2"][0] == {"baseline": [1, 0, None, 0]}
    collate.update_collated_runs(["passed\ttest1", "failed\ttest2"], "shuffle", 0, cp)
    assert cp[0]["test1"][0] == {"baseline": [1, 0, None, 0], "shuffle": [1, 0, None, 0]}
    assert cp[0]["test2"][0] == {"baseline": [1, 0, None, 0], "shuffle": [1, 1, 0, None]}
    collate.update_collated_runs(["failed\ttest1", "passed\ttest2"], "baseline", 1, cp)
    assert cp[0]["test1"][0] == {"baseline": [2, 1, 1, 0], "shuffle": [1, 0, None, 0]}
    assert cp[0]["test2"][0] == {"baseline": [2, 0, None, 0], "shuffle": [1, 1, 0, None]}
    collate.update_collated_runs(["failed\ttest1", "failed\ttest2"], "shuffle", 1, cp)
    assert cp[0]["test1"][0] == {"baseline": [2, 1, 1, 0], "shuffle": [2, 1, 1, 0]}
    assert cp[0]["test2"][0] == {"baseline": [2, 0, None, 0], "shuffle": [2, 2, 0, None]}


def test_update_collated_cov(tmp_path):
    """test_experiment.py:66-97 (the database is written by hand in coverage.py's schema)."""
    proj_dir = os.path.join(collate.SUBJECTS_DIR, "proj", "proj")
    con = sqlite3.connect(tmp_path / "cov.sqlite3")
    con.execute("CREATE TABLE context (id integer primary key, context text)")
    con.execute("CREATE TABLE file (id integer primary key, path text)")
    con.execute("CREATE TABLE line_bits (file_id integer, context_id integer, numbits blob)")
    for fid, name in ((1, "file1"), (2, "file2"), (3, "file3")):
        con.execute("INSERT INTO file VALUES (?, ?)", (fid, os.path.join(proj_dir, name)))
    con.execute("INSERT INTO context VALUES (1, 'test1')")
    con.execute("INSERT INTO context VALUES (2, 'test2')")
    for fid, cid, lines in ((1, 1, {1, 2}), (2, 1, {1, 2}), (2, 2, {2, 3}), (3, 2, {2, 3})):
        con.execute("INSERT INTO line_bits VALUES (?, ?, ?)", (fid, cid, collate.nums_to_numbits(lines)))
    con.commit()
    cp = [{}, None, None, None]
    collate.update_collated_cov(con, "proj", cp)
    assert cp[0]["test1"][1] == {"file1": {1, 2}, "file2": {1, 2}}
    assert cp[0]["test2"][1] == {"file2": {2, 3}, "file3": {2, 3}}


@pytest.mark.parametrize("runs_nid,expected", [                               # test_experiment.py:100-144
    ({"baseline": [N_RUNS["baseline"] - 1, 0, None, 0], "shuffle": [N_RUNS["shuffle"] - 1, 0, None, 0]}, (0, None)),
    ({"baseline": [N_RUNS["baseline"], 0, None, 0], "shuffle": [N_RUNS["shuffle"], 0, None, 0]}, (0, NON_FLAKY)),
    ({"baseline": [N_RUNS["baseline"], 0, None, 0], "shuffle": [N_RUNS["shuffle"], 1, 1, 0]}, (1, OD_FLAKY)),
    ({"baseline": [N_RUNS["baseline"], N_RUNS["baseline"], 0, None],
      "shuffle": [N_RUNS["shuffle"], N_RUNS["shuffle"], 0, None]}, (0, NON_FLAKY)),
    ({"baseline": [N_RUNS["baseline"], N_RUNS["baseline"], 0, None],
      "shuffle": [N_RUNS["shuffle"], N_RUNS["shuffle"] - 1, None, 1]}, (1, OD_FLAKY)),
    ({"baseline": [N_RUNS["baseline"], 1, 1, 0], "shuffle": [N_RUNS["shuffle"], 0, None, 0]}, (1, FLAKY)),
])
def test_get_req_runs_label_nid(runs_nid, expected):
    assert collate.get_req_runs_label_nid(runs_nid) == expected


@pytest.mark.parametrize("cov_nid,test_files,churn,expected", [               # test_experiment.py:147-170
    ({"file1.py": {1, 2, 3}, "file2.py": {1, 2, 3}}, {"file1.py"}, {"file1.py": {1: 1}, "file2.py": {1: 1, 2: 2}}, (6, 4, 3)),
    ({"file1.py": {1, 2, 3}, "file2.py": {1, 2, 3}}, set(), {"file1.py": {1: 1}, "file2.py": {1: 1, 2: 2}}, (6, 4, 6)),
    ({"file1.py": {1, 2, 3}, "file2.py": {1, 2, 3}}, set(), {"file1.py": {1: 10}, "file2.py": {1: 10, 2: 20}}, (6, 40, 6)),
])
def test_get_features_nid_cov(cov_nid, test_files, churn, expected):
    assert collate.get_features_nid_cov(cov_nid, test_files, churn) == expected


def test_numbits_vectors():
    """coverage/numbits.py: bit b of byte i <-> number 8 i + b (its docstring examples)."""
    assert collate.numbits_to_nums(b"") == []
    assert collate.numbits_to_nums(bytes([0b00000110])) == [1, 2]
    assert collate.numbits_to_nums(bytes([0x01, 0x80, 0x00, 0x10])) == [0, 15, 28]
    for nums in ([], [0], [7, 8], [1, 2, 3, 40, 41, 1000]):
        assert collate.numbits_to_nums(collate.nums_to_numbits(nums)) == nums


def test_write_tests_equals_reference_output(tmp_path, monkeypatch):
    """Same raw directory -> byte-identical tests.json as the reference's own write_tests()."""
    monkeypatch.setitem(collate.N_RUNS, "baseline", rawdata.N_RUNS_SMALL)
    monkeypatch.setitem(collate.N_RUNS, "shuffle", rawdata.N_RUNS_SMALL)
    data = rawdata.make_raw_data(str(tmp_path / "data"))
    out = tmp_path / "tests.json"
    tests = collate.write_tests(data, str(out))
    assert out.read_text() == open(GOLDEN).read()
    # labels by every route, the dropped tests and the dropped project
    assert list(tests) == ["proja", "ProjB"]
    rows = tests["proja"]
    assert "tests/test_g.py::test_incomplete" not in rows and "tests/test_h.py::test_no_rusage" not in rows
    assert "tests/test_i.py::test_fid_zero" not in rows
    assert [rows[k][:2] for k in rows] == [(0, 0), (2, 1), (0, 0), (1, 1), (1, 2), (3, 2)]


def test_cli_tests_command_feeds_the_hot_path(tmp_path, monkeypatch):
    """`python experiment.py tests` writes ./tests.json from ./data, and the result parses into the
    16-feature table the `scores` path reads (experiment.py:410-427)."""
    import importlib.util
    from flake16_framework_b200 import hostprep as hp
    monkeypatch.setitem(collate.N_RUNS, "baseline", rawdata.N_RUNS_SMALL)
    monkeypatch.setitem(collate.N_RUNS, "shuffle", rawdata.N_RUNS_SMALL)
    rawdata.make_raw_data(str(tmp_path / "data"))
    monkeypatch.chdir(tmp_path)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("f16_experiment_cli", os.path.join(root, "experiment.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    cli.main(["tests"])
    assert json.load(open(tmp_path / "tests.json")) == json.load(open(GOLDEN))
    feats, labels, projects = hp.parse_tests(str(tmp_path / "tests.json"))
    assert feats.shape == (12, 16) and sorted(set(labels.tolist())) == [0, 1, 2] and set(projects.tolist()) == {"proja", "ProjB"}
