"""Golden fixture for SURVEY.md 8(f) row N2: the reference's OWN ``write_tests()`` run here.

/root/reference/experiment.py is imported unmodified (coverage / shap / imblearn stubbed as in
make_golden.py); its module constants N_RUNS, DATA_DIR, SUBJECTS_DIR are set to the fixture's
(tests/rawdata.py builds the raw ``data/`` directory) and ``write_tests()`` writes tests.json,
committed as ``collated_tests.json``.  The one piece the reference delegates to coverage.py,
``numbits_to_nums``, is not installed here: the stub uses flake16_framework_b200.collate's decoder
(the bit layout documented in coverage/numbits.py), so that function alone is pinned only by the
hand-written vectors in tests/test_collate_cpu.py.

    python tests/golden/make_golden_collate.py
"""
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from make_golden import import_reference      # noqa: E402  (same stubbing of the missing packages)
import rawdata                                # noqa: E402
from flake16_framework_b200 import collate    # noqa: E402


def main():
    experiment = import_reference()
    experiment.numbits = types.SimpleNamespace(numbits_to_nums=collate.numbits_to_nums)
    experiment.N_RUNS["baseline"] = experiment.N_RUNS["shuffle"] = rawdata.N_RUNS_SMALL
    experiment.SUBJECTS_DIR = collate.SUBJECTS_DIR
    work = tempfile.mkdtemp(prefix="f16collate")
    os.chdir(work)
    rawdata.make_raw_data(os.path.join(work, experiment.DATA_DIR))
    experiment.write_tests()                  # the reference's code, unmodified
    shutil.copy(os.path.join(work, experiment.TESTS_FILE), os.path.join(HERE, "collated_tests.json"))
    print("wrote", os.path.join(HERE, "collated_tests.json"))


if __name__ == "__main__":
    main()
