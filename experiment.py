#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Drop-in for the reference's ``experiment.py`` CLI, restricted to its one compute-heavy
stage (reference experiment.py:693-714 dispatches setup|container|run|tests|scores|shap|figures).

    python experiment.py scores        # reads ./tests.json, writes ./scores.pkl
    python experiment.py tests         # reads ./data/, writes ./tests.json (the step before; host only)
    python experiment.py shap          # reads ./tests.json, writes ./shap.pkl (the step after: TreeSHAP on the GPU)

``scores`` keeps the reference's file names (experiment.py:34-35) and pickle schema
(experiment.py:488-490,498-501) but runs on B200 GPUs through libf16_b200.so.  Under
``torchrun`` (WORLD_SIZE > 1) the (dataset, fold) units are sharded over one process per GPU
and the integer counts are all-reduced over NCCL; rank 0 writes the pickle.

``tests`` is the collation step that produces the hot path's input; it is
host-side bookkeeping with the reference's function names (flake16_framework_b200/collate.py).
``shap`` (row N3) explains the reference's two configurations (experiment.py:520-530) with
path-dependent TreeSHAP over the forests fitted on the device.
The other reference commands (data collection in Docker, LaTeX figures) are
outside this repo's scope and raise the same ``ValueError`` the
reference raises for an unrecognised command (experiment.py:712-714).

Extra, non-reference commands used by the tests / bench:
    python experiment.py synth N [SEED]     # writes a synthetic tests.json
"""

import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")    # before any CUDA call (see flake16_framework_b200/__init__.py)

TESTS_FILE = "tests.json"       # experiment.py:34
SCORES_FILE = "scores.pkl"      # experiment.py:35


def write_scores():
    from flake16_framework_b200 import scores as S

    def progress(done, total):
        sys.stdout.write(f"{done}/{total - done}\r")
        sys.stdout.flush()

    n_streams = int(os.environ.get("F16_STREAMS", "4"))
    _, wall = S.write_scores(TESTS_FILE, SCORES_FILE, n_streams=n_streams, progress=progress)
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stdout.write(f"\n216 configs in {wall:.1f}s\n")


def write_tests():
    """`tests`: raw run data under ./data -> ./tests.json (reference experiment.py:376-407)."""
    from flake16_framework_b200 import collate
    collate.write_tests(collate.DATA_DIR, TESTS_FILE)


def main(argv):
    if not argv:
        raise ValueError("No command given.")
    command, *args = argv
    if command == "scores" and not args:
        write_scores()
    elif command == "tests" and not args:
        write_tests()
    elif command == "shap" and not args:
        from flake16_framework_b200 import explain
        explain.write_shap(TESTS_FILE, explain.SHAP_FILE)
    elif command == "synth" and args:
        from flake16_framework_b200 import synth
        synth.make_tests_json(TESTS_FILE, int(args[0]), int(args[1]) if len(args) > 1 else 16)
    else:
        raise ValueError("Unrecognized command given.")


if __name__ == "__main__":
    main(sys.argv[1:])
