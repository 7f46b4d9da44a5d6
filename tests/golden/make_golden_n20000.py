"""A second, larger golden pickle from THE REFERENCE's own write_scores() (see make_golden.py for how
the reference file is imported unmodified): the full 216-config grid on a 20 000-test synthetic
tests.json.  At this size the SMOTE'd training sets hold ~35 000 rows, so the k-NN calls of the
Tomek / ENN configs on StandardScaler / PCA data cross the size above which the product takes its
tensor-core candidate filter (n x nq >= 2.5e8) - the golden grid at 1 500 tests never does.

    OMP_NUM_THREADS=1 python tests/golden/make_golden_n20000.py      # ~1 h on 8 cores

Output (committed): scores_n20000_seed16.pkl   {config_keys: (per-project [fp,fn,tp], total [fp,fn,tp])}
"""
import os
import pickle
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, ROOT      # noqa: E402


def main():
    from flake16_framework_b200 import synth
    experiment = import_reference()
    work = tempfile.mkdtemp(prefix="f16golden20k")
    os.chdir(work)
    synth.make_tests_json("tests.json", 20000, 16)
    experiment.write_scores()
    with open("scores.pkl", "rb") as fd:
        scores = pickle.load(fd)
    gold = {}
    for keys, (t_train, t_test, per_proj, total) in scores.items():
        gold[tuple(keys)] = ({str(p): [int(v) for v in s[:3]] for p, s in per_proj.items()},
                             [int(v) for v in total[:3]])
    assert len(gold) == 216
    with open(os.path.join(HERE, "scores_n20000_seed16.pkl"), "wb") as fd:
        pickle.dump(gold, fd, protocol=4)
    print("wrote scores_n20000_seed16.pkl")


if __name__ == "__main__":
    main()
