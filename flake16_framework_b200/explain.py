"""The two analysis steps that follow ``scores`` in the reference and re-use its device state
(SURVEY.md 8(f) rows N3, N4):

* ``get_shap`` / ``write_shap`` (experiment.py:504-530): fit one forest on the whole (resampled)
  dataset and explain every row with path-dependent TreeSHAP - here ``f16_forest_shap`` over the node
  records the tree builders leave on the device;
* the Spearman table of ``write_figures`` (experiment.py:661-663) - ``f16_spearman``.

Host code mirrors the reference's functions; everything numeric runs on the GPU through the C ABI.
"""

import pickle

import numpy as np
import torch

from . import estimators as E
from . import hostprep as hp
from . import ops

SHAP_FILE = "shap.pkl"                                                    # experiment.py:36
SHAP_CONFIGS = (("NOD", "Flake16", "Scaling", "SMOTE Tomek", "Extra Trees"),          # experiment.py:523-526
                ("OD", "Flake16", "Scaling", "SMOTE", "Random Forest"))

_BALANCING = {"None": lambda: None, "Tomek Links": E.TomekLinks, "SMOTE": lambda: E.SMOTE(random_state=0),
              "ENN": E.EditedNearestNeighbours, "SMOTE ENN": lambda: E.SMOTEENN(random_state=0),
              "SMOTE Tomek": lambda: E.SMOTETomek(random_state=0)}
_MODEL = {"Extra Trees": lambda n: E.ExtraTreesClassifier(random_state=0, n_estimators=n),
          "Random Forest": lambda n: E.RandomForestClassifier(random_state=0, n_estimators=n),
          "Decision Tree": lambda n: E.DecisionTreeClassifier(random_state=0)}


def get_shap(config_keys, parsed, n_estimators=100, return_device=False):
    """experiment.py:504-517.  (The reference's no-balancing branch references an undefined name,
    ``feature``, :515; both configurations it actually runs have a balancer.  Here that branch fits
    on the unbalanced data, which is what the line evidently means.)"""
    ft, fs, pre, bal, model = config_keys
    features, labels, _ = hp.feat_lab_proj(parsed, hp.FLAKY_TYPES[ft], hp.FEATURE_SETS[fs])
    features = np.ascontiguousarray(hp.preprocess(features, pre))
    Xd = torch.from_numpy(features).cuda()
    yd = torch.from_numpy(labels.astype(np.uint8)).cuda()
    balancing = _BALANCING[bal]()
    m = _MODEL[model](n_estimators)
    if balancing is not None:
        Xb, yb = balancing.fit_resample(Xd, yd)
        m.fit(Xb.contiguous(), yb.contiguous())
    else:
        m.fit(Xd, yd)
    phi = m.forest_.shap_values(ops.rows_f32(Xd), 0)       # TreeExplainer(model).shap_values(features)[0]
    m.forest_.status()
    return phi if return_device else phi.cpu().numpy()


def write_shap(tests_file="tests.json", shap_file=SHAP_FILE, n_estimators=100):
    """experiment.py:520-530: the two explained configurations, pickled as a list."""
    parsed = hp.parse_tests(tests_file)
    shap = [get_shap(c, parsed, n_estimators) for c in SHAP_CONFIGS]
    with open(shap_file, "wb") as fd:
        pickle.dump(shap, fd)
    return shap


def spearman_table(parsed):
    """experiment.py:661: ``stats.spearmanr(features).correlation`` over all 16 raw features."""
    all_features = np.ascontiguousarray(np.asarray(parsed[0], dtype=np.float64))
    return ops.spearman(torch.from_numpy(all_features).cuda()).cpu().numpy()
