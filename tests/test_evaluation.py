"""Row N3: the evaluation metrics of evaluation.py.  The golden dictionary
(tests/golden/reference_eval.json) is the output of the reference's OWN, unmodified `evaluation.evaluate()` run on
seeded predictions (generator: tests/golden/make_reference_eval_golden.py); the CPU restatement and the device
kernel are both held to it."""
import json
import os

import numpy as np
import pytest

from oracle import evaluation_ref as E
from helpers import GOLDEN


def _load():
    g = np.load(os.path.join(GOLDEN, "reference_eval.npz"))
    ref = json.load(open(os.path.join(GOLDEN, "reference_eval.json")))
    return g, ref


def _max_diff(a, b, path=""):
    if isinstance(a, dict):
        assert set(a) == set(b), (path, set(a) ^ set(b))
        return max(_max_diff(a[k], b[k], path + "/" + k) for k in a)
    return float(np.abs(np.asarray(a, float) - np.asarray(b, float)).max())


@pytest.mark.parametrize("inverted", [False, True])
def test_oracle_matches_reference_evaluate(inverted):
    g, ref = _load()
    d = E.evaluate(g["pred_t"], g["pred_a"], g["pred_c"], g["gt_t"], g["gt_a"], g["gt_c1"], g["is_test"], inverted, 0.0125)
    key = "inverted" if inverted else "plain"
    assert _max_diff(d, ref[key]) < 1e-12
    assert ref[key]["num"] == 1499 and ref[key]["test"]["num"] == 500          # one transform beyond 10 km is skipped
    assert 0 < ref[key]["corr_levels"][0] < ref[key]["corr_levels"][1] < ref[key]["corr_levels"][2] < 1


def test_edge_cases_of_the_restatement():
    z3, z1 = np.zeros((0, 3)), np.zeros((0, 1))
    d = E.evaluate(z3, z1, z3, z3, z1, z3, np.zeros(0, bool))
    assert d["num"] == 0 and d["val"]["eval_5m"]["num"] == 0
    # the flipped prediction is accepted only with accept_inverted_angle
    one = np.zeros((1, 3))
    a = np.array([[np.pi - 0.01]])
    assert E.evaluate(one, a, one, one, np.zeros((1, 1)), one, [False], False)["corr_levels_angles"] == [0.0, 0.0, 0.0]
    assert E.evaluate(one, a, one, one, np.zeros((1, 1)), one, [False], True)["corr_levels_angles"] == [1.0, 1.0, 1.0]


@pytest.mark.gpu
@pytest.mark.parametrize("inverted", [False, True])
def test_device_evaluation_matches_reference_evaluate(inverted):
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import evaluation as DE
    g, ref = _load()
    d = DE.evaluate(g["pred_t"], g["pred_a"], g["gt_t"], g["gt_a"], g["pred_c"], g["gt_c1"], g["is_test"], inverted, 0.0125)
    key = "inverted" if inverted else "plain"
    assert _max_diff(d, ref[key]) < 1e-9
    acc = DE.accumulate(g["pred_t"], g["pred_a"], g["gt_t"], g["gt_a"], g["pred_c"], g["gt_c1"], g["is_test"], inverted)
    np.testing.assert_allclose(acc, E.accumulate(g["pred_t"], g["pred_a"], g["pred_c"], g["gt_t"], g["gt_a"], g["gt_c1"],
                                                 g["is_test"], inverted), rtol=1e-12, atol=1e-9)


@pytest.mark.gpu
def test_device_evaluation_edge_cases():
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import evaluation as DE
    z3, z1 = np.zeros((0, 3)), np.zeros((0,))
    assert DE.evaluate(z3, z1, z3, z1, z3, z3)["num"] == 0
    with pytest.raises(ValueError):
        DE.evaluate(np.zeros((2, 3)), np.zeros(3), np.zeros((2, 3)), np.zeros(2), np.zeros((2, 3)), np.zeros((2, 3)))
    # large n, no split flags: everything lands in 'val'
    rng = np.random.default_rng(0)
    n = 200_000
    t, c = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)) * 8
    a = rng.uniform(-3, 3, n)
    d = DE.evaluate(t, a, t, a, c, c)
    assert d["num"] == n and d["val"]["num"] == n and d["test"]["num"] == 0
    assert d["corr_levels"] == [1.0, 1.0, 1.0] and d["mean_dist_translation"] < 1e-12
