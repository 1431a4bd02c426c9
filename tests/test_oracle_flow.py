"""Self-checks of the per-frame flow/pose oracle (oracle/flow_lm.c): it has no reference golden vectors to pin it, so
check behaviour: recovers the true motion, flags the gross outliers, REF_QUIRK and CLEAN agree closely."""
import numpy as np

from oracle import pyoracle as po
from vdo_slam_b200.synth import make_flow_problem


def test_recovers_motion_and_flags_outliers():
    p = make_flow_problem(n=2000, seed=1234)
    for quirk in (0, 1):
        r = po.flow2(p, mode=1, quirk=quirk)
        assert 3 <= r["iters"] <= 200
        assert np.abs(r["T"] - p["T_true"]).max() < 2e-3           # 0.3 px flow noise on 2 000 points
        assert (~r["inlier"][p["outlier"]]).mean() > 0.95           # gross outliers rejected
        assert r["inlier"][~p["outlier"]].mean() > 0.9


def test_quirk_and_clean_modes_agree_to_1e5():
    p = make_flow_problem(n=1500, seed=7)
    a, b = po.flow2(p, 1, 0), po.flow2(p, 1, 1)
    assert np.abs(a["T"] - b["T"]).max() < 1e-5
    assert (a["inlier"] != b["inlier"]).mean() < 0.01


def test_camera_mode_and_small_inputs():
    p = make_flow_problem(n=800, seed=3, rot_deg=0.5, trans=(0.02, 0.0, 0.9))
    r = po.flow2(p, mode=0, quirk=1)
    assert r["iters"] >= 2 and np.isfinite(r["T"]).all()
    q = make_flow_problem(n=2, seed=1)
    assert po.flow2(q, 1, 1)["iters"] == -1
