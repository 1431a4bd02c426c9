"""the reference's .g2o text format (reader / writer / loader).  CPU-only: host code exercised through the
emulation library; the golden file is written by hand in the exact syntax of the reference's serialisers
(g2o/types/*.cpp read()/write(), optimizable_graph.cpp:817-860) with ostream's default 6 significant digits."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_batch_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul", "libvdo_emul.so")
GOLD = os.path.join(ROOT, "tests", "golden", "tiny_reference_syntax.g2o")


@pytest.fixture(scope="module")
def ectx():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "libvdo_emul.so"], stdout=subprocess.DEVNULL)
    return capi.Context(0, lib_path=EMUL)


def test_golden_file_in_reference_syntax(ectx):
    g = capi.g2o_read(ectx, GOLD)
    assert g["se3_id"].tolist() == [0, 1, 5] and g["pt_id"].tolist() == [2, 3, 4] and g["fixed_id"].tolist() == [0]
    # toVectorQT order is [t, qx qy qz qw]: vertex 1 is a 2 degree rotation about y
    R = g["se3"][1, :9].reshape(3, 3)
    c, s = np.cos(np.deg2rad(2.0)), np.sin(np.deg2rad(2.0))
    np.testing.assert_allclose(R, [[c, 0, s], [0, 1, 0], [-s, 0, c]], atol=2e-6)
    np.testing.assert_allclose(g["se3"][1, 9:], [0.5, 0.01, -0.02])
    assert g["prior_v"].tolist() == [0] and g["se3e_ij"].tolist() == [[0, 1]]
    assert g["obs_cp"].tolist() == [[0, 0], [1, 0], [0, 1], [1, 2]]          # compact indices: points 2,3,4 -> 0,1,2
    assert g["ter_pph"].tolist() == [[1, 2, 2]]                               # (p1=3, p2=4, H=5) -> (1, 2, 2)
    assert g["prior_w"].tolist() == [10000.0] and g["se3e_w"].tolist() == [100.0] and set(g["obs_w"]) == {16.0} and g["ter_w"].tolist() == [20.0]
    up = np.zeros((6, 6)); up[np.triu_indices(6)] = g["se3e_info"][0]
    np.testing.assert_array_equal(up, 100.0 * np.eye(6))
    # measurement quaternions are re-normalised on read (edge_se3.cpp:48)
    assert abs(np.linalg.det(g["se3e_Z"][0, :9].reshape(3, 3)) - 1.0) < 1e-12


def test_round_trip_is_lossless_and_solves_identically(ectx, tmp_path):
    g = make_batch_graph(n_frames=8, n_objects=1, n_static=60, n_dynamic=20, seed=3)
    path = str(tmp_path / "g.g2o")
    capi.g2o_write(ectx, path, g)
    h = capi.g2o_read(ectx, path)
    for k in ("pt", "prior_w", "se3e_w", "obs_z", "obs_w", "ter_w"):
        np.testing.assert_array_equal(h[k], g[k])
    for k in ("prior_v", "se3e_ij", "obs_cp", "ter_pph"):
        np.testing.assert_array_equal(h[k], g[k])
    np.testing.assert_allclose(h["se3"], g["se3"], atol=1e-15)              # rotation -> quaternion -> rotation
    np.testing.assert_allclose(h["se3e_Z"], g["se3e_Z"], atol=1e-15)
    # loader == arrays: same solve
    d_se3, d_obs, d_ter = float(g["se3e_delta"][0]), float(g["obs_delta"][0]), float(g["ter_delta"][0])
    Ga = capi.BatchGraph(ectx, g)
    Gb = capi.BatchGraph.from_g2o(ectx, path, d_se3, d_obs, d_ter)
    ra, rb = Ga.optimize(max_iterations=6, gain_threshold=0), Gb.optimize(max_iterations=6, gain_threshold=0)
    assert ra["iterations"] == rb["iterations"]
    np.testing.assert_allclose(Gb.vertices()[0], Ga.vertices()[0], atol=1e-10)
    np.testing.assert_allclose(Gb.vertices()[1], Ga.vertices()[1], atol=1e-10)
    # and the oracle on the parsed arrays agrees with the product
    h2 = dict(h); h2["se3e_delta"] = np.full(len(h["se3e_w"]), d_se3); h2["obs_delta"] = np.full(len(h["obs_w"]), d_obs); h2["ter_delta"] = np.full(len(h["ter_w"]), d_ter)
    ro = po.ba_optimize(h2, max_iters=6, gain_threshold=0)
    assert np.abs(Gb.vertices()[0] - ro["se3"]).max() < 1e-6


def test_reference_precision_and_custom_ids(ectx, tmp_path):
    g = make_batch_graph(n_frames=4, n_objects=0, n_static=10, n_dynamic=0, seed=1)
    g = dict(g); g["se3_id"] = np.arange(len(g["se3"]), dtype=np.int32) * 10 + 7; g["pt_id"] = 1000 + np.arange(len(g["pt"]), dtype=np.int32)
    path = str(tmp_path / "p6.g2o")
    capi.g2o_write(ectx, path, g, precision=6)                                # what the reference's own dumps look like
    first = open(path).read().splitlines()
    assert first[0].startswith("PARAMS_SE3OFFSET 0 ") and first[1].startswith("VERTEX_SE3:QUAT 7 ")
    h = capi.g2o_read(ectx, path)
    assert h["se3_id"].tolist() == g["se3_id"].tolist() and h["pt_id"].tolist() == g["pt_id"].tolist()
    np.testing.assert_allclose(h["pt"], g["pt"], rtol=1e-5)


def test_rejects_what_the_solver_does_not_cover(ectx, tmp_path):
    bad = open(GOLD).read().replace("EDGE_SE3_TRACKXYZ 0 2 0 1.5 -0.25 8 16 0 0 16 0 16", "EDGE_SE3_TRACKXYZ 0 2 0 1.5 -0.25 8 16 1 0 16 0 16")
    p = tmp_path / "full_info.g2o"; p.write_text(bad)
    capi.g2o_read(ectx, str(p))                                                # parses
    with pytest.raises(capi.VdoError):
        capi.BatchGraph.from_g2o(ectx, str(p), 0.0, 0.0, 0.0)                  # but a non-scalar information matrix is not silently approximated
    p2 = tmp_path / "unknown.g2o"; p2.write_text("VERTEX_SE2 0 0 0 0\n")
    with pytest.raises(capi.VdoError):
        capi.g2o_read(ectx, str(p2))
    p3 = tmp_path / "dangling.g2o"; p3.write_text("VERTEX_SE3:QUAT 0 0 0 0 0 0 0 1\nEDGE_SE3_TRACKXYZ 0 9 0 1 1 1 1 0 0 1 0 1\n")
    with pytest.raises(capi.VdoError):
        capi.g2o_read(ectx, str(p3))
