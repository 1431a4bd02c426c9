"""Whole per-frame path (config 3 shape): vdo_tracker on the GPU against the Python/C oracle pipeline on a synthetic,
geometry-consistent RGB-D sequence.  Track IDs / labels / index sets exact; poses and motions <= 1e-4."""
import numpy as np
import pytest

from oracle.tracking_pipeline import OracleTracker
from vdo_slam_b200 import capi
from vdo_slam_b200.synth import make_sequence_frame


def _compare(tr, orc, t):
    C = orc.cur
    assert np.abs(tr.get("Tcw").reshape(4, 4) - C.Tcw).max() <= 1e-4, f"frame {t}: camera pose"
    for name, ref in (("nModLabel", C.nModLabel), ("nSemPosition", C.nSemPosition), ("bObjStat", [int(b) for b in C.bObjStat])):
        assert tr.get(name).tolist() == [int(v) for v in ref], f"frame {t}: {name}"
    assert np.array_equal(tr.get("vObjLabel"), C.objLabel), f"frame {t}: vObjLabel"
    assert np.array_equal(tr.get("vSemObjLabel"), C.semObjLabel), f"frame {t}: vSemObjLabel"
    assert int(tr.get("max_id")[0]) == orc.max_id
    if t > 0:
        assert np.array_equal(tr.get("nStaInlierID"), C.staInlierID), f"frame {t}: nStaInlierID"
        assert np.array_equal(tr.get("nDynInlierID"), C.dynInlierID), f"frame {t}: nDynInlierID"
        assert np.array_equal(tr.get("TemperalMatch_subset"), orc.tm_sub), f"frame {t}: TemperalMatch_subset"
        mods = tr.get("vObjMod").reshape(-1, 4, 4)
        assert len(mods) == len(C.vObjMod)
        for a, b in zip(mods, C.vObjMod):
            assert np.abs(a - b).max() <= 1e-4, f"frame {t}: object motion"
    for name, ref in (("mvStatKeysTmp", C.statKeysTmp), ("mvCorres", C.corres), ("mvFlowNext", C.flowNext), ("mvObjKeys", C.objKeys), ("mvObjCorres", C.objCorres)):
        got = tr.get(name).reshape(-1, 2)
        assert got.shape == np.asarray(ref).reshape(-1, 2).shape, f"frame {t}: {name} count"
        assert np.abs(got - np.asarray(ref).reshape(-1, 2)).max() <= 1e-3, f"frame {t}: {name}"
    assert np.abs(tr.get("mvObjDepth") - C.objDepth).max() <= 1e-5 if len(C.objDepth) else True


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_frames,size", [(0, 6, (1242, 375)), (1, 5, (640, 240))])
def test_tracker_matches_oracle_pipeline(seed, n_frames, size):
    w, h = size
    K = np.array([721.5377, 721.5377, w * 0.49, h * 0.46], np.float32)
    ctx = capi.Context()
    tr = capi.Tracker(ctx, width=w, height=h, cx=float(K[2]), cy=float(K[3]))
    orc = OracleTracker(width=w, height=h, K4=K)
    for t in range(n_frames):
        f = make_sequence_frame(t, seed=seed, width=w, height=h, K=K)
        ids = f["obj_ids"] if t != 3 else f["obj_ids"][:1]          # frame 3: one object has no ground truth -> its motion is not estimated
        T_ref = orc.track(f["gray"], f["depth_raw"], f["flow"], f["mask"], ids)
        d, m = f["depth_raw"].copy(), f["mask"].copy()
        T = tr.track(f["gray"], d, f["flow"], m, ids, writeback=True)
        assert np.abs(T - T_ref).max() <= 1e-4
        assert np.array_equal(m, orc.mask), f"frame {t}: propagated mask"
        assert np.array_equal(d, orc.depth), f"frame {t}: prepared depth"
        _compare(tr, orc, t)
    # the estimated camera motion is close to the ground truth of the generator (sanity of the whole chain, not a parity claim)
    T_gt = np.linalg.inv(f["Twc"]) @ make_sequence_frame(0, seed=seed, width=w, height=h, K=K)["Twc"]
    assert np.abs(T[:3, 3] - T_gt[:3, 3]).max() < 0.1


@pytest.mark.gpu
def test_tracker_degenerate_frames():
    """empty masks (no objects), zero flow (no static candidates): the tracker must not fail and must keep identity / last pose."""
    w, h = 400, 240          # the smallest pyramid level must still hold one 30-px FAST cell (the reference divides by zero below that)
    ctx = capi.Context()
    tr = capi.Tracker(ctx, width=w, height=h, cx=200.0, cy=110.0)
    orc = OracleTracker(width=w, height=h, K4=np.array([721.5377, 721.5377, 200.0, 110.0], np.float32))
    rng = np.random.default_rng(0)
    for t in range(3):
        gray = rng.integers(0, 255, (h, w)).astype(np.uint8)
        depth = np.full((h, w), 9000.0, np.float32); flow = np.zeros((h, w, 2), np.float32); mask = np.zeros((h, w), np.int32)
        T_ref = orc.track(gray, depth, flow, mask, [])
        T = tr.track(gray, depth.copy(), flow, mask.copy(), [])
        assert np.abs(T - T_ref).max() <= 1e-6
        assert len(tr.get("mvStatKeysTmp")) == 2 * len(orc.cur.statKeysTmp)


@pytest.mark.gpu
def test_map_graph_builder_and_batch_optimisation():
    """Map -> graph builder (PartialBatchOptimization / FullBatchOptimization construction) against the Python restatement on the
    oracle pipeline's map: identical arrays; then the optimised camera poses against the CPU oracle LM."""
    from oracle import map_graph as mg
    from oracle import pyoracle as po
    w, h, n = 1242, 375, 9
    ctx = capi.Context()
    tr = capi.Tracker(ctx, window_size=8, overlap_size=4)
    orc = OracleTracker(width=w, height=h, window_size=8, overlap_size=4)      # both run the windowed pass at f_id = 7
    for t in range(n):
        f = make_sequence_frame(t, seed=3)
        orc.track(f["gray"], f["depth_raw"], f["flow"], f["mask"], f["obj_ids"])
        tr.track(f["gray"], f["depth_raw"].copy(), f["flow"], f["mask"].copy(), f["obj_ids"])
    metas = {}
    for mode, name in ((1, "full"), (0, "partial")):
        g_ref, metas[name] = mg.build_graph(orc.map, orc.K4, name, window=8)
        g = tr.graph_export(mode)
        for k in g_ref:
            assert g[k].shape == g_ref[k].shape, (name, k, g[k].shape, g_ref[k].shape)
            if g_ref[k].dtype == np.int32:
                assert np.array_equal(g[k], g_ref[k]), (name, k)
            elif g_ref[k].size:                                        # both maps hold the same f32 values; the quaternion round trip differs in the last bits
                assert np.abs(g[k] - g_ref[k]).max() <= 1e-12, (name, k)
    assert len(g_ref["prior_v"]) == 0 and len(g_ref["ter_pph"]) == 0   # partial: static only, no prior unless N == WINDOW (src/Optimizer.cc:240)

    def check(mode, name):
        g_in = tr.graph_export(mode)                                   # the builder's arrays at the map's current state
        meta = metas[name]
        r_ref = po.ba_optimize(g_in, max_iters=100 if mode == 0 else 300, gain_threshold=1e-3 if mode == 0 else 1e-4)
        which = "vmCameraPose" if mode == 0 else "vmCameraPose_RF"       # src/Optimizer.cc:1058-1101 vs :2094-2133
        before = tr.map_get(which).reshape(-1, 4, 4).copy()
        r = tr.batch_optimize(mode)
        assert r["iterations"] == r_ref["iters"], (name, r["iterations"], r_ref["iters"])
        assert r["sizes"]["n_se3"] == len(g_in["se3"]) and r["sizes"]["n_obs"] == len(g_in["obs_w"])
        poses = tr.map_get(which).reshape(-1, 4, 4)
        for i, v in enumerate(meta["cam_vid"]):
            if v == -1 or (mode == 1 and i == 0):                    # the full batch leaves vmCameraPose_RF[0] alone
                assert np.array_equal(poses[i], before[i]), (name, i)
                continue
            iso = r_ref["se3"][v]
            assert np.abs(poses[i][:3, :3] - iso[:9].reshape(3, 3)).max() <= 1e-4 and np.abs(poses[i][:3, 3] - iso[9:]).max() <= 1e-4, (name, i)
        return r

    r0 = check(0, "partial")          # Tracking.cc:1150-1160: the windowed pass runs first, on the live map
    r1 = check(1, "full")             # :1162-1176: the full batch runs on the map the windowed pass refined
    assert r1["sizes"]["n_ternary"] > 0 and r1["final_chi2"] <= r1["initial_chi2"]


@pytest.mark.gpu
def test_windowed_and_full_batch_inside_the_pipeline():
    """bLocalBatch schedule inside vdo_tracker_track (Tracking.cc:1150-1160) and the final FullBatchOptimization: same number of
    runs / LM iterations as the oracle pipeline, refined map poses <= 1e-4."""
    ctx = capi.Context()
    tr = capi.Tracker(ctx, window_size=6, overlap_size=2)
    orc = OracleTracker(window_size=6, overlap_size=2)
    for t in range(11):
        f = make_sequence_frame(t, seed=0)
        orc.track(f["gray"], f["depth_raw"], f["flow"], f["mask"], f["obj_ids"])
        tr.track(f["gray"], f["depth_raw"].copy(), f["flow"], f["mask"].copy(), f["obj_ids"])
    runs, iters = tr.get("local_ba").tolist()
    assert runs == len(orc.local_ba) == 2 and iters == sum(orc.local_ba)
    P = tr.map_get("vmCameraPose").reshape(-1, 4, 4)
    assert np.abs(P - np.array(orc.map["cameraPose"])).max() <= 1e-4
    ini = tr.map_get("vmCameraPose").copy()
    r_ref = orc.batch_optimize("full")
    r = tr.batch_optimize(1)
    assert r["iterations"] == r_ref["iters"]
    assert np.array_equal(tr.map_get("vmCameraPose"), ini)            # the full batch refines the _RF copies only (src/Optimizer.cc:2094-2133)
    P = tr.map_get("vmCameraPose_RF").reshape(-1, 4, 4)
    assert np.abs(P - np.array(orc.map["cameraPose_RF"])).max() <= 1e-4 and np.abs(P - ini.reshape(-1, 4, 4)).max() > 0
    M = tr.map_get("vmRigidMotion_RF").reshape(-1, 4, 4)
    M_ref = np.array([T for fr in orc.map["rigidMotion_RF"] for T in fr])
    assert M.shape == M_ref.shape and np.abs(M - M_ref).max() <= 1e-4
    M0 = tr.map_get("vmRigidMotion").reshape(-1, 4, 4)
    assert np.abs(M0 - np.array([T for fr in orc.map["rigidMotion"] for T in fr])).max() <= 1e-4
    assert tr.map_get("vnRMLabel").tolist() == [int(l) for fr in orc.map["rmLabel"] for l in fr]
