"""Host shim (VDO_SLAM::System::TrackRGBD over the C ABI): builds against the stub cv::Mat; on the GPU its poses and its
in-place depth / mask mutation are compared with the oracle pipeline."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "shim_stub")
EXE = os.path.join(SHIM, "shim_main")

YAML = """%YAML:1.0
Camera.fx: 721.5377
Camera.fy: 721.5377
Camera.cx: {cx}
Camera.cy: {cy}
Camera.k1: 0.0
Camera.width: {w}
Camera.height: {h}
Camera.fps: 10.0
Camera.bf: 387.5744
Camera.RGB: 1
ChooseData: 2
DepthMapFactor: 256.0
ThDepthBG: 40.0
ThDepthOBJ: 25.0
MaxTrackPointBG: 1200 # 1200 1500 2000
MaxTrackPointOBJ: 800 # 800
SFMgThres: 0.12 # 0.05
SFDsThres: 0.3 # 0.99
WINDOW_SIZE: 20
OVERLAP_SIZE: 4
UseSampleFeature: 0
ORBextractor.nFeatures: 2500
ORBextractor.scaleFactor: 1.2
ORBextractor.nLevels: 8
ORBextractor.iniThFAST: 20
ORBextractor.minThFAST: 7
"""


def _build():
    subprocess.check_call(["make", "-C", SHIM, "shim_main"], stdout=subprocess.DEVNULL)
    assert os.path.exists(EXE)


def test_shim_builds_and_fails_loudly(tmp_path):
    _build()
    r = subprocess.run([EXE, str(tmp_path / "missing.yaml"), str(tmp_path), "1", "640", "240"], capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to open settings file" in r.stderr        # src/System.cc:35-39 behaviour
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        y = tmp_path / "s.yaml"; y.write_text(YAML.format(cx=313.6, cy=110.4, w=640, h=240))
        r = subprocess.run([EXE, str(y), str(tmp_path), "1", "640", "240"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr                  # no device -> loud failure, never a CPU path


@pytest.mark.gpu
def test_shim_matches_oracle_pipeline(tmp_path):
    from oracle.tracking_pipeline import OracleTracker
    from vdo_slam_b200.synth import make_sequence_frame
    _build()
    w, h, n = 640, 240, 4
    K = np.array([721.5377, 721.5377, 313.6, 110.4], np.float32)
    (tmp_path / "s.yaml").write_text(YAML.format(cx=313.6, cy=110.4, w=w, h=h))
    orc = OracleTracker(width=w, height=h, K4=K)
    ref, ref_depth, ref_mask = [], [], []
    for t in range(n):
        f = make_sequence_frame(t, seed=2, width=w, height=h, K=K)
        b = str(tmp_path / f"f{t}")
        np.repeat(f["gray"][..., None], 3, -1).astype(np.uint8).tofile(b + ".rgb")
        f["depth_raw"].astype(np.float32).tofile(b + ".depth"); f["flow"].astype(np.float32).tofile(b + ".flow"); f["mask"].astype(np.int32).tofile(b + ".mask")
        np.array([len(f["obj_ids"])], np.int32).tofile(b + ".ngt"); np.array(f["obj_ids"], np.int32).tofile(b + ".gt")
        ref.append(orc.track(f["gray"], f["depth_raw"], f["flow"], f["mask"], f["obj_ids"]))
        ref_depth.append(orc.depth.copy()); ref_mask.append(orc.mask.copy())
    r = subprocess.run([EXE, str(tmp_path / "s.yaml"), str(tmp_path), str(n), str(w), str(h)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    poses = [np.array(l.split()[2:], np.float32).reshape(4, 4) for l in r.stdout.splitlines() if l.startswith("POSE")]
    assert len(poses) == n
    for t in range(n):
        assert np.abs(poses[t] - ref[t]).max() <= 1e-4
        assert np.array_equal(np.fromfile(str(tmp_path / f"f{t}.depth_out"), np.float32).reshape(h, w), ref_depth[t])
        assert np.array_equal(np.fromfile(str(tmp_path / f"f{t}.mask_out"), np.int32).reshape(h, w), ref_mask[t])
    # SaveResults: the reference's seven files (src/System.cc:74-77, 128, 148, 166), poses as `frame r00 .. r23 0 0 0 1` with 9 decimals
    names = ["obj_mot_stereo_new.txt", "obj_mot_stereo_rf_new.txt", "obj_mot_gt.txt", "obj_centre.txt", "initial_stereo_new.txt", "refined_stereo_new.txt",
             "cam_pose_gt_stereo.txt"]
    for nm in names:
        assert os.path.exists(tmp_path / ("out_" + nm)), nm
    ini = [l.split() for l in open(tmp_path / "out_initial_stereo_new.txt")]
    assert len(ini) == n and all(len(r) == 17 and r[-4:] == ["0.000000000", "0.000000000", "0.000000000", "1.000000000"] for r in ini)
    Twc = np.array(ini[1][1:13], np.float64).reshape(3, 4)
    Tcw1 = ref[1].astype(np.float64)
    assert np.abs(Twc[:, :3] - Tcw1[:3, :3].T).max() < 1e-4                                   # vmCameraPose = toInvMatrix(Tcw)
    gt = [l.split() for l in open(tmp_path / "out_cam_pose_gt_stereo.txt")]
    assert len(gt) == n and abs(float(gt[2][12]) - 1.6) < 1e-6                                 # Twc_gt of frame 2 relative to frame 0: z = 0.8 * 2
    mot = [l.split() for l in open(tmp_path / "out_obj_mot_stereo_new.txt")]
    mot_gt = [l.split() for l in open(tmp_path / "out_obj_mot_gt.txt")]
    assert len(mot) == len(mot_gt) and all(len(r) == 18 for r in mot)


def test_shim_rejects_mismatched_image_sizes(tmp_path):
    # a depth / flow / mask of another size than the image must be refused before anything is copied (heap over-read otherwise)
    _build()
    src = open(os.path.join(SHIM, "shim_main.cc")).read()
    assert "TrackRGBD" in src
    sys_cc = open(os.path.join(ROOT, "vdo_slam_b200", "host", "System.cc")).read()
    for needle in ("depthmap.cols != cols", "flowmap.cols != cols", "masksem.cols != cols", "frame size changed"):
        assert needle in sys_cc
