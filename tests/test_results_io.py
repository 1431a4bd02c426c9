"""the reference's result files (System::SaveResults, src/System.cc:66-244) and error metrics
(Tracking::GetMetricError, src/Tracking.cc:3243-3386).  The expected values are a statement-by-statement numpy / cv2
restatement: cv::Mat products through cv2.gemm (OpenCV's own float GEMM), iostream `fixed << setprecision(9)` through
'%.9f'.  CPU-only (host code, emulation library)."""
import math
import os
import subprocess

import cv2
import numpy as np
import pytest

from vdo_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul", "libvdo_emul.so")
f32 = np.float32


@pytest.fixture(scope="module")
def ectx():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "libvdo_emul.so"], stdout=subprocess.DEVNULL)
    return capi.Context(0, lib_path=EMUL)


def cvmul(A, B):
    return cv2.gemm(np.ascontiguousarray(A, f32), np.ascontiguousarray(B, f32), 1.0, None, 0.0)


def to_inv(T):                                    # Converter::toInvMatrix (src/Converter.cc:151-166)
    R, t = np.ascontiguousarray(T[:3, :3]), np.ascontiguousarray(T[:3, 3:4])
    out = np.eye(4, dtype=f32)
    out[:3, :3] = R.T
    out[:3, 3:4] = cv2.gemm(R, t, -1.0, None, 0.0, flags=cv2.GEMM_1_T)
    return out


def rand_T(rng, ang=0.2, tr=1.0):
    rv = rng.normal(0, ang, 3)
    R = cv2.Rodrigues(rv)[0]
    T = np.eye(4, dtype=f32); T[:3, :3] = R.astype(f32); T[:3, 3] = rng.normal(0, tr, 3).astype(f32)
    return T


def fmt12(T):
    return " ".join("%.9f" % float(v) for v in T[:3].reshape(-1)) + " 0.000000000 0.000000000 0.000000000 1.000000000"


def make_run(seed, n_frames=7, max_id=4):
    rng = np.random.default_rng(seed)
    cam = [np.eye(4, dtype=f32)]; cam_gt = [np.eye(4, dtype=f32)]
    for _ in range(n_frames):
        cam.append(cvmul(cam[-1], rand_T(rng, 0.02, 0.5))); cam_gt.append(cvmul(cam_gt[-1], rand_T(rng, 0.02, 0.5)))
    mot, pre, gt, lab, stat, cen = [], [], [], [], [], []
    for i in range(n_frames):
        k = int(rng.integers(0, max_id))          # objects in this frame (entry 0 = camera)
        ids = sorted(rng.choice(np.arange(1, max_id), size=min(k, max_id - 1), replace=False).tolist())
        mot.append([rand_T(rng) for _ in range(1 + len(ids))]); pre.append([rand_T(rng, 0.5, 5.0) for _ in range(1 + len(ids))])
        gt.append([rand_T(rng) for _ in range(1 + len(ids))]); lab.append([0] + ids)
        stat.append([1] + [int(rng.random() > 0.2) for _ in ids]); cen.append(rng.normal(0, 5, (1 + len(ids), 3)).astype(f32))
    return cam, cam_gt, mot, pre, gt, lab, stat, cen


def test_files_have_the_reference_text_format(ectx, tmp_path):
    cam, cam_gt, mot, pre, gt, lab, stat, cen = make_run(1)
    p = str(tmp_path / "initial_stereo_new.txt")
    capi.results_write_poses(ectx, p, cam, start_frame=0)
    assert open(p).read().splitlines() == [f"{i} " + fmt12(T) for i, T in enumerate(cam)]
    p = str(tmp_path / "obj_mot_stereo_new.txt")
    capi.results_write_object_motions(ectx, p, mot, lab, pose_pre=pre)
    exp = []
    for i in range(len(mot)):
        for j in range(1, len(mot[i])):
            body = cvmul(cvmul(to_inv(pre[i][j]), mot[i][j]), pre[i][j])      # System.cc:92
            exp.append(f"{i + 1} {lab[i][j]} " + fmt12(body))
    assert open(p).read().splitlines() == exp
    p = str(tmp_path / "obj_mot_gt.txt")
    capi.results_write_object_motions(ectx, p, gt, lab)
    assert open(p).read().splitlines() == [f"{i + 1} {lab[i][j]} " + fmt12(gt[i][j]) for i in range(len(gt)) for j in range(1, len(gt[i]))]
    p = str(tmp_path / "obj_centre.txt")
    capi.results_write_object_centres(ectx, p, cen, lab)
    assert open(p).read().splitlines() == [f"{i + 1} {lab[i][j]} " + " ".join("%.9f" % float(v) for v in cen[i][j]) for i in range(len(cen)) for j in range(1, len(cen[i]))]


def ref_err(E):
    t = f32(np.sqrt(f32(f32(E[0, 3] * E[0, 3]) + f32(E[1, 3] * E[1, 3])) + f32(E[2, 3] * E[2, 3])))
    tr = f32(0)
    for j in range(3):
        d = E[j, j]
        tr = f32(float(tr) + 1.0 - (float(d) - 1.0)) if float(d) > 1.0 else f32(tr + d)
    return t, f32(math.acos((float(tr) - 1.0) / 2.0) * 180.0 / 3.1415926)


@pytest.mark.parametrize("seed", [2, 3, 4])
def test_metric_error_follows_getmetricerror(ectx, seed):
    max_id = 4
    cam, cam_gt, mot, pre, gt, lab, stat, cen = make_run(seed, n_frames=9, max_id=max_id)
    r = capi.metric_error(ectx, cam, cam_gt, mot, pre, gt, lab, stat, max_id)
    ts = rs = f32(0)
    for i in range(1, len(cam)):
        E = cvmul(cvmul(cam[i], to_inv(cam[i - 1])), cvmul(cam_gt[i - 1], to_inv(cam_gt[i])))
        t, a = ref_err(E); ts = f32(ts + t); rs = f32(rs + a)
    n = len(cam) - 1
    assert r["cam_t"] == float(f32(ts / f32(n))) and r["cam_r"] == float(f32(rs / f32(n)))
    et, er, ec = np.zeros(max_id - 1, f32), np.zeros(max_id - 1, f32), np.zeros(max_id - 1, np.int32)
    tt = rr = f32(0); cnt = f32(0)
    for i in range(len(mot)):
        for j in range(1, len(mot[i])):
            if not stat[i][j]:
                continue
            body = cvmul(cvmul(to_inv(pre[i][j]), mot[i][j]), pre[i][j])
            t, a = ref_err(cvmul(to_inv(body), gt[i][j]))
            k = lab[i][j] - 1
            et[k] = f32(et[k] + t); er[k] = f32(er[k] + a); ec[k] += 1
            tt = f32(tt + t); rr = f32(rr + a); cnt = f32(cnt + 1)
    if cnt > 0:
        assert r["obj_t"] == float(f32(tt / cnt)) and r["obj_r"] == float(f32(rr / cnt))
    np.testing.assert_array_equal(r["each_count"], ec)
    with np.errstate(invalid="ignore", divide="ignore"):
        np.testing.assert_array_equal(r["each_t"], (et / ec).astype(f32)); np.testing.assert_array_equal(r["each_r"], (er / ec).astype(f32))
