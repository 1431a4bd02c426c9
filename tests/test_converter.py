"""A19: the float <-> double boundary (src/Converter.cc:25-41, 151-166) and the cv::Mat 4x4 product, as host functions of the C ABI.
toInvMatrix and the product are pinned against OpenCV's own gemm (cv2.gemm: the two roundings a cv::Mat expression can take);
toSE3Quat / toCvMat against an independent numpy restatement of Eigen's published quaternion conversions (Eigen is not in this image)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul", "libvdo_emul.so")        # the host-only translation units are linked into the emulation library too


@pytest.fixture(scope="module")
def L():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "libvdo_emul.so"], stdout=subprocess.DEVNULL)
    return C.CDLL(EMUL)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _rand_T(rng, noise=0.0):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = rng.uniform(-3.1, 3.1)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = (R + noise * rng.normal(size=(3, 3))).astype(np.float32); T[:3, 3] = rng.normal(scale=5, size=3).astype(np.float32)
    return T


def _eigen_quat(R):
    """Eigen::Quaternion(Matrix3d) (trace branch, else largest diagonal), as published."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t; q[1] = (R[0, 2] - R[2, 0]) * t; q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]: i = 1
        if R[2, 2] > R[i, i]: i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t; t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t; q[j] = (R[j, i] + R[i, j]) * t; q[k] = (R[k, i] + R[i, k]) * t
    return q


def test_inv_matrix_and_product_are_opencv_gemm_bit_for_bit(L):
    import cv2
    rng = np.random.default_rng(0)
    for _ in range(200):
        A, B = _rand_T(rng), _rand_T(rng)
        out = np.zeros((4, 4), np.float32)
        assert L.vdo_convert_mul4(_fp(A), _fp(B), _fp(out)) == 0
        assert np.array_equal(out, cv2.gemm(A, B, 1.0, None, 0.0))                       # small-matrix branch: float accumulation
        assert L.vdo_convert_inv_matrix(_fp(A), _fp(out)) == 0
        want = np.eye(4, dtype=np.float32)
        want[:3, :3] = A[:3, :3].T
        want[:3, 3:4] = cv2.gemm(np.ascontiguousarray(A[:3, :3]), np.ascontiguousarray(A[:3, 3:4]), -1.0, None, 0.0, flags=cv2.GEMM_1_T)
        assert np.array_equal(out, want)                                                   # -R.t()*t: generic branch, one rounding


def test_to_se3quat_and_back(L):
    rng = np.random.default_rng(1)
    for it in range(300):
        T = _rand_T(rng, noise=1e-4 if it % 3 == 0 else 0.0)          # float matrices are never exactly orthonormal; some are visibly off
        q, t = np.zeros(4), np.zeros(3)
        assert L.vdo_convert_to_se3quat(_fp(T), _dp(q), _dp(t)) == 0
        ref = _eigen_quat(T[:3, :3].astype(np.float64))
        if ref[3] < 0:
            ref = -ref
        ref /= np.linalg.norm(ref)
        np.testing.assert_allclose(q, ref, atol=1e-15)
        assert q[3] >= 0 and abs(np.linalg.norm(q) - 1) < 1e-15 and np.array_equal(t, T[:3, 3].astype(np.float64))
        back = np.zeros((4, 4), np.float32)
        assert L.vdo_convert_to_cvmat(_dp(q), _dp(t), _fp(back)) == 0
        x, y, z, w = q
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(back[:3, :3] - Rq.astype(np.float32)).max() <= 6e-8 and np.array_equal(back[:3, 3], T[:3, 3]) and back[3].tolist() == [0, 0, 0, 1]
        if it % 3:                                                      # orthonormal input: the round trip is the identity to float rounding
            assert np.abs(back - T).max() <= 3e-7


def test_shim_gray_conversion_follows_opencv_3_4_and_is_within_one_of_cv2_4():
    """System::TrackRGBD converts colour input with cvtColor(RGB2GRAY / BGR2GRAY) (src/Tracking.cc:209-222).  The shim restates the 8-bit
    fixed-point formula of the OpenCV the reference builds (3.4.0: 14-bit coefficients 4899 / 9617 / 1868); the cv2 in this image (4.13) uses
    15-bit coefficients (9798 / 19235 / 3735) -- version drift of at most one grey level, checked here against cv2.cvtColor itself."""
    import cv2
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (97, 211, 3), dtype=np.uint8)
    r, g, b = img[..., 0].astype(np.int64), img[..., 1].astype(np.int64), img[..., 2].astype(np.int64)
    v34 = ((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14)
    v4 = ((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15)
    ref = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY).astype(np.int64)
    assert np.array_equal(v4, ref)                                     # the 4.x arithmetic, restated, is cv2's
    assert np.abs(v34 - ref).max() <= 1 and (v34 != ref).mean() < 0.01  # 3.4 vs 4.13: at most one level, < 1 % of the pixels
    src = open(os.path.join(ROOT, "vdo_slam_b200", "host", "System.cc")).read()
    assert "R * 4899 + G * 9617 + B * 1868 + (1 << 13)) >> 14" in src
