"""CPU checks of the image-side oracle (oracle/image_ops.py): the formulas the CUDA kernels implement for the OpenCV-owned
arithmetic are verified bit-exact against the cv2 in this image (the pin), and the oracle's own pieces are sanity-checked."""
import math

import cv2
import numpy as np

from oracle import image_ops as io
from vdo_slam_b200.synth import make_frame


def _resize_fixed_point(src, dw, dh):
    sh, sw = src.shape

    def coeffs(dn, sn):
        scale = sn / dn
        idx, a = np.zeros(dn, np.int64), np.zeros((dn, 2), np.int32)
        for d in range(dn):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(math.floor(float(f))); f = np.float32(f - np.float32(s))
            if s < 0: s, f = 0, np.float32(0)
            if s >= sn - 1: s, f = sn - 1, np.float32(0)
            idx[d] = s
            a[d] = (int(np.rint(np.float32(1.0 - f) * np.float32(2048))), int(np.rint(f * np.float32(2048))))
        return idx, a

    xi, xa = coeffs(dw, sw); yi, ya = coeffs(dh, sh)
    S = src.astype(np.int32)
    H = S[:, xi] * xa[:, 0] + S[:, np.minimum(xi + 1, sw - 1)] * xa[:, 1]
    out = (((ya[:, 0:1] * (H[yi] >> 4)) >> 16) + ((ya[:, 1:2] * (H[np.minimum(yi + 1, sh - 1)] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def test_resize_formula_is_bit_exact_vs_cv2():
    rng = np.random.default_rng(0)
    for (h, w) in [(375, 1242), (313, 1035), (97, 131)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dw, dh = int(np.rint(w / 1.2)), int(np.rint(h / 1.2))
        assert np.array_equal(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), _resize_fixed_point(img, dw, dh))


_OFF = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def _fast_score(img):
    h, w = img.shape; I = img.astype(np.int32)
    c = I[3:h - 3, 3:w - 3]
    d = np.stack([c - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for (dx, dy) in _OFF], 0)
    best = np.full(c.shape, -10 ** 6)
    for k in range(16):
        idx = [(k + j) % 16 for j in range(9)]
        best = np.maximum(best, np.maximum(d[idx].min(0), (-d[idx]).min(0)))
    sc = np.zeros((h, w), np.int32)
    sc[3:h - 3, 3:w - 3] = np.maximum(best - 1, 0)
    return sc


def test_fast_score_and_nms_formula_matches_cv2():
    rng = np.random.default_rng(1)
    img = cv2.GaussianBlur(rng.integers(0, 256, (160, 220), dtype=np.uint8), (5, 5), 1.2)
    sc = _fast_score(img)
    for thr in (20, 7):
        kps = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(img)
        s = np.where(sc >= thr, sc, 0); p = np.pad(s, 1)
        nb = np.stack([p[1 + dy:p.shape[0] - 1 + dy, 1 + dx:p.shape[1] - 1 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)], 0).max(0)
        ys, xs = np.nonzero((sc >= thr) & (s > nb))
        assert [(int(k.pt[0]), int(k.pt[1])) for k in kps] == list(zip(xs.tolist(), ys.tolist()))      # same set, same order
        assert [int(k.response) for k in kps] == sc[ys, xs].tolist()                                      # same responses


def test_orb_oracle_quota_and_order():
    f = make_frame(0)
    prm = io.OrbParams()
    assert prm.per_level == [543, 452, 377, 314, 262, 218, 182, 152]
    r = io.orb_extract(f["gray"], prm, with_angle=False)
    assert 2000 < len(r["x"]) < 2700
    assert (np.diff(r["octave"]) >= 0).all()                                 # level-major output
    assert r["x"].min() >= 16 and r["x"].max() < 1242 - 16


def test_sampling_oracle_properties():
    f = make_frame(3)
    d = io.depth_prep(f["depth_raw"], 387.5744, 256.0)
    assert (d[f["depth_raw"] < 0] == 0).all() and np.isfinite(d[f["depth_raw"] > 0]).all()
    s = io.sample_objects(f["mask"], d, f["flow"], 25.0)
    assert len(s["x"]) > 100 and (s["label"] > 0).all() and (s["x"] % 4 == 0).all() and (s["y"] % 4 == 0).all()
    key = s["y"].astype(np.int64) * 10000 + s["x"]
    assert (np.diff(key) > 0).all()                                           # raster order
