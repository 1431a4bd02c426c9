"""decoders of the reference driver's per-frame input files, pinned against OpenCV (cv2 4.13) -- the library the
reference itself calls (example/vdo_slam.cc:105-117) -- and, for the text mask, against the statement of LoadMask."""
import os
import subprocess

import cv2
import numpy as np
import pytest

from vdo_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul", "libvdo_emul.so")


@pytest.fixture(scope="module")
def ectx():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "libvdo_emul.so"], stdout=subprocess.DEVNULL)
    return capi.Context(0, lib_path=EMUL)


@pytest.mark.parametrize("shape,dtype", [((375, 1242), np.uint16), ((375, 1242), np.uint8), ((120, 161, 3), np.uint8), ((33, 47, 4), np.uint8), ((21, 30, 3), np.uint16)])
def test_png_equals_cv2_imread_unchanged(ectx, tmp_path, shape, dtype):
    rng = np.random.default_rng(len(shape) * 7 + shape[1])
    # smooth + noisy content so that the encoder uses several scan-line filter types
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    base = (np.sin(xx / 17.0) * np.cos(yy / 11.0) * 0.5 + 0.5) * np.iinfo(dtype).max
    img = base[..., None] if len(shape) == 3 else base
    img = np.broadcast_to(img, shape) + rng.integers(0, 9, shape)
    img = np.clip(img, 0, np.iinfo(dtype).max).astype(dtype)
    for level in (1, 9):
        p = str(tmp_path / f"a{level}.png")
        assert cv2.imwrite(p, img, [cv2.IMWRITE_PNG_COMPRESSION, level])
        ref = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        got = capi.io_read_png(ectx, p)
        assert got.dtype == ref.dtype and got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)
    if len(shape) == 2:
        f = capi.io_read_png_gray_f32(ectx, p, shape[1], shape[0])
        np.testing.assert_array_equal(f, ref.astype(np.float32))        # imD.convertTo(imD_f, CV_32F)


def test_png_filter_types_are_all_exercised(ectx, tmp_path):
    """Hand-built PNG with one scan line per filter type (None, Sub, Up, Average, Paeth), checked against cv2."""
    import struct, zlib
    w, h = 13, 5
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    raw = bytearray()
    prev = np.zeros((w, 3), np.int32)
    for y in range(h):
        cur = img[y].astype(np.int32)
        left = np.vstack([np.zeros((1, 3), np.int32), cur[:-1]])
        ul = np.vstack([np.zeros((1, 3), np.int32), prev[:-1]])
        if y == 0: enc = cur
        elif y == 1: enc = cur - left
        elif y == 2: enc = cur - prev
        elif y == 3: enc = cur - ((left + prev) >> 1)
        else:
            p = left + prev - ul
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            enc = cur - pred
        raw += bytes([y]) + (enc & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b"")
    p = tmp_path / "filters.png"; p.write_bytes(png)
    ref = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
    np.testing.assert_array_equal(ref[:, :, ::-1], img)                  # cv2 agrees with the construction (BGR vs RGB)
    np.testing.assert_array_equal(capi.io_read_png(ectx, str(p)), ref)


def test_flo_equals_cv2_read_optical_flow(ectx, tmp_path):
    rng = np.random.default_rng(1)
    flow = rng.normal(0, 5, (375, 1242, 2)).astype(np.float32)
    p = str(tmp_path / "f.flo")
    assert cv2.writeOpticalFlow(p, flow)
    np.testing.assert_array_equal(capi.io_read_flo(ectx, p), cv2.readOpticalFlow(p))
    np.testing.assert_array_equal(capi.io_read_flo(ectx, p), flow)
    bad = tmp_path / "bad.flo"; bad.write_bytes(b"XXXX" + flow.tobytes()[:64])
    with pytest.raises(capi.VdoError):
        capi.io_read_flo(ectx, str(bad))


def test_mask_txt_follows_loadmask(ectx, tmp_path):
    rng = np.random.default_rng(2)
    h, w = 37, 59
    m = np.where(rng.random((h, w)) < 0.2, rng.integers(1, 70, (h, w)), 0).astype(np.int32)
    p = tmp_path / "m.txt"
    p.write_text("\n".join(" ".join(str(v) for v in row) + " " for row in m) + "\n\n")     # trailing blanks / empty lines as in the data set
    np.testing.assert_array_equal(capi.io_read_mask_txt(ectx, str(p), w, h), m)
    np.testing.assert_array_equal(capi.io_read_mask_txt(ectx, str(p), w, h), np.loadtxt(str(p), dtype=np.int32))
    with pytest.raises(capi.VdoError):
        capi.io_read_mask_txt(ectx, str(p), w, h - 1)                      # more text rows than image rows


def test_unsupported_png_flavours_are_refused(ectx, tmp_path):
    pal = np.zeros((8, 8), np.uint8)
    p = str(tmp_path / "interlaced.png")
    # a 1-bit image: bit depth < 8 is outside what the driver's inputs use
    assert cv2.imwrite(p, pal, [cv2.IMWRITE_PNG_BILEVEL, 1])
    with pytest.raises(capi.VdoError):
        capi.io_read_png(ectx, p)
    with pytest.raises(capi.VdoError):
        capi.io_read_png(ectx, str(tmp_path / "missing.png"))


def test_decoders_reject_corrupt_headers_and_directories(tmp_path):
    """Untrusted sizes behind the C ABI: a PNG whose IHDR claims a gigantic image, a truncated .flo and a directory must come back as an
    error code, not as std::bad_alloc through extern "C"."""
    import struct, zlib
    from vdo_slam_b200 import capi
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    bogus = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 60000, 60000, 16, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    p = tmp_path / "bogus.png"; p.write_bytes(bogus)
    L = capi.load(EMUL) if "EMUL" in globals() else capi.load()
    import ctypes as C
    buf = (C.c_ubyte * 16)()
    assert L.vdo_io_read_png(str(p).encode(), buf, C.c_size_t(16)) != 0
    assert L.vdo_io_read_png(str(tmp_path).encode(), buf, C.c_size_t(16)) != 0          # a directory
    out = C.c_void_p()
    assert L.vdo_g2o_read(str(tmp_path).encode(), C.byref(out)) != 0
