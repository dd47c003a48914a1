"""The CUDA JPEG decoder's cores (cylindertag_b200/csrc/jpeg_core.cuh: entropy decoder, integer inverse DCT, fancy chroma
upsampling, fixed-point colour conversion) compiled for the CPU, against OpenCV's decoder (libjpeg-turbo) -- CPU only.
The decoded pixels have to be the ones cv::imdecode gives, byte for byte: then a JPEG frame going through the GPU ingest
path is the frame the reference would have read with cv::imread."""
import ctypes

import cv2
import numpy as np
import pytest

from tests import configs
from tests.harness_api import lib, vp


def decode(buf):
    """Decodes with both compositions of the cores -- block by block (mode 0) and coefficients first, inverse DCT afterwards
    with the word-wise bit reader (mode 1, what the GPU kernels do) -- and requires them to agree."""
    L = lib()
    L.hh_jpeg_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    b = np.ascontiguousarray(buf, np.uint8).reshape(-1)
    w, h = ctypes.c_int(), ctypes.c_int()
    rc = L.hh_jpeg_decode(vp(b), b.size, None, 0, 0, 0, ctypes.byref(w), ctypes.byref(h), 0)
    if rc:
        return rc, None
    outs = []
    for mode in (0, 1):
        out = np.zeros((h.value, w.value, 3), np.uint8)
        rc = L.hh_jpeg_decode(vp(b), b.size, vp(out), w.value * 3, w.value, h.value, ctypes.byref(w), ctypes.byref(h), mode)
        if rc:
            return rc, None
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    return rc, outs[1]


def encode(img, quality=90, rst=16, sampling=None):
    params = [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    if sampling is not None:
        params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling]
    ok, buf = cv2.imencode(".jpg", img, params)
    assert ok
    return buf.reshape(-1)


@pytest.mark.parametrize("sampling", [None, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
def test_colour_frames_decode_like_imdecode(sampling):
    rng = np.random.default_rng(3)
    for (h, w, q, rst) in [(64, 96, 90, 4), (1080, 1920, 90, 16), (123, 211, 75, 1), (50, 34, 98, 7), (17, 17, 60, 2), (240, 320, 30, 40)]:
        if h * w > 1e6:
            img = configs.config3_frame(5)
        else:
            img = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 1.2)
            img[h // 4:h // 2, w // 4:w // 2] = (20, 30, 25)  # a dark block with sharp edges
        buf = encode(img, q, rst, sampling)
        rc, got = decode(buf)
        assert rc == 0, (h, w, rc)
        want = cv2.imdecode(buf, cv2.IMREAD_COLOR)
        assert got.shape == want.shape
        assert np.array_equal(got, want), (h, w, q, rst, int(np.abs(got.astype(int) - want).max()), int((got != want).sum()))


def test_gray_frame_and_extreme_values():
    rng = np.random.default_rng(4)
    g = cv2.GaussianBlur(rng.integers(0, 256, (200, 301), dtype=np.uint8), (0, 0), 1.0)
    buf = encode(g, 92, 8)
    rc, got = decode(buf)
    want = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    assert rc == 0 and np.array_equal(got, want)
    # saturated checkerboards: the clamps of the inverse DCT and of the colour conversion
    c = np.zeros((64, 64, 3), np.uint8)
    c[::2, ::2] = 255
    c[1::2, 1::2] = (255, 0, 255)
    buf = encode(c, 100, 3)
    rc, got = decode(buf)
    assert rc == 0 and np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_streams_outside_the_envelope_are_refused():
    img = np.zeros((32, 32, 3), np.uint8)
    assert decode(encode(img, 90, 0))[0] == 2                                   # no restart markers: one sequential chain
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_RST_INTERVAL, 4])
    assert decode(prog.reshape(-1))[0] == 2                                     # progressive
    assert decode(np.frombuffer(b"definitely not a jpeg", np.uint8))[0] == 1
    assert decode(encode(img, 90, 4)[:200])[0] in (1, 2)                        # truncated
