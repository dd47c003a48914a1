"""K1 parity (gpu): fused gray + 2x cubic decimation + adaptive threshold vs the oracle, bit-exact."""
import numpy as np
import pytest

from oracle import ctag_oracle as o

pytestmark = pytest.mark.gpu


def oracle_binary(gray):
    half = o.half_resize(gray)
    return o.adaptive_threshold(o.convert_to_float(half), 5)


def test_testbmp_binary_bit_exact(detector, test_gray, golden_testbmp):
    detector.detect_batch(test_gray[None], 5, True, 5)
    got = detector.debug_binary(0)
    assert np.array_equal(got, golden_testbmp["binary"])
    assert np.array_equal(detector.debug_gray(0), test_gray)


@pytest.mark.parametrize("shape", [(64, 96), (202, 326), (400, 640), (1080, 1920), (90, 2 * 163)])
def test_random_gray_binary_bit_exact(detector, shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    h, w = shape
    frames = np.stack([
        rng.integers(0, 256, (h, w), dtype=np.uint8),
        # smooth dark/bright field so that thresholds fall on both sides of 0.3
        np.clip(rng.normal(60, 40, (h, w)), 0, 255).astype(np.uint8),
        (rng.integers(0, 2, (h // 2, w // 2), dtype=np.uint8) * 200 + 20).repeat(2, 0).repeat(2, 1),
    ])
    detector.detect_batch(frames, 5, False, 3)
    for f in range(len(frames)):
        assert np.array_equal(detector.debug_binary(f), oracle_binary(frames[f])), (shape, f)


@pytest.mark.parametrize("shape", [(64, 96), (1200, 1920), (202, 326)])
def test_bgr_gray_and_binary_bit_exact(detector, shape):
    rng = np.random.default_rng(shape[0] + 31 * shape[1])
    h, w = shape
    frames = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    frames[1] = (frames[1] * 0.35).astype(np.uint8)
    detector.detect_batch(frames, 5, False, 3)
    for f in range(2):
        gray = o.bgr2gray(frames[f])
        assert np.array_equal(detector.debug_gray(f), gray), (shape, f)
        assert np.array_equal(detector.debug_binary(f), oracle_binary(gray)), (shape, f)


@pytest.mark.parametrize("win", [3, 4, 7, 11])
def test_generic_threshold_window(detector, test_gray, win):
    """adaptiveThresh other than 5 goes through the unfused kernels; same bit-exact bar."""
    rng = np.random.default_rng(win)
    frames = np.stack([test_gray[:400, :640], np.clip(rng.normal(70, 45, (400, 640)), 0, 255).astype(np.uint8)])
    detector.detect_batch(frames, win, False, 3)
    for f in range(2):
        half = o.half_resize(frames[f])
        assert np.array_equal(detector.debug_binary(f), o.adaptive_threshold(o.convert_to_float(half), win)), (win, f)
    bgr = rng.integers(0, 256, (1, 202, 326, 3), dtype=np.uint8)
    detector.detect_batch(bgr, win, False, 3)
    gray = o.bgr2gray(bgr[0])
    assert np.array_equal(detector.debug_gray(0), gray)
    assert np.array_equal(detector.debug_binary(0), o.adaptive_threshold(o.convert_to_float(o.half_resize(gray)), win))


def test_bad_window_is_loud(detector, test_gray):
    from cylindertag_b200 import CtagError
    with pytest.raises(CtagError):
        detector.detect_batch(test_gray[None, :64, :64], 0, False, 3)
