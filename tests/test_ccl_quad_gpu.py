"""CCL + quad extraction parity (gpu): component order/stats and quad candidate set vs the oracle."""
import numpy as np
import pytest

from oracle import ctag_oracle as o

pytestmark = pytest.mark.gpu


def oracle_components(gray):
    half = o.half_resize(gray)
    binary = o.adaptive_threshold(o.convert_to_float(half), 5)
    n, labels, comps = o.connected_components(binary)
    arr = np.array([[c.area, c.x0, c.y0, c.x1, c.y1] for c in comps], np.int32).reshape(-1, 5)
    return n, arr, comps, binary


def test_testbmp_components_and_quads(detector, test_gray, golden_testbmp):
    _, _, info = detector.detect_batch(test_gray[None], 5, True, 5)
    comps = detector.debug_components(0)
    g = golden_testbmp["comps"]
    assert info["n_labels"][0] == int(golden_testbmp["n_labels"])
    assert info["n_legal"][0] == len(g)
    assert np.array_equal(comps[:, 1:], g[:, 1:])  # area + bbox, in OpenCV label order
    idx, quads = detector.debug_quads(0)
    assert info["n_quads"][0] == len(golden_testbmp["quads"])
    assert np.array_equal(idx, golden_testbmp["quad_comp"])  # candidate set: bit-exact
    assert np.abs(quads - golden_testbmp["quads"]).max() <= 1e-3
    # the line fits reproduce the library arithmetic, so the corners are expected to be identical, not just close
    assert np.array_equal(quads, golden_testbmp["quads"])


def blobs(rng, h, w, n, dark=20, bright=200):
    """bright field with dark random convex-ish blobs (quads, ellipses), plus noise."""
    import cv2
    img = np.full((h, w), bright, np.uint8)
    for _ in range(n):
        cx, cy = rng.integers(20, w - 20), rng.integers(20, h - 20)
        if rng.random() < 0.6:
            a, b = rng.integers(6, 40), rng.integers(6, 40)
            ang = rng.uniform(0, np.pi)
            pts = np.array([[-a, -b], [a, -b], [a, b], [-a, b]], np.float64)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            pts = pts @ R.T + [cx, cy]
            cv2.fillConvexPoly(img, np.rint(pts).astype(np.int32), int(dark))
        else:
            cv2.ellipse(img, (int(cx), int(cy)), (int(rng.integers(4, 30)), int(rng.integers(4, 30))),
                        float(rng.uniform(0, 180)), 0, 360, int(dark), -1)
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    img = np.clip(img.astype(np.float32) + rng.normal(0, 2.0, img.shape), 0, 255).astype(np.uint8)
    return img


@pytest.mark.parametrize("shape,nblob,seed", [((240, 320), 12, 1), ((480, 640), 60, 2), ((1080, 1920), 250, 3),
                                              ((300, 2 * 211), 30, 4)])
def test_blob_frames_components_and_quads(detector, shape, nblob, seed):
    rng = np.random.default_rng(seed)
    frames = np.stack([blobs(rng, shape[0], shape[1], nblob) for _ in range(3)])
    _, _, info = detector.detect_batch(frames, 5, False, 3)
    for f in range(len(frames)):
        n, arr, comps, binary = oracle_components(frames[f])
        assert np.array_equal(detector.debug_binary(f), binary)
        got = detector.debug_components(f)
        assert info["n_labels"][f] == n, (shape, f)
        assert np.array_equal(got[:, 1:], arr), (shape, f)
        rows, cols = binary.shape
        quads, qc = o.edge_extraction(comps, cols, rows)
        idx, gq = detector.debug_quads(f)
        assert np.array_equal(idx, np.array(qc, np.int32)), (shape, f)
        if len(qc):
            assert np.abs(gq - np.array(quads)).max() <= 1e-3, (shape, f)


def test_noise_frame_component_order(detector):
    """salt-and-pepper: thousands of tiny components, exercises tile-border merges and the ordered compaction."""
    rng = np.random.default_rng(11)
    h, w = 600, 800
    base = (rng.random((h // 2, w // 2)) < 0.42).astype(np.uint8)
    img = np.where(base.repeat(2, 0).repeat(2, 1) > 0, 30, 220).astype(np.uint8)
    _, _, info = detector.detect_batch(img[None], 5, False, 3)
    n, arr, comps, binary = oracle_components(img)
    assert np.array_equal(detector.debug_binary(0), binary)
    assert info["n_labels"][0] == n
    assert np.array_equal(detector.debug_components(0)[:, 1:], arr)
