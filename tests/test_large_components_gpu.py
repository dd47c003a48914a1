"""Paths that ordinary frames do not reach: components too large for the shared-memory fast path (global scratch),
edges longer than the per-thread weight cache (128) and longer than the restart pick table (1024), exactly
axis-aligned edges (sub-EPS bookkeeping)."""
import cv2
import numpy as np
import pytest

from oracle import ctag_oracle as o

pytestmark = pytest.mark.gpu


def _frame(h, w, rects, blur=0.0, seed=0):
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 210, np.uint8)
    for (cx, cy, a, b, ang) in rects:
        pts = np.array([[-a, -b], [a, -b], [a, b], [-a, b]], np.float64)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        cv2.fillConvexPoly(img, np.rint(pts @ R.T + [cx, cy]).astype(np.int32), 20)
    if blur > 0:
        img = cv2.GaussianBlur(img, (0, 0), blur)
    img = np.clip(img.astype(np.float32) + rng.normal(0, 1.5, img.shape), 0, 255).astype(np.uint8)
    return img


def _compare(detector, frame):
    half = o.half_resize(frame)
    binary = o.adaptive_threshold(o.convert_to_float(half), 5)
    n, labels, comps = o.connected_components(binary)
    rows, cols = binary.shape
    dbg = []
    quads, qc = o.edge_extraction(comps, cols, rows, dbg)
    _, _, info = detector.detect_batch(frame[None], 5, True, 5)
    assert np.array_equal(detector.debug_binary(0), binary)
    got = detector.debug_components(0)
    assert np.array_equal(got[:, 1:], np.array([[c.area, c.x0, c.y0, c.x1, c.y1] for c in comps], np.int32).reshape(-1, 5))
    idx, gq = detector.debug_quads(0)
    assert np.array_equal(idx, np.array(qc, np.int32))
    if qc:
        assert np.abs(gq - np.array(quads)).max() <= 1e-3
    return comps, dbg, quads


def test_large_rotated_rectangles_4k(detector):
    # half-res boxes far beyond the 256-point / 256-word shared-memory fast path; edges of 150-400 boundary points
    frame = _frame(2160, 3840, [(900, 600, 380, 90, 0.3), (2600, 700, 300, 110, -0.7), (1800, 1500, 420, 60, 1.2),
                                (3200, 1600, 200, 150, 0.05)], blur=1.0, seed=1)
    comps, dbg, quads = _compare(detector, frame)
    # (large dark regions are hollowed out by the 5x5 adaptive threshold, so they need not yield quads: what matters
    # here is that the large-component code path agrees with the oracle)
    assert max(d.n_trace for d in dbg) > 512


def test_edge_longer_than_pick_table(detector):
    # a thin bar: ~1150 boundary points per long edge (> 1024) with an area still inside the 1 % limit
    frame = _frame(2160, 3840, [(1920, 1000, 1150, 14, 0.02), (1920, 1500, 1100, 12, -0.015)], blur=0.8, seed=2)
    comps, dbg, quads = _compare(detector, frame)
    assert max(d.n_trace for d in dbg) > 2200


def test_axis_aligned_rectangles_sub_eps_paths(detector):
    # exactly horizontal/vertical edges without noise: collinear clusters, errors below EPS, identical restarts
    img = np.full((1080, 1920), 215, np.uint8)
    for k, (x, y, w_, h_) in enumerate([(200, 200, 120, 40), (600, 300, 60, 200), (1000, 500, 300, 30), (1400, 700, 24, 24),
                                        (300, 700, 17, 91)]):
        img[y:y + h_, x:x + w_] = 15
    comps, dbg, quads = _compare(detector, img)
    assert len(quads) >= 1
