"""Pins the oracle to the reference's OWN code -- CPU only.

oracle/_ref/libctag_ref.so is /root/reference's corner_detector.cpp, CylinderTag.cpp and pose_estimation.cpp compiled
unmodified (oracle/build_ref.py) against a stand-in for the OpenCV / Ceres entry points they call (oracle/ref_shim/).
Three layers of evidence:
  1. every stand-in primitive against the real library (cv2 4.13) on random inputs -- known-answer tests;
  2. the compiled reference with its primitives routed to cv2 itself (callback backend) against the same binary using
     the restatements: identical stage dumps, so the result is "reference code + OpenCV arithmetic";
  3. the Python oracle (oracle/ctag_oracle.py, which the GPU parity tests use for stage-level comparisons) against the
     compiled reference on test.bmp, frames of the config-2 sequence and rendered config-3/4 frames: every integer
     stage output equal, floats equal to 1e-6 px.
The frozen files tests/golden/ref_*.npz are outputs of the compiled reference (tests/golden/make_golden_ref.py)."""
import ctypes
import os

import cv2
import numpy as np
import pytest

from cylindertag_b200 import synth
from oracle import ctag_oracle as o
from oracle import pose_oracle as po
from oracle import ref_api as R
from tests import configs

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref is not built and /root/reference is not present")
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def lib():
    return R.load()


@pytest.fixture(scope="module")
def ref(marker_path):
    r = R.RefDetector(marker_path=marker_path)
    yield r
    r.close()


# ---- 1. primitives against cv2 --------------------------------------------------------------------------------------
def test_resize_even_sizes_bit_exact(lib):
    rng = np.random.default_rng(5)
    for (h, w) in [(64, 96), (202, 326), (90, 74), (10, 18), (480, 642), (1080, 1920), (118, 122), (86, 152), (2, 2), (4, 6)]:
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dst = np.zeros((h // 2, w // 2), np.uint8)
        assert lib.shim_resize_cubic_u8(vp(src), w, h, vp(dst), w // 2, h // 2) == 0
        assert np.array_equal(dst, cv2.resize(src, (w // 2, h // 2), fx=0.5, fy=0.5, interpolation=cv2.INTER_CUBIC)), (h, w)


def test_resize_odd_sizes_equal_the_undispatched_library(lib):
    """Non-integer scales: cv2's own answer depends on cv::setUseOptimized (CPU dispatch); the stand-in follows the
    published generic path, which is what the library computes with dispatch off.  (Why the C ABI keeps to even sizes.)"""
    rng = np.random.default_rng(6)
    try:
        cv2.setUseOptimized(False)
        for (h, w) in [(61, 75), (101, 203), (33, 31), (255, 257), (7, 9)]:
            src = rng.integers(0, 256, (h, w), dtype=np.uint8)
            dst = np.zeros((h // 2, w // 2), np.uint8)
            assert lib.shim_resize_cubic_u8(vp(src), w, h, vp(dst), w // 2, h // 2) == 0
            assert np.array_equal(dst, cv2.resize(src, (w // 2, h // 2), fx=0.5, fy=0.5, interpolation=cv2.INTER_CUBIC)), (h, w)
    finally:
        cv2.setUseOptimized(True)


def test_convert_and_gray_bit_exact(lib):
    v = np.arange(256, dtype=np.uint8)
    out = np.zeros(256, np.float32)
    lib.shim_convert_u8_f32.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    lib.shim_convert_u8_f32(vp(v), 256, 1.0 / 255, vp(out))
    assert np.array_equal(out, o.convert_to_float(v.reshape(1, 256)).reshape(-1))
    rng = np.random.default_rng(4)
    bgr = rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)
    g = np.zeros((50, 70), np.uint8)
    lib.shim_bgr2gray(vp(bgr), 70, 50, vp(g))
    assert np.array_equal(g, cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))


def test_ccl_labels_equal_bbdt(lib):
    rng = np.random.default_rng(7)
    for (h, w, p) in [(301, 403, 0.5), (120, 96, 0.35), (64, 64, 0.62), (50, 51, 0.5), (1, 40, 0.5), (40, 1, 0.5), (200, 300, 0.05)]:
        img = ((rng.random((h, w)) < p) * 255).astype(np.uint8)
        if h > 100:
            img[10:60, 20:90] = 255  # a big blob with holes next to the noise
            img[30:40, 30:50] = 0
        lab = np.zeros((h, w), np.int32)
        stats = np.zeros((h * w + 1, 5), np.int32)
        n = lib.shim_ccl(vp(img), w, h, vp(lab), vp(stats), h * w + 1)
        n2, lab2, st2, _ = cv2.connectedComponentsWithStatsWithAlgorithm(img, 8, cv2.CV_32S, cv2.CCL_BBDT)
        assert n == n2 and np.array_equal(lab, lab2), (h, w)
        assert np.array_equal(stats[:n], st2)


def _fit(lib, pts, dist):
    xy = np.ascontiguousarray(pts, np.int32)
    out = np.zeros(4, np.float32)
    assert lib.shim_fit_line(vp(xy), len(xy), dist, vp(out)) == 0
    return out


def test_fitline_l2_and_welsch_bit_exact(lib):
    rng = np.random.default_rng(8)
    worst = 0.0
    exact = total = 0
    for it in range(700):
        n = int(rng.integers(2, 120))
        kind = it % 5
        if kind == 0:    # noisy line
            t = rng.uniform(0, np.pi)
            s = np.arange(n) - n / 2
            pts = np.stack([200 + s * np.cos(t) + rng.normal(0, 0.6, n), 150 + s * np.sin(t) + rng.normal(0, 0.6, n)], 1)
        elif kind == 1:  # line with outliers (corner of a quad)
            s = np.arange(n)
            pts = np.stack([50 + s, 80 + 0.3 * s], 1)
            k = max(1, n // 5)
            pts[-k:, 1] += np.arange(k) * 1.5
        elif kind == 2:  # exactly collinear (sub-EPS error bookkeeping)
            s = np.arange(n)
            pts = np.stack([10 + s, 20 + 2 * s], 1)
        elif kind == 3:  # axis aligned
            s = np.arange(n)
            pts = np.stack([np.full(n, 33), 5 + s], 1) if it % 2 else np.stack([7 + s, np.full(n, 91)], 1)
        else:            # blob
            pts = rng.integers(0, 40, (n, 2))
        pts = np.rint(pts).astype(np.int32)
        for dist in (cv2.DIST_L2, cv2.DIST_WELSCH):
            got = _fit(lib, pts, dist)
            want = cv2.fitLine(pts.astype(np.float32).reshape(-1, 1, 2), dist, 0, 0.01, 0.01).reshape(4)
            d = float(np.abs(got - want).max())
            worst = max(worst, d)
            exact += int(np.array_equal(got, want))
            total += 1
    assert exact == total, (exact, total, worst)


def test_solve_determinant_fastatan2(lib):
    rng = np.random.default_rng(9)
    lib.shim_fast_atan2.restype = ctypes.c_float
    lib.shim_fast_atan2.argtypes = [ctypes.c_float, ctypes.c_float]
    for _ in range(3000):
        a = (rng.normal(0, 10, 4)).astype(np.float32)
        b = (rng.normal(0, 300, 2)).astype(np.float32)
        x = np.zeros(2, np.float32)
        det = ctypes.c_double()
        rc = lib.shim_solve2x2(vp(a), vp(b), vp(x), ctypes.byref(det))
        assert det.value == cv2.determinant(a.reshape(2, 2))
        ok, want = cv2.solve(a.reshape(2, 2), b.reshape(2, 1))
        assert rc == 0 and ok and np.array_equal(x, want.reshape(2))
    for _ in range(3000):
        y, x = (float(v) for v in rng.normal(0, 50, 2).astype(np.float32))
        assert abs(lib.shim_fast_atan2(y, x) - cv2.fastAtan2(y, x)) <= 1e-4
    for (y, x) in [(0.0, 0.0), (1.0, 0.0), (-1.0, 0.0), (0.0, -1.0), (0.0, 1.0), (3.0, 3.0), (-2.0, 2.0)]:
        assert abs(lib.shim_fast_atan2(y, x) - cv2.fastAtan2(y, x)) <= 1e-4


def test_undistort_and_project_against_cv2(lib):
    K, D = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    rng = np.random.default_rng(10)
    n = 64
    xyz = np.stack([rng.uniform(-60, 60, n), rng.uniform(-60, 60, n), rng.uniform(-20, 20, n)], 1).astype(np.float32)
    rvec, tvec = np.array([0.3, -0.5, 0.2]), np.array([10.0, -20.0, 600.0])
    proj, _ = cv2.projectPoints(xyz, rvec, tvec, K, D)
    xy = proj.reshape(-1, 2).astype(np.float32)
    und, pr = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
    assert lib.shim_undistort_project(vp(xy), vp(xyz), n, vp(np.ascontiguousarray(K, np.float32)), vp(np.ascontiguousarray(D, np.float32)),
                                      vp(rvec), vp(tvec), vp(und), vp(pr)) == 0
    assert np.abs(pr - xy).max() <= 2e-3  # float32 outputs at ~1e3 px
    want = cv2.undistortPoints(xy.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    assert np.abs(und - want).max() <= 1e-3


# ---- 2. the compiled reference: restated primitives vs cv2 doing the arithmetic -------------------------------------
def _same_dump(a, b, tol=0.0):
    assert a.n_labels == b.n_labels and np.array_equal(a.labels, b.labels) and np.array_equal(a.comps, b.comps)
    assert a.status == b.status and a.flagged == b.flagged and a.n_groups == b.n_groups
    assert a.quads.shape == b.quads.shape and a.feats.shape == b.feats.shape and len(a.markers) == len(b.markers)
    if tol == 0.0:
        assert np.array_equal(a.quads, b.quads) and np.array_equal(a.feats, b.feats)
    else:
        assert np.abs(a.quads - b.quads).max(initial=0) <= tol and np.abs(a.feats - b.feats).max(initial=0) <= tol
    for m, w in zip(a.markers, b.markers):
        assert (m.markerID, m.featurePos, m.feature_ID, m.feature_ID_left, m.feature_ID_right) == \
               (w.markerID, w.featurePos, w.feature_ID, w.feature_ID_left, w.feature_ID_right)
        assert np.abs(m.cornerLists - w.cornerLists).max(initial=0) <= tol


def test_reference_with_cv2_backend_equals_reference_with_restatements(ref, test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = [test_gray, synth.video_sequence(test_gray, 120, 2024, first=7, count=1)[0],
              o.bgr2gray(configs.config3_frame(3)), o.bgr2gray(configs.config4_frame("2f12c", 1)[0])]
    for k, g in enumerate(frames):
        native = ref.detect(g, 5, True, 5)
        with R.cv2_backend() as b:
            real = ref.detect(g, 5, True, 5)
            assert b.stats["resize_cubic_u8"] == 1 and b.stats["ccl_bbdt"] == 1 and b.stats["fit_line"] > 50
        _same_dump(native, real)
        assert len(native.markers) >= 1, k


# ---- 3. the Python oracle against the compiled reference ------------------------------------------------------------
def _oracle_equals_ref(d, od, tol=1e-6):
    assert d.n_labels == od.n_labels
    assert np.array_equal(d.binary, od.binary)
    assert np.array_equal(d.comps, np.array([[c.area, c.x0, c.y0, c.x1, c.y1] for c in od.comps], np.int32).reshape(-1, 5))
    oq = np.array(od.quads, np.float32).reshape(-1, 4, 2)
    assert d.quads.shape == oq.shape and np.abs(d.quads - oq).max(initial=0) <= tol
    if od.flagged:
        assert d.flagged
        return
    assert d.status == od.status
    if od.status != "ok":
        assert len(d.markers) == 0
        return
    of = np.array([f.corners for f in od.feats], np.float32).reshape(-1, 8, 2)
    assert d.feats.shape == of.shape and np.abs(d.feats - of).max(initial=0) <= tol
    assert np.abs(d.feats_angle - np.array([f.angle for f in od.feats], np.float32)).max(initial=0) <= 1e-4
    assert d.n_groups == len(od.groups) and len(d.markers) == len(od.markers)
    state = None
    for m, om in zip(d.markers, od.markers):
        assert m.markerID == om.markerID and m.featurePos == list(om.featurePos) and m.feature_ID == list(om.feature_ID)
        assert m.feature_ID_left == list(om.feature_ID_left) and m.feature_ID_right == list(om.feature_ID_right)
        assert np.abs(m.cornerLists - np.array(om.cornerLists, np.float32)).max() <= tol
        assert np.abs(m.feature_center - np.array(om.feature_center, np.float32)).max() <= tol
        assert np.allclose(m.cr_left, np.array(om.cr_left, np.float32), rtol=0, atol=1e-6)
        assert np.allclose(m.cr_right, np.array(om.cr_right, np.float32), rtol=0, atol=1e-6)
        assert np.allclose(m.edge_length, np.array(om.edge_length, np.float32), rtol=0, atol=1e-5)


def test_oracle_equals_reference_on_testbmp_every_stage(ref, test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    assert ref.feature_size == fs and np.array_equal(ref.state, state)
    for subpix, dist in ((True, 5), (True, 3), (False, 3)):
        d = ref.detect(test_gray, 5, subpix, dist)
        od = o.detect(test_gray, state, fs, 5, subpix, dist)
        _oracle_equals_ref(d, od)
    d = ref.detect(test_gray, 5, True, 5)
    # SURVEY Appendix E, now from the reference's own code
    assert (d.n_labels, len(d.comps), len(d.quads), len(d.feats), d.n_groups) == (135, 90, 59, 26, 7)
    assert [m.markerID for m in d.markers] == [23, 0, 1, 17, 5]
    assert [R.inverse_flag(state, m.markerID, m.featurePos, m.feature_ID) for m in d.markers] == [1, 1, 0, 0, 1]
    assert d.markers[0].featurePos == [10, 9, 8, 7, 6] and d.markers[1].feature_ID == [52, 52, 11, 54, 3, 47, 27, 24, 19, 61]


def test_oracle_equals_reference_on_other_windows(ref, test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    crop = np.ascontiguousarray(test_gray[100:900, 200:1400])
    for win in (3, 4, 7, 10):
        _oracle_equals_ref(ref.detect(crop, win, True, 5), o.detect(crop, state, fs, win, True, 5))


def test_oracle_equals_reference_on_sequence_and_rendered_frames(ref, test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = list(synth.video_sequence(test_gray, 120, 2024, first=20, count=3))
    frames += [o.bgr2gray(configs.config3_frame(i)) for i in (0, 101, 255)]
    frames += [o.bgr2gray(configs.config4_frame("2f12c", 5)[0])]
    decoded = 0
    for g in frames:
        d = ref.detect(g, 5, True, 5)
        _oracle_equals_ref(d, o.detect(g, state, fs, 5, True, 5))
        decoded += len(d.markers)
    assert decoded >= 3 * 4 + 3 + 3


@pytest.mark.parametrize("name", ["15c3f", "18c4f"])
def test_oracle_equals_reference_on_generated_codebooks(name):
    state, fs = configs.codebook(name)
    r = R.RefDetector(state=state, feature_size=fs)
    frame, truth = configs.config4_frame(name, 2, w=1920, h=1080)
    g = o.bgr2gray(frame)
    d = r.detect(g, 5, True, 5)
    _oracle_equals_ref(d, o.detect(g, state, fs, 5, True, 5))
    assert len(d.markers) >= 1 and {m.markerID for m in d.markers} <= set(truth)
    r.close()


def test_early_exits_leave_the_output_untouched(ref, marker_path):
    state, fs = o.load_marker_file(marker_path)
    flat = np.full((400, 600), 180, np.uint8)
    d = ref.detect(flat, 5, True, 5)
    assert d.status == "no_corner" and not d.markers and o.detect(flat, state, fs, 5, True, 5).status == "no_corner"
    one = flat.copy()
    one[60:120, 100:130] = 20  # one dark quad: a corner list but no pair -> "No feature detected!"
    d = ref.detect(one, 5, True, 5)
    od = o.detect(one, state, fs, 5, True, 5)
    assert d.status == od.status == "no_feature" and len(d.quads) == len(od.quads) == 1


def test_constructor_errors_are_the_reference_strings(tmp_path):
    with pytest.raises(RuntimeError, match="could not open the file"):
        R.RefDetector(marker_path=str(tmp_path / "missing.marker"))
    bad = tmp_path / "bad.marker"
    bad.write_text("1 2 2\n5 64\n")
    with pytest.raises(RuntimeError, match="must between 0 to 63"):
        R.RefDetector(marker_path=str(bad))


def test_multithreaded_batch_leg_equals_single_calls(ref, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = np.stack([configs.config3_frame(i) for i in range(4)])
    counts, markers = R.detect_batch_mt(frames, state, fs, 5, True, 5, threads=4)
    for f in range(4):
        d = ref.detect(o.bgr2gray(frames[f]), 5, True, 5)
        assert list(counts[f]) == [d.n_labels, len(d.comps), len(d.quads), len(d.feats), d.n_groups, len(d.markers),
                                   ("ok", "no_corner", "no_feature").index(d.status), int(d.flagged)]
        for k, m in enumerate(d.markers):
            assert int(markers[f, k]["marker_id"]) == m.markerID
            assert np.array_equal(markers[f, k]["corners"][:len(m.cornerLists)], m.cornerLists)


# ---- pose: pose_estimation.cpp compiled unmodified (Ceres stand-in; EPnP / undistortPoints done by cv2) -------------
def test_reference_pose_equals_pose_oracle_and_the_product_pose_stage(ref, test_gray, marker_path):
    from cylindertag_b200 import _capi as C
    state, fs = o.load_marker_file(marker_path)
    ref.load_model_camera(os.path.join(DATA, "CTag_2f12c.model"), os.path.join(DATA, "cameraParams.yml"))
    ref.detect(test_gray, 5, True, 5)
    with R.cv2_backend(only=["solve_pnp_epnp", "undistort_points"]):
        poses = ref.estimate_pose()
    with R.cv2_backend(only=["solve_pnp_epnp"]):
        poses_native_undistort = ref.estimate_pose()
    od = o.detect(test_gray, state, fs, 5, True, 5)
    models = po.load_model(os.path.join(DATA, "CTag_2f12c.model"))
    K, D = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    want = po.estimate_pose(od.markers, models, K, D)
    assert [p[0] for p in poses] == [p[0] for p in want] == [5, 0, 1, 3, 2]
    for (i, r, t), (i2, r2, t2), (j, wr, wt, rms) in zip(poses, poses_native_undistort, want):
        assert np.abs(r - wr).max() <= 1e-5 and np.abs(t - wt).max() <= 1e-5 * np.linalg.norm(wt)
        assert np.abs(r - r2).max() <= 1e-6 and np.abs(t - t2).max() <= 1e-6 * np.linalg.norm(wt)
    # Appendix E: marker ID 0
    assert np.allclose(poses[1][1], [0.3942, 0.3277, 0.3045], atol=2e-4) and np.allclose(poses[1][2], [-258.47, 108.20, 282.04], atol=0.01)
    # the product's host pose stage (csrc/pose.cpp behind ctag_estimate_pose) on the reference's corners
    lib = C.load()
    Kf, Df = np.ascontiguousarray(K, np.float32), np.ascontiguousarray(D, np.float32).reshape(-1)
    _, recs = R.detect_batch_mt(test_gray[None], state, fs, 5, True, 5, threads=1)
    for k, (idx, wr, wt) in enumerate(poses):
        mc = np.ascontiguousarray(models[idx][3], np.float32)
        rv, tv, rms = np.zeros(3), np.zeros(3), ctypes.c_double()
        rec = recs[0, k:k + 1].copy()
        rc = lib.ctag_estimate_pose(vp(rec), vp(mc), len(mc), vp(Kf), vp(Df), len(Df), vp(rv), vp(tv), ctypes.byref(rms))
        assert rc == 0
        assert np.abs(rv - wr).max() <= 1e-4 and np.abs(tv - wt).max() <= 1e-4 * np.linalg.norm(wt)


# ---- the frozen goldens are what the compiled reference produces ------------------------------------------------------
def _golden_frame_equals(gold, f, d, state):
    assert list(gold["counts"][f]) == [d.n_labels, len(d.comps), len(d.quads), len(d.feats), d.n_groups, len(d.markers),
                                       ("ok", "no_corner", "no_feature").index(d.status), int(d.flagged)]
    import zlib
    assert int(gold["binary_crc"][f]) == zlib.crc32(np.ascontiguousarray(d.binary).tobytes())
    assert np.array_equal(gold["comps"][gold["comp_start"][f]:gold["comp_start"][f + 1]], d.comps)
    assert np.array_equal(gold["quads"][gold["quad_start"][f]:gold["quad_start"][f + 1]], d.quads)
    assert np.array_equal(gold["feats"][gold["feat_start"][f]:gold["feat_start"][f + 1]], d.feats)
    a = int(gold["marker_start"][f])
    assert int(gold["marker_start"][f + 1]) - a == len(d.markers)
    for k, m in enumerate(d.markers):
        n = len(m.cornerLists)
        assert int(gold["marker_id"][a + k]) == m.markerID and int(gold["n_features"][a + k]) == n
        assert int(gold["inverse"][a + k]) == R.inverse_flag(state, m.markerID, m.featurePos, m.feature_ID)
        assert list(gold["feature_pos"][a + k][:len(m.featurePos)]) == m.featurePos
        assert list(gold["feature_id"][a + k][:n]) == m.feature_ID
        assert np.array_equal(gold["corners"][a + k][:n], m.cornerLists[:20])


def test_frozen_goldens_are_reproduced_by_the_compiled_reference(ref, test_gray, marker_path):
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    state, fs = o.load_marker_file(marker_path)
    _golden_frame_equals(np.load(os.path.join(gdir, "ref_testbmp.npz")), 0, ref.detect(test_gray, 5, True, 5), state)
    gold = np.load(os.path.join(gdir, "ref_sequence.npz"))
    for f, g in enumerate(synth.video_sequence(test_gray, 120, 2024, first=0, count=4)):
        _golden_frame_equals(gold, f, ref.detect(g, 5, True, 5), state)
    gold = np.load(os.path.join(gdir, "ref_config3.npz"))
    for f in (0, 128):
        _golden_frame_equals(gold, f, ref.detect(o.bgr2gray(configs.config3_frame(f)), 5, True, 5), state)
    for name in ("15c3f",):
        st, f_s = configs.codebook(name)
        r = R.RefDetector(state=st, feature_size=f_s)
        gold = np.load(os.path.join(gdir, f"ref_config4_{name}.npz"))
        _golden_frame_equals(gold, 1, r.detect(o.bgr2gray(configs.config4_frame(name, 1)[0]), 5, True, 5), st)
        r.close()


def test_oracle_sequence_golden_equals_reference_golden():
    """The cv2-oracle results frozen in round 1 (sequence_detect.npz) equal the reference's (ref_sequence.npz) on all
    120 frames: counts, IDs, positions exactly; corners to 1e-6 px."""
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    a, b = np.load(os.path.join(gdir, "sequence_detect.npz")), np.load(os.path.join(gdir, "ref_sequence.npz"))
    assert np.array_equal(a["counts"], b["counts"][:, :6]) and np.array_equal(a["marker_start"], b["marker_start"])
    for k in ("marker_id", "inverse", "n_features", "feature_pos", "feature_id", "id_left", "id_right"):
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(a["corners"] - b["corners"]).max() <= 1e-6
