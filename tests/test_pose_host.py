"""Host-side pose stage (ctag_estimate_pose, csrc/pose.cpp) against the pose oracle (cv2 EPnP + scipy LM) -- CPU only:
the function does no GPU work, so it is called through the C ABI without a device."""
import ctypes
import os

import numpy as np
import pytest

from cylindertag_b200 import _capi as C
from cylindertag_b200.api import CamInfo, MarkerInfo, ModelInfo, marker_to_record, pnp_solver, select_pose_points
from oracle import ctag_oracle as o
from oracle import pose_oracle as po

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")

ROT_TOL = 1e-4   # rad   
TRANS_TOL = 1e-4  # relative to |t|


def _as_marker_info(m):
    return MarkerInfo(markerID=m.markerID, featurePos=list(m.featurePos), feature_ID=list(m.feature_ID),
                      feature_ID_left=list(m.feature_ID_left), feature_ID_right=list(m.feature_ID_right),
                      cornerLists=[np.asarray(c, np.float32).reshape(8, 2) for c in m.cornerLists])


def _models():
    raw = po.load_model(os.path.join(DATA, "CTag_2f12c.model"))
    return raw, [ModelInfo(MarkerID=m[0], base=m[1], axis=m[2], corners=m[3]) for m in raw]


def test_point_selection_matches_reference_rule(test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    d = o.detect(test_gray, state, fs, 5, True, 5)
    raw, models = _models()
    lib = C.load()
    for m in d.markers:
        mk = _as_marker_info(m)
        idx = next((j for j, mm in enumerate(models) if mm.MarkerID == mk.markerID), -1)
        if idx < 0:
            continue
        ip, op = select_pose_points(mk, models[idx])
        rec = marker_to_record(mk)
        f = np.zeros(8 * C.MAX_FEATURES, np.int32)
        k = np.zeros(8 * C.MAX_FEATURES, np.int32)
        n = lib.ctag_pose_select_points(rec.ctypes.data, f.ctypes.data, k.ctypes.data, f.size)
        assert n == len(ip)
        got = np.array([mk.cornerLists[f[i]][k[i]] for i in range(n)], np.float32)
        assert np.array_equal(got, ip)


def test_testbmp_poses_match_oracle(test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    d = o.detect(test_gray, state, fs, 5, True, 5)
    raw, models = _models()
    K, D = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    want = po.estimate_pose(d.markers, raw, K, D)
    cam = CamInfo(K, D)
    got = [pnp_solver(_as_marker_info(m), models, cam) for m in d.markers]
    got = [p for p in got if p.markerID != -1]
    assert [p.markerID for p in got] == [w[0] for w in want]
    for p, w in zip(got, want):
        assert np.abs(p.rvec - w[1]).max() < ROT_TOL, (p.rvec, w[1])
        assert np.linalg.norm(p.tvec - w[2]) < TRANS_TOL * np.linalg.norm(w[2]), (p.tvec, w[2])
        assert abs(p.rms_px - w[3]) < 1e-6


def _synthetic_marker(rng, corners3, K, D, n_feat, first):
    """Projects model corners of n_feat consecutive features with a random pose (+ pixel noise)."""
    import cv2
    rvec = rng.uniform(-0.5, 0.5, 3)
    cen = corners3[first * 8:(first + n_feat) * 8].mean(0)
    R, _ = cv2.Rodrigues(rvec)
    tvec = np.array([rng.uniform(-40, 40), rng.uniform(-30, 30), rng.uniform(250, 450)]) - R @ cen
    mk = MarkerInfo(markerID=0)
    for j in range(n_feat):
        pts3 = corners3[(first + j) * 8:(first + j + 1) * 8].astype(np.float64)
        ip, _ = cv2.projectPoints(pts3, rvec, tvec, K.astype(np.float64), D.astype(np.float64))
        mk.cornerLists.append((ip.reshape(8, 2) + rng.normal(0, 0.3, (8, 2))).astype(np.float32))
        mk.featurePos.append(first + j)
        idl = int(rng.integers(0, 8))
        mk.feature_ID_left.append(idl)
        mk.feature_ID_right.append(int(idl + rng.integers(-3, 4)) if rng.random() > 0.1 else -1)
        mk.feature_ID.append(idl * 8)
    return mk


@pytest.mark.parametrize("seed", range(12))
def test_random_poses_match_oracle(seed):
    rng = np.random.default_rng(seed)
    raw, models = _models()
    K, D = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    cam = CamInfo(K, D)
    n_feat = int(rng.integers(2, 7))
    first = int(rng.integers(0, 12 - n_feat))
    mk = _synthetic_marker(rng, raw[0][3], K, D, n_feat, first)

    class M:  # oracle-side view of the same marker
        pass
    om = M()
    om.markerID, om.cornerLists, om.featurePos = 0, mk.cornerLists, mk.featurePos
    om.feature_ID_left, om.feature_ID_right = mk.feature_ID_left, mk.feature_ID_right
    want = po.estimate_pose([om], raw, K, D)
    got = pnp_solver(mk, models, cam)
    if not want:
        assert got.markerID == -1
        return
    assert got.markerID == want[0][0]
    assert np.abs(got.rvec - want[0][1]).max() < ROT_TOL, (got.rvec, want[0][1])
    assert np.linalg.norm(got.tvec - want[0][2]) < TRANS_TOL * np.linalg.norm(want[0][2])


def test_too_few_points_is_an_error():
    lib = C.load()
    rec = np.zeros((), C.MARKER_DTYPE)
    corners3 = np.zeros((96, 3), np.float32)
    K = np.eye(3, dtype=np.float32).reshape(9)
    r, t = np.zeros(3), np.zeros(3)
    rc = lib.ctag_estimate_pose(rec.ctypes.data, corners3.ctypes.data, 96, K.ctypes.data, None, 0, r.ctypes.data,
                                t.ctypes.data, None)
    assert rc == C.ERR_ARG
