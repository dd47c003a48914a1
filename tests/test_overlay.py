"""drawAxis overlay (ctag_project_points / ctag_gray_to_3ch / ctag_draw_axis, csrc/overlay.cpp) against
the OpenCV calls the reference makes (CylinderTag.cpp:211-246: cvtColor, projectPoints, circle, arrowedLine) -- CPU
only, host code through the C ABI.  projectPoints must agree to float rounding; the rasterised overlay may differ from
OpenCV's on edge pixels only."""
import os

import cv2
import numpy as np

from cylindertag_b200 import _capi as C
from cylindertag_b200.api import CamInfo, CylinderTag, MarkerInfo, ModelInfo, PoseInfo
from oracle import ctag_oracle as o
from oracle import pose_oracle as po

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


def _cv2_overlay(img, markers, models, poses, camera, axis_length):
    """The reference's drawAxis body with cv2 (Point2f -> Point conversion rounds like cvRound)."""
    out = cv2.cvtColor(np.asarray(img), cv2.COLOR_GRAY2RGB)
    pts_all = []
    for i, pose in enumerate(poses):
        model = models[pose.markerID]
        pts3 = [model.corners[markers[i].featurePos[j] * 8 + k] for j in range(len(markers[i].cornerLists)) for k in range(8)]
        base = model.base.astype(np.float32)
        L = np.float32(axis_length)
        pts3 += [base, base + model.axis.astype(np.float32) * L, base + np.array([0.0372, 0.0372, 0.9986], np.float32) * L,
                 base + np.array([0.9980, -0.0520, -0.0353], np.float32) * L]
        ip, _ = cv2.projectPoints(np.array(pts3, np.float64), pose.rvec, pose.tvec, camera.Intrinsic.astype(np.float64),
                                  camera.distCoeffs.astype(np.float64))
        ip = ip.reshape(-1, 2).astype(np.float32)
        pts_all.append((np.array(pts3, np.float32), ip))
        rnd = lambda p: (int(np.rint(p[0])), int(np.rint(p[1])))
        for p in ip[:-5]:
            cv2.circle(out, rnd(p), 5, (255, 234, 32), -1)
        for k, col in ((-3, (255, 0, 0)), (-2, (0, 255, 0)), (-1, (0, 0, 255))):
            cv2.arrowedLine(out, rnd(ip[-4]), rnd(ip[k]), col, 10, cv2.LINE_AA, 0, 0.2)
        cv2.circle(out, rnd(ip[-4]), 8, (247, 235, 235), -1)
    return out, pts_all


def _scene(test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    d = o.detect(test_gray, state, fs, 5, True, 5)
    raw = po.load_model(os.path.join(DATA, "CTag_2f12c.model"))
    models = [ModelInfo(MarkerID=m[0], base=m[1], axis=m[2], corners=m[3]) for m in raw]
    K, D = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    cam = CamInfo(np.asarray(K, np.float32), np.asarray(D, np.float32))
    ref = po.estimate_pose(d.markers, raw, K, D)
    markers = [MarkerInfo(markerID=m.markerID, featurePos=list(m.featurePos), feature_ID=list(m.feature_ID),
                          feature_ID_left=list(m.feature_ID_left), feature_ID_right=list(m.feature_ID_right),
                          cornerLists=[np.asarray(c, np.float32).reshape(8, 2) for c in m.cornerLists]) for m in d.markers]
    poses = [PoseInfo(idx, np.asarray(r, np.float64).reshape(3), np.asarray(t, np.float64).reshape(3)) for idx, r, t, _ in ref]
    assert len(poses) == len(markers) == 5
    return markers, models, poses, cam


def test_project_points_matches_cv2(test_gray, marker_path):
    markers, models, poses, cam = _scene(test_gray, marker_path)
    _, pts_all = _cv2_overlay(test_gray, markers, models, poses, cam, 30)
    lib = C.load()
    K = np.ascontiguousarray(cam.Intrinsic, np.float32).reshape(9)
    D = np.ascontiguousarray(cam.distCoeffs, np.float32).reshape(-1)
    for pose, (p3, want) in zip(poses, pts_all):
        got = np.zeros((len(p3), 2), np.float32)
        p3 = np.ascontiguousarray(p3, np.float32)
        assert lib.ctag_project_points(p3.ctypes.data, len(p3), pose.rvec.ctypes.data, pose.tvec.ctypes.data, K.ctypes.data,
                                       D.ctypes.data, D.size, got.ctypes.data) == C.OK
        assert np.abs(got - want).max() <= 2e-3, np.abs(got - want).max()  # float32 pixels around 1e3


def test_overlay_agrees_with_opencv_drawing(test_gray, marker_path):
    markers, models, poses, cam = _scene(test_gray, marker_path)
    want, _ = _cv2_overlay(test_gray, markers, models, poses, cam, 30)
    tag = CylinderTag.__new__(CylinderTag)  # drawAxis is host code: no detector (and no GPU) needed
    got = tag.drawAxis(test_gray, markers, models, poses, cam, 30)
    assert got.shape == want.shape == test_gray.shape + (3,) and got.dtype == np.uint8
    drawn_want = np.any(want != test_gray[..., None], axis=2)
    drawn_got = np.any(got != test_gray[..., None], axis=2)
    assert drawn_want.sum() > 20000
    # the two masks differ on outline pixels only
    diff = drawn_want ^ drawn_got
    assert diff.sum() <= 0.06 * drawn_want.sum(), (int(diff.sum()), int(drawn_want.sum()))
    # interiors carry exactly the reference's colours
    core = cv2.erode((drawn_want & drawn_got).astype(np.uint8), np.ones((5, 5), np.uint8)).astype(bool)
    assert core.sum() > 8000
    assert (np.abs(got[core].astype(int) - want[core].astype(int)).max(axis=1) > 0).mean() <= 0.01
    # untouched pixels are the gray image on all three channels
    assert np.array_equal(got[~drawn_got], np.repeat(test_gray[~drawn_got][:, None], 3, axis=1))


def test_overlay_argument_checks(test_gray):
    lib = C.load()
    out = np.zeros((4, 4, 3), np.uint8)
    assert lib.ctag_gray_to_3ch(None, 4, 4, 4, out.ctypes.data, 12) == C.ERR_ARG
    g = np.arange(16, dtype=np.uint8).reshape(4, 4)
    assert lib.ctag_gray_to_3ch(g.ctypes.data, 4, 4, 4, out.ctypes.data, 8) == C.ERR_ARG  # output pitch too small
    assert lib.ctag_gray_to_3ch(g.ctypes.data, 4, 4, 4, out.ctypes.data, 12) == C.OK
    assert np.array_equal(out[..., 1], g)
