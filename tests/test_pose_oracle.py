"""Pose oracle known answers -- CPU only."""
import os

import numpy as np

from oracle import ctag_oracle as o
from oracle import pose_oracle as po

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


def test_testbmp_pose_known_answers(test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    d = o.detect(test_gray, state, fs, 5, True, 5)
    models = po.load_model(os.path.join(DATA, "CTag_2f12c.model"))
    K, D = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    assert [m[0] for m in models] == [0, 1, 5, 17, 21, 23]
    poses = po.estimate_pose(d.markers, models, K, D)
    assert [p[0] for p in poses] == [5, 0, 1, 3, 2]  # model INDEX per marker (pose_estimation.cpp:59,69)
    # marker ID 0: rvec (0.3942, 0.3277, 0.3045), tvec (-258.47, 108.20, 282.04)
    assert np.allclose(poses[1][1], [0.3942, 0.3277, 0.3045], atol=2e-3)
    assert np.allclose(poses[1][2], [-258.47, 108.20, 282.04], atol=0.1)
    assert all(p[3] < 0.6 for p in poses)  # sub-pixel reprojection against the shipped reconstruction
