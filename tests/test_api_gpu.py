"""The reference-facing class (cylindertag_b200.CylinderTag) on the GPU: call sequence of main.cpp:31-41 and pose parity."""
import os

import numpy as np
import pytest

from oracle import ctag_oracle as o
from oracle import pose_oracle as po

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


def test_reference_call_sequence_and_pose_parity(test_gray, marker_path):
    from cylindertag_b200 import CylinderTag
    tag = CylinderTag(marker_path)
    models = tag.loadModel(os.path.join(DATA, "CTag_2f12c.model"))
    cam = tag.loadCamera(os.path.join(DATA, "cameraParams.yml"))
    markers = []
    tag.detect(test_gray, markers, 5, True, 5)
    assert [m.markerID for m in markers] == [23, 0, 1, 17, 5]
    poses = tag.estimatePose(test_gray, markers, models, cam, False)
    assert len(poses) == 5
    # oracle: detect + pose on the CPU
    state, fs = o.load_marker_file(marker_path)
    dump = o.detect(test_gray, state, fs, 5, True, 5)
    ref = po.estimate_pose(dump.markers, po.load_model(os.path.join(DATA, "CTag_2f12c.model")), *po.load_camera(os.path.join(DATA, "cameraParams.yml")))
    for p, (idx, r, t, rms) in zip(poses, ref):
        assert p.markerID == idx
        assert np.abs(p.rvec - r).max() <= 1e-4, (p.rvec, r)                      # 1e-4 rad
        assert np.abs(p.tvec - t).max() <= 1e-4 * np.linalg.norm(t), (p.tvec, t)  # 1e-4 of translation scale
    overlay = tag.drawAxis(test_gray, markers, models, poses, cam, 30)
    assert overlay.shape == test_gray.shape + (3,)


def test_early_exit_leaves_output_untouched(marker_path, capsys):
    from cylindertag_b200 import CylinderTag
    tag = CylinderTag(marker_path)
    sentinel = ["untouched"]
    out = tag.detect(np.full((240, 320), 128, np.uint8), sentinel, 5, True, 5)
    assert out is None and sentinel == ["untouched"]
    assert "No corner detected!" in capsys.readouterr().out


def test_constructor_errors(tmp_path):
    from cylindertag_b200 import CylinderTag
    with pytest.raises(RuntimeError, match="could not open the file"):
        CylinderTag(str(tmp_path / "nope.marker"))
    with pytest.raises(RuntimeError, match="between 0 to 63"):
        CylinderTag(np.full((2, 3), 99, np.int32), 2)
