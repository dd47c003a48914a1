"""examples/main_image.cpp = the image branch of the reference's main.cpp on the C++ mirror (gpu): BMP in, markers and
poses out, compared with the oracle's known answers for test.bmp."""
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import ctag_oracle as o
from oracle import pose_oracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "golden", "data")


def test_main_image_flow(tmp_path, test_gray, marker_path):
    exe = tmp_path / "main_image"
    libdir = os.path.join(ROOT, "cylindertag_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "main_image.cpp"),
                    "-o", str(exe), "-L", libdir, "-lctag_b200", f"-Wl,-rpath,{libdir}"], check=True)
    bmp = tmp_path / "test.bmp"
    cv2.imwrite(str(bmp), cv2.cvtColor(test_gray, cv2.COLOR_GRAY2BGR))  # 24-bit like the reference's test.bmp
    ppm = tmp_path / "overlay.ppm"
    out = subprocess.run([str(exe), str(bmp), marker_path, os.path.join(DATA, "CTag_2f12c.model"),
                          os.path.join(DATA, "cameraParams.yml"), str(ppm)], capture_output=True, text=True, check=True).stdout
    lines = out.strip().splitlines()
    assert lines[0] == "markers 5 poses 5"
    assert [int(l.split()[1]) for l in lines if l.startswith("marker ")] == [23, 0, 1, 17, 5]
    state, fs = o.load_marker_file(marker_path)
    d = o.detect(test_gray, state, fs, 5, True, 5)
    want = po.estimate_pose(d.markers, po.load_model(os.path.join(DATA, "CTag_2f12c.model")),
                            *po.load_camera(os.path.join(DATA, "cameraParams.yml")))
    got = [l.split() for l in lines if l.startswith("pose ")]
    assert [int(g[2]) for g in got] == [w[0] for w in want]
    for g, w in zip(got, want):
        rvec = np.array([float(v) for v in g[6:9]])
        tvec = np.array([float(v) for v in g[10:13]])
        assert np.abs(rvec - w[1]).max() < 2e-4            # printed with 5 decimals
        assert np.linalg.norm(tvec - w[2]) < 2e-4 * np.linalg.norm(w[2]) + 1e-3
    # drawAxis(..., 30) of the demo (main.cpp:41) went to a file: same picture as the Python mirror draws from its own
    # detections and poses (both call ctag_draw_axis; poses agree to ~1e-6, so at most a few edge pixels may differ)
    from cylindertag_b200 import CylinderTag
    overlay = cv2.imread(str(ppm), cv2.IMREAD_UNCHANGED)[..., ::-1]  # P6 is read as RGB -> BGR by OpenCV: undo
    tag = CylinderTag(marker_path)
    models = tag.loadModel(os.path.join(DATA, "CTag_2f12c.model"))
    cam = tag.loadCamera(os.path.join(DATA, "cameraParams.yml"))
    markers = []
    tag.detect(test_gray, markers, 5, True, 5)
    poses = tag.estimatePose(test_gray, markers, models, cam, False)
    mine = tag.drawAxis(test_gray, markers, models, poses, cam, 30)
    assert overlay.shape == mine.shape
    assert (np.any(overlay != mine, axis=2)).sum() <= 200
    assert np.any(mine != test_gray[..., None], axis=2).sum() > 20000


def test_main_video_flow(tmp_path, test_gray, marker_path):
    """examples/main_video.cpp = the video branch of main.cpp (:43-60) on an image sequence: six frames of the config-2
    substitute sequence, frame loop + the batched call, IDs and poses against the oracles."""
    from cylindertag_b200 import synth
    exe = tmp_path / "main_video"
    libdir = os.path.join(ROOT, "cylindertag_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "main_video.cpp"),
                    "-o", str(exe), "-L", libdir, "-lctag_b200", f"-Wl,-rpath,{libdir}"], check=True)
    frames = synth.video_sequence(test_gray, 120, 2024, first=10, count=6)
    paths = []
    for i, fr in enumerate(frames):
        paths.append(str(tmp_path / f"f{i}.pgm"))
        cv2.imwrite(paths[-1], fr)
    out = subprocess.run([str(exe), marker_path, os.path.join(DATA, "CTag_2f12c.model"), os.path.join(DATA, "cameraParams.yml")] + paths,
                         capture_output=True, text=True, check=True).stdout
    lines = out.strip().splitlines()
    assert lines[-1] == "batched call equals the frame loop on 6 of 6 frames"
    state, fs = o.load_marker_file(marker_path)
    models = po.load_model(os.path.join(DATA, "CTag_2f12c.model"))
    cam = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    for i, fr in enumerate(frames):
        d = o.detect(fr, state, fs, 5, True, 5)
        head = next(l for l in lines if l.startswith(f"frame {i} markers"))
        assert [int(v) for v in head.split("ids")[1].split()] == [m.markerID for m in d.markers]
        want = po.estimate_pose(d.markers, models, *cam)
        got = [l.split() for l in lines if l.startswith(f"frame {i} pose")]
        assert [int(g[4]) for g in got] == [w[0] for w in want]
        for g, w in zip(got, want):
            rvec = np.array([float(v) for v in g[6:9]])
            tvec = np.array([float(v) for v in g[10:13]])
            assert np.abs(rvec - w[1]).max() < 2e-4
            assert np.linalg.norm(tvec - w[2]) < 2e-4 * np.linalg.norm(w[2]) + 1e-3
