"""BASELINE.json config 2 ("test.avi full video, frame-by-frame detect + pose vs reference per frame").  test.avi is not
shipped with the reference, so the sequence is the documented substitute: 120 frames made from test.bmp
by seeded homographies / blur / noise (cylindertag_b200.synth.video_sequence, default_rng(2024)) -- said here and in
DESIGN.md.  The caller loop is main.cpp:48-60: per frame detect(gray, 5, true, 5) -> estimatePose, outputs cleared
between frames.

CPU part: the generator is deterministic; the compiled reference (oracle/_ref) agrees with the cv2 oracle on sequence
frames (live on three frames, against the frozen oracle results of tests/golden/sequence_detect.npz on all 120).
GPU part: every frame of the sequence against the frozen results (tests/golden/ref_sequence.npz, made by the compiled
reference, and sequence_detect.npz, made by the cv2 oracle) and against the compiled reference run live, a subset stage
by stage against the live cv2 oracle, poses against the pose oracle."""
import os

import numpy as np
import pytest

from cylindertag_b200 import synth
from oracle import ctag_oracle as o
from oracle import pose_oracle as po
from oracle import ref_api as cpu
from tests.parity import assert_frame_matches, assert_markers_match

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
N_FRAMES, SEED = 120, 2024
GOLDEN_SUMS = [172982, 573768, 835821]  # sum of pixels mod 1000003 of frames 0..2 (cv2 4.13)
COUNT_KEYS = ("n_labels", "n_legal", "n_quads", "n_features", "n_groups", "n_markers")


@pytest.fixture(scope="module")
def dictionary(marker_path):
    return o.load_marker_file(marker_path)


@pytest.fixture(scope="module")
def golden_sequence():
    """cv2-oracle results on all 120 frames (tests/golden/make_golden_sequence.py)."""
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sequence_detect.npz"))


def _assert_matches_golden(gold, counts6, n_markers, markers, tol, ctx):
    """counts6: [frame][6] stage counts, markers[f][k]: ctag_marker records of the implementation under test."""
    total = 0
    for f in range(N_FRAMES):
        if gold["flagged"][f]:
            continue
        assert list(counts6[f]) == list(gold["counts"][f]), f"{ctx} frame {f}"
        a, b = int(gold["marker_start"][f]), int(gold["marker_start"][f + 1])
        assert int(n_markers[f]) == b - a, f"{ctx} frame {f}"
        for k in range(b - a):
            g = markers[f][k]
            nf = int(gold["n_features"][a + k])
            assert int(g["marker_id"]) == int(gold["marker_id"][a + k]) and int(g["inverse"]) == int(gold["inverse"][a + k])
            assert int(g["n_features"]) == nf, (ctx, f, k)
            for name in ("feature_pos", "feature_id", "id_left", "id_right"):
                npos = nf if name != "feature_pos" else int((gold["feature_pos"][a + k] >= 0).sum())
                assert list(g[name][:npos]) == list(gold[name][a + k][:npos]), (ctx, f, k, name)
            assert np.abs(g["corners"][:nf] - gold["corners"][a + k][:nf]).max() <= tol, (ctx, f, k)
            total += 1
    return total


def test_compiled_reference_matches_oracle_golden_on_all_frames(test_gray, dictionary, golden_sequence):
    """Every frame of the sequence: the reference's own code against the frozen cv2-oracle results."""
    state, fs = dictionary
    seq = synth.video_sequence(test_gray, N_FRAMES, SEED)
    counts, markers = cpu.detect_batch(seq, state, fs, True, 5, threads=os.cpu_count() or 1, cap=32)
    total = _assert_matches_golden(golden_sequence, counts[:, :6], counts[:, 5], markers, 1e-6, "oracle/_ref")
    assert total == int(golden_sequence["marker_start"][-1]) >= 4 * N_FRAMES


def test_sequence_is_deterministic_and_sliceable(test_gray):
    a = synth.video_sequence(test_gray, N_FRAMES, SEED, first=0, count=3)
    b = synth.video_sequence(test_gray, N_FRAMES, SEED, first=1, count=2)
    assert a.shape == (3,) + test_gray.shape and a.dtype == np.uint8
    assert np.array_equal(a[1:], b)
    assert not np.array_equal(a[0], a[1])
    # frozen checksum of the first frames: the sequence every report quotes does not drift with the code
    assert [int(f.astype(np.uint64).sum() % 1000003) for f in a] == GOLDEN_SUMS


def test_compiled_reference_matches_oracle_on_sequence_frames(test_gray, dictionary):
    state, fs = dictionary
    frames = synth.video_sequence(test_gray, N_FRAMES, SEED, first=0, count=3)
    counts, markers = cpu.detect_batch(frames, state, fs, True, 5, threads=3)
    for f in range(len(frames)):
        d = o.detect(frames[f], state, fs, 5, True, 5)
        assert list(counts[f][:6]) == [d.n_labels, len(d.comps), len(d.quads), len(d.feats), len(d.groups), len(d.markers)]
        assert_markers_match(markers[f], int(counts[f][5]), d.markers, tol=1e-6)
        assert {m.markerID for m in d.markers} <= {0, 1, 5, 17, 21, 23}  # the shipped .model's IDs


@pytest.mark.gpu
def test_full_sequence_matches_compiled_reference(detector, test_gray, dictionary):
    state, fs = dictionary
    seq = synth.video_sequence(test_gray, N_FRAMES, SEED)
    markers, counts, info = detector.detect_batch(seq, 5, True, 5, cap_per_frame=32)
    ref_counts, ref_markers = cpu.detect_batch(seq, state, fs, True, 5, threads=os.cpu_count() or 1, cap=32)
    decoded = 0
    for f in range(N_FRAMES):
        assert [int(info[k][f]) for k in COUNT_KEYS] == list(ref_counts[f][:6]), f"frame {f}"
        n = int(counts[f])
        decoded += n
        for k in range(n):
            g, w = markers[f][k], ref_markers[f][k]
            for name in ("marker_id", "n_features", "inverse"):
                assert int(g[name]) == int(w[name]), (f, k, name)
            nf = int(g["n_features"])
            for name in ("feature_pos", "feature_id", "id_left", "id_right"):
                assert list(g[name][:nf]) == list(w[name][:nf]), (f, k, name)
            assert np.abs(g["corners"][:nf] - w["corners"][:nf]).max() <= 1e-3, (f, k)
    assert decoded >= 4 * N_FRAMES  # five markers in the photo, a warped frame loses one now and then


@pytest.mark.gpu
def test_full_sequence_matches_golden(detector, test_gray, golden_sequence):
    """Every frame of the sequence: the CUDA path against the frozen cv2-oracle results."""
    seq = synth.video_sequence(test_gray, N_FRAMES, SEED)
    markers, counts, info = detector.detect_batch(seq, 5, True, 5, cap_per_frame=32)
    counts6 = np.stack([info[k] for k in COUNT_KEYS], axis=1)
    total = _assert_matches_golden(golden_sequence, counts6, counts, markers, 1e-3, "gpu")
    assert total == int(golden_sequence["marker_start"][-1])


@pytest.mark.gpu
def test_sequence_frames_stage_parity_with_cv2_oracle(detector, test_gray, dictionary):
    state, fs = dictionary
    frames = synth.video_sequence(test_gray, N_FRAMES, SEED, first=40, count=4)
    markers, counts, info = detector.detect_batch(frames, 5, True, 5, cap_per_frame=32)
    for f in range(len(frames)):
        dump = o.detect(frames[f], state, fs, 5, True, 5)
        worst = assert_frame_matches(detector, f, info, markers, counts, dump, True, ctx=f"sequence frame {40 + f}")
        assert worst <= 1e-3


@pytest.mark.gpu
def test_video_loop_detect_and_pose_per_frame(test_gray, marker_path, dictionary):
    """main.cpp:48-60 on the mirror class, one frame at a time; poses against the pose oracle fed with oracle corners."""
    from cylindertag_b200 import CylinderTag
    state, fs = dictionary
    tag = CylinderTag(marker_path)
    models = tag.loadModel(os.path.join(DATA, "CTag_2f12c.model"))
    cam = tag.loadCamera(os.path.join(DATA, "cameraParams.yml"))
    ref_models = po.load_model(os.path.join(DATA, "CTag_2f12c.model"))
    ref_cam = po.load_camera(os.path.join(DATA, "cameraParams.yml"))
    frames = synth.video_sequence(test_gray, N_FRAMES, SEED, first=0, count=5)
    checked = 0
    for fr in frames:
        markers = []          # main.cpp:55-56 clears both lists between frames
        tag.detect(fr, markers, 5, True, 5)
        poses = tag.estimatePose(fr, markers, models, cam, False)
        dump = o.detect(fr, state, fs, 5, True, 5)
        assert [m.markerID for m in markers] == [m.markerID for m in dump.markers]
        ref = po.estimate_pose(dump.markers, ref_models, *ref_cam)
        assert len(poses) == len(ref)
        for p, (idx, r, t, rms) in zip(poses, ref):
            assert p.markerID == idx
            assert np.abs(p.rvec - r).max() <= 1e-4, (p.rvec, r)
            assert np.abs(p.tvec - t).max() <= 1e-4 * np.linalg.norm(t), (p.tvec, t)
            checked += 1
    assert checked >= 15
