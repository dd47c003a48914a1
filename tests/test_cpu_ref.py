"""The C++ CPU restatement (bench baseline) against the Python/cv2 oracle -- CPU only."""
import numpy as np

from cylindertag_b200 import synth
from oracle import ctag_oracle as o
from oracle.cpu_ref import api as cpu
from tests.parity import assert_markers_match


def test_testbmp_matches_oracle(test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    d = o.detect(test_gray, state, fs, 5, True, 5)
    counts, markers = cpu.detect_batch(test_gray[None], state, fs, True, 5)
    assert list(counts[0][:6]) == [d.n_labels, len(d.comps), len(d.quads), len(d.feats), len(d.groups), len(d.markers)]
    assert_markers_match(markers[0], int(counts[0][5]), d.markers, tol=1e-6)


def test_synthetic_bgr_frames_match_oracle_multithreaded(marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = np.stack([synth.synthetic_frame(1000 + i, 1920, 1080, state, 1, channels=3)[0] for i in range(3)])
    counts, markers = cpu.detect_batch(frames, state, fs, True, 5, threads=3)
    for f in range(3):
        d = o.detect(o.bgr2gray(frames[f]), state, fs, 5, True, 5)
        assert list(counts[f][:6]) == [d.n_labels, len(d.comps), len(d.quads), len(d.feats), len(d.groups), len(d.markers)]
        assert_markers_match(markers[f], int(counts[f][5]), d.markers, tol=1e-6)
