"""BASELINE.json's full size (4K BGR frames, six markers each): the CUDA path against the reference's own code
(oracle/_ref: corner_detector.cpp + CylinderTag.cpp compiled unmodified, run here on the host cores through
ref_detect_batch_mt) on every frame of a multi-chunk batch, plus size-independent properties: a frame's result does not
depend on its position in the batch or on the batch around it."""
import numpy as np
import pytest

from cylindertag_b200 import synth
from oracle import ctag_oracle as o
from oracle import ref_api as cpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames_4k(marker_path):
    state, fs = o.load_marker_file(marker_path)
    return state, fs, np.stack([synth.synthetic_frame(2100 + i, 3840, 2160, state, 6, channels=3)[0] for i in range(8)])


def test_4k_bgr_batch_matches_compiled_reference(detector, frames_4k):
    state, fs, frames = frames_4k
    batch = np.concatenate([frames, frames[::-1]])  # 16 frames: chunked over all workspaces, each frame twice
    markers, counts, info = detector.detect_batch(batch, 5, True, 5, cap_per_frame=32)
    ref_counts, ref_markers = cpu.detect_batch(frames, state, fs, True, 5, threads=8, cap=32)
    total = 0
    for f in range(16):
        r = f if f < 8 else 15 - f
        assert [int(info[k][f]) for k in ("n_labels", "n_legal", "n_quads", "n_features", "n_groups", "n_markers")] == \
            list(ref_counts[r][:6]), f"frame {f}"
        n = int(counts[f])
        assert n == int(ref_counts[r][5])
        total += n
        for k in range(n):
            g, w = markers[f][k], ref_markers[r][k]
            for field in ("marker_id", "n_features", "inverse"):
                assert int(g[field]) == int(w[field]), (f, k, field)
            nf = int(g["n_features"])
            for field in ("feature_pos", "feature_id", "id_left", "id_right"):
                assert list(g[field][:nf]) == list(w[field][:nf]), (f, k, field)
            assert np.abs(g["corners"][:nf] - w["corners"][:nf]).max() <= 1e-3, (f, k)
    assert total >= 40  # most of the 6 x 16 rendered markers decode
    # position independence: the mirrored half of the batch repeats the first half exactly
    for f in range(8):
        assert int(counts[f]) == int(counts[15 - f])
        a, b = markers[f][:int(counts[f])].copy(), markers[15 - f][:int(counts[f])].copy()
        a["frame"], b["frame"] = 0, 0
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"frame {f} vs {15 - f}"
