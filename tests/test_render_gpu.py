"""The CUDA renderer (ctag_render_frames): frames rendered straight into device memory decode to the rendered
dictionary rows, and the detections on them equal the reference's own code (oracle/_ref) on the same pixels."""
import numpy as np
import pytest
import torch

from cylindertag_b200 import Detector, synth
from oracle import ctag_oracle as o
from oracle import ref_api as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,channels,markers", [(1920, 1080, 3, 2), (3840, 2160, 1, 5)])
def test_rendered_frames_decode_and_match_the_reference(marker_path, w, h, channels, markers):
    state, fs = o.load_marker_file(marker_path)
    det = Detector(state=state, feature_size=fs)
    n = 6
    shape = (n, h, w, 3) if channels == 3 else (n, h, w)
    frames = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    truth = synth.render_frames_gpu(det, frames.data_ptr(), range(5000, 5000 + n), w, h, markers, channels)
    det.enqueue_device(frames.data_ptr(), n, w, h, w * channels, w * h * channels, channels, 5, True, 5)
    m, c, info = det.collect(16)
    host = frames.cpu().numpy()
    assert host.std() > 5  # something was drawn
    rc, rm = R.detect_batch_mt(host, state, fs, 5, True, 5, threads=n, cap=16)
    found = 0
    for f in range(n):
        assert [int(info[k][f]) for k in ("n_labels", "n_legal", "n_quads", "n_features", "n_groups", "n_markers")] == list(rc[f][:6]), f
        for k in range(int(c[f])):
            nf = int(rm[f][k]["n_features"])
            assert int(m[f][k]["marker_id"]) == int(rm[f][k]["marker_id"]) and int(m[f][k]["inverse"]) == int(rm[f][k]["inverse"])
            assert list(m[f][k]["feature_pos"][:nf]) == list(rm[f][k]["feature_pos"][:nf])
            assert np.abs(m[f][k]["corners"][:nf] - rm[f][k]["corners"][:nf]).max() <= 1e-3
            assert int(m[f][k]["marker_id"]) in truth[f]  # a decoded ID is one of the rendered rows
            found += 1
    assert found >= 0.6 * n * markers  # most rendered markers decode
    det.close()


def test_render_is_deterministic_and_checks_its_arguments(marker_path):
    from cylindertag_b200 import CtagError
    det = Detector(marker_path=marker_path)
    a = torch.zeros((2, 480, 640), dtype=torch.uint8, device="cuda")
    b = torch.zeros_like(a)
    synth.render_frames_gpu(det, a.data_ptr(), [1, 2], 640, 480, 1, 1)
    synth.render_frames_gpu(det, b.data_ptr(), [1, 2], 640, 480, 1, 1)
    assert torch.equal(a, b) and not torch.equal(a[0], a[1])
    with pytest.raises(CtagError):
        synth.render_frames_gpu(det, 0, [1], 640, 480, 1, 1)
    det.close()
