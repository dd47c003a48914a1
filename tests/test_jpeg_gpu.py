"""Compressed ingest: ctag_detect_batch_jpeg decodes JPEG frames on the GPU (nvJPEG) and runs the detect
path on them.  Parity is defined on the DECODED pixels: they are copied back (ctag_debug_get_input) and the reference's
own code (oracle/_ref) runs on exactly those pixels -- JPEG decoders differ from each other by a level or two, which is
not the detector's business."""
import cv2
import numpy as np
import pytest

from cylindertag_b200 import CtagError
from oracle import ctag_oracle as o
from oracle import ref_api as R
from tests import configs
from tests.parity import assert_markers_match

pytestmark = pytest.mark.gpu


def _encode(frames, quality=92):
    out = []
    for f in frames:
        ok, buf = cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, quality])
        assert ok
        out.append(buf.reshape(-1).copy())
    return out


def test_jpeg_batch_matches_reference_on_the_decoded_pixels(detector, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = [configs.config3_frame(i) for i in range(6)]
    jpegs = _encode(frames)
    markers, counts, info = detector.detect_batch_jpeg(jpegs, 5, True, 5, cap_per_frame=16)
    assert detector.jpeg_backend() != "none"
    decoded = np.stack([detector.debug_input(f) for f in range(6)])
    assert decoded.shape == (6, 1080, 1920, 3)
    # the decoder's output is the original up to JPEG loss (the rendered frames carry +-6 levels of chroma noise per pixel)
    assert np.abs(decoded.astype(np.int16) - np.stack(frames).astype(np.int16)).mean() < 5.0
    rc, rm = R.detect_batch_mt(decoded, state, fs, 5, True, 5, threads=6, cap=16)
    ref = R.RefDetector(state=state, feature_size=fs)
    for f in range(6):
        assert [int(info[k][f]) for k in ("n_labels", "n_legal", "n_quads", "n_features", "n_groups", "n_markers")] == list(rc[f][:6]), f
        d = ref.detect(o.bgr2gray(decoded[f]), 5, True, 5)
        assert np.array_equal(detector.debug_binary(f), d.binary), f
        n = int(counts[f])
        for k in range(n):
            g, w = markers[f][k], rm[f][k]
            nf = int(w["n_features"])
            assert (int(g["marker_id"]), int(g["inverse"]), int(g["n_features"])) == (int(w["marker_id"]), int(w["inverse"]), nf)
            assert list(g["feature_pos"][:nf]) == list(w["feature_pos"][:nf]) and list(g["feature_id"][:nf]) == list(w["feature_id"][:nf])
            assert np.abs(g["corners"][:nf] - w["corners"][:nf]).max() <= 1e-3
    ref.close()
    assert counts.sum() >= 5  # the markers survive JPEG quality 92


def test_jpeg_batch_is_chunked_like_a_host_batch(detector):
    frames = [configs.config3_frame(i) for i in range(5)]
    jpegs = _encode(frames)
    batch = [jpegs[i % 5] for i in range(22)]
    m, c, info = detector.detect_batch_jpeg(batch, 5, True, 5, cap_per_frame=8)  # 22 frames -> 4 chunks
    m5, c5, i5 = detector.detect_batch_jpeg(jpegs, 5, True, 5, cap_per_frame=8)  # 5 frames -> one chunk
    for f in range(22):
        assert int(c[f]) == int(c5[f % 5]) and info[f] == i5[f % 5]
        a, b = m[f][:int(c[f])].copy(), m5[f % 5][:int(c[f])].copy()
        assert all(int(v) == f for v in a["frame"])
        a["frame"], b["frame"] = 0, 0
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f


def test_jpeg_gray_scale_and_bad_input(detector, test_gray):
    # a baseline gray JPEG decodes to three equal channels (nvJPEG BGRI output), i.e. the gray frame itself
    ok, buf = cv2.imencode(".jpg", test_gray, [cv2.IMWRITE_JPEG_QUALITY, 95])
    m, c, _ = detector.detect_batch_jpeg([buf.reshape(-1)], 5, True, 5, cap_per_frame=16)
    dec = detector.debug_input(0)
    assert np.array_equal(dec[..., 0], dec[..., 1]) and np.array_equal(dec[..., 1], dec[..., 2])
    assert int(c[0]) >= 4
    # garbage, and frames of different sizes in one batch, are refused; the handle stays usable
    with pytest.raises(CtagError):
        detector.detect_batch_jpeg([np.frombuffer(b"not a jpeg at all", np.uint8)], 5, True, 5)
    small = cv2.imencode(".jpg", test_gray[:600, :800])[1].reshape(-1)
    with pytest.raises(CtagError):
        detector.detect_batch_jpeg([buf.reshape(-1), small], 5, True, 5)
    m2, c2, _ = detector.detect_batch_jpeg([buf.reshape(-1)], 5, True, 5, cap_per_frame=16)
    assert int(c2[0]) == int(c[0])
