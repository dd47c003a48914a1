"""Compressed ingest: ctag_detect_batch_jpeg decodes JPEG frames on the GPU and runs the detect path on
them.  Frames with restart markers go through the library's own CUDA decoder (csrc/jpeg.cu), whose pixels have to equal
cv::imdecode's byte for byte -- i.e. the frame the reference would have read with cv::imread; frames without restart
markers fall back to nvJPEG, whose output differs from libjpeg's by a level or two, so there parity is defined on the
DECODED pixels: they are copied back (ctag_debug_get_input) and the reference's own code (oracle/_ref) runs on exactly
those pixels."""
import cv2
import numpy as np
import pytest

from cylindertag_b200 import CtagError
from oracle import ctag_oracle as o
from oracle import ref_api as R
from tests import configs
from tests.parity import assert_markers_match

pytestmark = pytest.mark.gpu


def _encode(frames, quality=92, rst=16):
    out = []
    for f in frames:
        ok, buf = cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        assert ok
        out.append(buf.reshape(-1).copy())
    return out


def test_jpeg_batch_matches_reference_on_the_decoded_pixels(detector, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = [configs.config3_frame(i) for i in range(6)]
    jpegs = _encode(frames)
    markers, counts, info = detector.detect_batch_jpeg(jpegs, 5, True, 5, cap_per_frame=16)
    assert detector.jpeg_backend().startswith("cuda decoder")
    decoded = np.stack([detector.debug_input(f) for f in range(6)])
    assert decoded.shape == (6, 1080, 1920, 3)
    # the CUDA decoder's pixels are cv::imdecode's (1080 rows: the last MCU row is cut, chroma rows replicate at the edge)
    for f in range(6):
        assert np.array_equal(decoded[f], cv2.imdecode(jpegs[f], cv2.IMREAD_COLOR)), f
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


@pytest.mark.parametrize("sampling", [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422])
def test_cuda_decoder_other_samplings_and_restart_lengths(detector, sampling):
    frames = [configs.config3_frame(i) for i in (7, 8)]
    for rst in (1, 5, 120, 2000):
        jpegs = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, rst,
                                           cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling])[1].reshape(-1) for f in frames]
        m, c, _ = detector.detect_batch_jpeg(jpegs, 5, True, 5, cap_per_frame=8)
        assert detector.jpeg_backend().startswith("cuda decoder")
        for f in range(2):
            assert np.array_equal(detector.debug_input(f), cv2.imdecode(jpegs[f], cv2.IMREAD_COLOR)), (rst, f)


def test_frames_without_restart_markers_fall_back_to_nvjpeg(detector, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = [configs.config3_frame(i) for i in range(2)]
    jpegs = _encode(frames, rst=0)
    markers, counts, info = detector.detect_batch_jpeg(jpegs, 5, True, 5, cap_per_frame=16)
    assert detector.jpeg_backend().startswith("nvjpeg")
    decoded = np.stack([detector.debug_input(f) for f in range(2)])
    assert np.abs(decoded.astype(np.int16) - np.stack([cv2.imdecode(j, cv2.IMREAD_COLOR) for j in jpegs]).astype(np.int16)).max() <= 8
    rc, rm = R.detect_batch_mt(decoded, state, fs, 5, True, 5, threads=2, cap=16)
    for f in range(2):
        assert int(counts[f]) == int(rc[f][5])
        for k in range(int(counts[f])):
            nf = int(rm[f][k]["n_features"])
            assert int(markers[f][k]["marker_id"]) == int(rm[f][k]["marker_id"])
            assert np.abs(markers[f][k]["corners"][:nf] - rm[f][k]["corners"][:nf]).max() <= 1e-3
    # and nvJPEG can be forced
    detector.set_option("jpeg_decoder", 1)
    try:
        detector.detect_batch_jpeg(_encode(frames), 5, True, 5)
        assert detector.jpeg_backend().startswith("nvjpeg")
    finally:
        detector.set_option("jpeg_decoder", 0)


def test_jpeg_gray_scale_and_bad_input(detector, test_gray):
    # a gray JPEG decodes to three equal channels, i.e. the gray frame itself (as cv::imread would give it)
    ok, buf = cv2.imencode(".jpg", test_gray, [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_RST_INTERVAL, 30])
    m, c, _ = detector.detect_batch_jpeg([buf.reshape(-1)], 5, True, 5, cap_per_frame=16)
    dec = detector.debug_input(0)
    assert np.array_equal(dec, cv2.imdecode(buf, cv2.IMREAD_COLOR))
    assert int(c[0]) >= 4
    # a stream cut in the middle of its entropy-coded data: the restart markers no longer add up
    cut = np.concatenate([buf.reshape(-1)[:len(buf) // 2], np.array([0xFF, 0xD9], np.uint8)])
    with pytest.raises(CtagError):
        detector.detect_batch_jpeg([cut], 5, True, 5)
    # garbage, and frames of different sizes in one batch, are refused; the handle stays usable
    with pytest.raises(CtagError):
        detector.detect_batch_jpeg([np.frombuffer(b"not a jpeg at all", np.uint8)], 5, True, 5)
    small = cv2.imencode(".jpg", test_gray[:600, :800], [cv2.IMWRITE_JPEG_RST_INTERVAL, 30])[1].reshape(-1)
    with pytest.raises(CtagError):
        detector.detect_batch_jpeg([buf.reshape(-1), small], 5, True, 5)
    m2, c2, _ = detector.detect_batch_jpeg([buf.reshape(-1)], 5, True, 5, cap_per_frame=16)
    assert int(c2[0]) == int(c[0])
