"""Parity on rendered frames (BASELINE.json configs 3-5): 1080p single-marker frames, 4K multi-marker frames, generated
15c3f / 18c4f codebooks, BGR input, and a batch large enough to exercise the chunked host pipeline."""
import numpy as np
import pytest

from cylindertag_b200 import Detector, synth
from oracle import ctag_oracle as o
from tests.parity import assert_frame_matches, assert_markers_match

pytestmark = pytest.mark.gpu


def _check(det, frames, state, fs, channels=1, subpix=True, dist=5):
    markers, counts, info = det.detect_batch(frames, 5, subpix, dist, cap_per_frame=32)
    worst = 0.0
    for f in range(len(frames)):
        gray = frames[f] if channels == 1 else o.bgr2gray(frames[f])
        dump = o.detect(gray, state, fs, 5, subpix, dist)
        if len(frames) < 8:
            worst = max(worst, assert_frame_matches(det, f, info, markers, counts, dump, subpix, ctx=f"frame {f}"))
        else:  # chunked host path: the stage dumps only cover the last chunk, compare the results
            if dump.flagged:
                assert int(info["flagged"][f]) == 1
                continue
            assert int(info["n_quads"][f]) == len(dump.quads) and int(info["n_features"][f]) == len(dump.feats_half)
            if dump.status == "ok":
                worst = max(worst, assert_markers_match(markers[f], int(counts[f]), dump.markers, ctx=f"frame {f}"))
    return worst, counts


def test_config3_1080p_single_marker(detector, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames, truth = [], []
    for seed in range(1000, 1006):
        fr, specs = synth.synthetic_frame(seed, 1920, 1080, state, 1)
        frames.append(fr)
        truth.append(specs[0][0])
    markers, counts, info = detector.detect_batch(np.stack(frames), 5, True, 5)
    worst, _ = _check(detector, np.stack(frames), state, fs)
    assert worst <= 1e-3
    hits = sum(int(any(int(markers[f][k]["marker_id"]) == truth[f] for k in range(int(counts[f])))) for f in range(len(frames)))
    assert hits == len(frames)  # every rendered ID is decoded


def test_config4_4k_multi_marker_bgr(detector, marker_path):
    state, fs = o.load_marker_file(marker_path)
    frames = np.stack([synth.synthetic_frame(2000 + i, 3840, 2160, state, 6, channels=3)[0] for i in range(2)])
    worst, counts = _check(detector, frames, state, fs, channels=3)
    assert worst <= 1e-3 and counts.sum() >= 8


@pytest.mark.parametrize("cols,fsz", [(15, 3), (18, 4)])
def test_generated_codebooks(cols, fsz):
    state = synth.generate_codebook(cols, fsz, 30, seed=7)
    assert synth.check_codebook(state, fsz)
    det = Detector(state=state, feature_size=fsz)
    frames = np.stack([synth.synthetic_frame(3000 + i, 1920, 1080, state, 2)[0] for i in range(3)])
    worst, counts = _check(det, frames, state, fsz)
    assert worst <= 1e-3
    det.close()


def test_large_host_batch_is_chunked(detector, marker_path):
    state, fs = o.load_marker_file(marker_path)
    base = [synth.synthetic_frame(1000 + i, 1280, 720, state, 1)[0] for i in range(4)]
    frames = np.stack([base[i % 4] for i in range(18)])  # 18 frames -> 4 chunks, three in flight
    worst, counts = _check(detector, frames, state, fs)
    assert worst <= 1e-3
    assert all(counts[i] == counts[i % 4] for i in range(18))


def test_more_chunks_than_workspaces(detector, marker_path):
    """A host batch cut into more chunks than there are workspaces (what a 64-frame 4K batch does): chunks queue up
    behind the in-flight ones and the results equal those of the default chunking, frame for frame."""
    state, fs = o.load_marker_file(marker_path)
    base = [synth.synthetic_frame(1000 + i, 1280, 720, state, 1)[0] for i in range(3)]
    frames = np.stack([base[i % 3] for i in range(19)])
    want_m, want_c, want_i = detector.detect_batch(frames, 5, True, 5, cap_per_frame=8)
    detector.set_option("chunk_frames", 2)  # 10 chunks, the last one a single frame
    try:
        got_m, got_c, got_i = detector.detect_batch(frames, 5, True, 5, cap_per_frame=8)
    finally:
        detector.set_option("chunk_frames", 0)
    assert np.array_equal(got_c, want_c) and got_c.sum() >= 19
    assert np.array_equal(got_i, want_i)
    assert np.array_equal(got_m.view(np.uint8), want_m.view(np.uint8))


def test_failed_host_batch_leaves_the_detector_usable(detector, marker_path):
    """A failure in the middle of the chunked host pipeline (injected: the third chunk) must not wedge the handle: pending
    chunks are drained and the next call gives the same result as before."""
    from cylindertag_b200 import CtagError
    state, fs = o.load_marker_file(marker_path)
    base = [synth.synthetic_frame(1000 + i, 1280, 720, state, 1)[0] for i in range(3)]
    frames = np.stack([base[i % 3] for i in range(19)])
    want_m, want_c, _ = detector.detect_batch(frames, 5, True, 5, cap_per_frame=8)
    detector.set_option("chunk_frames", 2)
    detector.set_option("debug_fail_chunk", 2)
    try:
        with pytest.raises(CtagError):
            detector.detect_batch(frames, 5, True, 5, cap_per_frame=8)
        got_m, got_c, _ = detector.detect_batch(frames, 5, True, 5, cap_per_frame=8)  # same handle, next call
    finally:
        detector.set_option("chunk_frames", 0)
    assert np.array_equal(got_c, want_c) and np.array_equal(got_m.view(np.uint8), want_m.view(np.uint8))


def test_batch_without_count_array_still_numbers_its_frames(detector, marker_path):
    """n_out == NULL is allowed by ctag.h; the records of a chunked batch still carry the batch-global frame index."""
    import ctypes
    from cylindertag_b200 import _capi as C
    state, fs = o.load_marker_file(marker_path)
    base = [synth.synthetic_frame(1000 + i, 1280, 720, state, 1)[0] for i in range(3)]
    frames = np.ascontiguousarray(np.stack([base[i % 3] for i in range(12)]))
    out = np.zeros((12, 4), C.MARKER_DTYPE)
    rc = C.load().ctag_detect_batch(detector._h, frames.ctypes.data_as(ctypes.c_void_p), 12, 1280, 720, 1280, 0, 1, 0, 5, 1, 5,
                                    out.ctypes.data_as(ctypes.c_void_p), 4, None, None)
    assert rc == 0
    assert [int(out[f][0]["frame"]) for f in range(12)] == list(range(12)) and all(int(out[f][0]["n_features"]) >= 2 for f in range(12))


def test_multi_detector_call_equals_single_detector(marker_path):
    """ctag_detect_batch_multi (one detector per GPU, one host thread each) against the single call; with one
    GPU in the box the detectors share the device, which exercises the same sharding / threading / index code."""
    import torch
    from cylindertag_b200 import detect_batch_multi
    state, fs = o.load_marker_file(marker_path)
    ngpu = max(1, torch.cuda.device_count())
    dets = [Detector(state=state, feature_size=fs, device=g % ngpu) for g in range(max(2, min(ngpu, 4)))]
    base = [synth.synthetic_frame(1000 + i, 1280, 720, state, 1)[0] for i in range(5)]
    frames = np.stack([base[i % 5] for i in range(23)])
    want_m, want_c, want_i = dets[0].detect_batch(frames, 5, True, 5, cap_per_frame=8)
    got_m, got_c, got_i = detect_batch_multi(dets, frames, 5, True, 5, cap_per_frame=8)
    assert np.array_equal(got_c, want_c) and got_c.sum() >= 23 and np.array_equal(got_i, want_i)
    assert np.array_equal(got_m.view(np.uint8), want_m.view(np.uint8))
    from cylindertag_b200 import CtagError
    with pytest.raises(CtagError):  # the same detector for two blocks: refused, a detector is not re-entrant
        detect_batch_multi([dets[0], dets[0]], frames, 5, True, 5, cap_per_frame=8)
    for d in dets:
        d.close()
