"""End-to-end detect() parity through the C ABI (gpu)."""
import numpy as np
import pytest

from oracle import ctag_oracle as o
from tests.parity import assert_frame_matches, assert_markers_match

pytestmark = pytest.mark.gpu


def test_testbmp_known_answers(detector, test_gray, golden_testbmp):
    markers, counts, info = detector.detect_batch(test_gray[None], 5, True, 5)
    i = info[0]
    assert (i["n_labels"], i["n_legal"], i["n_quads"], i["n_features"], i["n_groups"], i["n_markers"]) == (135, 90, 59, 26, 7, 5)
    assert i["status"] == 0 and i["flagged"] == 0 and i["stale_ids"] == 0
    # SURVEY Appendix E
    assert list(markers[0]["marker_id"][:5]) == [23, 0, 1, 17, 5]
    assert list(markers[0]["inverse"][:5]) == [1, 1, 0, 0, 1]
    g = golden_testbmp
    cor, cen, ang, qp = detector.debug_features(0)
    assert np.array_equal(qp, g["feats_quads"])
    assert np.abs(cor - g["feats_refined"]).max() <= 1e-3
    n = g["mk_nfeat"]
    off = np.concatenate([[0], np.cumsum(n)])
    for k in range(5):
        m = markers[0][k]
        assert m["n_features"] == n[k]
        sl = slice(off[k], off[k + 1])
        assert list(m["feature_pos"][:n[k]]) == list(g["mk_feature_pos"][sl])
        assert list(m["feature_id"][:n[k]]) == list(g["mk_feature_id"][sl])
        assert np.abs(m["corners"][:n[k]] - g["mk_corners"][sl]).max() <= 1e-3


@pytest.mark.parametrize("win", [4, 7])
def test_other_threshold_windows_full_dump_parity(detector, test_gray, marker_path, win):
    """adaptiveThresh != 5: the unfused front kernels write no tile flags, so the component stage lists every tile and
    finds the empty ones itself -- every stage must still equal the reference's."""
    state, fs = o.load_marker_file(marker_path)
    dump = o.detect(test_gray, state, fs, win, True, 5)
    markers, counts, info = detector.detect_batch(test_gray[None], win, True, 5)
    worst = assert_frame_matches(detector, 0, info, markers, counts, dump, True, ctx=f"window={win}")
    assert worst <= 1e-3


def test_testbmp_full_dump_parity(detector, test_gray, marker_path):
    state, fs = o.load_marker_file(marker_path)
    for subpix, dist in ((True, 5), (False, 3), (True, 3)):
        dump = o.detect(test_gray, state, fs, 5, subpix, dist)
        markers, counts, info = detector.detect_batch(test_gray[None], 5, subpix, dist)
        worst = assert_frame_matches(detector, 0, info, markers, counts, dump, subpix, ctx=f"subpix={subpix},{dist}")
        assert worst <= 1e-3


def test_single_image_entry_point(detector, test_gray):
    mk, status = detector.detect(test_gray, 5, True, 5)
    assert status == 0 and list(mk["marker_id"]) == [23, 0, 1, 17, 5]


def test_empty_and_featureless_frames(detector):
    flat = np.full((2, 240, 320), 128, np.uint8)
    flat[1, 100:140, 100:160] = 10  # one dark blob: a quad but no pair
    markers, counts, info = detector.detect_batch(flat, 5, True, 5)
    assert list(counts) == [0, 0]
    assert info["status"][0] == 1  # "No corner detected!"
    assert info["status"][1] in (1, 2)


def test_batch_equals_single(detector, test_gray):
    batch = np.stack([test_gray, test_gray[::-1].copy(), test_gray])
    markers, counts, info = detector.detect_batch(batch, 5, True, 5)
    assert counts[0] == counts[2] == 5
    assert np.array_equal(markers[0][:5]["corners"], markers[2][:5]["corners"])
    assert list(markers[0]["frame"][:5]) == [0] * 5 and list(markers[2]["frame"][:5]) == [2] * 5
    single, _, _ = detector.detect_batch(batch[1:2], 5, True, 5)
    assert np.array_equal(single[0][:counts[1]]["corners"], markers[1][:counts[1]]["corners"])
