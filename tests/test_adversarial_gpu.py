"""Hostile inputs through the C ABI (gpu): noise, extremes, tiny and ragged sizes, dense clutter.  The CUDA path has to
agree with the oracle where the oracle is cheap, stay deterministic everywhere and flag (never overrun) its capacity
limits."""
import numpy as np
import pytest

from oracle import ctag_oracle as o
from tests.parity import assert_frame_matches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dictionary(marker_path):
    return o.load_marker_file(marker_path)


@pytest.mark.parametrize("shape", [(8, 8), (10, 14), (18, 34), (64, 48), (122, 76), (250, 314)])
def test_tiny_and_ragged_frames_match_oracle(detector, dictionary, shape):
    state, fs = dictionary
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    dump = o.detect(img, state, fs, 5, True, 5)
    markers, counts, info = detector.detect_batch(img[None], 5, True, 5)
    assert_frame_matches(detector, 0, info, markers, counts, dump, True, ctx=f"noise {shape}")


@pytest.mark.parametrize("seed", range(3))
def test_blocky_noise_matches_oracle(detector, dictionary, seed):
    """Noise blown up 6x: thousands of pixels of clutter components above the area threshold, most of them not quads."""
    state, fs = dictionary
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (50, 70), dtype=np.uint8)
    img = np.kron(small, np.ones((6, 6), np.uint8))
    dump = o.detect(img, state, fs, 5, True, 5)
    markers, counts, info = detector.detect_batch(img[None], 5, True, 5)
    assert_frame_matches(detector, 0, info, markers, counts, dump, True, ctx=f"blocky noise {seed}")


def test_odd_sizes_are_refused_loudly(detector):
    """2x decimation is only exact for even sizes: the oracle asserts, the C ABI returns CTAG_ERR_ARG."""
    from cylindertag_b200 import _capi as C
    with pytest.raises(C.CtagError) as e:
        detector.detect_batch(np.zeros((1, 17, 33), np.uint8), 5, True, 5)
    assert e.value.code == C.ERR_ARG


def test_extreme_frames(detector):
    frames = np.stack([np.zeros((300, 420), np.uint8), np.full((300, 420), 255, np.uint8),
                       np.tile(np.array([[0, 255], [255, 0]], np.uint8), (150, 210))])
    markers, counts, info = detector.detect_batch(frames, 5, True, 5)
    assert list(counts) == [0, 0, 0]
    assert all(int(s) in (1, 2) for s in info["status"])


def test_dense_clutter_flags_only_itself(detector):
    """A 4K frame tiled with 4,000 dark squares is far past the reference's envelope (1000 quads per frame, SURVEY C-4):
    it must come back flagged, and it must not touch the calm frame that shares its batch (per-frame fit quotas): that
    frame's results are bit-identical to a run on its own, twice."""
    h, w = 2160, 3840
    clutter = np.full((h, w), 200, np.uint8)
    for y in range(8, h - 40, 44):
        for x in range(8, w - 40, 44):
            clutter[y:y + 30, x:x + 30] = 20
    calm = np.full((h, w), 200, np.uint8)
    calm[500:560, 700:790] = 15
    calm[500:560, 800:890] = 15
    batch = np.stack([clutter, calm, clutter])
    runs = []
    for _ in range(2):
        m, c, i = detector.detect_batch(batch, 5, True, 5)
        runs.append((m[1].copy(), int(c[1]), i[1].copy(), detector.debug_components(1).copy(), detector.debug_quads(1)))
        assert i["n_labels"][0] == i["n_labels"][2] and i["n_labels"][0] > 3000
        assert i["n_legal"][0] == i["n_legal"][2] and i["n_legal"][0] > 3000
        assert int(i["flagged"][0]) == 1 and int(i["flagged"][2]) == 1
        assert int(i["flagged"][1]) == 0
    sm, sc, si = detector.detect_batch(calm[None], 5, True, 5)
    alone = (sm[0].copy(), int(sc[0]), si[0].copy(), detector.debug_components(0).copy(), detector.debug_quads(0))
    for r in runs:
        assert r[1] == alone[1] and r[2] == alone[2]
        assert np.array_equal(r[0].view(np.uint8), alone[0].view(np.uint8))
        assert np.array_equal(r[3], alone[3])
        assert np.array_equal(r[4][0], alone[4][0]) and np.array_equal(r[4][1], alone[4][1])
    assert int(alone[2]["n_legal"]) >= 1
