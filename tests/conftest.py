import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def test_gray():
    import cv2
    img = cv2.imread(os.path.join(DATA, "test_gray.png"), cv2.IMREAD_UNCHANGED)
    assert img is not None and img.shape == (1200, 1920)
    return img


@pytest.fixture(scope="session")
def golden_testbmp():
    return np.load(os.path.join(GOLDEN, "testbmp_detect.npz"))


@pytest.fixture(scope="session")
def marker_path():
    return os.path.join(DATA, "CTag_2f12c.marker")


@pytest.fixture(scope="session")
def detector(marker_path):
    """CUDA detector; fails (not skips) if the library or the GPU is missing -- there is no fallback to hide behind."""
    from cylindertag_b200 import Detector
    d = Detector(marker_path=marker_path)
    yield d
    d.close()
