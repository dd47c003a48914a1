"""include/cylindertag/imageio.h (BMP / PGM / PPM ingest for the C++ mirror) against cv2 -- CPU only."""
import os
import subprocess

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fnv(a):
    h = 1469598103934665603
    for v in a.reshape(-1).tolist():
        h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def reader(tmp_path_factory):
    exe = tmp_path_factory.mktemp("imageio") / "imageio_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cxx", "test_imageio.cpp"), "-o", str(exe)], check=True)
    return str(exe)


def run(reader, path):
    return subprocess.run([reader, str(path)], capture_output=True, text=True, check=True).stdout.split()


@pytest.mark.parametrize("ext,channels", [("bmp", 1), ("bmp", 3), ("pgm", 1), ("ppm", 3)])
def test_reads_what_cv2_writes(reader, tmp_path, ext, channels):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (37, 53) if channels == 1 else (37, 53, 3), dtype=np.uint8)  # odd width: BMP row padding
    p = tmp_path / f"img{channels}.{ext}"
    assert cv2.imwrite(str(p), img)
    rows, cols, ch, h = run(reader, p)
    gray = img if channels == 1 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert (int(rows), int(cols), int(ch)) == (37, 53, channels)
    assert int(h) == fnv(gray)


def test_missing_or_foreign_file_is_empty(reader, tmp_path):
    assert run(reader, tmp_path / "nope.bmp") == ["empty"]
    p = tmp_path / "x.bmp"
    p.write_bytes(b"not an image at all")
    assert run(reader, p) == ["empty"]
