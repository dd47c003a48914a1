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
    subprocess.run(["g++", "-std=c++17", "-O1", "-DCTAG_WITH_ZLIB", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cxx", "test_imageio.cpp"), "-o", str(exe), "-lz"], check=True)
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


@pytest.mark.parametrize("kind", ["gray", "bgr", "bgra", "gray_noise", "bgr_big"])
def test_reads_png(reader, tmp_path, kind):
    """PNG through zlib (CTAG_WITH_ZLIB): every row filter occurs in what cv2 writes for smooth and noisy content."""
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:61, 0:83]
    smooth = ((yy * 3 + xx * 2) % 256).astype(np.uint8)
    if kind == "gray":
        img = smooth
    elif kind == "gray_noise":
        img = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    elif kind == "bgr":
        img = np.stack([smooth, smooth[::-1], rng.integers(0, 256, (61, 83), dtype=np.uint8)], axis=2)
    elif kind == "bgra":
        img = np.stack([smooth, smooth[::-1], smooth.T[:61, :83] if False else smooth // 2, np.full((61, 83), 200, np.uint8)], axis=2)
    else:
        img = rng.integers(0, 256, (300, 517, 3), dtype=np.uint8)
        img[:150] = (np.arange(517)[None, :, None] % 256).astype(np.uint8)
    p = tmp_path / f"{kind}.png"
    assert cv2.imwrite(str(p), img)
    rows, cols, ch, h = run(reader, p)
    bgr = img if img.ndim == 2 else img[..., :3]
    gray = bgr if bgr.ndim == 2 else cv2.cvtColor(np.ascontiguousarray(bgr), cv2.COLOR_BGR2GRAY)
    assert (int(rows), int(cols), int(ch)) == (img.shape[0], img.shape[1], 1 if img.ndim == 2 else 3)
    assert int(h) == fnv(gray)


def test_reads_the_repo_fixture_png(reader):
    """tests/golden/data/test_gray.png = the reference's test.bmp as gray: what the demo programs are run on."""
    p = os.path.join(ROOT, "tests", "golden", "data", "test_gray.png")
    rows, cols, ch, h = run(reader, p)
    want = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert (int(rows), int(cols), int(ch)) == (1200, 1920, 1)
    assert int(h) == fnv(want)


def test_missing_or_foreign_file_is_empty(reader, tmp_path):
    assert run(reader, tmp_path / "nope.bmp") == ["empty"]
    p = tmp_path / "x.bmp"
    p.write_bytes(b"not an image at all")
    assert run(reader, p) == ["empty"]
