"""Regenerates the committed golden fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted).  Nothing in tests/, smoke() or bench.py reads /root/reference at run time.

  python tests/golden/make_golden.py

Outputs
  data/test_gray.png          cvtColor(imread(test.bmp), BGR2GRAY)  (lossless PNG; main.cpp:29,36)
  data/CTag_2f12c.marker|.model, data/cameraParams.yml   fixture data files
  testbmp_detect.npz          oracle stage dumps of detect(gray, 5, true, 5) on test.bmp
"""
import os
import shutil
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ctag_oracle as o  # noqa: E402
from oracle.dump import dump_to_dict  # noqa: E402

REF = "/root/reference"


def main():
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    bgr = cv2.imread(os.path.join(REF, "test.bmp"))
    gray = o.bgr2gray(bgr)
    cv2.imwrite(os.path.join(HERE, "data", "test_gray.png"), gray, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    assert np.array_equal(cv2.imread(os.path.join(HERE, "data", "test_gray.png"), cv2.IMREAD_UNCHANGED), gray)
    for f in ("CTag_2f12c.marker", "CTag_2f12c.model", "cameraParams.yml"):
        shutil.copyfile(os.path.join(REF, f), os.path.join(HERE, "data", f))
    state, fs = o.load_marker_file(os.path.join(HERE, "data", "CTag_2f12c.marker"))
    d = o.detect(gray, state, fs, 5, True, 5)
    np.savez_compressed(os.path.join(HERE, "testbmp_detect.npz"), **dump_to_dict(d))
    print("markers:", [(m.markerID, m.inverse, m.featurePos) for m in d.markers])


if __name__ == "__main__":
    main()
