"""Warm single-frame latency of ctag_detect (host gray frame in, markers out; wall clock around the synchronous call).
Usage: python tools/latency.py [--res 4k|1080p|testbmp] [--iters 200]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2  # noqa: E402

from cylindertag_b200 import Detector, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--res", default="testbmp")
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
data = os.path.join(ROOT, "tests", "golden", "data")
if a.res == "testbmp":
    gray = cv2.imread(os.path.join(data, "test_gray.png"), cv2.IMREAD_GRAYSCALE)
    det = Detector(marker_path=os.path.join(data, "CTag_2f12c.marker"))
else:
    w, h = {"4k": (3840, 2160), "1080p": (1920, 1080)}[a.res]
    from cylindertag_b200 import workloads
    state, fs = workloads.codebook("2f12c")
    gray, _ = synth.synthetic_frame(2000, w, h, state, 6, channels=1)
    det = Detector(state=state, feature_size=fs)
for _ in range(10):
    recs, status = det.detect(gray, 5, True, 5, cap=64)
ts = []
for _ in range(a.iters):
    t0 = time.perf_counter()
    recs, status = det.detect(gray, 5, True, 5, cap=64)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"{a.res} {gray.shape[1]}x{gray.shape[0]}: markers {len(recs)} median {np.median(ts):.3f} ms p10 {np.percentile(ts, 10):.3f} p90 {np.percentile(ts, 90):.3f}"
      f" stages {det.stage_times_ms()}")
