"""Diagnostic (GPU box): per-feature corner differences between the CUDA path and the compiled reference on frames of the
config-2 sequence.  python tools/diag_features.py 28 [more frames]"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cylindertag_b200 import Detector, synth  # noqa: E402
from oracle import ref_api as R  # noqa: E402

data = os.path.join(ROOT, "tests", "golden", "data")
gray = cv2.imread(os.path.join(data, "test_gray.png"), cv2.IMREAD_UNCHANGED)
det = Detector(marker_path=os.path.join(data, "CTag_2f12c.marker"))
ref = R.RefDetector(marker_path=os.path.join(data, "CTag_2f12c.marker"))
for f in [int(v) for v in sys.argv[1:]] or [28]:
    fr = synth.video_sequence(gray, 120, 2024, first=f, count=1)[0]
    det.detect_batch(fr[None], 5, True, 5, cap_per_frame=32)
    cor = det.debug_features(0)[0]
    det.detect_batch(fr[None], 5, False, 5, cap_per_frame=32)
    cor0 = det.debug_features(0)[0]
    d = ref.detect(fr, 5, True, 5)
    d0 = ref.detect(fr, 5, False, 5)
    print("frame", f, "features", len(cor), len(d.feats), "unrefined max diff", float(np.abs(cor0 - d0.feats).max()))
    diff = np.abs(cor - d.feats)
    for i in range(len(cor)):
        if diff[i].max() > 1e-4:
            print(" feature", i, "max diff", float(diff[i].max()))
            for k in range(8):
                e = np.linalg.norm(cor0[i][(k + 1) % 4 + 4 * (k // 4)] - cor0[i][k])
                print("   corner", k, "gpu", cor[i][k], "ref", d.feats[i][k], "diff", diff[i][k], "unrefined", cor0[i][k], "edge len to next", float(e))
