import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import ctag_oracle as o
from cylindertag_b200 import Detector
d = Detector(marker_path='tests/golden/data/CTag_2f12c.marker')
gray = cv2.imread('tests/golden/data/test_gray.png', cv2.IMREAD_UNCHANGED)
d.detect_batch(gray[None], 5, True, 5)
got = d.debug_binary(0)
ref = np.load('tests/golden/testbmp_detect.npz')['binary']
mism = got != ref
print('mismatch', mism.sum(), 'of', mism.size, 'got fg', (got>0).sum(), 'ref fg', (ref>0).sum())
ys, xs = np.nonzero(mism)
print('y range', ys.min() if len(ys) else None, ys.max() if len(ys) else None, 'x range', xs.min() if len(xs) else None, xs.max() if len(xs) else None)
print('x mod 80 hist', np.bincount(xs % 80, minlength=80))
print('y mod 40 hist', np.bincount(ys % 40, minlength=40))
print('first', list(zip(ys[:20], xs[:20])))
print('got vals uniq', np.unique(got))
