"""Where the host spends its time in the pipelined loop of bench.py: wall clock of the enqueue and collect calls
(4K BGR, batch 64).  A loop whose enqueue time approaches the step time is launch bound, one that waits in collect is
GPU bound.  Usage: python tools/host_overhead.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cylindertag_b200 import Detector, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
marker = os.path.join(ROOT, "tests", "golden", "data", "CTag_2f12c.marker")
det = Detector(marker_path=marker)
state = det.dictionary()[0] if hasattr(det, "dictionary") else None
if state is None:
    from oracle import ctag_oracle as o
    state, _ = o.load_marker_file(marker)
w, h, batch = 3840, 2160, 64
distinct = np.stack([synth.synthetic_frame(2000 + i, w, h, state, 6, channels=3)[0] for i in range(4)])
frames = torch.from_numpy(np.concatenate([distinct] * (batch // 4))).cuda()
pitch, fstride = w * 3, w * 3 * h
depth = det.max_in_flight()
for _ in range(depth):
    det.enqueue_device(frames.data_ptr(), batch, w, h, pitch, fstride, 3, 5, True, 5)
for _ in range(depth):
    det.collect(16)
torch.cuda.synchronize()
te, tc = [], []
t00 = time.perf_counter()
queued = 0
while queued < depth - 1:
    t0 = time.perf_counter()
    det.enqueue_device(frames.data_ptr(), batch, w, h, pitch, fstride, 3, 5, True, 5)
    te.append(time.perf_counter() - t0)
    queued += 1
for i in range(steps):
    if queued < steps:
        t0 = time.perf_counter()
        det.enqueue_device(frames.data_ptr(), batch, w, h, pitch, fstride, 3, 5, True, 5)
        te.append(time.perf_counter() - t0)
        queued += 1
    t0 = time.perf_counter()
    det.collect(16)
    tc.append(time.perf_counter() - t0)
total = time.perf_counter() - t00
print(f"steps {steps}: {1e3 * total / steps:.3f} ms/step wall; enqueue mean {1e3 * np.mean(te):.3f} ms (max {1e3 * np.max(te):.3f}), "
      f"collect mean {1e3 * np.mean(tc):.3f} ms (min {1e3 * np.min(tc):.3f})")
