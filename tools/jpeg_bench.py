"""Compressed-ingest timing on the config-5 ring (GPU box): python tools/jpeg_bench.py [frames] [reps] [rst_interval]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    import torch

    from cylindertag_b200 import Detector, workloads as wl
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rst = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    state, fs = wl.codebook("2f12c")
    frames = wl.render_many([(4, "2f12c", i) for i in range(min(n, 8))], workers=1)
    enc = [cv2.imencode(".jpg", frames[i % len(frames)], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])[1].reshape(-1)
           for i in range(n)]
    pinned = torch.empty(sum(e.size for e in enc), dtype=torch.uint8).pin_memory()
    jpegs, pos = [], 0
    for e in enc:
        pinned[pos:pos + e.size] = torch.from_numpy(e)
        jpegs.append(pinned[pos:pos + e.size].numpy())
        pos += e.size
    det = Detector(state=state, feature_size=fs)
    det.set_option("chunk_frames", chunk)
    det.detect_batch_jpeg(jpegs, 5, True, 5)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        m, c, _ = det.detect_batch_jpeg(jpegs, 5, True, 5)
        ts.append(time.perf_counter() - t0)
    print(f"{n} frames, rst {rst}, chunk {chunk}, {pos / n / 1e6:.2f} MB/frame: best {min(ts) * 1e3:.2f} ms = {n / min(ts):.0f} frames/s, markers {int(c.sum())}, decoder {det.jpeg_backend()}")


if __name__ == "__main__":
    main()
