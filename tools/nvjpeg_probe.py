"""Which nvJPEG backends work on this GPU and how fast they decode a batch of 4K JPEGs (decode only, device output).
python tools/nvjpeg_probe.py [n_frames]"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import torch  # noqa: E402

from cylindertag_b200 import workloads as wl  # noqa: E402

class Img(ctypes.Structure):
    _fields_ = [("channel", ctypes.c_void_p * 4), ("pitch", ctypes.c_size_t * 4)]




def main():
    nvj = ctypes.CDLL("libnvjpeg.so.12")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    backends = [int(v) for v in sys.argv[2:]] or [2, 1, 0]
    frames = wl.render_many([(4, "2f12c", i) for i in range(min(n, 4))], workers=1)
    jpegs = [cv2.imencode(".jpg", frames[i % len(frames)], [cv2.IMWRITE_JPEG_QUALITY, 90])[1].reshape(-1).copy() for i in range(n)]
    h, w = frames[0].shape[:2]
    pitch = (w * 3 + 15) // 16 * 16
    out = torch.empty((n, h, pitch), dtype=torch.uint8, device="cuda")


    names = {0: "DEFAULT", 1: "HYBRID", 2: "GPU_HYBRID", 3: "HARDWARE"}
    stream = torch.cuda.current_stream().cuda_stream
    for backend in backends:
        for threads in ((1, 8) if backend in (0, 1) else (1,)):
            print("creating", names[backend], flush=True)
            handle = ctypes.c_void_p()
            st = nvj.nvjpegCreateEx(backend, None, None, 0, ctypes.byref(handle))
            if st != 0:
                print(f"backend {names[backend]}: nvjpegCreateEx status {st}", flush=True)
                break
            state = ctypes.c_void_p()
            st = nvj.nvjpegJpegStateCreate(handle, ctypes.byref(state))
            print("created; initialising", flush=True)
            st2 = nvj.nvjpegDecodeBatchedInitialize(handle, state, n, threads, 6)
            print("initialised", st2, flush=True)
            ptrs = (ctypes.c_void_p * n)(*[j.ctypes.data for j in jpegs])
            sizes = (ctypes.c_size_t * n)(*[j.size for j in jpegs])
            imgs = (Img * n)()
            for i in range(n):
                imgs[i].channel[0] = out[i].data_ptr()
                imgs[i].pitch[0] = pitch
            sts = []
            ts = []
            for rep in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                s = nvj.nvjpegDecodeBatched(handle, state, ptrs, sizes, imgs, ctypes.c_void_p(stream))
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
                sts.append(s)
            ref = cv2.imdecode(jpegs[0], cv2.IMREAD_COLOR)
            got = out[0, :, :w * 3].reshape(h, w, 3).cpu().numpy()
            print(f"backend {names[backend]} cpu_threads {threads}: state {st} init {st2} decode {sts} best {min(ts) * 1e3:.1f} ms for {n} frames = "
                  f"{n / min(ts):.0f} frames/s; max |nvjpeg - libjpeg| = {int(np.abs(got.astype(int) - ref.astype(int)).max())}", flush=True)
            nvj.nvjpegJpegStateDestroy(state)
            nvj.nvjpegDestroy(handle)


if __name__ == "__main__":
    main()
