"""ctypes access to the C++ CPU restatement (TEST / BASELINE INFRASTRUCTURE)."""
import ctypes
import os

import numpy as np

from cylindertag_b200 import _capi as C
from .build import LIB, build_cpu_ref

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build_cpu_ref()
        _lib = ctypes.CDLL(LIB)
    return _lib


def detect_batch(frames, state, fs, subpix=True, dist=5, threads=1, cap=32):
    fr = np.ascontiguousarray(frames, np.uint8)
    ch = 1 if fr.ndim == 3 else 3
    n, h, w = fr.shape[:3]
    st = np.ascontiguousarray(state, np.int32)
    counts = np.zeros((n, 8), np.int32)
    markers = np.zeros((n, cap), C.MARKER_DTYPE)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    load().cpu_ref_detect_batch(vp(fr), n, w, h, ch, vp(st), st.shape[0], st.shape[1], int(fs), int(bool(subpix)), int(dist),
                                int(threads), vp(counts), vp(markers), cap)
    return counts, markers


def detect_batch_bgr(frames, state, fs, adaptive_thresh, subpix, dist, threads):
    assert adaptive_thresh == 5
    counts, _ = detect_batch(frames, state, fs, subpix, dist, threads)
    return int(counts[:, 5].sum())
