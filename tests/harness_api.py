"""ctypes access to the host logic harness (tests/host_harness): the CUDA kernels' __host__ __device__ cores compiled
for the CPU.  TEST INFRASTRUCTURE -- used to check kernel logic against the oracle where no GPU is available."""
import ctypes

import numpy as np

from cylindertag_b200 import _capi as C
from tests.host_harness.build import build_harness

FEAT = np.dtype([("c", "<f4", (8, 2)), ("cx", "<f4"), ("cy", "<f4"), ("angle", "<f4"), ("qi", "<i4"), ("qj", "<i4")])
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_harness())
        assert FEAT.itemsize == _lib.hh_sizeof_feature()
    return _lib


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fit_l2(p):
    p = np.ascontiguousarray(p, np.int32)
    out = np.zeros(4, np.float32)
    lib().hh_fit_l2(vp(p), len(p), vp(out))
    return out


def fit_welsch(p, mode=1):
    p = np.ascontiguousarray(p, np.int32)
    out = np.zeros(4, np.float32)
    it = ctypes.c_int()
    lib().hh_fit_welsch(vp(p), len(p), mode, vp(out), ctypes.byref(it))
    return out


def block_labels(labels):
    """cv2 label image -> one label per 2x2 block (-1 = background), as the CCL kernels produce (up to naming)."""
    rows, cols = labels.shape
    bw, bh = (cols + 1) // 2, (rows + 1) // 2
    pad = np.zeros((bh * 2, bw * 2), np.int32)
    pad[:rows, :cols] = labels
    blk = pad.reshape(bh, 2, bw, 2).max(axis=(1, 3)).astype(np.int32)
    blk[blk == 0] = -1
    return np.ascontiguousarray(blk), bw


def quad_extract(binary, blk, bw, comp):
    cor = np.zeros(8, np.float32)
    info = np.zeros(3, np.int32)
    rows, cols = binary.shape
    lib().hh_quad_extract(vp(binary), cols, vp(blk), bw, cols, rows, comp.label, comp.area, comp.x0, comp.y0, comp.x1,
                          comp.y1, vp(cor), vp(info))
    return cor.reshape(4, 2), info


def detect_tail(quads, gray, state, fs, subpix, dist):
    """features -> refine -> markers with the host cores. Returns (feats[FEAT], markers, n_markers, info)."""
    q = np.ascontiguousarray(np.array(quads, np.float32).reshape(-1, 8))
    feats = np.zeros(256, FEAT)
    nf = lib().hh_feature_recovery(vp(q), len(q), vp(feats), 256)
    half = feats["c"][:nf].copy()
    if nf < fs or nf > 100:
        return feats[:nf], half, None, 0, None
    lib().hh_corner_obtain(vp(feats), nf)
    g = np.ascontiguousarray(gray)
    if subpix:
        lib().hh_edge_refine(vp(g), g.shape[1], g.shape[1], g.shape[0], vp(feats), nf, dist)
    out = np.zeros(64, C.MARKER_DTYPE)
    info = np.zeros(3, np.int32)
    st = np.ascontiguousarray(state, np.int32)
    nm = lib().hh_organize_decode(vp(feats), nf, vp(st), st.shape[0], st.shape[1], fs, vp(out), 64, vp(info))
    return feats[:nf], half, out, nm, info
