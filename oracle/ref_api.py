"""TEST INFRASTRUCTURE ONLY.  ctypes access to oracle/_ref/libctag_ref.so -- the reference's own sources
(corner_detector.cpp, CylinderTag.cpp, pose_estimation.cpp) compiled unmodified against oracle/ref_shim/ (see
oracle/build_ref.py) -- plus the optional cv2 backend that routes the shim's OpenCV primitives to the real library.

    ref = RefDetector(marker_path=...)            # or state=..., feature_size=...
    dump = ref.detect(gray, 5, True, 5)           # RefDump: labels, comps, quads, feats, markers (golden-npz layout)
    with cv2_backend():  dump2 = ref.detect(...)  # same reference code, OpenCV arithmetic done by cv2 itself
"""
import contextlib
import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

from .build_ref import LIB, build_ref

_lib = None
c_int_p = ctypes.POINTER(ctypes.c_int)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_ubyte_p = ctypes.POINTER(ctypes.c_ubyte)


def load():
    global _lib
    if _lib is None:
        path = build_ref()
        lib = ctypes.CDLL(path)
        lib.ref_create_from_file.restype = ctypes.c_void_p
        lib.ref_create_from_file.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        lib.ref_create_from_state.restype = ctypes.c_void_p
        lib.ref_create_from_state.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        lib.ref_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_set_quiet.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.ref_last_error.restype = ctypes.c_char_p
        lib.ref_last_error.argtypes = [ctypes.c_void_p]
        lib.ref_feature_size.argtypes = [ctypes.c_void_p]
        lib.ref_dictionary_shape.argtypes = [ctypes.c_void_p, c_int_p, c_int_p]
        lib.ref_dictionary.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_detect.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int]
        lib.ref_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_labels.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_components.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_quads.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_features.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_marker_summary.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_marker_data.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 9
        lib.ref_load_model_camera.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        lib.ref_estimate_pose.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_detect_batch_mt.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 8 + \
            [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.shim_set_backend.argtypes = [ctypes.c_void_p]
        _lib = lib
    return _lib


def available():
    """True when the library exists or can be built (i.e. /root/reference is present)."""
    try:
        load()
        return True
    except (RuntimeError, OSError):
        return False


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@dataclass
class RefMarker:
    markerID: int
    featurePos: list
    feature_ID: list
    feature_ID_left: list
    feature_ID_right: list
    cr_left: list
    cr_right: list
    edge_length: list
    feature_center: np.ndarray  # (n, 2)
    cornerLists: np.ndarray     # (n, 8, 2)


@dataclass
class RefDump:
    n_labels: int
    labels: np.ndarray       # int32 (h/2, w/2); binary = labels > 0
    comps: np.ndarray        # (n_legal, 5): area, x0, y0, x1, y1 in list (= OpenCV label) order
    quads: np.ndarray        # (n_quads, 4, 2) float32
    feats: np.ndarray        # (n_features, 8, 2) float32, as detect() left them
    feats_center: np.ndarray
    feats_angle: np.ndarray
    status: str              # ok | no_corner | no_feature
    flagged: bool
    markers: list = field(default_factory=list)
    n_groups: int = 0

    @property
    def binary(self):
        return ((self.labels > 0) * 255).astype(np.uint8)


class RefDetector:
    """The reference's CylinderTag, compiled from /root/reference (header/CylinderTag.h:12-52)."""

    def __init__(self, marker_path=None, state=None, feature_size=None):
        lib = load()
        err = ctypes.create_string_buffer(512)
        if marker_path is not None:
            self._h = lib.ref_create_from_file(os.fsencode(marker_path), err, 512)
        else:
            st = np.ascontiguousarray(state, np.int32)
            self._h = lib.ref_create_from_state(_vp(st), st.shape[0], st.shape[1], int(feature_size), err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode(errors="replace"))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def feature_size(self):
        return self._lib.ref_feature_size(self._h)

    @property
    def state(self):
        r, c = ctypes.c_int(), ctypes.c_int()
        self._lib.ref_dictionary_shape(self._h, ctypes.byref(r), ctypes.byref(c))
        out = np.zeros((r.value, c.value), np.int32)
        self._lib.ref_dictionary(self._h, _vp(out))
        return out

    def detect(self, gray, adaptive_thresh=5, subpix=False, dist=3, stages=True, reset_ids=True) -> RefDump:
        g = np.ascontiguousarray(gray, np.uint8)
        assert g.ndim == 2
        h, w = g.shape
        lib = self._lib
        rc = lib.ref_detect(self._h, _vp(g), w, h, g.strides[0], int(adaptive_thresh), int(bool(subpix)), int(dist), int(bool(reset_ids)))
        if rc < 0:
            raise RuntimeError("reference detect failed: " + lib.ref_last_error(self._h).decode(errors="replace"))
        c = np.zeros(12, np.int32)
        lib.ref_counts(self._h, _vp(c))
        labels = np.zeros((h // 2, w // 2), np.int32)
        comps = np.zeros((max(int(c[1]), 1), 5), np.int32)
        if stages:
            lib.ref_labels(self._h, _vp(labels))
            lib.ref_components(self._h, _vp(comps), int(c[1]))
        quads = np.zeros((max(int(c[2]), 1), 4, 2), np.float32)
        lib.ref_quads(self._h, _vp(quads), int(c[2]))
        nf = int(c[3])
        fc, fcen, fang = np.zeros((max(nf, 1), 8, 2), np.float32), np.zeros((max(nf, 1), 2), np.float32), np.zeros(max(nf, 1), np.float32)
        lib.ref_features(self._h, _vp(fc), _vp(fcen), _vp(fang), nf)
        nm = int(c[4])
        summ = np.zeros((max(nm, 1), 3), np.int32)
        lib.ref_marker_summary(self._h, _vp(summ), nm)
        summ = summ[:nm]
        tf, tp = int(summ[:, 1].sum()), int(summ[:, 2].sum())
        fpos = np.zeros(max(tp, 1), np.int32)
        fid, idl, idr = (np.zeros(max(tf, 1), np.int32) for _ in range(3))
        crl, crr, el = (np.zeros(max(tf, 1), np.float32) for _ in range(3))
        cen, cor = np.zeros((max(tf, 1), 2), np.float32), np.zeros((max(tf, 1), 8, 2), np.float32)
        lib.ref_marker_data(self._h, _vp(fpos), _vp(fid), _vp(idl), _vp(idr), _vp(crl), _vp(crr), _vp(el), _vp(cen), _vp(cor))
        markers, a, p = [], 0, 0
        for mid, n, npos in summ:
            n, npos = int(n), int(npos)
            markers.append(RefMarker(int(mid), fpos[p:p + npos].tolist(), fid[a:a + n].tolist(), idl[a:a + n].tolist(), idr[a:a + n].tolist(),
                                     crl[a:a + n].tolist(), crr[a:a + n].tolist(), el[a:a + n].tolist(), cen[a:a + n].copy(), cor[a:a + n].copy()))
            a += n
            p += npos
        return RefDump(int(c[0]), labels, comps[:int(c[1])], quads[:int(c[2])], fc[:nf], fcen[:nf], fang[:nf],
                       ("ok", "no_corner", "no_feature")[int(c[5])], bool(c[6]), markers, int(c[8]))

    def load_model_camera(self, model_path, camera_path):
        if self._lib.ref_load_model_camera(self._h, os.fsencode(model_path), os.fsencode(camera_path)) != 0:
            raise RuntimeError(self._lib.ref_last_error(self._h).decode(errors="replace"))

    def estimate_pose(self, cap=64):
        """estimatePose on the markers of the last detect(); returns [(model_index, rvec(3), tvec(3))]."""
        ids = np.zeros(cap, np.int32)
        rt = np.zeros((cap, 6), np.float64)
        n = self._lib.ref_estimate_pose(self._h, _vp(ids), _vp(rt), cap)
        if n < 0:
            raise RuntimeError(self._lib.ref_last_error(self._h).decode(errors="replace"))
        return [(int(ids[i]), rt[i, :3].copy(), rt[i, 3:].copy()) for i in range(n)]


def inverse_flag(state, marker_id, feature_pos, feature_id):
    """pos_with_ID::inverse is not kept in MarkerInfo (header/corner_detector.h:16-22); it follows from the output.
    featurePos lists (y + direc * i) mod cols for the occupied code slots i in ascending order
    (corner_detector.cpp:1317-1321), so consecutive entries step forwards (direc = 1) or backwards (direc = -1, the
    inverse reading).  Fallback for a single occupied slot: a legal state never equals its inverse reading (its two
    digits lie in the same half, the inverse swaps halves), so a matching feature tells the direction.
    Returns 0 / 1, or -1 if undecidable."""
    cols = state.shape[1]
    pos = [p for p in feature_pos if p >= 0]
    steps = [(b - a) % cols for a, b in zip(pos, pos[1:])]
    fwd = sum(1 for s in steps if 0 < s < cols / 2)
    inv = sum(1 for s in steps if s > cols / 2)
    if fwd != inv:
        return int(inv > fwd)
    if len(pos) != len(feature_id):
        return -1
    fwd = inv = 0
    for p, c in zip(pos, feature_id):
        if c < 0:
            continue
        s = int(state[marker_id, p])
        fwd += s == c
        inv += s == (7 - c // 8) + (7 - c % 8) * 8
    return -1 if fwd == inv else int(inv > fwd)


def detect_batch_mt(frames, state, feature_size, adaptive_thresh=5, subpix=True, dist=5, threads=1, cap=32):
    """main.cpp's per-frame loop (cvtColor for BGR input, then detect) over a batch, one CylinderTag per host thread.
    Returns (counts [n][8]: n_labels, n_legal, n_quads, n_features, n_groups, n_markers, status, flagged;
             markers [n][cap] in the ctag_marker record layout, `inverse` derived from the dictionary)."""
    from cylindertag_b200 import _capi as C
    fr = np.ascontiguousarray(frames, np.uint8)
    ch = 1 if fr.ndim == 3 else 3
    n, h, w = fr.shape[:3]
    st = np.ascontiguousarray(state, np.int32)
    nmk = np.zeros(n, np.int32)
    counts = np.zeros((n, 8), np.int32)
    markers = np.zeros((n, cap), C.MARKER_DTYPE)
    rc = load().ref_detect_batch_mt(_vp(st), st.shape[0], st.shape[1], int(feature_size), _vp(fr), n, w, h, ch, int(adaptive_thresh),
                                    int(bool(subpix)), int(dist), int(threads), _vp(nmk), None, 0, _vp(counts), _vp(markers), cap)
    if rc != 0:
        raise RuntimeError("reference batch detect failed")
    for f in range(n):
        for k in range(min(int(nmk[f]), cap)):
            m = markers[f, k]
            nf = min(int(m["n_features"]), len(m["feature_id"]))
            m["inverse"] = inverse_flag(st, int(m["marker_id"]), m["feature_pos"].tolist(), m["feature_id"][:nf].tolist())
    return counts, markers


def detect_batch(frames, state, fs, subpix=True, dist=5, threads=1, cap=32):
    """detect_batch_mt with adaptiveThresh = 5 (the demo's call, main.cpp:57)."""
    return detect_batch_mt(frames, state, fs, 5, subpix, dist, threads, cap)


def detect_batch_bgr(frames, state, fs, adaptive_thresh, subpix, dist, threads):
    """bench.py's CPU arm: total markers found."""
    fr = np.ascontiguousarray(frames, np.uint8)
    ch = 1 if fr.ndim == 3 else 3
    n, h, w = fr.shape[:3]
    st = np.ascontiguousarray(state, np.int32)
    nmk = np.zeros(n, np.int32)
    rc = load().ref_detect_batch_mt(_vp(st), st.shape[0], st.shape[1], int(fs), _vp(fr), n, w, h, ch, int(adaptive_thresh),
                                    int(bool(subpix)), int(dist), int(threads), _vp(nmk), None, 0, None, None, 0)
    if rc != 0:
        raise RuntimeError("reference batch detect failed")
    return int(nmk.sum())


# ---- cv2 backend ---------------------------------------------------------------------------------------------------
class _Backend(ctypes.Structure):
    _fields_ = [
        ("resize_cubic_u8", ctypes.CFUNCTYPE(ctypes.c_int, c_ubyte_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, c_ubyte_p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_size_t)),
        ("ccl_bbdt", ctypes.CFUNCTYPE(ctypes.c_int, c_ubyte_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, c_int_p)),
        ("fit_line", ctypes.CFUNCTYPE(ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_float_p)),
        ("fast_atan2", ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float, ctypes.c_float)),
        ("solve_pnp_epnp", ctypes.CFUNCTYPE(ctypes.c_int, c_float_p, c_float_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int, c_double_p, c_double_p)),
        ("undistort_points", ctypes.CFUNCTYPE(ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int, c_float_p)),
        ("project_points", ctypes.CFUNCTYPE(ctypes.c_int, c_float_p, ctypes.c_int, c_double_p, c_double_p, c_float_p, c_float_p, ctypes.c_int, c_float_p)),
        ("convert_u8_f32", ctypes.CFUNCTYPE(ctypes.c_int, c_ubyte_p, ctypes.c_int, ctypes.c_double, c_float_p)),
    ]


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape) if n else np.zeros(shape, dtype)


def _strided_u8(ptr, w, h, step):
    buf = np.ctypeslib.as_array(ptr, shape=(h * step,))
    return np.lib.stride_tricks.as_strided(buf, shape=(h, w), strides=(step, 1))


def make_cv2_backend(only=None):
    """Builds the callback table; `only` = iterable of member names to route (default: all)."""
    import cv2
    stats = {k: 0 for k, _ in _Backend._fields_}

    def resize_cb(src, sw, sh, sstep, dst, dw, dh, dstep):
        stats["resize_cubic_u8"] += 1
        s = _strided_u8(src, sw, sh, sstep)
        out = cv2.resize(np.ascontiguousarray(s), (dw, dh), fx=0.5, fy=0.5, interpolation=cv2.INTER_CUBIC)
        _strided_u8(dst, dw, dh, dstep)[:] = out
        return 0

    def ccl_cb(img, w, h, step, labels):
        stats["ccl_bbdt"] += 1
        s = np.ascontiguousarray(_strided_u8(img, w, h, step))
        n, lab, _, _ = cv2.connectedComponentsWithStatsWithAlgorithm(s, 8, cv2.CV_32S, cv2.CCL_BBDT)
        np.ctypeslib.as_array(labels, shape=(h * w,))[:] = lab.reshape(-1)
        return n

    def fit_cb(xy, n, dist, param, reps, aeps, line):
        stats["fit_line"] += 1
        pts = np.ctypeslib.as_array(xy, shape=(2 * n,)).reshape(n, 1, 2).astype(np.float32)
        out = cv2.fitLine(pts, dist, param, reps, aeps).reshape(4)
        for k in range(4):
            line[k] = float(out[k])
        return 0

    def atan_cb(y, x):
        stats["fast_atan2"] += 1
        return float(cv2.fastAtan2(y, x))

    def kd(K, D, nd):
        Km = np.ctypeslib.as_array(K, shape=(9,)).reshape(3, 3).astype(np.float32)
        Dm = np.ctypeslib.as_array(D, shape=(nd,)).astype(np.float32).reshape(-1, 1) if nd else np.zeros((0, 1), np.float32)
        return Km, Dm

    def pnp_cb(obj, img, n, K, D, nd, rvec, tvec):
        stats["solve_pnp_epnp"] += 1
        o = np.ctypeslib.as_array(obj, shape=(3 * n,)).reshape(n, 3).astype(np.float32)
        i = np.ctypeslib.as_array(img, shape=(2 * n,)).reshape(n, 2).astype(np.float32)
        Km, Dm = kd(K, D, nd)
        ok, r, t = cv2.solvePnP(o, i, Km, Dm, flags=cv2.SOLVEPNP_EPNP)
        if not ok:
            return 1
        for k in range(3):
            rvec[k] = float(r[k, 0])
            tvec[k] = float(t[k, 0])
        return 0

    def undist_cb(xy, n, K, D, nd, out):
        stats["undistort_points"] += 1
        p = np.ctypeslib.as_array(xy, shape=(2 * n,)).reshape(n, 1, 2).astype(np.float32)
        Km, Dm = kd(K, D, nd)
        u = cv2.undistortPoints(p, Km, Dm, None, Km).reshape(-1)
        np.ctypeslib.as_array(out, shape=(2 * n,))[:] = u
        return 0

    def proj_cb(obj, n, rvec, tvec, K, D, nd, out):
        stats["project_points"] += 1
        o = np.ctypeslib.as_array(obj, shape=(3 * n,)).reshape(n, 3).astype(np.float32)
        r = np.ctypeslib.as_array(rvec, shape=(3,)).astype(np.float64)
        t = np.ctypeslib.as_array(tvec, shape=(3,)).astype(np.float64)
        Km, Dm = kd(K, D, nd)
        p, _ = cv2.projectPoints(o, r, t, Km, Dm)
        np.ctypeslib.as_array(out, shape=(2 * n,))[:] = p.reshape(-1).astype(np.float32)
        return 0

    def conv_cb(src, n, alpha, dst):
        stats["convert_u8_f32"] += 1
        s = np.ctypeslib.as_array(src, shape=(n,)).reshape(1, n)
        # Mat::convertTo(CV_32F, alpha) -- reached through cv2.convertScaleAbs' sibling: numpy-free OpenCV entry point
        out = cv2.multiply(s, 1.0, scale=alpha, dtype=cv2.CV_32F) if False else _convert_to(s, alpha)
        np.ctypeslib.as_array(dst, shape=(n,))[:] = out.reshape(-1)
        return 0

    cbs = {"resize_cubic_u8": resize_cb, "ccl_bbdt": ccl_cb, "fit_line": fit_cb, "fast_atan2": atan_cb, "solve_pnp_epnp": pnp_cb,
           "undistort_points": undist_cb, "project_points": proj_cb, "convert_u8_f32": conv_cb}
    b = _Backend()
    keep = []
    for name, ftype in _Backend._fields_:
        if only is None or name in only:
            f = ftype(cbs[name])
            keep.append(f)
            setattr(b, name, f)
    b._keep = keep
    b.stats = stats
    return b


def _convert_to(u8, alpha):
    """cv::Mat::convertTo(CV_32F, alpha) through the real library (the same entry oracle/ctag_oracle.py uses)."""
    from . import ctag_oracle as o
    if abs(alpha - 1.0 / 255) < 1e-18:
        return o.convert_to_float(np.ascontiguousarray(u8))
    import cv2
    return cv2.multiply(u8.astype(np.float32), np.float32(alpha))


@contextlib.contextmanager
def cv2_backend(only=None):
    """Routes the shim's OpenCV primitives to cv2 for the duration of the block (process-wide, single-threaded use)."""
    b = make_cv2_backend(only)
    lib = load()
    lib.shim_set_backend(ctypes.byref(b))
    try:
        yield b
    finally:
        lib.shim_set_backend(None)
