"""Python host-side mirror of the reference detect interface over the C ABI (include/ctag.h).

`Detector` is the thin handle wrapper the parity tests and bench drive; `CylinderTag` (api.py) keeps the reference's
class surface (header/CylinderTag.h:12-52).  Every call goes to the CUDA library; nothing here computes detections.
"""
import ctypes

import numpy as np

from . import _capi as C


def _ptr(arr):
    return ctypes.c_void_p(arr.ctypes.data)


class Detector:
    """One detector per CUDA device / host thread (the reference object is not re-entrant either)."""

    def __init__(self, state=None, feature_size=None, marker_path=None, device=-1):
        self._lib = C.load()
        self._h = ctypes.c_void_p()
        if marker_path is not None:
            C.check(self._lib.ctag_create_from_file(ctypes.byref(self._h), str(marker_path).encode(), device), "ctag_create_from_file")
        else:
            st = np.ascontiguousarray(state, dtype=np.int32)
            if st.ndim != 2 or feature_size is None:
                raise ValueError("state must be a 2-D int matrix and feature_size must be given")
            C.check(self._lib.ctag_create(ctypes.byref(self._h), _ptr(st), st.shape[0], st.shape[1], int(feature_size), device), "ctag_create")
        r, c, f = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        C.check(self._lib.ctag_get_dictionary(self._h, ctypes.byref(r), ctypes.byref(c), ctypes.byref(f), None, 0), "ctag_get_dictionary")
        self.rows, self.cols, self.feature_size = r.value, c.value, f.value
        self._shape = None
        self._n = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ctag_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def state(self):
        out = np.zeros((self.rows, self.cols), np.int32)
        C.check(self._lib.ctag_get_dictionary(self._h, None, None, None, _ptr(out), out.size), "ctag_get_dictionary")
        return out

    def set_option(self, key, value):
        """ctag_set_option: "chunk_frames" (frames per chunk of a host batch, 0 = automatic), "debug_fail_chunk"."""
        C.check(self._lib.ctag_set_option(self._h, key.encode(), int(value)), "ctag_set_option")

    # ---- detection ------------------------------------------------------------------------------------------
    def detect_batch(self, frames, adaptive_thresh=5, corner_subpix=False, subpix_dist=3, cap_per_frame=16):
        """frames: host uint8 array [n,h,w] (gray) or [n,h,w,3] (BGR).  Returns (markers, counts, info) numpy
        structured arrays: markers[n, cap_per_frame], counts[n], info[n]."""
        fr = np.ascontiguousarray(frames, dtype=np.uint8)
        if fr.ndim == 3:
            n, h, w = fr.shape
            ch = 1
        elif fr.ndim == 4 and fr.shape[3] == 3:
            n, h, w, ch = fr.shape
        else:
            raise ValueError("frames must be [n,h,w] or [n,h,w,3] uint8")
        out = np.zeros((n, cap_per_frame), C.MARKER_DTYPE)
        cnt = np.zeros(n, np.int32)
        info = np.zeros(n, C.INFO_DTYPE)
        C.check(self._lib.ctag_detect_batch(self._h, _ptr(fr), n, w, h, w * ch, 0, ch, 0, int(adaptive_thresh),
                                            int(bool(corner_subpix)), int(subpix_dist), _ptr(out), cap_per_frame,
                                            _ptr(cnt), _ptr(info)), "ctag_detect_batch")
        self._shape, self._n = (h, w), n
        return out, cnt, info

    def detect_batch_jpeg(self, jpegs, adaptive_thresh=5, corner_subpix=False, subpix_dist=3, cap_per_frame=16):
        """jpegs: list of JPEG byte strings (bytes / uint8 arrays) of equally sized frames; decoded on the GPU by nvJPEG
        (ctag_detect_batch_jpeg).  Returns (markers, counts, info)."""
        bufs = [np.frombuffer(j, np.uint8) if isinstance(j, (bytes, bytearray, memoryview)) else np.ascontiguousarray(j, np.uint8).reshape(-1)
                for j in jpegs]
        n = len(bufs)
        ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (ctypes.c_size_t * n)(*[b.size for b in bufs])
        out = np.zeros((n, cap_per_frame), C.MARKER_DTYPE)
        cnt = np.zeros(n, np.int32)
        info = np.zeros(n, C.INFO_DTYPE)
        w, h = ctypes.c_int(), ctypes.c_int()
        C.check(self._lib.ctag_detect_batch_jpeg(self._h, ptrs, sizes, n, int(adaptive_thresh), int(bool(corner_subpix)), int(subpix_dist),
                                                 _ptr(out), cap_per_frame, _ptr(cnt), _ptr(info), ctypes.byref(w), ctypes.byref(h)),
                "ctag_detect_batch_jpeg")
        self._shape, self._n, self._channels = (h.value, w.value), n, 3
        return out, cnt, info

    def jpeg_backend(self):
        return {0: "nvjpeg hardware engine (NVJPG)", 1: "nvjpeg hybrid backend", 2: "cuda decoder (csrc/jpeg.cu)"}.get(
            int(self._lib.ctag_jpeg_backend(self._h)), "none")

    def debug_input(self, frame=0, channels=3):
        """Staged input of the last host / JPEG batch as the kernels saw it (last chunk)."""
        h, w = self._shape
        out = np.zeros((h, w, channels) if channels > 1 else (h, w), np.uint8)
        C.check(self._lib.ctag_debug_get_input(self._h, frame, _ptr(out), w * channels), "ctag_debug_get_input")
        return out

    def detect(self, gray, adaptive_thresh=5, corner_subpix=False, subpix_dist=3, cap=16):
        """Single 8-bit gray host image through ctag_detect.  Returns (markers[count], frame_status)."""
        g = np.ascontiguousarray(gray, dtype=np.uint8)
        if g.ndim != 2:
            raise ValueError("detect expects an 8-bit single-channel image (CylinderTag.cpp:67)")
        h, w = g.shape
        out = np.zeros(cap, C.MARKER_DTYPE)
        n, st = ctypes.c_int(), ctypes.c_int()
        C.check(self._lib.ctag_detect(self._h, _ptr(g), w, h, w, int(adaptive_thresh), int(bool(corner_subpix)),
                                      int(subpix_dist), _ptr(out), cap, ctypes.byref(n), ctypes.byref(st)), "ctag_detect")
        self._shape, self._n = (h, w), 1
        return out[:min(n.value, cap)], st.value

    def enqueue_device(self, dev_ptr, n, w, h, pitch, frame_stride, channels, adaptive_thresh=5, corner_subpix=False,
                       subpix_dist=3):
        C.check(self._lib.ctag_detect_batch_enqueue(self._h, ctypes.c_void_p(dev_ptr), n, w, h, pitch, frame_stride,
                                                    channels, int(adaptive_thresh), int(bool(corner_subpix)),
                                                    int(subpix_dist)), "ctag_detect_batch_enqueue")
        self._shape, self._n = (h, w), n
        self._pending = getattr(self, "_pending", [])
        self._pending.append(n)

    def collect(self, cap_per_frame=16):
        """Results of the oldest enqueued batch (up to max_in_flight() batches may be pending)."""
        n = self._pending.pop(0) if getattr(self, "_pending", None) else self._n
        out = np.zeros((n, cap_per_frame), C.MARKER_DTYPE)
        cnt = np.zeros(n, np.int32)
        info = np.zeros(n, C.INFO_DTYPE)
        C.check(self._lib.ctag_detect_batch_collect(self._h, _ptr(out), cap_per_frame, _ptr(cnt), _ptr(info)), "ctag_detect_batch_collect")
        return out, cnt, info

    # ---- instrumentation ------------------------------------------------------------------------------------
    def stage_times_ms(self):
        ms = (ctypes.c_float * len(C.STAGE_NAMES))()
        C.check(self._lib.ctag_stage_time_ms(self._h, ms), "ctag_stage_time_ms")
        return dict(zip(C.STAGE_NAMES, [float(v) for v in ms]))

    def stage_timeline_ms(self):
        """Stage boundaries (ms since the detector was created) of the most recent collected batch."""
        ms = (ctypes.c_float * (len(C.STAGE_NAMES) + 1))()
        C.check(self._lib.ctag_stage_timeline_ms(self._h, ms), "ctag_stage_timeline_ms")
        return [float(v) for v in ms]

    def launch_count(self):
        return int(self._lib.ctag_last_launch_count(self._h))

    def max_in_flight(self):
        return int(self._lib.ctag_max_in_flight())

    def stream(self):
        return self._lib.ctag_stream(self._h)

    def debug_gray(self, frame=0):
        h, w = self._shape
        out = np.zeros((h, w), np.uint8)
        C.check(self._lib.ctag_debug_get_gray(self._h, frame, _ptr(out), w), "ctag_debug_get_gray")
        return out

    def debug_binary(self, frame=0):
        h, w = self._shape
        out = np.zeros((h // 2, w // 2), np.uint8)
        C.check(self._lib.ctag_debug_get_binary(self._h, frame, _ptr(out), w // 2), "ctag_debug_get_binary")
        return out

    def debug_quad_counters(self):
        out = np.zeros(16, np.int32)
        C.check(self._lib.ctag_debug_get_quad_counters(self._h, _ptr(out)), "ctag_debug_get_quad_counters")
        return {"fit_components": int(out[2]), "exact_edges": int(out[1]), "restarts_total": 80 * int(out[2]),
                "restarts_one_per_warp": int(out[6]), "restarts_parked_in_tail": int(out[7]), "sequential_fallback_passes": int(out[10])}

    def debug_components(self, frame=0, cap=65536):
        out = np.zeros((cap, 6), np.int32)
        n = ctypes.c_int()
        C.check(self._lib.ctag_debug_get_components(self._h, frame, _ptr(out), cap, ctypes.byref(n)), "ctag_debug_get_components")
        return out[:n.value]

    def debug_quads(self, frame=0, cap=4096):
        idx = np.zeros(cap, np.int32)
        cor = np.zeros((cap, 4, 2), np.float32)
        n = ctypes.c_int()
        C.check(self._lib.ctag_debug_get_quads(self._h, frame, _ptr(idx), _ptr(cor), cap, ctypes.byref(n)), "ctag_debug_get_quads")
        return idx[:n.value], cor[:n.value]

    def debug_features(self, frame=0, cap=256):
        cor = np.zeros((cap, 8, 2), np.float32)
        cen = np.zeros((cap, 2), np.float32)
        ang = np.zeros(cap, np.float32)
        qp = np.zeros((cap, 2), np.int32)
        n = ctypes.c_int()
        C.check(self._lib.ctag_debug_get_features(self._h, frame, _ptr(cor), _ptr(cen), _ptr(ang), _ptr(qp), cap, ctypes.byref(n)), "ctag_debug_get_features")
        k = n.value
        return cor[:k], cen[:k], ang[:k], qp[:k]


def detect_batch_multi(detectors, frames, adaptive_thresh=5, corner_subpix=False, subpix_dist=3, cap_per_frame=16):
    """ctag_detect_batch_multi: one host batch sharded over several detectors (one per GPU, contiguous frame blocks, one
    host thread each inside the library).  Returns (markers, counts, info) indexed by the global frame number."""
    fr = np.ascontiguousarray(frames, dtype=np.uint8)
    if fr.ndim == 3:
        n, h, w = fr.shape
        ch = 1
    elif fr.ndim == 4 and fr.shape[3] == 3:
        n, h, w, ch = fr.shape
    else:
        raise ValueError("frames must be [n,h,w] or [n,h,w,3] uint8")
    lib = C.load()
    handles = (ctypes.c_void_p * len(detectors))(*[d._h for d in detectors])
    out = np.zeros((n, cap_per_frame), C.MARKER_DTYPE)
    cnt = np.zeros(n, np.int32)
    info = np.zeros(n, C.INFO_DTYPE)
    C.check(lib.ctag_detect_batch_multi(handles, len(detectors), _ptr(fr), n, w, h, w * ch, 0, ch, int(adaptive_thresh),
                                        int(bool(corner_subpix)), int(subpix_dist), _ptr(out), cap_per_frame, _ptr(cnt), _ptr(info)),
            "ctag_detect_batch_multi")
    return out, cnt, info
