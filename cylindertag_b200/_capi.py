"""ctypes binding of include/ctag.h.  Loading fails loudly when the CUDA library has not been built; there is no
CPU fallback behind these calls."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTAG_LIB") or os.path.join(HERE, "lib", "libctag_b200.so")  # CTAG_LIB: A/B builds of the same ABI

MAX_FEATURES = 20
STAGE_NAMES = ("front", "ccl", "quad", "feature", "decode")

OK = 0
ERR_ARG, ERR_FILE, ERR_DICTIONARY, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_ALIGNMENT = range(-1, -9, -1)
FRAME_OK, FRAME_NO_CORNER, FRAME_NO_FEATURE = 0, 1, 2


class CtagMarker(ctypes.Structure):
    _fields_ = [
        ("marker_id", ctypes.c_int32),
        ("n_features", ctypes.c_int32),
        ("inverse", ctypes.c_int32),
        ("frame", ctypes.c_int32),
        ("feature_pos", ctypes.c_int32 * MAX_FEATURES),
        ("feature_id", ctypes.c_int32 * MAX_FEATURES),
        ("id_left", ctypes.c_int32 * MAX_FEATURES),
        ("id_right", ctypes.c_int32 * MAX_FEATURES),
        ("cr_left", ctypes.c_float * MAX_FEATURES),
        ("cr_right", ctypes.c_float * MAX_FEATURES),
        ("edge_length", ctypes.c_float * MAX_FEATURES),
        ("center", ctypes.c_float * 2 * MAX_FEATURES),
        ("corners", ctypes.c_float * 2 * 8 * MAX_FEATURES),
    ]


class CtagFrameInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("status", "n_labels", "n_legal", "n_quads", "n_features", "n_groups", "n_markers", "flagged", "stale_ids")]
    _fields_ += [("reserved", ctypes.c_int32 * 3)]


# numpy view of ctag_marker for zero-copy array handling
MARKER_DTYPE = np.dtype([
    ("marker_id", "<i4"), ("n_features", "<i4"), ("inverse", "<i4"), ("frame", "<i4"),
    ("feature_pos", "<i4", (MAX_FEATURES,)), ("feature_id", "<i4", (MAX_FEATURES,)),
    ("id_left", "<i4", (MAX_FEATURES,)), ("id_right", "<i4", (MAX_FEATURES,)),
    ("cr_left", "<f4", (MAX_FEATURES,)), ("cr_right", "<f4", (MAX_FEATURES,)), ("edge_length", "<f4", (MAX_FEATURES,)),
    ("center", "<f4", (MAX_FEATURES, 2)), ("corners", "<f4", (MAX_FEATURES, 8, 2)),
])
INFO_DTYPE = np.dtype([(n, "<i4") for n in
                       ("status", "n_labels", "n_legal", "n_quads", "n_features", "n_groups", "n_markers", "flagged",
                        "stale_ids")] + [("reserved", "<i4", (3,))])
assert MARKER_DTYPE.itemsize == ctypes.sizeof(CtagMarker)
assert INFO_DTYPE.itemsize == ctypes.sizeof(CtagFrameInfo)

# symbol -> (restype, argtypes); must list every function include/ctag.h declares (checked by tests/test_capi.py)
_P = ctypes.c_void_p
_I = ctypes.c_int
_SZ = ctypes.c_size_t
SIGNATURES = {
    "ctag_create": (_I, [ctypes.POINTER(_P), _P, _I, _I, _I, _I]),
    "ctag_create_from_file": (_I, [ctypes.POINTER(_P), ctypes.c_char_p, _I]),
    "ctag_destroy": (None, [_P]),
    "ctag_set_option": (_I, [_P, ctypes.c_char_p, _I]),
    "ctag_detect_batch_multi": (_I, [_P, _I, _P, _I, _I, _I, ctypes.c_size_t, ctypes.c_size_t, _I, _I, _I, _I, _P, _I, _P, _P]),
    "ctag_get_dictionary": (_I, [_P, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), _P, _I]),
    "ctag_detect": (_I, [_P, _P, _I, _I, _SZ, _I, _I, _I, _P, _I, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "ctag_detect_batch": (_I, [_P, _P, _I, _I, _I, _SZ, _SZ, _I, _I, _I, _I, _I, _P, _I, _P, _P]),
    "ctag_detect_batch_jpeg": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "ctag_jpeg_backend": (_I, [_P]),
    "ctag_debug_get_input": (_I, [_P, _I, _P, _SZ]),
    "ctag_detect_batch_enqueue": (_I, [_P, _P, _I, _I, _I, _SZ, _SZ, _I, _I, _I, _I]),
    "ctag_detect_batch_collect": (_I, [_P, _P, _I, _P, _P]),
    "ctag_max_in_flight": (_I, []),
    "ctag_stage_time_ms": (_I, [_P, ctypes.POINTER(ctypes.c_float)]),
    "ctag_stage_timeline_ms": (_I, [_P, ctypes.POINTER(ctypes.c_float)]),
    "ctag_estimate_pose": (_I, [_P, _P, _I, _P, _P, _I, _P, _P, _P]),
    "ctag_pose_select_points": (_I, [_P, _P, _P, _I]),
    "ctag_project_points": (_I, [_P, _I, _P, _P, _P, _P, _I, _P]),
    "ctag_gray_to_3ch": (_I, [_P, _I, _I, _SZ, _P, _SZ]),
    "ctag_draw_axis": (_I, [_P, _I, _I, _SZ, _P, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I]),
    "ctag_codebook_capacity": (_I, [_I, _I]),
    "ctag_generate_codebook": (_I, [_I, _I, _I, ctypes.c_uint64, _P, _I, ctypes.POINTER(_I)]),
    "ctag_check_codebook": (_I, [_P, _I, _I, _I]),
    "ctag_render_frames": (_I, [_P, _P, _I, _I, _I, _SZ, _SZ, _I, _P, _P, _P]),
    "ctag_last_launch_count": (_I, [_P]),
    "ctag_stream": (_P, [_P]),
    "ctag_debug_get_gray": (_I, [_P, _I, _P, _SZ]),
    "ctag_debug_get_binary": (_I, [_P, _I, _P, _SZ]),
    "ctag_debug_get_quad_counters": (_I, [_P, _P]),
    "ctag_debug_get_components": (_I, [_P, _I, _P, _I, ctypes.POINTER(_I)]),
    "ctag_debug_get_quads": (_I, [_P, _I, _P, _P, _I, ctypes.POINTER(_I)]),
    "ctag_debug_get_features": (_I, [_P, _I, _P, _P, _P, _P, _I, ctypes.POINTER(_I)]),
    "ctag_strerror": (ctypes.c_char_p, [_I]),
    "ctag_last_error": (ctypes.c_char_p, []),
    "ctag_version": (ctypes.c_char_p, []),
}

_lib = None


def load():
    """Loads libctag_b200.so (built in-tree by cylindertag_b200.build).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m cylindertag_b200.build` "
                "(the detection path is CUDA-only; there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class CtagError(RuntimeError):
    def __init__(self, code, where=""):
        lib = load()
        msg = lib.ctag_strerror(code).decode()
        detail = lib.ctag_last_error().decode()
        super().__init__(f"{where}: {msg} (code {code})" + (f" [{detail}]" if detail and code in (ERR_CUDA, ERR_NO_DEVICE) else ""))
        self.code = code


def check(code, where=""):
    if code != OK:
        raise CtagError(code, where)
