"""C-ABI surface (no GPU): the library loads, exports every symbol include/ctag.h declares, the ctypes mirror of the
POD types has the C layout, and calls fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctag.h")).read()
    return sorted(set(re.findall(r"CTAG_API\s+[\w\s\*]+?\b(ctag_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("ctag_create", "ctag_create_from_file", "ctag_destroy", "ctag_detect", "ctag_detect_batch",
                 "ctag_detect_batch_enqueue", "ctag_detect_batch_collect", "ctag_strerror"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from cylindertag_b200 import _capi
    lib = _capi.load()
    for s in declared_symbols():
        assert hasattr(lib, s), s
        assert s in _capi.SIGNATURES, f"{s} missing from the ctypes signature table"
    assert set(_capi.SIGNATURES) == set(declared_symbols())
    assert b"sm_100a" in lib.ctag_version()


def test_pod_layout_matches_c(tmp_path):
    """sizeof/offsetof from a C compile of the header vs the numpy/ctypes mirrors."""
    from cylindertag_b200 import _capi
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ctag.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                 'sizeof(ctag_marker), offsetof(ctag_marker, corners), offsetof(ctag_marker, center),'
                 'sizeof(ctag_frame_info), offsetof(ctag_frame_info, stale_ids));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == _capi.MARKER_DTYPE.itemsize == ctypes.sizeof(_capi.CtagMarker)
    assert vals[1] == _capi.MARKER_DTYPE.fields["corners"][1]
    assert vals[2] == _capi.MARKER_DTYPE.fields["center"][1]
    assert vals[3] == _capi.INFO_DTYPE.itemsize
    assert vals[4] == _capi.INFO_DTYPE.fields["stale_ids"][1]


def test_no_device_fails_loudly(marker_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    from cylindertag_b200 import CtagError, Detector, _capi
    with pytest.raises(CtagError) as ei:
        Detector(marker_path=marker_path)
    assert ei.value.code == _capi.ERR_NO_DEVICE


def test_argument_errors_precede_device_selection(tmp_path):
    from cylindertag_b200 import _capi
    lib = _capi.load()
    h = ctypes.c_void_p()
    bad = np.full((2, 3), 64, np.int32)  # out of the 0..63 range (CylinderTag.cpp:56-65)
    assert lib.ctag_create(ctypes.byref(h), bad.ctypes.data_as(ctypes.c_void_p), 2, 3, 2, -1) == _capi.ERR_DICTIONARY
    assert lib.ctag_create_from_file(ctypes.byref(h), str(tmp_path / "missing.marker").encode(), -1) == _capi.ERR_FILE
    assert lib.ctag_create(ctypes.byref(h), None, 2, 3, 2, -1) == _capi.ERR_ARG
    assert lib.ctag_strerror(_capi.ERR_FILE) == b"could not open the file"


def test_product_package_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "cylindertag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)


def test_detection_modules_do_not_need_opencv():
    """cv2 is the ORACLE's library.  In the product package only synth.py (input synthesis for tests / bench; workloads.py
    calls it) may use it; the detector, the mirror class and the C ABI binding must not."""
    pkg = os.path.join(ROOT, "cylindertag_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py") and f != "synth.py":
            assert "cv2" not in open(os.path.join(pkg, f)).read(), f


def test_camera_loader_matches_filestorage(tmp_path):
    """CylinderTag.loadCamera parses OpenCV YAML 1.0 itself: same matrices as cv::FileStorage (CylinderTag.cpp:192-196)."""
    import cv2
    import numpy as np
    from cylindertag_b200.api import CylinderTag, _read_opencv_matrix
    path = os.path.join(ROOT, "tests", "golden", "data", "cameraParams.yml")
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    tag = CylinderTag.__new__(CylinderTag)  # the loader is host code: no detector needed
    cam = tag.loadCamera(path)
    for got, name in ((cam.Intrinsic, "cameraMatrix"), (cam.distCoeffs, "distCoeffs")):
        want = fs.getNode(name).mat()
        assert got.dtype == want.dtype and np.array_equal(got, want), name
    # a written-by-OpenCV file with another dt / layout, and a missing node
    out = str(tmp_path / "cam.yml")
    w = cv2.FileStorage(out, cv2.FILE_STORAGE_WRITE)
    K = np.array([[2200.5, 0, 960], [0, 2201.25, 540], [0, 0, 1]], np.float64)
    w.write("cameraMatrix", K)
    w.write("distCoeffs", np.zeros((5, 1), np.float32))
    w.release()
    assert np.array_equal(_read_opencv_matrix(out, "cameraMatrix"), K)
    assert _read_opencv_matrix(out, "distCoeffs").shape == (5, 1)
    assert _read_opencv_matrix(out, "nothing").size == 0


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback behind the binding: with the shared library absent, loading raises (in a fresh interpreter, with
    CTAG_LIB pointing at a path that does not exist)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cylindertag_b200 import _capi\n"
            "try:\n    _capi.load()\nexcept ImportError as e:\n    print('ImportError', 'no CPU fallback' in str(e))\n"
            "else:\n    print('loaded')\n") % ROOT
    env = dict(os.environ, CTAG_LIB=str(tmp_path / "absent" / "libctag_b200.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120).stdout
    assert out.strip() == "ImportError True"
