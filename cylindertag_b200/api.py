"""Python mirror of the reference's `CylinderTag` class (header/CylinderTag.h:12-52) for the detect path.

Same method names and argument meaning as the reference: construct from a `.marker` file or a state matrix, `detect`,
`loadModel`, `loadCamera`, `estimatePose`, `drawAxis`.  `detect` runs entirely in the CUDA library (C ABI, include/ctag.h);
`estimatePose` stays on the host as in the reference (CylinderTag.cpp:198-209, pose_estimation.cpp:50-143) and runs in
the same library (ctag_estimate_pose, csrc/pose.cpp): EPnP for the start, then a Levenberg-Marquardt refinement of the
pinhole reprojection error on undistorted points (the reference uses OpenCV's EPnP and Ceres on the same cost).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _capi as C
from .detector import Detector


@dataclass
class MarkerInfo:  # header/corner_detector.h:16-22
    markerID: int = -1
    featurePos: list = field(default_factory=list)
    feature_ID: list = field(default_factory=list)
    feature_ID_left: list = field(default_factory=list)
    feature_ID_right: list = field(default_factory=list)
    cornerLists: list = field(default_factory=list)  # each [8][2] float32, full-res pixel coordinates
    feature_center: list = field(default_factory=list)
    edge_length: list = field(default_factory=list)
    cr_left: list = field(default_factory=list)
    cr_right: list = field(default_factory=list)


@dataclass
class CamInfo:  # header/pose_estimation.h:12-14
    Intrinsic: np.ndarray = None
    distCoeffs: np.ndarray = None


@dataclass
class ModelInfo:  # header/pose_estimation.h:16-20
    MarkerID: int = -1
    axis: np.ndarray = None
    base: np.ndarray = None
    corners: np.ndarray = None  # [8 * model_size, 3]


@dataclass
class PoseInfo:  # header/pose_estimation.h:22-25
    markerID: int = -1  # INDEX into the model list, not the dictionary row (pose_estimation.cpp:59,69)
    rvec: np.ndarray = None
    tvec: np.ndarray = None
    rms_px: float = float("nan")  # reprojection RMS of the refined pose (not in the reference struct)


def marker_from_record(rec) -> MarkerInfo:
    n = int(rec["n_features"])
    m = MarkerInfo(markerID=int(rec["marker_id"]))
    m.feature_ID = [int(v) for v in rec["feature_id"][:n]]
    m.feature_ID_left = [int(v) for v in rec["id_left"][:n]]
    m.feature_ID_right = [int(v) for v in rec["id_right"][:n]]
    m.featurePos = [int(v) for v in rec["feature_pos"][:n] if v >= 0]
    m.cornerLists = [rec["corners"][k].copy() for k in range(n)]
    m.feature_center = [tuple(rec["center"][k]) for k in range(n)]
    m.edge_length = [float(v) for v in rec["edge_length"][:n]]
    m.cr_left = [float(v) for v in rec["cr_left"][:n]]
    m.cr_right = [float(v) for v in rec["cr_right"][:n]]
    return m


class CylinderTag:
    def __init__(self, path_or_state, feature_size=None, device=-1):
        """CylinderTag(path) loads a .marker file (CylinderTag.cpp:16-41); CylinderTag(state, feature_size) takes the
        state matrix (CylinderTag.cpp:43-54; feature_size must be given, the reference leaves it unset, SURVEY C-3).
        Errors surface as the reference's messages (it throws std::string)."""
        try:
            if isinstance(path_or_state, (str, bytes)) or hasattr(path_or_state, "__fspath__"):
                self._det = Detector(marker_path=path_or_state, device=device)
            else:
                self._det = Detector(state=path_or_state, feature_size=feature_size, device=device)
        except C.CtagError as e:
            if e.code == C.ERR_FILE:
                raise RuntimeError("load_from_file, could not open the file\n") from e
            if e.code == C.ERR_DICTIONARY:
                raise RuntimeError("check_dictionary, the number in state matrix must between 0 to 63\n") from e
            raise

    @property
    def detector(self) -> Detector:
        return self._det

    # ---- Marker Detector (CylinderTag.cpp:67-128) ----
    def detect(self, img, cornerList=None, adaptiveThresh=5, cornerSubPix=False, cornerSubPixDist=3):
        """Returns the marker list; if `cornerList` (a list) is given it is assigned in place like the reference's
        output argument: replaced on success, left untouched on the two early exits (which print the reference's
        messages and return None)."""
        img = np.asarray(img)
        if img.ndim != 2 or img.dtype != np.uint8:
            raise ValueError("detect expects an 8-bit single-channel image")
        recs, status = self._det.detect(img, adaptiveThresh, cornerSubPix, cornerSubPixDist, cap=64)
        if status == C.FRAME_NO_CORNER:
            print("No corner detected!")
            return None
        if status == C.FRAME_NO_FEATURE:
            print("No feature detected!")
            return None
        markers = [marker_from_record(r) for r in recs]
        if cornerList is not None:
            cornerList[:] = markers
        return markers

    # ---- loaders ----
    def loadModel(self, path) -> list:
        """CylinderTag.cpp:161-190."""
        try:
            toks = open(path).read().split()
        except OSError as e:
            raise RuntimeError("loadModel, could not open the model file\n") from e
        it = iter(toks)
        n, size = int(next(it)), int(next(it))
        out = []
        for _ in range(n):
            m = ModelInfo(MarkerID=int(next(it)))
            m.base = np.array([float(next(it)) for _ in range(3)], np.float32)
            m.axis = np.array([float(next(it)) for _ in range(3)], np.float32)
            m.corners = np.zeros((8 * size, 3), np.float32)
            for _ in range(8 * size):
                cid = int(next(it))
                m.corners[cid] = [float(next(it)) for _ in range(3)]
            out.append(m)
        return out

    def loadCamera(self, path) -> CamInfo:
        """CylinderTag.cpp:192-196 (cv::FileStorage: cameraMatrix, distCoeffs, both dt: f).  The OpenCV YAML 1.0 matrices
        are parsed here (no OpenCV in the detection modules); dt: f values are float32 like FileStorage returns them."""
        cam = CamInfo(_read_opencv_matrix(path, "cameraMatrix"), _read_opencv_matrix(path, "distCoeffs"))
        return cam

    # ---- Marker Localization (CylinderTag.cpp:198-209, pose_estimation.cpp:50-143) ----
    def estimatePose(self, img, markers, reconstruct_model, camera, useDensePoseRefine=False) -> list:
        poses = []
        for mk in markers:
            p = pnp_solver(mk, reconstruct_model, camera)
            if p.markerID != -1:  # poses with markerID == -1 are erased (CylinderTag.cpp:206-208)
                poses.append(p)
        return poses

    def drawAxis(self, img, markers, reconstruct_model, poses, camera, axisLength=5):
        """CylinderTag.cpp:211-246 without the imshow window: returns the 3-channel overlay image (host code in the
        library: ctag_gray_to_3ch + ctag_draw_axis per pose; pose[i] is paired with markers[i] like the reference)."""
        lib = C.load()
        gray = np.ascontiguousarray(img, np.uint8)
        h, w = gray.shape
        out = np.empty((h, w, 3), np.uint8)
        rc = lib.ctag_gray_to_3ch(gray.ctypes.data, w, h, gray.strides[0], out.ctypes.data, out.strides[0])
        C.check(rc, "drawAxis")
        K = np.ascontiguousarray(camera.Intrinsic, np.float32).reshape(9)
        D = np.ascontiguousarray(camera.distCoeffs, np.float32).reshape(-1)
        for i, pose in enumerate(poses):
            if i >= len(markers) or not 0 <= pose.markerID < len(reconstruct_model):
                continue
            model = reconstruct_model[pose.markerID]
            rec = marker_to_record(markers[i])
            corners3 = np.ascontiguousarray(model.corners, np.float32)
            base = np.ascontiguousarray(model.base, np.float32)
            axis = np.ascontiguousarray(model.axis, np.float32)
            rvec = np.ascontiguousarray(pose.rvec, np.float64).reshape(3)
            tvec = np.ascontiguousarray(pose.tvec, np.float64).reshape(3)
            rc = lib.ctag_draw_axis(out.ctypes.data, w, h, out.strides[0], rec.ctypes.data, corners3.ctypes.data,
                                    corners3.shape[0], base.ctypes.data, axis.ctypes.data, K.ctypes.data, D.ctypes.data,
                                    int(D.size), rvec.ctypes.data, tvec.ctypes.data, int(axisLength))
            C.check(rc, "drawAxis")
        return out


def _read_opencv_matrix(path, name):
    """One `name: !!opencv-matrix` node of an OpenCV YAML 1.0 file -> ndarray (rows x cols; dt f -> float32, d ->
    float64, i -> int32).  A missing node gives an empty array, like an empty FileNode."""
    import re
    try:
        text = open(path).read()
    except OSError:
        return np.zeros((0, 0), np.float32)  # cv::FileStorage on a missing file: loadCamera does not check
    m = re.search(r"^%s\s*:\s*!!opencv-matrix\s*\n(.*?)data\s*:\s*\[(.*?)\]" % re.escape(name), text, re.S | re.M)
    if not m:
        return np.zeros((0, 0), np.float32)
    head = m.group(1)
    rows = int(re.search(r"rows\s*:\s*(\d+)", head).group(1))
    cols = int(re.search(r"cols\s*:\s*(\d+)", head).group(1))
    dt = re.search(r"dt\s*:\s*\"?(\w+)", head).group(1)
    vals = [float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
    dtype = {"f": np.float32, "d": np.float64, "i": np.int32}.get(dt[-1], np.float64)
    return np.array(vals, np.float64).astype(dtype).reshape(rows, cols)


def select_pose_points(mk: MarkerInfo, model: ModelInfo):
    """pose_estimation.cpp:72-95: which corners of which features enter the PnP problem."""
    img_pts, obj_pts = [], []
    n = len(mk.cornerLists)
    for j in range(n):
        bad = abs(mk.feature_ID_left[j] - mk.feature_ID_right[j]) > 1 or mk.feature_ID_right[j] == -1
        if n > 3 and (j == 0 or j == len(mk.feature_ID_left) - 1) and bad:
            continue
        ks = [0, 1, 4, 5]
        if abs(mk.feature_ID_left[j] - mk.feature_ID_right[j]) < 3 and mk.feature_ID_right[j] != -1:
            ks += [2, 3, 6, 7]
        for k in ks:
            img_pts.append(mk.cornerLists[j][k])
            obj_pts.append(model.corners[mk.featurePos[j] * 8 + k])
    return np.array(img_pts, np.float32).reshape(-1, 2), np.array(obj_pts, np.float32).reshape(-1, 3)


def marker_to_record(mk: MarkerInfo):
    """MarkerInfo -> ctag_marker record (the fields the pose stage reads)."""
    rec = np.zeros((), C.MARKER_DTYPE)
    n = len(mk.cornerLists)
    rec["marker_id"], rec["n_features"] = mk.markerID, n
    rec["feature_pos"][:] = -1
    rec["feature_pos"][:n] = mk.featurePos[:n]
    rec["feature_id"][:n] = mk.feature_ID[:n]
    rec["id_left"][:n] = mk.feature_ID_left[:n]
    rec["id_right"][:n] = mk.feature_ID_right[:n]
    rec["corners"][:n] = np.asarray(mk.cornerLists, np.float32).reshape(n, 8, 2)
    return rec


def pnp_solver(mk: MarkerInfo, models, camera: CamInfo) -> PoseInfo:
    """PoseEstimator::PnPSolver + PoseBA (pose_estimation.cpp:50-128) through the library's host-side pose stage
    (ctag_estimate_pose: corner selection, undistortion, EPnP, Levenberg-Marquardt; csrc/pose.cpp)."""
    import ctypes
    idx = next((j for j, m in enumerate(models) if m.MarkerID == mk.markerID), -1)
    if idx < 0:
        return PoseInfo(-1)
    lib = C.load()
    rec = marker_to_record(mk)
    corners3 = np.ascontiguousarray(models[idx].corners, np.float32)
    K = np.ascontiguousarray(camera.Intrinsic, np.float32).reshape(9)
    D = np.ascontiguousarray(camera.distCoeffs, np.float32).reshape(-1)
    rvec, tvec = np.zeros(3, np.float64), np.zeros(3, np.float64)
    rms = ctypes.c_double(0)
    rc = lib.ctag_estimate_pose(rec.ctypes.data, corners3.ctypes.data, corners3.shape[0], K.ctypes.data, D.ctypes.data,
                                int(D.size), rvec.ctypes.data, tvec.ctypes.data, ctypes.addressof(rms))
    if rc != C.OK:
        return PoseInfo(-1)
    p = PoseInfo(idx, rvec, tvec)
    p.rms_px = float(rms.value)
    return p
