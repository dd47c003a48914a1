"""estimatePose ORACLE (TEST INFRASTRUCTURE ONLY): PoseEstimator::PnPSolver + PoseBA (pose_estimation.cpp:50-143) on
the oracle's Marker objects, with the real OpenCV calls (solvePnP EPNP, undistortPoints with P=K) and scipy's
Levenberg-Marquardt in place of Ceres (same 2-residual pinhole cost, pose_estimation.cpp:14-41).  The reference has no
pose test vectors; pinned only by reprojection RMS against the shipped .model (0.12-0.32 px)."""
import cv2
import numpy as np
from scipy.optimize import least_squares


def load_model(path):
    toks = open(path).read().split()
    it = iter(toks)
    n, size = int(next(it)), int(next(it))
    out = []
    for _ in range(n):
        mid = int(next(it))
        base = [float(next(it)) for _ in range(3)]
        axis = [float(next(it)) for _ in range(3)]
        corners = np.zeros((8 * size, 3), np.float32)
        for _ in range(8 * size):
            cid = int(next(it))
            corners[cid] = [float(next(it)) for _ in range(3)]
        out.append((mid, np.array(base, np.float32), np.array(axis, np.float32), corners))
    return out


def load_camera(path):
    fs = cv2.FileStorage(str(path), cv2.FILE_STORAGE_READ)
    K, D = fs.getNode("cameraMatrix").mat(), fs.getNode("distCoeffs").mat()
    fs.release()
    return K, D


def estimate_pose(markers, models, K, D):
    """Returns [(model_index, rvec, tvec, rms_px)] for markers whose ID has a model (others are erased)."""
    out = []
    for mk in markers:
        idx = next((j for j, m in enumerate(models) if m[0] == mk.markerID), -1)
        if idx < 0:
            continue
        corners3 = models[idx][3]
        ip, op = [], []
        n = len(mk.cornerLists)
        for j in range(n):
            bad = abs(mk.feature_ID_left[j] - mk.feature_ID_right[j]) > 1 or mk.feature_ID_right[j] == -1
            if n > 3 and (j == 0 or j == n - 1) and bad:
                continue
            ks = [0, 1, 4, 5]
            if abs(mk.feature_ID_left[j] - mk.feature_ID_right[j]) < 3 and mk.feature_ID_right[j] != -1:
                ks += [2, 3, 6, 7]
            for k in ks:
                ip.append(mk.cornerLists[j][k])
                op.append(corners3[mk.featurePos[j] * 8 + k])
        ip = np.array(ip, np.float64).reshape(-1, 2)
        op = np.array(op, np.float64).reshape(-1, 3)
        Kd, Dd = K.astype(np.float64), D.astype(np.float64)
        ok, rvec, tvec = cv2.solvePnP(op, ip, Kd, Dd, flags=cv2.SOLVEPNP_EPNP)
        und = cv2.undistortPoints(ip.reshape(-1, 1, 2), Kd, Dd, P=Kd).reshape(-1, 2)
        fx, fy, cx, cy = (float(np.float32(Kd[0, 0])), float(np.float32(Kd[1, 1])), float(np.float32(Kd[0, 2])),
                          float(np.float32(Kd[1, 2])))

        def resid(p):
            R, _ = cv2.Rodrigues(p[:3])
            Pc = op @ R.T + p[3:]
            return np.concatenate([fx * Pc[:, 0] / Pc[:, 2] + cx - und[:, 0], fy * Pc[:, 1] / Pc[:, 2] + cy - und[:, 1]])

        sol = least_squares(resid, np.concatenate([rvec.reshape(3), tvec.reshape(3)]), method="lm", xtol=1e-12, ftol=1e-15, gtol=1e-15)
        r = resid(sol.x)
        rms = float(np.sqrt(np.mean(r[:len(ip)] ** 2 + r[len(ip):] ** 2)))
        out.append((idx, sol.x[:3].copy(), sol.x[3:].copy(), rms))
    return out
