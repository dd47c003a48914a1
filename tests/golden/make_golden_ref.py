"""Golden fixtures produced by the REFERENCE'S OWN CODE: /root/reference's corner_detector.cpp + CylinderTag.cpp compiled
unmodified into oracle/_ref/libctag_ref.so (oracle/build_ref.py) and run on every frame of BASELINE.json configs 1-4.
Run in the build container (the only place /root/reference exists):

  python tests/golden/make_golden_ref.py [--no-crosscheck]          (a few minutes)

Each frame is also run through the Python oracle (oracle/ctag_oracle.py) and the two must agree (integers exactly,
floats to 1e-6 px) -- the script fails otherwise -- so the goldens double as the record that the oracle used for the
stage-level GPU comparisons is pinned to the reference on all of these frames.

Outputs (tests/golden/):
  ref_testbmp.npz     config 1: test.bmp, detect(gray, 5, true, 5) -- plus the label image (packed) and refined features
  ref_sequence.npz    config 2 substitute: synth.video_sequence(test_gray, 120, seed=2024)
  ref_config3.npz     config 3: 256 frames 1080p BGR, seeds 1000..1255 (tests/configs.py)
  ref_config4_<codebook>.npz   config 4: 8 frames 4K BGR, 4..8 markers, codebooks 2f12c / 15c3f / 18c4f
Layout of every file ("results layout"):
  counts       [F][8]  n_labels, n_legal, n_quads, n_features, n_groups, n_markers, status (0 ok, 1 no corner, 2 no
                       feature), flagged
  binary_crc   [F]     zlib.crc32 of the half-resolution binary image bytes ({0,255}, row-major)
  comp_start   [F+1], comps [C][5]        legal components in reference order: area, x0, y0, x1, y1
  quad_start   [F+1], quads [Q][4][2]     quad candidates in reference order (half-res coordinates)
  feat_start   [F+1], feats [N][8][2]     features as detect() left them (refined)
  marker_start [F+1]; marker_id, inverse, n_features [M]; feature_pos, feature_id, id_left, id_right [M][20];
  cr_left, cr_right, edge_length [M][20]; center [M][20][2]; corners [M][20][8][2]
Frames are regenerated (seeded) at test time; nothing here is needed at run time except the .npz files."""
import os
import sys
import time
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from cylindertag_b200 import synth  # noqa: E402
from oracle import ctag_oracle as o  # noqa: E402
from oracle import ref_api as R  # noqa: E402
from tests import configs  # noqa: E402
from tests.test_ref_pinning import _oracle_equals_ref  # noqa: E402

PAD = 20


def pad(v, fill, dt):
    a = np.full(PAD, fill, dt)
    v = list(v)[:PAD]
    a[:len(v)] = v
    return a


class Results:
    def __init__(self):
        self.counts, self.crc = [], []
        self.comp_start, self.comps = [0], []
        self.quad_start, self.quads = [0], []
        self.feat_start, self.feats = [0], []
        self.marker_start = [0]
        self.m = {k: [] for k in ("marker_id", "inverse", "n_features", "feature_pos", "feature_id", "id_left", "id_right", "cr_left",
                                  "cr_right", "edge_length", "center", "corners")}

    def add(self, d, state):
        self.counts.append([d.n_labels, len(d.comps), len(d.quads), len(d.feats), d.n_groups, len(d.markers),
                            ("ok", "no_corner", "no_feature").index(d.status), int(d.flagged)])
        self.crc.append(zlib.crc32(np.ascontiguousarray(d.binary).tobytes()))
        self.comps += d.comps.tolist()
        self.comp_start.append(len(self.comps))
        self.quads += d.quads.tolist()
        self.quad_start.append(len(self.quads))
        self.feats += d.feats.tolist()
        self.feat_start.append(len(self.feats))
        for mk in d.markers:
            n = len(mk.cornerLists)
            self.m["marker_id"].append(mk.markerID)
            self.m["inverse"].append(R.inverse_flag(state, mk.markerID, mk.featurePos, mk.feature_ID))
            self.m["n_features"].append(n)
            self.m["feature_pos"].append(pad(mk.featurePos, -1, np.int32))
            self.m["feature_id"].append(pad(mk.feature_ID, 0, np.int32))
            self.m["id_left"].append(pad(mk.feature_ID_left, 0, np.int32))
            self.m["id_right"].append(pad(mk.feature_ID_right, 0, np.int32))
            self.m["cr_left"].append(pad(mk.cr_left, 0, np.float32))
            self.m["cr_right"].append(pad(mk.cr_right, 0, np.float32))
            self.m["edge_length"].append(pad(mk.edge_length, 0, np.float32))
            c = np.zeros((PAD, 2), np.float32)
            c[:min(n, PAD)] = mk.feature_center[:PAD]
            self.m["center"].append(c)
            c = np.zeros((PAD, 8, 2), np.float32)
            c[:min(n, PAD)] = mk.cornerLists[:PAD]
            self.m["corners"].append(c)
        self.marker_start.append(len(self.m["marker_id"]))

    def save(self, path, **extra):
        dt = {"marker_id": np.int32, "inverse": np.int32, "n_features": np.int32, "feature_pos": np.int32, "feature_id": np.int32,
              "id_left": np.int32, "id_right": np.int32}
        out = {k: np.array(v, dt.get(k, np.float32)) for k, v in self.m.items()}
        if not len(self.m["marker_id"]):
            out["feature_pos"] = np.zeros((0, PAD), np.int32)
        np.savez_compressed(path, counts=np.array(self.counts, np.int32), binary_crc=np.array(self.crc, np.uint32),
                            comp_start=np.array(self.comp_start, np.int32), comps=np.array(self.comps, np.int32).reshape(-1, 5),
                            quad_start=np.array(self.quad_start, np.int32), quads=np.array(self.quads, np.float32).reshape(-1, 4, 2),
                            feat_start=np.array(self.feat_start, np.int32), feats=np.array(self.feats, np.float32).reshape(-1, 8, 2),
                            marker_start=np.array(self.marker_start, np.int32), **out, **extra)


def run_set(name, grays, state, fs, crosscheck, **extra):
    ref = R.RefDetector(state=state, feature_size=fs)
    res = Results()
    t0 = time.time()
    for f, g in enumerate(grays):
        d = ref.detect(g, 5, True, 5)
        if crosscheck:
            _oracle_equals_ref(d, o.detect(g, state, fs, 5, True, 5))
        res.add(d, state)
    ref.close()
    res.save(os.path.join(HERE, name), **extra)
    c = np.array(res.counts)
    print(f"{name}: {len(grays)} frames, {int(c[:, 5].sum())} markers, {int(c[:, 7].sum())} flagged, "
          f"{'cross-checked against the Python oracle, ' if crosscheck else ''}{time.time() - t0:.0f} s", flush=True)


def main():
    crosscheck = "--no-crosscheck" not in sys.argv
    gray = cv2.imread(os.path.join(HERE, "data", "test_gray.png"), cv2.IMREAD_UNCHANGED)
    state, fs = configs.codebook("2f12c")
    # config 1, with the full label image
    ref = R.RefDetector(marker_path=os.path.join(HERE, "data", "CTag_2f12c.marker"))
    d = ref.detect(gray, 5, True, 5)
    d0 = ref.detect(gray, 5, False, 3)
    ref.load_model_camera(os.path.join(HERE, "data", "CTag_2f12c.model"), os.path.join(HERE, "data", "cameraParams.yml"))
    ref.detect(gray, 5, True, 5)
    with R.cv2_backend(only=["solve_pnp_epnp", "undistort_points"]):
        poses = ref.estimate_pose()
    ref.close()
    run_set("ref_testbmp.npz", [gray], state, fs, crosscheck, labels_packed=np.packbits(d.labels > 0),
            labels_max=np.int32(d.labels.max()), feats_unrefined=d0.feats,
            pose_model_index=np.array([p[0] for p in poses], np.int32), pose_rvec=np.array([p[1] for p in poses]),
            pose_tvec=np.array([p[2] for p in poses]))
    run_set("ref_sequence.npz", list(synth.video_sequence(gray, 120, 2024)), state, fs, crosscheck)
    frames = configs.render_many([(3, None, i) for i in range(configs.CONFIG3_FRAMES)])
    run_set("ref_config3.npz", [o.bgr2gray(f) for f in frames], state, fs, crosscheck)
    for name in configs.CODEBOOKS:
        st, f_s = configs.codebook(name)
        frames = configs.render_many([(4, name, i) for i in range(configs.CONFIG4_FRAMES)])
        run_set(f"ref_config4_{name}.npz", [o.bgr2gray(f) for f in frames], st, f_s, crosscheck, dictionary=st, feature_size=np.int32(f_s))


if __name__ == "__main__":
    main()
