"""Golden fixture for BASELINE.json config 2 (the substitute for the missing test.avi): the Python/cv2 oracle on every
frame of cylindertag_b200.synth.video_sequence(test_gray, 120, seed=2024).

  python tests/golden/make_golden_sequence.py          (about a minute)

Output: sequence_detect.npz
  counts        [120][6]   n_labels, legal components, quads, features, groups, markers
  flagged       [120]      frame exceeds a reference fixed-array limit (excluded from parity)
  marker_start  [121]      markers of frame f are rows marker_start[f] .. marker_start[f+1] of the arrays below
  marker_id, inverse, n_features   [M]
  feature_pos, feature_id, id_left, id_right   [M][20]  (-1 / 0 padded)
  corners       [M][20][8][2] float32 (zero padded)
Only the frames are regenerated at test time (seeded); nothing here needs /root/reference."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from cylindertag_b200 import synth  # noqa: E402
from oracle import ctag_oracle as o  # noqa: E402

N_FRAMES, SEED = 120, 2024


def main():
    gray = cv2.imread(os.path.join(HERE, "data", "test_gray.png"), cv2.IMREAD_UNCHANGED)
    state, fs = o.load_marker_file(os.path.join(HERE, "data", "CTag_2f12c.marker"))
    seq = synth.video_sequence(gray, N_FRAMES, SEED)
    counts, flagged, start = [], [], [0]
    mid, inv, nfe, fpos, fid, il, ir, cor = [], [], [], [], [], [], [], []
    for f in range(N_FRAMES):
        d = o.detect(seq[f], state, fs, 5, True, 5)
        counts.append([d.n_labels, len(d.comps), len(d.quads), len(d.feats), len(d.groups), len(d.markers)])
        flagged.append(int(bool(d.flagged)))
        for m in d.markers:
            n = len(m.cornerLists)
            mid.append(m.markerID), inv.append(int(m.inverse)), nfe.append(n)
            pad = lambda v, fill: list(v)[:20] + [fill] * (20 - len(list(v)[:20]))
            fpos.append(pad(m.featurePos, -1)), fid.append(pad(m.feature_ID, 0))
            il.append(pad(m.feature_ID_left, 0)), ir.append(pad(m.feature_ID_right, 0))
            c = np.zeros((20, 8, 2), np.float32)
            c[:n] = np.asarray(m.cornerLists, np.float32).reshape(n, 8, 2)
            cor.append(c)
        start.append(len(mid))
        print(f, counts[-1], [m.markerID for m in d.markers], flush=True)
    np.savez_compressed(os.path.join(HERE, "sequence_detect.npz"), counts=np.array(counts, np.int32),
                        flagged=np.array(flagged, np.int32), marker_start=np.array(start, np.int32),
                        marker_id=np.array(mid, np.int32), inverse=np.array(inv, np.int32), n_features=np.array(nfe, np.int32),
                        feature_pos=np.array(fpos, np.int32), feature_id=np.array(fid, np.int32), id_left=np.array(il, np.int32),
                        id_right=np.array(ir, np.int32), corners=np.array(cor, np.float32))


if __name__ == "__main__":
    main()
