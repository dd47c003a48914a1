"""Flatten an oracle DetectDump into plain numpy arrays (golden .npz layout).  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def markers_to_arrays(markers, prefix):
    out = {}
    out[prefix + "id"] = np.array([m.markerID for m in markers], np.int32)
    out[prefix + "inverse"] = np.array([int(m.inverse) for m in markers], np.int32)
    out[prefix + "nfeat"] = np.array([len(m.cornerLists) for m in markers], np.int32)
    cat = lambda key, dt: np.array([v for m in markers for v in getattr(m, key)], dt)
    out[prefix + "feature_pos"] = cat("featurePos", np.int32)
    out[prefix + "feature_id"] = cat("feature_ID", np.int32)
    out[prefix + "id_left"] = cat("feature_ID_left", np.int32)
    out[prefix + "id_right"] = cat("feature_ID_right", np.int32)
    out[prefix + "cr_left"] = cat("cr_left", np.float32)
    out[prefix + "cr_right"] = cat("cr_right", np.float32)
    out[prefix + "edge_length"] = cat("edge_length", np.float32)
    out[prefix + "center"] = np.array([c for m in markers for c in m.feature_center], np.float32).reshape(-1, 2)
    out[prefix + "corners"] = np.array([c for m in markers for c in m.cornerLists], np.float32).reshape(-1, 8, 2)
    return out


def dump_to_dict(d):
    out = {
        "half": d.half,
        "binary": d.binary,
        "n_labels": np.int32(d.n_labels),
        "comps": np.array([[c.label, c.area, c.x0, c.y0, c.x1, c.y1] for c in d.comps], np.int32).reshape(-1, 6),
        "quads": np.array(d.quads, np.float32).reshape(-1, 4, 2),
        "quad_comp": np.array(d.quad_comp, np.int32),
        "feats_half": np.array(d.feats_half, np.float32).reshape(-1, 8, 2),
        "feats_init": np.array(d.feats_init, np.float32).reshape(-1, 8, 2),
        "feats_refined": np.array([f.corners for f in d.feats], np.float32).reshape(-1, 8, 2),
        "feats_center": np.array([f.center for f in d.feats], np.float32).reshape(-1, 2),
        "feats_angle": np.array([f.angle for f in d.feats], np.float32),
        "feats_quads": np.array([[f.quad_i, f.quad_j] for f in d.feats], np.int32).reshape(-1, 2),
        "status": np.array(d.status),
        "flagged": np.int32(d.flagged),
    }
    out.update(markers_to_arrays(d.groups, "grp_"))
    out.update(markers_to_arrays(d.markers, "mk_"))
    return out
