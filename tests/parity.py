"""Shared comparison helpers: CUDA path (through the C ABI) vs oracle DetectDump.  Bars (BASELINE.json north_star):
candidate sets, IDs, positions: bit-exact; sub-pixel corners: <= 1e-3 px."""
import numpy as np

CORNER_TOL = 1e-3


def marker_rows(markers, count):
    return [markers[k] for k in range(count)]


def assert_markers_match(got_markers, got_count, oracle_markers, tol=CORNER_TOL, ctx=""):
    assert got_count == len(oracle_markers), f"{ctx}: marker count {got_count} vs oracle {len(oracle_markers)}"
    worst = 0.0
    for k, m in enumerate(oracle_markers):
        g = got_markers[k]
        n = int(g["n_features"])
        assert n == len(m.cornerLists), f"{ctx}: marker {k} feature count"
        assert int(g["marker_id"]) == m.markerID, f"{ctx}: marker {k} id {int(g['marker_id'])} vs {m.markerID}"
        assert bool(g["inverse"]) == bool(m.inverse), f"{ctx}: marker {k} inverse flag"
        assert list(g["feature_pos"][:len(m.featurePos)]) == list(m.featurePos), f"{ctx}: marker {k} featurePos"
        assert list(g["feature_id"][:n]) == list(m.feature_ID), f"{ctx}: marker {k} feature_ID"
        assert list(g["id_left"][:n]) == list(m.feature_ID_left), f"{ctx}: marker {k} ID_left"
        assert list(g["id_right"][:n]) == list(m.feature_ID_right), f"{ctx}: marker {k} ID_right"
        ref = np.array(m.cornerLists, np.float32)
        d = float(np.abs(g["corners"][:n] - ref).max())
        worst = max(worst, d)
        assert d <= tol, f"{ctx}: marker {k} corner error {d}"
        assert np.allclose(g["center"][:n], np.array(m.feature_center, np.float32), atol=tol)
        assert np.allclose(g["cr_left"][:n], np.array(m.cr_left, np.float32), rtol=1e-4, atol=1e-4)
        assert np.allclose(g["cr_right"][:n], np.array(m.cr_right, np.float32), rtol=1e-4, atol=1e-4)
        assert np.allclose(g["edge_length"][:n], np.array(m.edge_length, np.float32), rtol=1e-5, atol=tol)
    return worst


def assert_frame_matches(det, f, info, markers, counts, dump, subpix, ctx=""):
    """det: Detector after a batch; dump: oracle DetectDump of frame f."""
    assert np.array_equal(det.debug_binary(f), dump.binary), f"{ctx}: binary"
    assert int(info["n_labels"][f]) == dump.n_labels, f"{ctx}: n_labels"
    comps = det.debug_components(f)
    ref = np.array([[c.area, c.x0, c.y0, c.x1, c.y1] for c in dump.comps], np.int32).reshape(-1, 5)
    assert np.array_equal(comps[:, 1:], ref), f"{ctx}: legal components"
    idx, quads = det.debug_quads(f)
    assert np.array_equal(idx, np.array(dump.quad_comp, np.int32)), f"{ctx}: quad candidate set"
    if len(dump.quads):
        assert np.abs(quads - np.array(dump.quads)).max() <= CORNER_TOL, f"{ctx}: quad corners"
    if dump.flagged:
        assert int(info["flagged"][f]) == 1, f"{ctx}: flagged"
        return 0.0
    status = {"ok": 0, "no_corner": 1, "no_feature": 2}[dump.status]
    assert int(info["status"][f]) == status, f"{ctx}: status {int(info['status'][f])} vs {dump.status}"
    if status != 0:
        assert int(counts[f]) == 0
        return 0.0
    cor, cen, ang, qp = det.debug_features(f)
    assert len(cor) == len(dump.feats), f"{ctx}: feature count"
    assert np.array_equal(qp, np.array([[ft.quad_i, ft.quad_j] for ft in dump.feats], np.int32).reshape(-1, 2)), f"{ctx}: pairing"
    ref = np.array([ft.corners for ft in dump.feats], np.float32).reshape(-1, 8, 2)
    assert np.abs(cor - ref).max() <= CORNER_TOL, f"{ctx}: feature corners {np.abs(cor - ref).max()}"
    assert int(info["n_groups"][f]) == len(dump.groups), f"{ctx}: groups"
    return assert_markers_match(markers[f], int(counts[f]), dump.markers, ctx=ctx)
