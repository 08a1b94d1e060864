"""Runs a firing stream through continuous_clustering_b200.ContinuousClustering and records the same things the
oracle's recording driver (oracle/cc_driver.h) records, so tests/parity.py can compare them."""
from __future__ import annotations

import numpy as np

from oracle import drvlib

_MAP = {  # drv_cell_t field -> cc_read_columns field
    "continuous_azimuth_angle": "continuous_azimuth_angle", "global_column_index": "global_column_index",
    "globally_unique_point_index": "globally_unique_point_index", "stamp": "stamp", "firing_index": "firing_index",
    "id": "id", "tree_root_gcol": "tree_root_gcol", "distance": "distance", "azimuth_angle": "azimuth_angle",
    "inclination_angle": "inclination_angle", "tree_root_row": "tree_root_row", "intensity": "intensity",
    "ground_point_label": "ground_point_label", "debug_ground_point_label": "debug_ground_point_label",
    "is_ignored": "is_ignored", "number_of_visited_neighbors": "number_of_visited_neighbors",
    "finished_at_continuous_azimuth_angle": "finished_at_continuous_azimuth_angle", "tree_num_points": "tree_num_points",
    "cluster_width": "cluster_width", "x": "x", "y": "y", "z": "z",
}


def _to_cells(cells, ring_buffer_max_columns):
    """cc_cell_t records (cc_export_columns) -> the recording driver's drv_cell_t."""
    out = np.zeros(cells.shape, dtype=drvlib.CELL_DTYPE)
    for k, v in _MAP.items():
        out[k] = cells[v]
    # what the facade derives for `Point` (facade/src/continuous_clustering.cpp materialise())
    g = cells["global_column_index"]
    out["local_column_index"] = np.where(g >= 0, g % ring_buffer_max_columns, -1)
    out["row_index"] = np.where(np.isnan(cells["distance"]), -1, np.arange(cells.shape[1])[None, :])
    return out


def child_counts(cc, lo, hi, max_steps_in_row=20):
    """child_points.size() of the cells of columns [lo, hi]: children name their parent (first_parent_*), and sit at most
    max_steps_in_row columns ahead of it."""
    ext = cc.export_columns(lo, hi + max_steps_in_row).copy()
    rows = ext.shape[1]
    cnt = np.zeros((hi - lo + 1, rows), dtype=np.int32)
    pg, pr = ext["first_parent_gcol"], ext["first_parent_row"]
    sel = (pg >= lo) & (pg <= hi)
    np.add.at(cnt, (pg[sel] - lo, pr[sel]), 1)
    return cnt


def record(cc, pts, poses, chunk, pipelined=False, hooks=None):
    """Feeds the stream in pushes of `chunk` firings; returns the dict layout of tests/parity.record().
    pipelined=True uses the asynchronous API with two pushes in flight (submit(k + 1); wait(k)), pipelined=2 keeps a third
    push staged (submit(k + 2); wait(k))."""
    rows = pts.shape[1]
    events, gcols, gcells, ccols, ccells, clusters, cpoints = [], [], [], [], [], [], []
    n_events = 0
    used_exact = 0
    slow_firings = 0
    starts = list(range(0, pts.shape[0], chunk))
    depth = int(pipelined) if pipelined else 0  # True: submit(k + 1); wait(k). 2: submit(k + 2); wait(k) (third push staged)
    for b in starts[:depth]:
        cc.submitFirings(pts[b : b + chunk], poses[b : b + chunk])
    for i, a in enumerate(starts):
        if pipelined:
            if i + depth < len(starts):
                b = starts[i + depth]
                cc.submitFirings(pts[b : b + chunk], poses[b : b + chunk])
            res = cc.wait()
        else:
            if hooks and a in hooks:  # e.g. setConfiguration between two pushes (synchronous pushes only)
                hooks[a](cc)
            res = cc.addFirings(pts[a : a + chunk], poses[a : a + chunk])
        used_exact += int(res.info.used_exact_path)
        slow_firings += int(res.info.slow_insert_firings)
        ev = res.events.copy()  # results are views of the handle's buffers, valid until the next push
        if len(ev) == 0:
            continue
        g = ev[ev["ground_points_only"] == 1]
        if len(g):
            lo, hi = int(g["from_gcol"].min()), int(g["to_gcol"].max())
            cells = _to_cells(cc.export_columns(lo, hi), cc.ring_buffer_max_columns)
            gcols.append(np.arange(lo, hi + 1))
            gcells.append(cells)
        c = ev[(ev["ground_points_only"] == 0) & (ev["to_gcol"] >= ev["from_gcol"])]
        if len(c):
            lo, hi = int(c["from_gcol"].min()), int(c["to_gcol"].max())
            cells = _to_cells(cc.export_columns(lo, hi), cc.ring_buffer_max_columns)
            cells["num_child_points"] = child_counts(cc, lo, hi)
            ccols.append(np.arange(lo, hi + 1))
            ccells.append(cells)
        # finished clusters the reference would hand to the callback (> 20 points, cpp:1023), with the number of
        # column events delivered before each
        nxt = 0
        for i, e in enumerate(ev):
            while nxt < int(e["n_clusters_before"]):
                cl = res.clusters[nxt]
                nxt += 1
                if cl["num_points"] > 20:
                    p = res.cluster_points[int(cl["point_offset"]) : int(cl["point_offset"]) + int(cl["num_points"])]
                    off = sum(len(x) for x in cpoints)
                    rec = np.zeros(1, dtype=drvlib.CLUSTER_DTYPE)
                    rec["stamp"], rec["id"], rec["point_offset"], rec["num_points"] = cl["stamp"], cl["id"], off, len(p)
                    rec["event_index"] = n_events + i
                    clusters.append(rec)
                    q = np.zeros(len(p), dtype=drvlib.CLUSTER_POINT_DTYPE)
                    q["gcol"], q["row"] = p["gcol"], p["row"]
                    cpoints.append(q)
        n_events += len(ev)
        events.append(ev)

    def cat(xs, dt, shape=(0,)):
        return np.concatenate(xs) if xs else np.zeros(shape, dtype=dt)

    return dict(
        events=cat(events, drvlib.EVENT_DTYPE),
        ground_cols=cat(gcols, np.int64), ground_cells=cat(gcells, drvlib.CELL_DTYPE, (0, rows)),
        cluster_cols=cat(ccols, np.int64), cluster_cells=cat(ccells, drvlib.CELL_DTYPE, (0, rows)),
        clusters=cat(clusters, drvlib.CLUSTER_DTYPE), cluster_points=cat(cpoints, drvlib.CLUSTER_POINT_DTYPE),
        reset_required=cc.resetRequired(), used_exact_path=used_exact, slow_insert_firings=slow_firings,
    )
