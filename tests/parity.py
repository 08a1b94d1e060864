"""Shared parity helpers: run a firing stream through a recording driver and compare two recordings.

Bar (BASELINE.json north_star): ground masks and every per-cell field bit-exact; cluster labelling identical up to
a permutation of the ids (id 0 = "not in a published cluster" must match exactly); finished-column events identical
and in the same order; finished clusters identical as multisets of (stamp, set of points).
"""
from __future__ import annotations

import numpy as np

EXACT_CELL_FIELDS = [
    "continuous_azimuth_angle", "global_column_index", "globally_unique_point_index", "stamp", "firing_index",
    "x", "y", "z", "distance", "azimuth_angle", "inclination_angle", "intensity",
    "ground_point_label", "debug_ground_point_label", "is_ignored",
]
TREE_FIELDS = ["tree_root_gcol", "tree_root_row", "number_of_visited_neighbors"]
# per-tree bookkeeping the ROS node publishes (ros_utils.cpp:287-298), final once a column is published as clustered
PUBLISHED_TREE_FIELDS = ["tree_root_gcol", "tree_root_row", "finished_at_continuous_azimuth_angle", "tree_num_points",
                         "cluster_width", "num_child_points", "local_column_index", "row_index"]


def record(driver, pts, poses, chunk=None, hooks=None):
    """Feeds the stream and returns everything the driver recorded. hooks: {first firing of a chunk: callable(driver)}."""
    n = pts.shape[0]
    chunk = chunk or n
    for a in range(0, n, chunk):
        if hooks and a in hooks:
            hooks[a](driver)
        driver.add_firings(pts[a : a + chunk], poses[a : a + chunk])
    gcols, gcells = driver.ground_columns()
    ccols, ccells = driver.cluster_columns()
    clusters, cpoints = driver.clusters()
    return dict(events=driver.events(), ground_cols=gcols, ground_cells=gcells, cluster_cols=ccols,
                cluster_cells=ccells, clusters=clusters, cluster_points=cpoints,
                reset_required=driver.reset_required())


def _bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        return a.view("u%d" % a.dtype.itemsize)
    return a


def canonical_partition(ids, keys):
    """Relabels cluster ids by the smallest key (globally unique point index) of each cluster; 0 stays 0."""
    ids = np.asarray(ids).ravel()
    keys = np.asarray(keys).ravel()
    out = np.zeros(ids.shape, dtype=np.uint64)
    nz = ids != 0
    if nz.any():
        order = np.lexsort((keys[nz], ids[nz]))
        sid = ids[nz][order]
        skey = keys[nz][order]
        first = np.r_[True, sid[1:] != sid[:-1]]
        rep = skey[first]  # min key per id
        idx = np.cumsum(first) - 1
        canon = np.empty(sid.shape, dtype=np.uint64)
        canon[:] = rep[idx] + 1
        tmp = np.empty(sid.shape, dtype=np.uint64)
        tmp[order] = canon
        out[nz] = tmp
    return out


def cluster_multiset(rec):
    out = []
    cl, pt = rec["clusters"], rec["cluster_points"]
    for c in cl:
        p = pt[int(c["point_offset"]) : int(c["point_offset"]) + int(c["num_points"])]
        key = tuple(sorted(zip(p["gcol"].tolist(), p["row"].tolist())))
        out.append((int(c["stamp"]), int(c["event_index"]), key))
    return sorted(out)


def compare(a, b, check_tree_fields=False, name_a="a", name_b="b", check_published_tree_fields=False):
    """Raises AssertionError with a readable message on the first difference."""
    assert a["reset_required"] == b["reset_required"], "reset_required differs"
    ea, eb = a["events"], b["events"]
    assert len(ea) == len(eb), f"event count {len(ea)} vs {len(eb)}"
    for f in ("from_gcol", "to_gcol", "ground_points_only"):
        if not np.array_equal(ea[f], eb[f]):
            i = int(np.nonzero(ea[f] != eb[f])[0][0])
            raise AssertionError(f"event {i} field {f}: {ea[i]} vs {eb[i]}")
    for kind in ("ground", "cluster"):
        ca, cb = a[kind + "_cols"], b[kind + "_cols"]
        assert np.array_equal(ca, cb), f"{kind} column list differs"
        xa, xb = a[kind + "_cells"], b[kind + "_cells"]
        fields = list(EXACT_CELL_FIELDS) + (TREE_FIELDS if check_tree_fields and kind == "cluster" else [])
        if check_published_tree_fields and kind == "cluster":
            fields += [f for f in PUBLISHED_TREE_FIELDS if f not in fields]
        for f in fields:
            ba, bb = _bits(xa[f]), _bits(xb[f])
            if f in ("continuous_azimuth_angle", "x", "y", "z", "distance", "azimuth_angle", "inclination_angle"):
                # NaN payloads may differ; compare NaN-ness + bits of the rest
                na, nb = np.isnan(xa[f]), np.isnan(xb[f])
                ok = (na == nb) & (na | (ba == bb))
            else:
                ok = ba == bb
            if not ok.all():
                i = np.argwhere(~ok)[0]
                raise AssertionError(
                    f"{kind} column {int(ca[i[0]])} row {int(i[1])} field {f}: "
                    f"{name_a}={xa[f][tuple(i)]!r} {name_b}={xb[f][tuple(i)]!r} ({int((~ok).sum())} cells differ)")
        if kind == "cluster" and xa.size:
            pa = canonical_partition(xa["id"], xa["globally_unique_point_index"])
            pb = canonical_partition(xb["id"], xb["globally_unique_point_index"])
            if not np.array_equal(pa, pb):
                i = np.nonzero(pa != pb)[0][0]
                raise AssertionError(f"cluster partition differs at flat cell {int(i)} ({int((pa != pb).sum())} cells)")
    ma, mb = cluster_multiset(a), cluster_multiset(b)
    assert len(ma) == len(mb), f"finished cluster count {len(ma)} vs {len(mb)}"
    for i, (x, y) in enumerate(zip(ma, mb)):
        assert x == y, f"finished cluster {i} differs: stamp/event {x[:2]} vs {y[:2]}, sizes {len(x[2])} vs {len(y[2])}"
    return True
