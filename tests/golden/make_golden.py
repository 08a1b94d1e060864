"""Generates tests/golden/*.npz from the REFERENCE's own sources (oracle/_ref/libcc_ref.so, built by
`make -C oracle ref` from /root/reference): the reference has no tests or golden vectors of its own (SURVEY.md
section 4), so these recordings of its deterministic single-threaded mode are what pins the oracle and the CUDA path.

    python tests/golden/make_golden.py

Each fixture stores the INPUT stream parameters (the stream itself is regenerated from the seed by
continuous_clustering_b200.synth) and the recorded outputs: finished-column events, per-cell fields of every
column reported by a callback (bit patterns), cluster ids, finished clusters."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import parity  # noqa: E402
from continuous_clustering_b200 import synth  # noqa: E402
from oracle import drvlib  # noqa: E402

FIXTURES = {
    "tiny16_static": ("tiny16", dict(n_rotations=2.0, seed=7), {}),
    "tiny16_moving_dropout": ("tiny16", dict(n_rotations=2.5, seed=11, moving=True, dropout=0.08), {}),
    "tiny16_wall_forced_finish": ("tiny16", dict(n_rotations=3.0, seed=3, n_boxes=0, wall_radius=8.0), {}),
    "tiny16_every_3rd_column": ("tiny16", dict(n_rotations=2.0, seed=5), dict(cluster_point_trees_every_nth_column=3)),
    "os32_left_short": ("os32_left", dict(n_firings=700, seed=21, moving=True), {}),
}

KEEP_CELL_FIELDS = parity.EXACT_CELL_FIELDS + ["id", "tree_root_gcol", "tree_root_row"]


def pack(rec):
    out = {"events": rec["events"], "ground_cols": rec["ground_cols"], "cluster_cols": rec["cluster_cols"],
           "clusters": rec["clusters"], "cluster_points": rec["cluster_points"],
           "reset_required": np.array(rec["reset_required"])}
    for kind in ("ground", "cluster"):
        cells = rec[kind + "_cells"]
        for f in KEEP_CELL_FIELDS:
            a = np.ascontiguousarray(cells[f])
            out[f"{kind}__{f}"] = a.view("u%d" % a.dtype.itemsize) if a.dtype.kind == "f" else a
    return out


def unpack(npz):
    rec = {"events": npz["events"], "ground_cols": npz["ground_cols"], "cluster_cols": npz["cluster_cols"],
           "clusters": npz["clusters"], "cluster_points": npz["cluster_points"],
           "reset_required": bool(npz["reset_required"])}
    for kind in ("ground", "cluster"):
        shape = npz[f"{kind}__id"].shape
        cells = np.zeros(shape, dtype=drvlib.CELL_DTYPE)
        for f in KEEP_CELL_FIELDS:
            a = npz[f"{kind}__{f}"]
            dt = drvlib.CELL_DTYPE[f]
            cells[f] = a.view(dt) if dt.kind == "f" else a
        rec[kind + "_cells"] = cells
    return rec


def stream_for(name):
    spec, kw, cfg_over = FIXTURES[name]
    pts, poses, sp = synth.make_stream(spec, **kw)
    return pts, poses, sp, drvlib.stream_config(spec, **cfg_over)


if __name__ == "__main__":
    assert drvlib.have_ref(), "build oracle/_ref first: make -C oracle ref"
    for name in FIXTURES:
        pts, poses, sp, cfg = stream_for(name)
        d = drvlib.Driver(drvlib.REF_LIB)
        d.configure(cfg, sp.rows)
        rec = parity.record(d, pts, poses)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **pack(rec))
        print(name, "events", len(rec["events"]), "clusters", len(rec["clusters"]), os.path.getsize(path) // 1024, "KiB")
