"""Randomised scene sweep of the CPU emulation of the kernels against the oracle restatement (not part of the default test run:
minutes of CPU time). Covers what the committed cases sample: closed walls at several ranges (forced finish once per rotation),
two concentric walls broken by boxes (several ring-spanning components with forced finishes at unrelated columns), finish passes
every n-th column, azimuth jitter, moving sensor, drop-outs, push sizes from 1 to more than two rotations.
Usage: python tests/tools/fuzz_scenes.py [walls|two_walls|nth|all]"""
import itertools
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

import parity  # noqa: E402
import recorder  # noqa: E402
from continuous_clustering_b200 import _lib, synth  # noqa: E402
from oracle import drvlib  # noqa: E402
from test_emu_parity import make_cc  # noqa: E402

REPO = os.path.dirname(os.path.dirname(HERE))
EMU = _lib.load_library(os.path.join(REPO, "tests", "emu", "libcc_b200_emu_test.so"))
ORC = os.path.join(REPO, "oracle", "libcc_oracle.so")


def run(pts, poses, sp, cfg, chunk, what):
    d = drvlib.Driver(ORC)
    d.configure(cfg, sp.rows)
    want = parity.record(d, pts, poses)
    d.close()
    cc = make_cc(EMU, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, chunk)
    cc.close()
    try:
        parity.compare(want, got, name_a="oracle", name_b="emulated kernels", check_tree_fields=True, check_published_tree_fields=True)
    except AssertionError as e:
        print("MISMATCH", what, str(e)[:400])
        sys.exit(1)
    return got["used_exact_path"]


def merge_nearer(a, b):
    da = np.sqrt(a["x"] ** 2 + a["y"] ** 2 + a["z"] ** 2)
    db = np.sqrt(b["x"] ** 2 + b["y"] ** 2 + b["z"] ** 2)
    use_a = (~np.isnan(da)) & (np.isnan(db) | (da < db))
    out = b.copy()
    for f in ("x", "y", "z"):
        out[f] = np.where(use_a, a[f], b[f])
    return out


def walls():
    n = ex = 0
    for seed, r, nb, chunk, mov, drop in itertools.product((1, 2, 3), (3.0, 7.0, 12.0), (0, 25), (48, 200, 520), (False, True), (0.0, 0.3)):
        kw = dict(n_rotations=3.2, seed=seed, n_boxes=nb, wall_radius=r, moving=mov, dropout=drop, extent=r * 0.9, min_box_dist=1.5)
        pts, poses, sp = synth.make_stream("tiny16", **kw)
        ex += run(pts, poses, sp, drvlib.stream_config("tiny16"), chunk, (kw, chunk))
        n += 1
    print("walls: ok", n, "scenes,", ex, "pushes through the split path")


def two_walls():
    n = ex = 0
    for seed, r1, h1, r2, nb, chunk in itertools.product((1, 2, 3), (3.0, 5.0), (0.5, 0.9), (9.0, 14.0), (3, 8), (40, 128, 400)):
        a, poses, sp = synth.make_stream("tiny16", n_rotations=4.2, seed=seed, n_boxes=0, wall_radius=r1, wall_height=h1)
        b, _, _ = synth.make_stream("tiny16", n_rotations=4.2, seed=seed + 10, n_boxes=nb, wall_radius=r2, wall_height=3.0, extent=r2 * 0.8,
                                    min_box_dist=r1 + 1.0, box_height_range=(2.0, 3.0))
        ex += run(merge_nearer(a, b), poses, sp, drvlib.stream_config("tiny16"), chunk, (seed, r1, h1, r2, nb, chunk))
        n += 1
    print("two walls: ok", n, "scenes,", ex, "pushes through the split path")


def nth():
    n = ex = 0
    scenes = [dict(n_rotations=3.0), dict(n_rotations=3.0, moving=True, dropout=0.1), dict(n_rotations=3.2, n_boxes=0, wall_radius=8.0),
              dict(n_rotations=3.0, az_jitter=0.4, az_step_scale=0.95)]
    for seed, k, chunk, si in itertools.product((1, 2), (2, 3, 5, 16, 100), (1, 37, 64, 300, 700), range(4)):
        if chunk == 1 and si != 0:
            continue
        kw = dict(scenes[si], seed=seed)
        pts, poses, sp = synth.make_stream("tiny16", **kw)
        ex += run(pts, poses, sp, drvlib.stream_config("tiny16", cluster_point_trees_every_nth_column=k), chunk, (kw, k, chunk))
        n += 1
    print("every n-th column: ok", n, "scenes,", ex, "pushes through the split path")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for name, fn in (("walls", walls), ("two_walls", two_walls), ("nth", nth)):
        if which in (name, "all"):
            fn()
