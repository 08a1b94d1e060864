"""The oracle (CPU restatement, oracle/cc_oracle.cpp) against the fixtures recorded from the reference's own
sources (tests/golden/make_golden.py) and, where the reference build is available, against that build live."""
import os

import numpy as np
import pytest

import parity
from continuous_clustering_b200 import synth
from golden import make_golden
from oracle import drvlib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def run_driver(lib, pts, poses, sp, cfg, chunk=None):
    d = drvlib.Driver(lib)
    d.configure(cfg, sp.rows)
    rec = parity.record(d, pts, poses, chunk)
    d.close()
    return rec


@pytest.mark.parametrize("name", sorted(make_golden.FIXTURES))
def test_oracle_matches_golden(oracle_lib, name):
    pts, poses, sp, cfg = make_golden.stream_for(name)
    want = make_golden.unpack(np.load(os.path.join(GOLDEN, name + ".npz")))
    got = run_driver(oracle_lib, pts, poses, sp, cfg)
    parity.compare(want, got, name_a="reference(golden)", name_b="oracle")
    # the restatement reproduces even the id numbering and the tree roots of the reference
    assert np.array_equal(want["cluster_cells"]["id"], got["cluster_cells"]["id"])
    assert np.array_equal(want["cluster_cells"]["tree_root_gcol"], got["cluster_cells"]["tree_root_gcol"])
    assert np.array_equal(want["cluster_cells"]["tree_root_row"], got["cluster_cells"]["tree_root_row"])


STREAMS = [
    ("velodyne64", dict(n_rotations=1.2), {}),
    ("velodyne64", dict(n_rotations=1.2, moving=True, dropout=0.05), {}),
    ("kitti64", dict(n_rotations=1.1), {}),
    ("vls128", dict(n_rotations=1.1, moving=True), {}),  # first firings straddle the -x axis: reset_required
    ("vls128", dict(n_rotations=1.1, moving=True, start_firing=40), {}),
    ("os32_right", dict(n_rotations=2.0, moving=True), {}),
    ("tiny16", dict(n_rotations=3.0, n_boxes=0, wall_radius=6.0), {}),
    ("tiny16", dict(n_rotations=3.0, az_jitter=0.7), {}),
    ("tiny16", dict(n_rotations=2.0, az_jitter=3.0, dropout=0.05), {}),
    ("velodyne64", dict(n_rotations=1.3, az_jitter=0.5, az_step_scale=1.04), {}),
    ("vls128", dict(n_rotations=1.1, az_jitter=1.5, start_firing=40, moving=True), {}),
    ("tiny16", dict(n_rotations=2.0, dropout=0.3), dict(stop_after_association_enabled=0)),
    ("tiny16", dict(n_rotations=2.0), dict(sensor_is_clockwise=0)),
    ("tiny16", dict(n_rotations=2.0), dict(fog_filtering_enabled=1, fog_filtering_intensity_below=120,
                                           fog_filtering_distance_below=30.0, fog_filtering_inclination_above=-0.2)),
    ("tiny16", dict(n_rotations=2.0), dict(use_last_point_for_cluster_stamp=1, supplement_inclination_angle_for_nan_cells=0,
                                           ignore_points_in_chessboard_pattern=0)),
]


@pytest.mark.parametrize("spec,kw,cfg_over", STREAMS)
def test_oracle_matches_reference_build(oracle_lib, ref_lib, spec, kw, cfg_over):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = run_driver(ref_lib, pts, poses, sp, cfg)
    got = run_driver(oracle_lib, pts, poses, sp, cfg)
    parity.compare(want, got, check_tree_fields=True, name_a="reference", name_b="oracle", check_published_tree_fields=True)
    assert np.array_equal(want["cluster_cells"]["id"], got["cluster_cells"]["id"])


def test_oracle_errors_like_reference(oracle_lib):
    pts, poses, sp = synth.make_stream("tiny16", n_firings=300)
    cfg = drvlib.stream_config("tiny16")
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, sp.rows, identity_robot_tf=False)  # no robot transform: cpp:298-299
    assert d.add_firings(pts, poses, raise_on_error=False) == 1
    assert "Transform robot frame from sensor frame was not set yet" in d.error()
    d2 = drvlib.Driver(oracle_lib)
    d2.configure(cfg, sp.rows)
    assert d2.add_firings(pts[:, :8], poses, raise_on_error=False) == 1  # cpp:90-91
    assert "number of points in a firing has changed" in d2.error()


def test_partition_canonicalisation():
    ids = np.array([0, 5, 5, 9, 0, 9, 7])
    keys = np.array([10, 11, 12, 13, 14, 15, 16])
    other = np.array([0, 2, 2, 1, 0, 1, 3])
    assert np.array_equal(parity.canonical_partition(ids, keys), parity.canonical_partition(other, keys))
    assert not np.array_equal(parity.canonical_partition(ids, keys),
                              parity.canonical_partition(np.array([0, 2, 2, 1, 0, 1, 1]), keys))
