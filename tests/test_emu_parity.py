"""Kernel + host logic of the CUDA library, executed by the CPU emulation build (tests/emu), against the oracle.
This does not replace the -m gpu parity tests (no races can show up here); it keeps the logic checked in CI."""
import numpy as np
import pytest

import parity
import recorder
from continuous_clustering_b200 import ClusteringError, ContinuousClustering, synth
from golden import make_golden
from oracle import drvlib

IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]


def oracle_record(lib, pts, poses, sp, cfg):
    d = drvlib.Driver(lib)
    d.configure(cfg, sp.rows)
    rec = parity.record(d, pts, poses)
    d.close()
    return rec


def make_cc(library, cfg, rows, max_push=4096, tf=True):
    cc = ContinuousClustering(_library=library, max_firings_per_push=max_push)
    cc.setConfiguration(cfg)
    cc.reset(rows)
    if tf:
        cc.setTransformRobotFrameFromSensorFrame(IDENTITY)
    return cc


CASES = [
    ("tiny16", dict(n_rotations=2.0), {}, 64, 0),
    ("tiny16", dict(n_rotations=1.5, moving=True, dropout=0.1), {}, 1, 0),  # one firing per push, like addFiring
    ("tiny16", dict(n_rotations=3.0), {}, 700, 0),                           # more than two rotations per push
    ("tiny16", dict(n_rotations=2.0), {}, 37, 5),                            # every 5th column through the exact path
    ("tiny16", dict(n_rotations=2.0, moving=True), {}, 300, 1),              # every column through the exact path
    ("tiny16", dict(n_rotations=3.0, n_boxes=0, wall_radius=8.0), {}, 128, 0),  # forced finish (cpp:909-919)
    ("tiny16", dict(n_rotations=2.0), dict(cluster_point_trees_every_nth_column=3), 64, 0),
    ("tiny16", dict(n_rotations=2.0, dropout=0.3), dict(stop_after_association_enabled=0), 200, 0),
    ("tiny16", dict(n_rotations=2.0), dict(sensor_is_clockwise=0), 64, 0),
    ("tiny16", dict(n_rotations=2.0), dict(fog_filtering_enabled=1, fog_filtering_intensity_below=120,
                                           fog_filtering_distance_below=30.0, fog_filtering_inclination_above=-0.2), 64, 0),
    ("tiny16", dict(n_rotations=2.0), dict(use_last_point_for_cluster_stamp=1, supplement_inclination_angle_for_nan_cells=0,
                                           ignore_points_in_chessboard_pattern=0), 64, 0),
    ("velodyne64", dict(n_rotations=1.2, moving=True, dropout=0.03), {}, 700, 0),
    ("os32_left", dict(n_firings=900, moving=True), {}, 256, 0),
    ("vls128", dict(n_firings=600, moving=True), {}, 256, 0),  # begins with firings that straddle the -x axis
    # azimuth jitter: firings land twice in a column / skip columns / run backwards (collision rule cpp:188-208,
    # "too far behind" cpp:210-221): regular and irregular chunks of the insertion scan are mixed
    ("tiny16", dict(n_rotations=3.0, az_jitter=0.7), {}, 64, 0),
    ("tiny16", dict(n_rotations=3.0, az_jitter=0.3, az_step_scale=0.93, moving=True, dropout=0.05), {}, 200, 0),
    ("tiny16", dict(n_rotations=2.0, az_jitter=3.0), {}, 33, 0),
    ("velodyne64", dict(n_rotations=1.3, az_jitter=0.5, az_step_scale=1.04), {}, 1024, 0),
    ("vls128", dict(n_rotations=1.1, az_jitter=1.5, start_firing=40, moving=True), {}, 512, 0),
    # rough ground (range noise) and low boxes with ground visible behind them: the label rules that carry state through
    # a column -- YELLOW / YELLOWGREEN, the last-certain-ground updates and the DARKRED relabel walk (cpp:433-565) --
    # fire on a large share of the points instead of on a handful
    ("velodyne64", dict(n_rotations=1.2, range_noise=0.08, box_height_range=(0.2, 0.8), min_box_dist=3.0), {}, 700, 0),
    ("velodyne64", dict(n_rotations=1.1, range_noise=0.05, box_height_range=(0.2, 1.0), min_box_dist=3.0, n_boxes=300,
                        moving=True, dropout=0.03), dict(max_slope=0.08), 512, 0),
    ("tiny16", dict(n_rotations=3.0, range_noise=0.1, box_height_range=(0.2, 0.8), min_box_dist=3.0, n_boxes=300), {}, 100, 0),
    ("tiny16", dict(n_rotations=3.0, range_noise=0.1, box_height_range=(0.2, 0.8), min_box_dist=3.0, n_boxes=300, moving=True),
     dict(first_ring_as_ground_max_allowed_z_diff=0.03, first_ring_as_ground_min_allowed_z_diff=-0.03), 64, 0),  # ORANGE
    ("tiny16", dict(n_rotations=2.0, range_noise=0.1, box_height_range=(0.2, 0.8), min_box_dist=3.0), dict(use_terrain=1), 64, 0),
    ("vls128", dict(n_rotations=1.1, range_noise=0.1, box_height_range=(0.2, 0.8), min_box_dist=3.0, n_boxes=300, moving=True),
     {}, 600, 0),
]


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", CASES)
def test_emulated_kernels_match_oracle(emu_library, oracle_lib, spec, kw, cfg_over, chunk, flag_period):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(emu_library, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="emulated kernels", check_tree_fields=True,
                   check_published_tree_fields=True)
    assert np.array_equal(want["cluster_cells"]["tree_root_gcol"], got["cluster_cells"]["tree_root_gcol"])
    if flag_period:
        assert got["used_exact_path"] > 0


@pytest.mark.parametrize("name", ["tiny16_static", "tiny16_wall_forced_finish", "os32_left_short"])
def test_emulated_kernels_match_golden(emu_library, name):
    import os

    pts, poses, sp, cfg = make_golden.stream_for(name)
    want = make_golden.unpack(np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz")))
    cc = make_cc(emu_library, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, 200)
    parity.compare(want, got, name_a="reference(golden)", name_b="emulated kernels")


ASYNC_CASES = [
    ("tiny16", dict(n_rotations=3.0, moving=True, dropout=0.05), {}, 64, 0),
    ("tiny16", dict(n_rotations=3.0), {}, 50, 11),                                # flagged columns: halt + replay
    ("tiny16", dict(n_rotations=3.0, n_boxes=0, wall_radius=8.0), {}, 100, 0),   # forced finish: abort + rollback + replay
    ("tiny16", dict(n_rotations=2.0), dict(cluster_point_trees_every_nth_column=3), 64, 0),
    ("tiny16", dict(n_rotations=3.0, az_jitter=0.7), {}, 40, 0),
    ("velodyne64", dict(n_rotations=1.3, moving=True), {}, 512, 0),
]


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", ASYNC_CASES)
def test_pipelined_pushes_match_oracle(emu_library, oracle_lib, spec, kw, cfg_over, chunk, flag_period):
    """Two pushes in flight (cc_submit_firings / cc_wait): same results, including when a push in flight has to
    fall back to the column-sequential path and the one queued behind it is re-run."""
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(emu_library, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk, pipelined=True)
    parity.compare(want, got, name_a="oracle", name_b="emulated kernels, pipelined")
    assert cc.pending == 0


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", ASYNC_CASES)
def test_staged_pushes_match_oracle(emu_library, oracle_lib, spec, kw, cfg_over, chunk, flag_period):
    """submit(k + 2); wait(k): a third push is staged behind the two in flight and launched by the first call that finds
    room; the columns push k reported are read right after wait(k). Includes the halt + replay paths."""
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(emu_library, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk, pipelined=2)
    parity.compare(want, got, name_a="oracle", name_b="emulated kernels, staged")
    assert cc.pending == 0


def test_staged_push_limits(emu_library):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=2.0)
    cc = make_cc(emu_library, drvlib.stream_config("tiny16"), sp.rows)
    for k in range(3):
        cc.submitFirings(pts[k * 32:(k + 1) * 32], poses[k * 32:(k + 1) * 32])
    assert cc.pending == 3
    with pytest.raises(ClusteringError):  # two in flight + one staged: the fourth has to wait
        cc.submitFirings(pts[96:128], poses[96:128])
    for k in range(3):
        cc.wait()
    assert cc.pending == 0


def check_label_prefetch(library):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=2.0, moving=True)
    cc = make_cc(library, drvlib.stream_config("tiny16"), sp.rows)
    cc.set_label_prefetch(True)
    for a in range(0, pts.shape[0], 128):
        r = cc.addFirings(pts[a:a + 128], poses[a:a + 128])
        lab = cc.column_labels()
        lo, hi = int(r.info.ground_from_gcol), int(r.info.ground_to_gcol) - 1
        ref = cc.read_columns(lo, hi, fields=["ground_point_label", "debug_ground_point_label", "is_ignored", "intensity"])
        assert lab.shape == (hi - lo + 1, sp.rows, 4)
        for j, f in enumerate(("ground_point_label", "debug_ground_point_label", "is_ignored", "intensity")):
            assert np.array_equal(lab[..., j], ref[f])


def test_label_prefetch_equals_read_columns(emu_library):
    check_label_prefetch(emu_library)


def test_results_do_not_depend_on_push_size(emu_library):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=2.0, moving=True, dropout=0.05)
    cfg = drvlib.stream_config("tiny16")
    recs = []
    for chunk in (1, 13, 256, 512):
        cc = make_cc(emu_library, cfg, sp.rows)
        recs.append(recorder.record(cc, pts, poses, chunk))
    for r in recs[1:]:
        parity.compare(recs[0], r)


def test_error_behaviour_mirrors_reference(emu_library):
    pts, poses, sp = synth.make_stream("tiny16", n_firings=300)
    cfg = drvlib.stream_config("tiny16")
    cc = ContinuousClustering(_library=emu_library)
    cc.setConfiguration(cfg)
    with pytest.raises(ClusteringError, match="cc_reset"):
        cc.addFirings(pts, poses)
    cc.reset(sp.rows)
    assert not cc.hasTransformRobotFrameFromSensorFrame()
    with pytest.raises(ClusteringError, match="Transform robot frame from sensor frame was not set yet"):  # cpp:298-299
        cc.addFirings(pts, poses)
    cc.reset(sp.rows)
    cc.setTransformRobotFrameFromSensorFrame(IDENTITY)
    with pytest.raises(ClusteringError, match="number of points in a firing has changed"):  # cpp:90-91
        cc.addFirings(pts[:, :8], poses)
    cc.addFirings(pts[:0], poses[:0])  # empty push is a no-op
    assert cc.addFirings(pts, poses).info.n_events > 0


def test_reset_required_semantics(emu_library, oracle_lib):
    cfg = drvlib.stream_config("tiny16")
    cc = make_cc(emu_library, cfg, 16)
    assert not cc.resetRequired()
    cfg2 = drvlib.stream_config("tiny16", num_columns=512)  # cpp:69-74
    cc.setConfiguration(cfg2)
    assert cc.resetRequired()
    cc.reset(16)
    assert not cc.resetRequired() and cc.num_columns_ == 512 and cc.ring_buffer_max_columns == 5120
    # a first firing that straddles the negative x axis (cpp:252-261)
    pts, poses, sp = synth.make_stream("vls128", n_firings=64, moving=True)
    cfgv = drvlib.stream_config("vls128")
    ccv = make_cc(emu_library, cfgv, sp.rows)
    ccv.addFirings(pts, poses)
    d = drvlib.Driver(oracle_lib)
    d.configure(cfgv, sp.rows)
    d.add_firings(pts, poses)
    assert ccv.resetRequired() == d.reset_required() is True


def test_reset_mid_stream_and_callbacks(emu_library, oracle_lib):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=2.0)
    cfg = drvlib.stream_config("tiny16")
    cc = make_cc(emu_library, cfg, sp.rows)
    cc.addFirings(pts[:300], poses[:300])
    cc.reset(sp.rows)
    cc.setTransformRobotFrameFromSensorFrame(IDENTITY)
    got = recorder.record(cc, pts, poses, 128)
    # the reference keeps sc_inclination_angles_between_lasers_ across a reset with the same row count (cpp:46), so
    # the oracle has to live through the same history
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, sp.rows)
    d.add_firings(pts[:300], poses[:300])
    d.configure(cfg, sp.rows)
    d.clear_records()
    want = parity.record(d, pts, poses)
    parity.compare(want, got)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    # callback interface (hpp:217-218): same sequence of column callbacks, clusters with > 20 points
    cc2 = make_cc(emu_library, cfg, sp.rows)
    cols, clusters = [], []
    cc2.setFinishedColumnCallback(lambda a, b, g: cols.append((a, b, g)))
    cc2.setFinishedClusterCallback(lambda p, stamp: clusters.append((stamp, len(p), set(p["id"].tolist()))))
    cc2.batch_firings = 100
    for k in range(pts.shape[0]):
        cc2.addFiring(pts[k], poses[k])
    cc2.flush()
    ev = want["events"]
    assert cols == [(int(e["from_gcol"]), int(e["to_gcol"]), bool(e["ground_points_only"])) for e in ev]
    assert sorted((s, n) for s, n, _ in clusters) == sorted((int(c["stamp"]), int(c["num_points"])) for c in want["clusters"])
    assert all(len(ids) == 1 and 0 not in ids for _, _, ids in clusters)


def test_debug_hooks_are_inert_in_the_emulation(emu_library):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=1.0)
    cc = make_cc(emu_library, drvlib.stream_config("tiny16"), sp.rows)
    cc.debug_trace(True)
    cc.addFirings(pts[:128], poses[:128])
    assert cc.get_trace() == []  # the stamps are compiled out of the CPU build
    cc.debug_trace(False)


CHAIN_CASES = [CASES[0], CASES[3], CASES[5], CASES[14], CASES[12]]


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", CHAIN_CASES)
def test_short_pushes_through_the_kernel_chain(emu_library, oracle_lib, monkeypatch, spec, kw, cfg_over, chunk, flag_period):
    """Short pushes normally take the fused single-launch kernel (k_push_fused); CC_B200_FUSED_MAX=0 sends them through
    the kernel chain the long pushes use. Same results either way."""
    monkeypatch.setenv("CC_B200_FUSED_MAX", "0")
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    probe = make_cc(emu_library, cfg, sp.rows)
    assert probe.addFirings(pts[:chunk], poses[:chunk]).info.gpu_launches > 1
    probe.close()
    cc = make_cc(emu_library, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="emulated kernels (chain)")


def test_short_pushes_are_one_launch(emu_library):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=2.0)
    cc = make_cc(emu_library, drvlib.stream_config("tiny16"), sp.rows)
    launches = [int(cc.addFirings(pts[a:a + 64], poses[a:a + 64]).info.gpu_launches) for a in range(0, 384, 64)]
    assert launches == [1] * len(launches), launches


TILE_CASES = [CASES[0], CASES[2], CASES[5], CASES[11], CASES[12], CASES[13], CASES[17]]


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", TILE_CASES)
def test_tiled_probe_matches_oracle(emu_library, oracle_lib, monkeypatch, spec, kw, cfg_over, chunk, flag_period):
    """The alternative association probe (k_probe_tile: one thread per cell of a tile of columns, the tile's field of view
    staged in shared memory; CC_B200_TUNE bit 2) gives the same results as the list-driven one, through the kernel chain."""
    monkeypatch.setenv("CC_B200_TUNE", "4")
    monkeypatch.setenv("CC_B200_FUSED_MAX", "0")
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(emu_library, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="emulated kernels (tiled probe)", check_tree_fields=True)
