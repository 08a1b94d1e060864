"""Edge cases of the hot path the reference handles with exceptions or refusals:
  * the ring-overrun throw of performGroundPointSegmentationForColumn (cpp:318-344): a column that was never cleared is
    reached again after ten rotations,
  * associations the reference REFUSES on an organic (no closed wall) scene -- joining a finished tree (cpp:654-659),
    linking finished trees (cpp:688-690): tall boxes right beside an OS-32 whose steep beams see them under a wide azimuth,
  * setConfiguration between two firings of a running stream (cpp:66-81, node.cpp:234) without a reset.
Each on the CPU emulation of the kernels and, marked gpu, on the device."""
import numpy as np
import pytest

import parity
import recorder
from continuous_clustering_b200 import ClusteringError, synth
from oracle import drvlib
from test_emu_parity import make_cc

ORGANIC = ("os32_left", dict(n_rotations=1.5, seed=2, min_box_dist=1.7, n_boxes=60, extent=8.0, box_height_range=(3.5, 6.0)))


def overrun(library, oracle_lib):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=11.5)
    cfg = drvlib.stream_config("tiny16", cluster_point_trees_every_nth_column=1 << 30)  # nothing is ever published or cleared
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, sp.rows)
    chunk, failed_at = 128, None
    for a in range(0, pts.shape[0], chunk):
        try:
            d.add_firings(pts[a:a + chunk], poses[a:a + chunk])
        except RuntimeError as e:
            assert "This column is not cleared" in str(e)
            failed_at = a
            break
    assert failed_at is not None and failed_at >= 9 * sp.num_columns
    cc = make_cc(library, cfg, sp.rows)
    for a in range(0, failed_at, chunk):
        cc.addFirings(pts[a:a + chunk], poses[a:a + chunk])
    with pytest.raises(ClusteringError, match="This column is not cleared"):
        cc.addFirings(pts[failed_at:failed_at + chunk], poses[failed_at:failed_at + chunk])
    cc.reset(sp.rows)  # the handle is usable again after a reset
    cc.setTransformRobotFrameFromSensorFrame([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0])
    assert cc.addFirings(pts[:chunk], poses[:chunk]).info.n_events >= 0


def organic_refusals(library, oracle_lib, chunk):
    spec, kw = ORGANIC
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec)
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, sp.rows)
    want = parity.record(d, pts, poses)
    joins, links = d.refusals()
    assert joins > 0 and links > 0, "the scene must make the restatement refuse joins and links"
    cc = make_cc(library, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="product", check_tree_fields=True, check_published_tree_fields=True)
    return got


def config_mid_stream(library, oracle_lib, chunk=100):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=3.0, moving=True, dropout=0.05)
    cfg_a = drvlib.stream_config("tiny16")
    cfg_b = drvlib.stream_config("tiny16", max_distance=0.4, max_slope=0.1, cluster_point_trees_every_nth_column=2,
                                 ignore_points_in_chessboard_pattern=0)
    cfg_c = drvlib.stream_config("tiny16", max_steps_in_row=5, use_last_point_for_cluster_stamp=1)
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg_a, sp.rows)
    want = parity.record(d, pts, poses, chunk, hooks={300: lambda x: x.set_config(cfg_b), 600: lambda x: x.set_config(cfg_c)})
    assert not d.reset_required()
    cc = make_cc(library, cfg_a, sp.rows)
    got = recorder.record(cc, pts, poses, chunk, hooks={300: lambda x: x.setConfiguration(cfg_b), 600: lambda x: x.setConfiguration(cfg_c)})
    parity.compare(want, got, name_a="oracle", name_b="product")
    plain = parity.record(_fresh(oracle_lib, cfg_a, sp.rows), pts, poses, chunk)
    assert not np.array_equal(plain["cluster_cells"]["id"], want["cluster_cells"]["id"]), "the configuration change must matter"


def _fresh(oracle_lib, cfg, rows):
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, rows)
    return d


def test_ring_overrun_throws_emulation(emu_library, oracle_lib):
    overrun(emu_library, oracle_lib)


def test_organic_refusals_emulation(emu_library, oracle_lib):
    organic_refusals(emu_library, oracle_lib, 256)


def test_set_configuration_mid_stream_emulation(emu_library, oracle_lib):
    config_mid_stream(emu_library, oracle_lib)


def test_reference_build_agrees_on_the_organic_scene(oracle_lib):
    """The restatement's refusals are the reference's: same recording from the reference build on the same scene."""
    import os

    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libcc_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/libcc_ref.so not built (needs /root/reference at build time)")
    spec, kw = ORGANIC
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec)
    parity.compare(parity.record(_fresh(ref, cfg, sp.rows), pts, poses), parity.record(_fresh(oracle_lib, cfg, sp.rows), pts, poses),
                   name_a="reference", name_b="oracle")


@pytest.mark.gpu
def test_ring_overrun_throws_cuda(cuda_library, oracle_lib):
    overrun(None, oracle_lib)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [64, 1024])
def test_organic_refusals_cuda(cuda_library, oracle_lib, chunk):
    organic_refusals(None, oracle_lib, chunk)


@pytest.mark.gpu
def test_set_configuration_mid_stream_cuda(cuda_library, oracle_lib):
    config_mid_stream(None, oracle_lib)


# Forced finish (cpp:909-919): a closed wall makes one cluster span a rotation, again and again. Only the columns whose walk can
# reach the force-finished component go through the exact path; the ranges behind them are committed speculatively again --
# pushes that contain several dangerous columns, walls at different ranges (different window sizes), walls + boxes.
WALL_CASES = [
    ("tiny16", dict(n_rotations=4.0, n_boxes=0, wall_radius=4.0), 64),
    ("tiny16", dict(n_rotations=4.0, n_boxes=0, wall_radius=8.0), 300),
    ("tiny16", dict(n_rotations=5.0, n_boxes=0, wall_radius=15.0, moving=False), 700),   # > 2 rotations per push: several forced finishes in one
    ("tiny16", dict(n_rotations=4.0, n_boxes=60, wall_radius=10.0, extent=9.0, min_box_dist=2.0), 128),
    ("tiny16", dict(n_rotations=4.0, n_boxes=0, wall_radius=6.0, dropout=0.2), 100),
]


def wall_case(library, oracle_lib, spec, kw, chunk):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec)
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, sp.rows)
    want = parity.record(d, pts, poses)
    cc = make_cc(library, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="product", check_tree_fields=True, check_published_tree_fields=True)
    if kw.get("n_boxes", 150) == 0:  # (boxes in front of the wall can break the ring)
        assert got["used_exact_path"] > 0, "the scene must send pushes through the split path"


@pytest.mark.parametrize("spec,kw,chunk", WALL_CASES)
def test_forced_finish_scenes_emulation(emu_library, oracle_lib, spec, kw, chunk):
    wall_case(emu_library, oracle_lib, spec, kw, chunk)


@pytest.mark.gpu
@pytest.mark.parametrize("spec,kw,chunk", WALL_CASES + [
    ("velodyne64", dict(n_rotations=2.6, n_boxes=0, wall_radius=12.0), 1024),
    ("velodyne64", dict(n_rotations=2.6, n_boxes=100, wall_radius=20.0), 4096),
    ("os32_left", dict(n_rotations=3.0, n_boxes=0, wall_radius=6.0), 512),
])
def test_forced_finish_scenes_cuda(cuda_library, oracle_lib, spec, kw, chunk):
    wall_case(None, oracle_lib, spec, kw, chunk)


# cluster_point_trees_every_nth_column > 1 (cpp:841): finish passes only at every n-th column. The speculative whole-push
# commit handles it (finish columns rounded up to pass columns, columns between two passes keep the first unpublished column
# of the pass before them); until round 2 such configurations went through the exact path column by column.
NTH_CASES = [
    ("tiny16", dict(n_rotations=3.0), 2, 64),
    ("tiny16", dict(n_rotations=3.0, moving=True, dropout=0.1), 5, 37),
    ("tiny16", dict(n_rotations=3.0), 16, 1),                                   # one firing per push (fused kernel)
    ("tiny16", dict(n_rotations=3.2, n_boxes=0, wall_radius=8.0), 3, 300),      # with forced finishes
    ("tiny16", dict(n_rotations=3.0, az_jitter=0.4, az_step_scale=0.95), 100, 700),
]


def nth_case(library, oracle_lib, spec, kw, nth, chunk):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, cluster_point_trees_every_nth_column=nth)
    want = parity.record(_fresh(oracle_lib, cfg, sp.rows), pts, poses)
    cc = make_cc(library, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="product", check_tree_fields=True, check_published_tree_fields=True)
    if "wall_radius" not in kw:
        assert got["used_exact_path"] == 0, "every-nth-column passes must not need the exact path on an ordinary scene"


@pytest.mark.parametrize("spec,kw,nth,chunk", NTH_CASES)
def test_every_nth_column_emulation(emu_library, oracle_lib, spec, kw, nth, chunk):
    nth_case(emu_library, oracle_lib, spec, kw, nth, chunk)


@pytest.mark.gpu
@pytest.mark.parametrize("spec,kw,nth,chunk", NTH_CASES + [
    ("velodyne64", dict(n_rotations=2.3, moving=True, dropout=0.02), 4, 1024),
    ("velodyne64", dict(n_rotations=2.3), 7, 4096),
    ("os32_left", dict(n_rotations=2.5, moving=True), 3, 256),
])
def test_every_nth_column_cuda(cuda_library, oracle_lib, spec, kw, nth, chunk):
    nth_case(None, oracle_lib, spec, kw, nth, chunk)


def test_two_ring_components_emulation(emu_library, oracle_lib):
    """Two concentric walls (a low inner one, a tall outer one broken by boxes): several ring-spanning components whose forced
    finishes fall at unrelated columns -- the retry at the next dangerous column and the extension of a run of exact columns
    (forced_col). A sample of tests/tools/fuzz_scenes.py two_walls."""
    exact_pushes = 0
    for seed, r1, h1, r2, nb, chunk in ((1, 5.0, 0.5, 9.0, 3, 40), (1, 5.0, 0.5, 9.0, 8, 128), (2, 5.0, 0.9, 14.0, 8, 400)):
        a, poses, sp = synth.make_stream("tiny16", n_rotations=4.2, seed=seed, n_boxes=0, wall_radius=r1, wall_height=h1)
        b, _, _ = synth.make_stream("tiny16", n_rotations=4.2, seed=seed + 10, n_boxes=nb, wall_radius=r2, wall_height=3.0, extent=r2 * 0.8,
                                    min_box_dist=r1 + 1.0, box_height_range=(2.0, 3.0))
        da = np.sqrt(a["x"] ** 2 + a["y"] ** 2 + a["z"] ** 2)
        db = np.sqrt(b["x"] ** 2 + b["y"] ** 2 + b["z"] ** 2)
        use_a = (~np.isnan(da)) & (np.isnan(db) | (da < db))
        pts = b.copy()
        for f in ("x", "y", "z"):
            pts[f] = np.where(use_a, a[f], b[f])
        cfg = drvlib.stream_config("tiny16")
        want = parity.record(_fresh(oracle_lib, cfg, sp.rows), pts, poses)
        cc = make_cc(emu_library, cfg, sp.rows)
        got = recorder.record(cc, pts, poses, chunk)
        parity.compare(want, got, name_a="oracle", name_b="product", check_tree_fields=True, check_published_tree_fields=True)
        exact_pushes += got["used_exact_path"]
    assert exact_pushes > 0
