"""Parity tests proper: the sm_100a kernels through the C ABI on a real B200 against the oracle (and the golden
recordings of the reference). Bar: every per-cell field and ground label bit-exact, finished-column events identical
and in order, cluster partition identical up to a permutation of ids, finished clusters identical."""
import os

import numpy as np
import pytest

import parity
import recorder
from continuous_clustering_b200 import ContinuousClustering, synth
from golden import make_golden
from oracle import drvlib
from test_emu_parity import ASYNC_CASES, CASES, IDENTITY, check_label_prefetch, make_cc, oracle_record

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", CASES)
def test_cuda_matches_oracle_small(cuda_library, oracle_lib, spec, kw, cfg_over, chunk, flag_period):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="cuda", check_tree_fields=True, check_published_tree_fields=True)
    assert np.array_equal(want["cluster_cells"]["tree_root_gcol"], got["cluster_cells"]["tree_root_gcol"])


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", ASYNC_CASES + [
    ("velodyne64", dict(n_rotations=4.2, moving=True, dropout=0.02), {}, 2048, 0),
    ("velodyne64", dict(n_rotations=2.2), {}, 1024, 97),
])
def test_cuda_pipelined_pushes_match_oracle(cuda_library, oracle_lib, spec, kw, cfg_over, chunk, flag_period):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk, pipelined=True)
    parity.compare(want, got, name_a="oracle", name_b="cuda, pipelined")


@pytest.mark.parametrize("spec,kw,chunk", [
    ("tiny16", dict(n_rotations=4.0, moving=True, dropout=0.05), 100),
    ("tiny16", dict(n_rotations=6.0), 700),                                     # pushes of more than two rotations
    ("velodyne64", dict(n_rotations=4.2, moving=True, dropout=0.02), 4096),   # bench.py's operating point
])
def test_cuda_staged_pushes_match_oracle(cuda_library, oracle_lib, spec, kw, chunk):
    """submit(k + 2); wait(k): two pushes in flight and a third one staged (its input copy starts at once, its kernels
    when a later call finds room). Columns reported by push k are read after wait(k), as a caller would."""
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows, max_push=max(4096, chunk))
    got = recorder.record(cc, pts, poses, chunk, pipelined=2)
    parity.compare(want, got, name_a="oracle", name_b="cuda, staged")


@pytest.mark.parametrize("name", sorted(make_golden.FIXTURES))
def test_cuda_matches_golden(cuda_library, name):
    pts, poses, sp, cfg = make_golden.stream_for(name)
    want = make_golden.unpack(np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz")))
    cc = make_cc(None, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, 256)
    parity.compare(want, got, name_a="reference(golden)", name_b="cuda")


FULL = [  # BASELINE.json configs at full size (several rotations), oracle finishes these in seconds
    ("velodyne64", dict(n_rotations=4.2), 2048),
    ("velodyne64", dict(n_rotations=3.1, moving=True, dropout=0.02), 4096),
    ("velodyne64", dict(n_rotations=6.2, moving=True, dropout=0.01), 6144),  # the largest push the ring allows (3 rotations)
    ("kitti64", dict(n_rotations=2.2, moving=True), 2200),
    ("vls128", dict(n_rotations=3.2, moving=True, start_firing=40), 1700),
    ("os32_left", dict(n_rotations=4.0, moving=True), 1024),
    ("os32_right", dict(n_rotations=4.0, moving=True, min_box_dist=2.0, box_height_range=(3.0, 10.0), extent=20.0), 512),
    ("velodyne64", dict(n_rotations=2.5, n_boxes=0, wall_radius=12.0), 1024),  # forced finish at full size
]


@pytest.mark.parametrize("spec,kw,chunk", FULL)
def test_cuda_matches_oracle_full_size(cuda_library, oracle_lib, spec, kw, chunk):
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows, max_push=max(4096, chunk))
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="cuda", check_tree_fields=True, check_published_tree_fields=True)


def test_cuda_matches_reference_build_when_present(cuda_library):
    if not drvlib.have_ref():
        pytest.skip("oracle/_ref/libcc_ref.so not shipped")
    pts, poses, sp = synth.make_stream("velodyne64", n_rotations=2.2, moving=True)
    cfg = drvlib.stream_config("velodyne64")
    want = oracle_record(drvlib.REF_LIB, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows)
    got = recorder.record(cc, pts, poses, 2048)
    parity.compare(want, got, name_a="reference", name_b="cuda")


def test_size_independent_properties(cuda_library):
    """Determinism across push sizes and repeated runs (the union-find is lock-free and racy by design, the
    partition must not be), idempotence of canonical labelling, id 0 exactly where the oracle has it."""
    pts, poses, sp = synth.make_stream("velodyne64", n_rotations=3.0, moving=True, dropout=0.01)
    cfg = drvlib.stream_config("velodyne64")
    recs = []
    for chunk in (256, 2048, 4096, 2048):
        cc = make_cc(None, cfg, sp.rows)
        recs.append(recorder.record(cc, pts, poses, chunk))
    for r in recs[1:]:
        parity.compare(recs[0], r, name_a="chunk 256", name_b="other chunk")
    cells = recs[0]["cluster_cells"]
    canon = parity.canonical_partition(cells["id"], cells["globally_unique_point_index"])
    assert np.array_equal(canon, parity.canonical_partition(canon, cells["globally_unique_point_index"]))
    # every published cluster has more than 5 points (cpp:936-940) and only obstacle, non-ignored cells carry an id
    ids, counts = np.unique(cells["id"][cells["id"] != 0], return_counts=True)
    assert counts.min() > 5
    assert (cells["is_ignored"][cells["id"] != 0] == 0).all()
    assert (cells["ground_point_label"][cells["id"] != 0] == 119).all()


def test_cuda_label_prefetch(cuda_library):
    check_label_prefetch(None)


def test_device_resident_push_equals_host_push(cuda_library):
    import torch

    pts, poses, sp = synth.make_stream("velodyne64", n_rotations=1.5)
    cfg = drvlib.stream_config("velodyne64")
    a = make_cc(None, cfg, sp.rows)
    b = make_cc(None, cfg, sp.rows)
    d_pts = torch.from_numpy(pts.view(np.uint8).reshape(pts.shape[0], -1)).cuda()
    d_poses = torch.from_numpy(poses).cuda()
    B = 1024
    for s in range(0, pts.shape[0] - B + 1, B):
        ra = a.addFirings(pts[s:s + B], poses[s:s + B])
        rb = b.addFiringsDevice(d_pts.data_ptr() + s * sp.rows * 48, d_poses.data_ptr() + s * 96, B, sp.rows)
        assert np.array_equal(ra.events, rb.events)
        assert ra.info.n_cluster_points == rb.info.n_cluster_points
        assert rb.info.gpu_launches > 0


@pytest.mark.gpu
def test_device_timeline_does_not_change_results(cuda_library, oracle_lib):
    """cc_debug_trace: the %globaltimer stamps of every CTA give per-kernel spans of a push as the kernels overlap in
    normal operation; results stay identical with the hook on."""
    pts, poses, sp = synth.make_stream("velodyne64", n_rotations=2.2, moving=True)
    cfg = drvlib.stream_config("velodyne64")
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows)
    cc.debug_trace(True)
    got = recorder.record(cc, pts, poses, 1024)
    parity.compare(want, got, name_a="oracle", name_b="cuda, traced")
    tr = {name: (a, z, longest, blocks) for name, a, z, longest, blocks in cc.get_trace()}
    for k in ("k_prep", "k_ground", "k_probe", "k_fin_all", "k_fin_label"):
        assert k in tr and tr[k][3] > 0 and tr[k][1] > tr[k][0]
    assert tr["k_ground"][0] >= tr["k_prep"][0]


@pytest.mark.parametrize("spec,kw,cfg_over,chunk,flag_period", [
    ("velodyne64", dict(n_rotations=2.2, moving=True, dropout=0.03), {}, 1024, 0),
    ("vls128", dict(n_rotations=1.3, moving=True), {}, 700, 0),
    ("os32_left", dict(n_rotations=2.5, moving=True), {}, 512, 0),
    ("tiny16", dict(n_rotations=3.0, n_boxes=0, wall_radius=8.0), {}, 128, 0),
    ("velodyne64", dict(n_rotations=1.3, az_jitter=0.5, az_step_scale=1.04), {}, 1024, 0),
])
def test_cuda_tiled_probe_matches_oracle(cuda_library, oracle_lib, monkeypatch, spec, kw, cfg_over, chunk, flag_period):
    """k_probe_tile (field of view of a tile of columns staged in shared memory by cp.async.bulk + mbarrier, one thread per
    cell; CC_B200_TUNE bit 2): same results as the default list-driven probe."""
    monkeypatch.setenv("CC_B200_TUNE", "4")
    monkeypatch.setenv("CC_B200_FUSED_MAX", "0")
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want = oracle_record(oracle_lib, pts, poses, sp, cfg)
    cc = make_cc(None, cfg, sp.rows)
    cc.debug_flag_columns(flag_period)
    got = recorder.record(cc, pts, poses, chunk)
    parity.compare(want, got, name_a="oracle", name_b="cuda (tiled probe)", check_tree_fields=True)
