"""The reference's own CALLER code against the drop-in class: ros_utils.cpp's PointCloud2 conversion
(columnToPointCloud / clusterToPointCloud / addPointToMessage, ros_utils.cpp:11-77, 108-298) and kitti_demo.cpp's
evaluation callback (kitti_demo.cpp:173-224) are cut out of /root/reference at build time, compiled UNMODIFIED against
the reference's class (oracle/_ref/libcc_ref.so) and against the facade's headers + library
(oracle/_ref/libcc_facade_callers*.so), and run on the same streams: the messages the ROS node would publish must be
byte-identical (cluster ids up to a permutation, points of a cluster up to their order)."""
import os
import subprocess

import numpy as np
import pytest

from continuous_clustering_b200 import synth
from oracle import drvlib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def _need(*libs):
    for lib in libs:
        if not os.path.exists(lib):
            pytest.skip(os.path.relpath(lib, REPO) + " not built (needs /root/reference at build time)")


def test_reference_callers_compile_against_the_facade_headers():
    """The compile-time half of the drop-in claim: ros_utils.cpp:245-298 reads every Point member the ROS node publishes
    (incl. BLUE of the full colour table), kitti_demo.cpp:173-224 is the evaluation callback."""
    if not os.path.exists(os.path.join(REFERENCE, "src/ros/ros_utils.cpp")):
        _need(drvlib.FACADE_CALLERS_EMU_LIB)  # the GPU box: only the prebuilt library can be checked
        d = drvlib.Driver(drvlib.FACADE_CALLERS_EMU_LIB)
        assert d.has_caller_excerpts()
        return
    inc = os.path.join(REPO, "oracle", "_ref")
    subprocess.run(["python3", os.path.join(REPO, "oracle", "extract_caller_excerpts.py"), REFERENCE, inc], check=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-DDRV_WITH_CALLER_EXCERPTS", "-DCC_B200_FACADE",
           "-I" + os.path.join(REPO, "facade", "include"), "-I" + os.path.join(REPO, "oracle", "eigen_standin"),
           "-I" + os.path.join(REPO, "oracle", "ros_standin"), "-I" + inc, os.path.join(REPO, "oracle", "cc_driver.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


def record_clouds(lib, pts, poses, sp, cfg, batch=64, kitti_frames=None):
    os.environ["CC_B200_BATCH"] = str(batch)
    os.environ["CC_B200_PIPELINE"] = "0"
    d = drvlib.Driver(lib)
    assert d.has_caller_excerpts()
    d.configure(cfg, sp.rows)
    d.set_record(drvlib_record_events())
    d.set_cloud_record(True)
    if kitti_frames is not None:
        d.kitti_begin(0, kitti_frames)
    for a in range(0, pts.shape[0], 500):
        d.add_firings(pts[a:a + 500], poses[a:a + 500])
    clouds = d.clouds()
    kitti = [d.kitti_get(f) for f in range(len(kitti_frames))] if kitti_frames is not None else None
    d.close()
    return clouds, kitti


def drvlib_record_events():
    return 1  # DRV_RECORD_EVENTS


def id_bijection(a, b):
    """ids are a running counter in both implementations: equal up to a permutation; 0 = no cluster must match."""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    assert np.array_equal(a == 0, b == 0)
    pairs = np.unique(np.stack([a, b], axis=1), axis=0)
    assert len(np.unique(pairs[:, 0])) == len(pairs) and len(np.unique(pairs[:, 1])) == len(pairs), "id mapping is not 1:1"


def canonical_order(clouds):
    """Clusters finished by the same pass are delivered back to back in an unspecified order (the reference's BFS order
    over its list of unfinished trees): runs of consecutive cluster messages are sorted by (stamp, size, smallest point)."""
    out, run = [], []
    for d, c in clouds:
        if d["kind"] == 2:
            run.append((d, c))
            continue
        out += sorted(run, key=lambda x: (int(x[0]["stamp_ns"]), int(x[0]["width"]), float(x[1]["globally_unique_point_index"].min())))
        run = []
        out.append((d, c))
    out += sorted(run, key=lambda x: (int(x[0]["stamp_ns"]), int(x[0]["width"]), float(x[1]["globally_unique_point_index"].min())))
    return out


def compare_clouds(want, got, check_visited=True):
    assert len(want) == len(got), f"{len(want)} vs {len(got)} messages"
    want, got = canonical_order(want), canonical_order(got)
    ids_w, ids_g = [], []
    for i, ((dw, cw), (dg, cg)) in enumerate(zip(want, got)):
        for f in ("from_gcol", "to_gcol", "kind", "width", "height", "point_step", "stamp_ns"):
            assert dw[f] == dg[f], f"message {i} {f}: {dw[f]} vs {dg[f]}"
        cw, cg = cw.copy(), cg.copy()
        if "id" in cw.dtype.names:
            ids_w.append(cw["id"].ravel().copy())
            ids_g.append(cg["id"].ravel().copy())
            cw["id"] = 0
            cg["id"] = 0
            if not check_visited:
                cw["number_of_visited_neighbors"] = 0
                cg["number_of_visited_neighbors"] = 0
        if dw["kind"] == 2:  # finished cluster: the order of the points inside a cluster is unspecified (BFS order)
            cw = np.sort(cw.ravel(), order=["globally_unique_point_index"])
            cg = np.sort(cg.ravel(), order=["globally_unique_point_index"])
        if cw.tobytes() != cg.tobytes():
            for f in cw.dtype.names:
                x, y = cw[f], cg[f]
                same = (x == y) | ((x != x) & (y != y)) if x.dtype.kind == "f" else (x == y)
                assert same.all(), f"message {i} (kind {dw['kind']}, columns {dw['from_gcol']}..{dw['to_gcol']}) field {f}: " \
                                   f"{x[~same][:4]} vs {y[~same][:4]} ({int((~same).sum())} points)"
            raise AssertionError(f"message {i}: bytes differ although every field compares equal (NaN payloads?)")
    if ids_w:
        id_bijection(np.concatenate(ids_w), np.concatenate(ids_g))


CASES = [
    ("tiny16", dict(n_rotations=2.0, moving=True), {}, 64),
    ("tiny16", dict(n_rotations=2.0, dropout=0.1), dict(ignore_points_in_chessboard_pattern=0), 100),
    ("velodyne64", dict(n_rotations=1.1), {}, 256),
]


@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES)
def test_published_messages_identical_emulation(emu_library, spec, kw, cfg_over, batch):
    _need(drvlib.REF_LIB, drvlib.FACADE_CALLERS_EMU_LIB)
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want, _ = record_clouds(drvlib.REF_LIB, pts, poses, sp, cfg)
    got, _ = record_clouds(drvlib.FACADE_CALLERS_EMU_LIB, pts, poses, sp, cfg, batch)
    compare_clouds(want, got)


def kitti_like_stream(spec, n_rotations):
    """guid = sequence << 48 | frame << 32 | point index, as kitti_demo.cpp:146-149 encodes it (one frame per rotation)."""
    pts, poses, sp = synth.make_stream(spec, n_rotations=n_rotations, moving=True)
    pts = pts.copy()
    n, rows = pts.shape
    k = np.arange(n, dtype=np.uint64)[:, None]
    frame = k // np.uint64(sp.num_columns)
    idx = (k % np.uint64(sp.num_columns)) * np.uint64(rows) + np.arange(rows, dtype=np.uint64)[None, :]
    pts["globally_unique_point_index"] = (frame << np.uint64(32)) | idx
    n_frames = int(frame.max()) + 1
    return pts, poses, sp, [sp.num_columns * rows] * n_frames


def check_kitti(lib, emu):
    pts, poses, sp, frames = kitti_like_stream("tiny16", 3.0)
    cfg = drvlib.stream_config("tiny16")
    _, want = record_clouds(drvlib.REF_LIB, pts, poses, sp, cfg, kitti_frames=frames)
    _, got = record_clouds(lib, pts, poses, sp, cfg, 64, kitti_frames=frames)
    assert any(w is not None and w[0].any() for w in want)
    for f, (w, g) in enumerate(zip(want, got)):
        assert (w is None) == (g is None), f"frame {f}"
        if w is None:
            continue
        assert np.array_equal(w[0], g[0]), f"frame {f}: has_corresponding_point / is_ground_point flags differ"
        id_bijection(w[1], g[1])


def test_kitti_demo_callback_identical_emulation(emu_library):
    _need(drvlib.REF_LIB, drvlib.FACADE_CALLERS_EMU_LIB)
    check_kitti(drvlib.FACADE_CALLERS_EMU_LIB, True)


@pytest.mark.gpu
@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES + [("velodyne64", dict(n_rotations=2.2, moving=True), {}, 1024)])
def test_published_messages_identical_cuda(cuda_library, spec, kw, cfg_over, batch):
    _need(drvlib.REF_LIB, drvlib.FACADE_CALLERS_LIB)
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want, _ = record_clouds(drvlib.REF_LIB, pts, poses, sp, cfg)
    got, _ = record_clouds(drvlib.FACADE_CALLERS_LIB, pts, poses, sp, cfg, batch)
    compare_clouds(want, got)


@pytest.mark.gpu
def test_kitti_demo_callback_identical_cuda(cuda_library):
    _need(drvlib.REF_LIB, drvlib.FACADE_CALLERS_LIB)
    check_kitti(drvlib.FACADE_CALLERS_LIB, False)


# ---- the same messages packed ON THE DEVICE (cc_pack_*_pointcloud2) ----
def device_clouds(cc, pts, poses, chunk):
    """Runs the stream through the C ABI and builds, for every callback the reference's node would get, the message
    payload with the device-side packer; same descriptor layout as Driver.clouds()."""
    out = []
    for a in range(0, pts.shape[0], chunk):
        res = cc.addFirings(pts[a:a + chunk], poses[a:a + chunk])
        nxt = 0
        for e in res.events.copy():
            while nxt < int(e["n_clusters_before"]):
                cl = res.clusters[nxt]
                if cl["num_points"] > 20:  # cpp:1023
                    c = cc.pack_cluster_pointcloud2(nxt)
                    out.append((2, -1, -1, dict(c, data=c["data"].copy())))  # the payload is a view: keep a copy
                nxt += 1
            if e["to_gcol"] >= e["from_gcol"]:  # columnToPointCloud: no message for an empty range
                c = cc.pack_columns_pointcloud2(int(e["from_gcol"]), int(e["to_gcol"]), bool(e["ground_points_only"]))
                out.append((0 if e["ground_points_only"] else 1, int(e["from_gcol"]), int(e["to_gcol"]), dict(c, data=c["data"].copy())))
    clouds = []
    for kind, f, t, c in out:
        d = np.zeros((), dtype=drvlib.CLOUD_DTYPE)
        d["from_gcol"], d["to_gcol"], d["kind"] = f, t, kind
        d["width"], d["height"], d["point_step"], d["stamp_ns"] = c["width"], c["height"], c["point_step"], c["stamp_ns"]
        dt = drvlib.pointcloud2_dtype(c["n_fields"])
        clouds.append((d, c["data"].view(dt).reshape(c["height"], c["width"])))
    return clouds


def device_clouds_batched(cc, pts, poses, chunk):
    """Same as device_clouds, but every message of a push comes from ONE launch (cc_pack_requests_pointcloud2)."""
    clouds = []
    for a in range(0, pts.shape[0], chunk):
        res = cc.addFirings(pts[a:a + chunk], poses[a:a + chunk])
        req, nxt = [], 0
        for e in res.events.copy():
            while nxt < int(e["n_clusters_before"]):
                if res.clusters[nxt]["num_points"] > 20:
                    req.append((2, nxt, nxt))
                nxt += 1
            if e["to_gcol"] >= e["from_gcol"]:
                req.append((0 if e["ground_points_only"] else 1, int(e["from_gcol"]), int(e["to_gcol"])))
        for (kind, f, t), c in zip(req, cc.pack_requests_pointcloud2(req)):
            d = np.zeros((), dtype=drvlib.CLOUD_DTYPE)
            d["from_gcol"], d["to_gcol"], d["kind"] = (-1, -1, 2) if kind == 2 else (f, t, kind)
            d["width"], d["height"], d["point_step"], d["stamp_ns"] = c["width"], c["height"], c["point_step"], c["stamp_ns"]
            clouds.append((d, c["data"].copy().view(drvlib.pointcloud2_dtype(c["n_fields"])).reshape(c["height"], c["width"])))
    return clouds


def check_device_packer(library, spec, kw, cfg_over, chunk):
    from test_emu_parity import make_cc

    _need(drvlib.REF_LIB)
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    want, _ = record_clouds(drvlib.REF_LIB, pts, poses, sp, cfg)
    cc = make_cc(library, cfg, sp.rows)
    got = device_clouds(cc, pts, poses, chunk)
    compare_clouds(want, got)
    cc = make_cc(library, cfg, sp.rows)
    got = device_clouds_batched(cc, pts, poses, chunk)
    compare_clouds(want, got)


@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES)
def test_device_packed_messages_identical_emulation(emu_library, spec, kw, cfg_over, batch):
    check_device_packer(emu_library, spec, kw, cfg_over, batch)


@pytest.mark.gpu
@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES + [("velodyne64", dict(n_rotations=2.2, moving=True), {}, 1024)])
def test_device_packed_messages_identical_cuda(cuda_library, spec, kw, cfg_over, batch):
    check_device_packer(None, spec, kw, cfg_over, batch)
