"""the KITTI replay front-end on the device (csrc/cc_kitti.cuh, cc_kitti_* C ABI) against the reference's own
code -- KittiLoader::recoverLaserIndices / undoEgoMotionCorrection / generateRangeImage / interpolate and kitti_demo's
makePseudoFiringFromRangeImageColumn, cut out of /root/reference at build time and compiled unmodified
(oracle/cc_eval_driver.cpp -> oracle/_ref/libcc_eval_ref.so) -- and against digests of that build's output committed under
tests/golden/kitti_golden.json (tests/golden/make_kitti_golden.py). Everything is compared bit for bit: laser indices,
the point of every range-image cell, the un-corrected coordinates, the 2200 x 64 RawPoint records, the 2200 poses."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from continuous_clustering_b200 import KittiReplay
from continuous_clustering_b200.synth import RAW_POINT_DTYPE, make_kitti_frame

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EVAL_REF = os.path.join(REPO, "oracle", "_ref", "libcc_eval_ref.so")
GOLDEN = os.path.join(REPO, "tests", "golden", "kitti_golden.json")
W, H = 2200, 64
FIELDS = ["x", "y", "z", "firing_index", "intensity", "stamp", "globally_unique_point_index"]

# (name, make_kitti_frame arguments, sequence index)
CASES = [
    ("street", dict(seed=7, frame_index=3), 0),
    ("dense", dict(seed=11, frame_index=5, dropout=0.0, n_boxes=300), 4),
    ("sparse", dict(seed=13, frame_index=2, dropout=0.6, n_boxes=20), 10),
    ("fewer_rows", dict(seed=17, frame_index=4, top_rows_empty=6), 1),
]


def make_case(name):
    for n, kw, seq in CASES:
        if n == name:
            xyzi, s0, s1, pstamps, poses, mid = make_kitti_frame(**kw)
            if name == "sparse":
                # more than 64 azimuth wraps: the points behind the last row keep laser index 0 (kitti_loader.cpp:74-76)
                xyzi = np.concatenate([xyzi, xyzi[: 3 * 1500]])
            return xyzi, s0, s1, pstamps, poses, mid, seq, kw["frame_index"]
    raise KeyError(name)


def reference(xyzi, s0, s1, pstamps, poses, mid, seq, frame):
    lib = C.CDLL(EVAL_REF)
    vp = C.c_void_p
    lib.ev_frame_to_firings.argtypes = [C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    n = xyzi.shape[0]
    firings = np.zeros(W * H, dtype=RAW_POINT_DTYPE)
    fposes = np.zeros((W, 12))
    laser = np.zeros(n, np.uint8)
    cell_of_point = np.zeros(n, np.int32)
    unc = np.zeros((n, 3), np.float32)
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    rc = lib.ev_frame_to_firings(n, xyzi.ctypes.data, s0, s1, mid.ctypes.data, len(pstamps), pstamps.ctypes.data, poses.ctypes.data,
                                 seq, frame, firings.ctypes.data, fposes.ctypes.data, laser.ctypes.data, cell_of_point.ctypes.data,
                                 unc.ctypes.data)
    assert rc == 0
    cell_point = np.full(W * H, -1, np.int32)
    idx = np.nonzero(cell_of_point >= 0)[0]
    cell_point[cell_of_point[idx]] = idx
    return {"firings": firings.reshape(W, H), "poses": fposes, "laser_index": laser, "cell_point": cell_point.reshape(H, W),
            "uncorrected": unc}


def product(library, xyzi, s0, s1, pstamps, poses, mid, seq, frame):
    kr = KittiReplay(max_points_per_frame=1 << 18, _library=library)
    kr.set_poses(pstamps, poses)
    info = kr.frame(xyzi, s0, s1, mid, seq, frame)
    out = kr.read_debug()
    out["poses"] = info["poses"]
    out["info"] = info
    kr.close()
    return out


def digest(res):
    h = hashlib.sha256()
    for f in FIELDS:  # (the padding bytes of RawPoint are not part of the contract)
        h.update(np.ascontiguousarray(res["firings"][f]).tobytes())
    h.update(np.ascontiguousarray(res["poses"]).tobytes())
    h.update(res["laser_index"].tobytes())
    h.update(np.ascontiguousarray(res["cell_point"]).tobytes())
    h.update(res["uncorrected"].tobytes())
    return h.hexdigest()


def compare(ref, got, what):
    assert np.array_equal(ref["laser_index"], got["laser_index"]), f"{what}: laser indices"
    assert np.array_equal(ref["uncorrected"].view(np.uint32), got["uncorrected"].view(np.uint32)), f"{what}: un-corrected points"
    assert np.array_equal(ref["cell_point"], got["cell_point"]), f"{what}: range-image cells"
    for f in FIELDS:
        a, b = ref["firings"][f], got["firings"][f]
        if a.dtype.kind == "f":
            a, b = a.view(np.uint32), b.view(np.uint32)
        assert np.array_equal(a, b), f"{what}: firing field {f}"
    assert np.array_equal(ref["poses"].view(np.uint64), got["poses"].view(np.uint64)), f"{what}: firing poses"


def need_ref():
    if not os.path.exists(EVAL_REF):
        pytest.skip("oracle/_ref/libcc_eval_ref.so not built (needs /root/reference at build time)")


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_reference_build_matches_golden(name):
    need_ref()
    golden = json.load(open(GOLDEN))
    assert digest(reference(*make_case(name))) == golden[name]["sha256"]


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_emulated_kernels_match_reference(emu_library, name):
    case = make_case(name)
    got = product(emu_library, *case)
    golden = json.load(open(GOLDEN))
    assert digest(got) == golden[name]["sha256"], "emulated kernels vs golden digest of the reference build"
    assert got["info"]["rows_found"] == golden[name]["rows_found"]
    if os.path.exists(EVAL_REF):
        compare(reference(*case), got, "emulated kernels vs reference")


def test_row_overflow_and_missing_rows_are_covered():
    golden = json.load(open(GOLDEN))
    assert golden["sparse"]["rows_found"] == 65 and golden["fewer_rows"]["rows_found"] < 64 and golden["street"]["rows_found"] == 64


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_cuda_matches_reference(cuda_library, name):
    case = make_case(name)
    got = product(None, *case)
    golden = json.load(open(GOLDEN))
    assert digest(got) == golden[name]["sha256"], "cuda vs golden digest of the reference build"
    if os.path.exists(EVAL_REF):
        compare(reference(*case), got, "cuda vs reference")


@pytest.mark.gpu
def test_cuda_frame_feeds_the_clustering(cuda_library):
    """The device-resident firings of a frame go straight into the hot path (cc_push_firings_device) and give the same
    result as the host copy of the same records through cc_push_firings."""
    from continuous_clustering_b200 import ContinuousClustering
    from continuous_clustering_b200.presets import stream_configuration

    xyzi, s0, s1, pstamps, poses, mid, seq, frame = make_case("street")
    kr = KittiReplay()
    kr.set_poses(pstamps, poses)
    info = kr.frame(xyzi, s0, s1, mid, seq, frame)
    host = kr.read_debug()["firings"]
    cfg = stream_configuration("kitti64")
    results = []
    for device_input in (True, False):
        cc = ContinuousClustering(max_firings_per_push=1100)
        cc.setConfiguration(cfg)
        cc.reset(H)
        cc.setTransformRobotFrameFromSensorFrame(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1.73], dtype=np.float64))
        res = []
        for k in range(0, W, 1100):
            if device_input:
                r = cc.addFiringsDevice(info["d_firings"] + k * H * 48, info["d_poses"] + k * 96, 1100, H)
            else:
                r = cc.addFirings(np.ascontiguousarray(host[k:k + 1100]), np.ascontiguousarray(info["poses"][k:k + 1100]))
            res.append((r.events.tobytes(), int(r.info.n_clusters), int(r.info.n_cluster_points)))
        results.append(res)
        del cc
    assert results[0] == results[1]
    assert sum(r[1] for r in results[0]) >= 0
    kr.close()


def test_empty_and_tiny_frames(emu_library):
    """Degenerate inputs: a frame without points gives 2200 firings of empty cells (NaN coordinates, guid of an empty cell);
    a frame with a single point puts it into exactly one cell. Same on the reference build where it is present."""
    _, s0, s1, pstamps, poses, mid, seq, frame = make_case("street")
    for pts in (np.zeros((0, 4), np.float32), np.array([[5.0, -2.0, -1.0, 0.5]], np.float32)):
        got = product(emu_library, pts, s0, s1, pstamps, poses, mid, seq, frame)
        f = got["firings"]
        assert int((~np.isnan(f["x"])).sum()) == pts.shape[0]
        assert int((f["globally_unique_point_index"] != np.uint64(0xFFFFFFFFFFFFFFFF)).sum()) == pts.shape[0]
        assert got["info"]["rows_found"] == (1 if pts.shape[0] else 0)
        assert np.array_equal(f["firing_index"][:, 0], np.arange(W, dtype=np.uint64))
        if os.path.exists(EVAL_REF) and pts.shape[0]:
            compare(reference(pts, s0, s1, pstamps, poses, mid, seq, frame), got, "single point vs reference")
