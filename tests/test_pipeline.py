"""The whole dataset-replay pipeline of the reference's kitti_demo (kitti_demo.cpp:173-224, 369-403) on the device, end to end,
against the reference's own code for every stage:
    frame (x, y, z, intensity in file order)
      -> cc_kitti_frame          (recoverLaserIndices / undoEgoMotionCorrection / generateRangeImage / pseudo firings / poses)
      -> cc_push_firings_device  (the hot path; the firings never leave the device)
      -> clustered-column events + cc_export_columns  (what kitti_demo's evaluation callback reads: guid, ground label, id)
      -> cc_eval_frame           (ground confusion counts, over- / under-segmentation entropy)
Reference chain: oracle/_ref/libcc_eval_ref.so (kitti_loader.cpp / kitti_demo.cpp / kitti_evaluation.cpp excerpts) around
oracle/_ref/libcc_ref.so (the reference's continuous_clustering.cpp with kitti_demo's callback compiled in)."""
import ctypes as C
import os

import numpy as np
import pytest

from continuous_clustering_b200 import ContinuousClustering, KittiEvaluation, KittiReplay
from continuous_clustering_b200.presets import stream_configuration
from continuous_clustering_b200.synth import make_kitti_frame
from oracle import drvlib
from test_callers import id_bijection
from test_evaluation import check as check_eval
from test_evaluation import reference as reference_eval
from test_kitti import EVAL_REF, H, W
from test_kitti import reference as reference_front_end

GP_GROUND = 54  # PointCloudColors GREEN (general.hpp:208-357)
FRAMES = [2, 3, 4]
ROBOT_FROM_SENSOR = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1.73], dtype=np.float64)


def frames():
    out = []
    for f in FRAMES:
        xyzi, s0, s1, pstamps, poses, mid = make_kitti_frame(seed=21, frame_index=f, n_poses=8)
        out.append((f - FRAMES[0], xyzi, s0, s1, pstamps, poses, mid))
    return out


def labels_for_eval(n, seed):
    rng = np.random.RandomState(seed)
    sem = rng.choice(np.array([60, 40, 44, 48, 49, 72, 0, 10, 30, 50, 70, 80], dtype=np.uint16), size=n).astype(np.uint16)
    gt = rng.randint(0, 40, size=n).astype(np.uint32)
    return sem, gt


def product_chain(library):
    cfg = stream_configuration("kitti64")
    cc = ContinuousClustering(max_firings_per_push=1100, _library=library)
    cc.setConfiguration(cfg)
    cc.reset(H)
    cc.setTransformRobotFrameFromSensorFrame(ROBOT_FROM_SENSOR)
    kr = KittiReplay(_library=library)
    fr = frames()
    flags = {f: np.zeros(x.shape[0], np.uint8) for f, x, *_ in fr}
    det = {f: np.zeros(x.shape[0], np.uint32) for f, x, *_ in fr}
    for f, xyzi, s0, s1, pstamps, poses, mid in fr:
        kr.set_poses(pstamps, poses)
        info = kr.frame(xyzi, s0, s1, mid, 0, f)
        for k in range(0, W, 1100):
            res = cc.addFiringsDevice(info["d_firings"] + k * H * 48, info["d_poses"] + k * 96, 1100, H)
            ev = res.events.copy()
            for e in ev[(ev["ground_points_only"] == 0) & (ev["to_gcol"] >= ev["from_gcol"])]:
                cells = cc.export_columns(int(e["from_gcol"]), int(e["to_gcol"]))
                guid = cells["globally_unique_point_index"].ravel()
                ok = guid != np.uint64(0xFFFFFFFFFFFFFFFF)  # kitti_demo.cpp:196
                fi = ((guid[ok] >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64)
                pi = (guid[ok] & np.uint64(0xFFFFFFFF)).astype(np.int64)
                ground = cells["ground_point_label"].ravel()[ok] == GP_GROUND
                ids = cells["id"].ravel()[ok].astype(np.uint32)
                for frame_index in np.unique(fi):
                    m = fi == frame_index
                    flags[int(frame_index)][pi[m]] = 1 | (ground[m].astype(np.uint8) << 1)
                    det[int(frame_index)][pi[m]] = ids[m]
    ev_dev = KittiEvaluation(_library=library)
    metrics = {}
    for f, xyzi, *_ in fr:
        sem, gt = labels_for_eval(xyzi.shape[0], f)
        metrics[f] = ev_dev.evaluate(sem, (flags[f] >> 1) & 1, gt, det[f])
    ev_dev.close()
    kr.close()
    cc.close()
    return flags, det, metrics


def reference_chain():
    cfg = drvlib.stream_config("kitti64")
    d = drvlib.Driver(drvlib.REF_LIB)
    d.configure(cfg, H, robot_from_sensor=ROBOT_FROM_SENSOR)
    d.set_record(0)
    fr = frames()
    d.kitti_begin(0, [x.shape[0] for _, x, *_ in fr])
    for f, xyzi, s0, s1, pstamps, poses, mid in fr:
        ref = reference_front_end(xyzi, s0, s1, pstamps, poses, mid, 0, f)
        d.add_firings(np.ascontiguousarray(ref["firings"]), np.ascontiguousarray(ref["poses"]))
    flags, det, metrics = {}, {}, {}
    for f, xyzi, *_ in fr:
        fl, lab = d.kitti_get(f)
        flags[f], det[f] = fl, lab
        sem, gt = labels_for_eval(xyzi.shape[0], f)
        metrics[f] = reference_eval(sem, ((fl >> 1) & 1).astype(np.uint8), gt, lab.astype(np.uint32))
    d.close()
    return flags, det, metrics


def compare_chains(library):
    if not (os.path.exists(EVAL_REF) and os.path.exists(drvlib.REF_LIB)):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    want_flags, want_det, want_metrics = reference_chain()
    got_flags, got_det, got_metrics = product_chain(library)
    seen = 0
    for f in want_flags:
        assert np.array_equal(want_flags[f], got_flags[f]), f"frame {f}: has_corresponding_point / is_ground_point flags differ"
        id_bijection(want_det[f], got_det[f])
        check_eval(want_metrics[f], got_metrics[f], f"frame {f} metrics")
        seen += int((want_flags[f] & 1).sum())
    assert seen > 100000, "the replay must have delivered most points of the frames through clustered-column events"


def test_replay_pipeline_matches_the_reference_emulation(emu_library):
    compare_chains(emu_library)


@pytest.mark.gpu
def test_replay_pipeline_matches_the_reference_cuda(cuda_library):
    compare_chains(None)
