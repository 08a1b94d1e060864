"""Diagnostic (GPU box): device time per push when the sensor does not deliver exactly one firing per column (az_step_scale) or
jitters (az_jitter): how much of a push leaves the regular insertion path and what it costs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from continuous_clustering_b200 import ContinuousClustering, synth
from continuous_clustering_b200.presets import stream_configuration
IDENT = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
B = 4096
for kw in ({}, {"az_step_scale": 0.95}, {"az_step_scale": 1.05}, {"az_jitter": 0.2}, {"az_jitter": 0.6}, {"az_step_scale": 0.97, "az_jitter": 0.1}):
    pts, poses, sp = synth.make_stream("velodyne64", n_firings=10 * B, **kw)
    R = sp.rows
    cc = ContinuousClustering(max_firings_per_push=B)
    cc.setConfiguration(stream_configuration("velodyne64")); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(IDENT)
    d_pts = torch.from_numpy(pts.view(np.uint8).reshape(-1, R * 48)).cuda(); d_poses = torch.from_numpy(poses).cuda()
    dev, slow, ex = [], [], 0
    for s in range(10):
        r = cc.addFiringsDevice(d_pts.data_ptr() + s * B * R * 48, d_poses.data_ptr() + s * B * 96, B, R)
        dev.append(r.info.device_ms); slow.append(int(r.info.slow_insert_firings)); ex += int(r.info.used_exact_path)
    print(kw, "device ms per push (median of last 6): %.3f" % float(np.median(dev[4:])), "slow firings per push:", slow[4:], "exact pushes", ex,
          "reset required", cc.resetRequired(), flush=True)
    cc.close()
