#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_edge_cases.py tests/test_gpu_parity.py tests/test_facade.py -x -q -m gpu 2>&1 | tail -3 )
python - <<PY
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration
base_pts, base_poses, sp = bench.make_rotations(); R = sp.rows; B = 4096
for nth in (1, 4):
    cc = ContinuousClustering(device=0, max_firings_per_push=B)
    cc.setConfiguration(stream_configuration(bench.SPEC, cluster_point_trees_every_nth_column=nth)); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
    n = 12
    pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
    d_pts = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).cuda(); d_poses = torch.from_numpy(poses).cuda()
    dev = []; ex = 0
    for s in range(n):
        r = cc.addFiringsDevice(d_pts.data_ptr() + s * B * R * 48, d_poses.data_ptr() + s * B * 96, B, R)
        dev.append(r.info.device_ms); ex += int(r.info.used_exact_path)
    print('nth', nth, 'device ms per 4096-firing push (median of last 8):', round(float(np.median(dev[4:])), 4), 'exact pushes', ex)
    cc.close()
PY
