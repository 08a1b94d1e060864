"""Diagnostic (GPU box): wall-clock split of small synchronous host pushes (latency mode)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base_pts, base_poses, sp = bench.make_rotations()
cfg = stream_configuration(bench.SPEC)
R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=max(B, 256))
cc.setConfiguration(cfg); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
n = 400
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
hp = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).pin_memory()
hq = torch.from_numpy(poses).pin_memory()
wall, dev = [], []
for s in range(n):
    t0 = time.perf_counter()
    res = cc.addFirings(hp[s * B:(s + 1) * B].numpy().view(np.uint8), hq[s * B:(s + 1) * B].numpy(), B, R) if False else None
    break
import inspect
print([m for m in dir(cc) if not m.startswith('__')])
