"""Diagnostic (GPU box): wall clock of short synchronous host pushes (latency mode), fused kernel vs kernel chain.
usage: lat_probe.py [firings_per_push] ; CC_B200_FUSED_MAX=0 selects the chain."""
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
cc.set_label_prefetch(True)
n = 600
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
hp = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).pin_memory()
hq = torch.from_numpy(poses).pin_memory()
p0, q0 = hp.data_ptr(), hq.data_ptr()
L, h = cc._L, cc._h
raw, full, dev, launches = [], [], [], []
for s in range(n):
    t0 = time.perf_counter()
    rc = L.cc_push_firings(h, B, R, p0 + s * B * R * 48, q0 + s * B * 96)
    t1 = time.perf_counter()
    assert rc == 0, cc._L.cc_last_error(h)
    res = cc._collect()
    t2 = time.perf_counter()
    raw.append(t1 - t0); full.append(t2 - t0); dev.append(res.info.device_ms); launches.append(res.info.gpu_launches)
raw, full, dev = np.array(raw[100:]) * 1e6, np.array(full[100:]) * 1e6, np.array(dev[100:]) * 1e3
print(f"B={B} fused_max={os.environ.get('CC_B200_FUSED_MAX','default')} launches/push={np.median(launches):.0f}: "
      f"cc_push_firings p50 {np.median(raw):.1f} us p99 {np.percentile(raw,99):.1f} us | +collect p50 {np.median(full):.1f} | device p50 {np.median(dev):.1f} us "
      f"| exact pushes {0}")
cc.close()
