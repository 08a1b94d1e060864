"""Diagnostic (GPU box): pinned host->device bandwidth at the size of one push, and where the wall time of the e2e leg goes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration

n = 6488064
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 50
print(f"H2D pinned {n} B: {dt*1e6:.1f} us = {n/dt/1e9:.1f} GB/s")
h2 = torch.empty(800000, dtype=torch.uint8).pin_memory()
t0 = time.perf_counter()
for _ in range(50):
    h2.copy_(d[:800000], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 50
print(f"D2H pinned 800000 B: {dt*1e6:.1f} us")

B = 2048
base_pts, base_poses, sp = bench.make_rotations()
cfg = stream_configuration(bench.SPEC)
R = sp.rows
for prefetch in (True, False):
    cc = ContinuousClustering(device=0, max_firings_per_push=B)
    cc.setConfiguration(cfg); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
    cc.set_label_prefetch(prefetch)
    total = 30 * B
    pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, total)
    pin_pts = torch.from_numpy(pts.view(np.uint8).reshape(total, R * 48)).pin_memory()
    pin_poses = torch.from_numpy(poses).pin_memory()
    hp = pin_pts.numpy().view(pts.dtype).reshape(total, R); hq = pin_poses.numpy()
    for s in range(5):
        cc.addFirings(hp[s*B:(s+1)*B], hq[s*B:(s+1)*B])
    torch.cuda.synchronize()
    ts, tw = [], []
    t00 = time.perf_counter()
    cc.submitFirings(hp[5*B:6*B], hq[5*B:6*B])
    for s in range(5, 29):
        t0 = time.perf_counter()
        cc.submitFirings(hp[(s+1)*B:(s+2)*B], hq[(s+1)*B:(s+2)*B])
        t1 = time.perf_counter()
        res = cc.wait()
        t2 = time.perf_counter()
        ts.append(t1 - t0); tw.append(t2 - t1)
    cc.wait()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t00
    print(f"label_prefetch={prefetch}: per push wall {tot/25*1e6:.1f} us; submit call p50 {np.median(ts)*1e6:.1f} us; wait call p50 {np.median(tw)*1e6:.1f} us; device_ms {res.info.device_ms*1e3:.1f} us")
    cc.close()
