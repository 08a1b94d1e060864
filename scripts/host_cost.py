"""Diagnostic (GPU box): host time of the submit / wait calls of the device-resident pipelined loop."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base_pts, base_poses, sp = bench.make_rotations()
cfg = stream_configuration(bench.SPEC)
R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=B)
cc.setConfiguration(cfg); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
n = 30
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
d_pts = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).cuda()
d_poses = torch.from_numpy(poses).cuda()
sub = lambda s: cc.submitFiringsDevice(d_pts.data_ptr() + s * B * R * 48, d_poses.data_ptr() + s * B * 96, B, R)
for s in range(4):
    cc.addFiringsDevice(d_pts.data_ptr() + s * B * R * 48, d_poses.data_ptr() + s * B * 96, B, R)
torch.cuda.synchronize()
import ctypes as C
cc._L.cc_debug_slot_base(cc._h)
out = (C.c_float * 5)()
ts, tw, dev, marks = [], [], [], []
t00 = time.perf_counter()
sub(4)
for s in range(4, n - 1):
    t0 = time.perf_counter(); sub(s + 1); t1 = time.perf_counter(); r = cc.wait(); t2 = time.perf_counter()
    cc._L.cc_debug_slot_times(cc._h, s % 2, out)
    marks.append((s, (t0 - t00) * 1e3, (t1 - t00) * 1e3, (t2 - t00) * 1e3, out[2], out[3], out[4]))
    ts.append(t1 - t0); tw.append(t2 - t1); dev.append(r.info.device_ms)
cc.wait(); torch.cuda.synchronize()
tot = time.perf_counter() - t00
print(f"batch {B}: per step wall {tot/(n-5)*1e6:.1f} us; submit p50 {np.median(ts)*1e6:.1f} us; wait p50 {np.median(tw)*1e6:.1f} us; device p50 {np.median(dev)*1e3:.1f} us")
for m in marks[8:16]:
    print('push %d: host submit(next) %.3f->%.3f wait ->%.3f | kernels %.3f->%.3f results %.3f' % m)
#
L = cc._L
t0 = time.perf_counter()
for s in range(5):
    pass
cc.close()
