"""Diagnostic (GPU box): host and device milestones of the end-to-end leg, push by push."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base_pts, base_poses, sp = bench.make_rotations()
cfg = stream_configuration(bench.SPEC)
R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=max(B, 256))
cc.setConfiguration(cfg); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
cc.set_label_prefetch(True)
n = 16
total = n * B
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, total)
pin_pts = torch.from_numpy(pts.view(np.uint8).reshape(total, R * 48)).pin_memory()
pin_poses = torch.from_numpy(poses).pin_memory()
hp = pin_pts.numpy().view(pts.dtype).reshape(total, R); hq = pin_poses.numpy()
for _ in range(3):  # DMA-warm staging memory (see bench.py)
    _w = pin_pts.cuda(); _w2 = pin_poses.cuda(); torch.cuda.synchronize(); del _w, _w2
for s in range(4):
    cc.addFirings(hp[s*B:(s+1)*B], hq[s*B:(s+1)*B])
L = cc._L
L.cc_debug_slot_base(cc._h)
t_base = time.perf_counter()
out = (C.c_float * 5)()
host = []
cc.submitFirings(hp[4*B:5*B], hq[4*B:5*B])
cc.submitFirings(hp[5*B:6*B], hq[5*B:6*B])
for s in range(4, n - 2):
    a = time.perf_counter()
    cc.submitFirings(hp[(s+2)*B:(s+3)*B], hq[(s+2)*B:(s+3)*B])
    b = time.perf_counter()
    L.cc_debug_slot_times(cc._h, s % 2, out)   # slot of push s (slots alternate; push 4 used slot 0), before it is reused
    cc.wait()
    c = time.perf_counter()
    o1 = list(out)
    host.append((s, (a - t_base) * 1e3, (b - t_base) * 1e3, (c - t_base) * 1e3, o1))
cc.wait(); cc.wait()
for s, a, b, c, o in host:
    print(f"push {s}: host submit(s+2) {a:7.3f}->{b:7.3f} wait(s) ->{c:7.3f} | dev h2d {o[0]:7.3f}->{o[1]:7.3f} kernels {o[2]:7.3f}->{o[3]:7.3f} results {o[4]:7.3f}")
cc.close()
