"""Diagnostic (GPU box): timeline of the kernels of one push as they overlap in normal operation (cc_debug_trace)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
FLUSH = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
WALL = len(sys.argv) > 3 and sys.argv[3] == "wall"  # closed wall around the sensor: the split / exact path
base_pts, base_poses, sp = bench.make_rotations()
if WALL:
    from continuous_clustering_b200 import synth
    base_pts, base_poses, sp = synth.make_stream(bench.SPEC, n_rotations=2.0, seed=3, n_boxes=0, wall_radius=12.0)
cfg = stream_configuration(bench.SPEC)
R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=max(B, 256))
cc.setConfiguration(cfg); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
n = 12
total = n * B
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, total)
d_pts = torch.from_numpy(pts.view(np.uint8).reshape(total, R * 48)).cuda()
d_poses = torch.from_numpy(poses).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(cc.stream)
for s in range(n):
    if s == n - 3:
        cc.debug_trace(True)
    if FLUSH:
        with torch.cuda.stream(stream):
            flush.fill_(s)
    res = cc.addFiringsDevice(d_pts.data_ptr() + s * B * R * 48, d_poses.data_ptr() + s * B * 96, B, R)
    if s >= n - 3:
        tr = sorted(cc.get_trace(), key=lambda t: t[1])
        t0 = tr[0][1]
        print(f"push {s}: device_ms {res.info.device_ms*1e3:.1f} us, n_clusters {res.info.n_clusters}, visited recounts {res.info.visited_recounts}")
        prev_end = t0
        for name, a, z, longest, blocks in tr:
            print(f"  {name:18s} start {(a-t0)/1e3:7.1f}  end {(z-t0)/1e3:7.1f}  span {(z-a)/1e3:6.1f}  gap_after_prev {(a-prev_end)/1e3:6.1f}  longest_block {longest/1e3:6.1f}  blocks {blocks}")
            prev_end = max(prev_end, z)
cc.close()
