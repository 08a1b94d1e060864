"""Diagnostic (GPU box): end-to-end throughput (host buffers, DMA-warm) under variations of the host side."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
labels = int(sys.argv[2]) if len(sys.argv) > 2 else 1
base_pts, base_poses, sp = bench.make_rotations(); R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=max(B, 256))
cc.setConfiguration(stream_configuration(bench.SPEC)); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
cc.set_label_prefetch(bool(labels))
n = 24
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
pp = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).pin_memory(); pq = torch.from_numpy(poses).pin_memory()
hp = pp.numpy().view(pts.dtype).reshape(n * B, R); hq = pq.numpy()
if os.environ.get('E2E_WARM'):
    for _ in range(3):
        _w = pp.cuda(); torch.cuda.synchronize(); del _w
for rep in range(3):
    cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
    for s in range(3):
        cc.addFirings(hp[s * B:(s + 1) * B], hq[s * B:(s + 1) * B])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cc.submitFirings(hp[3 * B:4 * B], hq[3 * B:4 * B]); cc.submitFirings(hp[4 * B:5 * B], hq[4 * B:5 * B])
    for s in range(3, n):
        if s + 2 < n:
            cc.submitFirings(hp[(s + 2) * B:(s + 3) * B], hq[(s + 2) * B:(s + 3) * B])
        r = cc.wait()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"B={B} labels={labels} split={os.environ.get('CC_B200_H2D_SPLIT','1')} block={os.environ.get('CC_B200_BLOCKING_WAIT','0')} rep {rep}: {(n - 3) * B / dt / 1e6:.2f} M col/s = {(n - 3) * B * R * 48 / dt / 1e9:.1f} GB/s")
cc.close()
