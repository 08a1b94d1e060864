"""Diagnostic (GPU box): does the kernel chain of a push slow the host->device copy engine down? 20 back-to-back copies of
one push's input (12.6 MB, page-locked, DMA-warm) alone, and while device-resident pushes run on the handle's stream."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration
B = 4096
base_pts, base_poses, sp = bench.make_rotations(); R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=B)
cc.setConfiguration(stream_configuration(bench.SPEC)); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
n = 40
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
d_pts = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).cuda(); d_poses = torch.from_numpy(poses).cuda()
N = B * R * 48
tot = 20
src = torch.empty(tot * N, dtype=torch.uint8).pin_memory(); src.numpy()[:] = 1
dst = torch.empty(3 * N, dtype=torch.uint8, device="cuda")
s_copy = torch.cuda.Stream()
warm = src.cuda(); torch.cuda.synchronize(); del warm
def copies():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_copy):
        e0.record(s_copy)
        for i in range(tot):
            dst[(i % 3) * N:(i % 3 + 1) * N].copy_(src[i * N:(i + 1) * N], non_blocking=True)
        e1.record(s_copy)
    return e0, e1
def pushes(k):
    cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
    p, q = d_pts.data_ptr(), d_poses.data_ptr()
    cc.submitFiringsDevice(p, q, B, R)
    for s in range(1, k):
        cc.submitFiringsDevice(p + s * B * R * 48, q + s * B * 96, B, R)
        cc.wait()
    cc.wait()
for rep in range(2):
    e0, e1 = copies(); torch.cuda.synchronize()
    print(f"copies alone: {tot * N / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
    pushes(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); pushes(30); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"pushes alone: {30 * B / dt / 1e6:.2f} M col/s")
    e0, e1 = copies()
    t0 = time.perf_counter(); pushes(38); dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"copies while pushes run: {tot * N / e0.elapsed_time(e1) / 1e6:.1f} GB/s ({e0.elapsed_time(e1):.2f} ms of copies, {dt * 1e3:.2f} ms of pushes at {38 * B / dt / 1e6:.2f} M col/s)")
    # the same with a plain bandwidth-light torch kernel stream instead of pushes
    x = torch.zeros(1 << 20, device="cuda")
    e0, e1 = copies()
    for i in range(400):
        x.add_(1.0)
    torch.cuda.synchronize()
    print(f"copies while 400 tiny kernels run: {tot * N / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
cc.close()
