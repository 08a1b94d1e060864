"""Diagnostic (GPU box): (a) repeated passes of host->device copies over the same page-locked buffer, (b) PCIe link
generation / clocks while the end-to-end leg runs, (c) visited-fix statistics of the bench stream."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.cuda.init()
N = 13 << 20
tot = 20
src = torch.empty(tot * N, dtype=torch.uint8).pin_memory(); src.numpy()[:] = 1
dst = torch.empty(3 * N, dtype=torch.uint8, device="cuda")
def q():
    return subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,clocks.sm,pstate", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True).stdout.strip()
print("idle:", q())
for p in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(tot):
        dst[(i % 3) * N:(i % 3 + 1) * N].copy_(src[i * N:(i + 1) * N], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"pass {p}: {tot * N / dt / 1e9:.1f} GB/s", q())
time.sleep(2.0)
print("after 2 s idle:", q())
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(tot):
    dst[(i % 3) * N:(i % 3 + 1) * N].copy_(src[i * N:(i + 1) * N], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pass after idle: {tot * N / dt / 1e9:.1f} GB/s", q())

import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration
B = 4096
base_pts, base_poses, sp = bench.make_rotations()
R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=B)
cc.setConfiguration(stream_configuration(bench.SPEC)); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
cc.set_label_prefetch(True)
n = 24
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * B)
pp = torch.from_numpy(pts.view(np.uint8).reshape(n * B, R * 48)).pin_memory(); pq = torch.from_numpy(poses).pin_memory()
hp = pp.numpy().view(pts.dtype).reshape(n * B, R); hq = pq.numpy()
samples = []
stop = False
def sampler():
    while not stop:
        samples.append(q())
th = threading.Thread(target=sampler); th.start()
for rep in range(3):
    cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
    for s in range(3):
        cc.addFirings(hp[s * B:(s + 1) * B], hq[s * B:(s + 1) * B])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cc.submitFirings(hp[3 * B:4 * B], hq[3 * B:4 * B]); cc.submitFirings(hp[4 * B:5 * B], hq[4 * B:5 * B])
    rc = []
    for s in range(3, n):
        if s + 2 < n:
            cc.submitFirings(hp[(s + 2) * B:(s + 3) * B], hq[(s + 2) * B:(s + 3) * B])
        r = cc.wait(); rc.append(int(r.info.visited_recounts))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"e2e rep {rep}: {(n - 3) * B / dt / 1e6:.2f} M col/s = {(n - 3) * B * R * 48 / dt / 1e9:.1f} GB/s; visited recounts per push {rc[:6]}")
stop = True; th.join()
print("link samples during e2e:", sorted(set(samples)))
cc.close()
