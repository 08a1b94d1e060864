"""Diagnostic (GPU box): which 64-firing pushes are the slow ones? Device time of every push beside what it produced."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration
LB = 64
base_pts, base_poses, sp = bench.make_rotations(); R = sp.rows
cc = ContinuousClustering(device=0, max_firings_per_push=256)
cc.setConfiguration(stream_configuration(bench.SPEC)); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
cc.set_label_prefetch(True)
n = 260
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, n * LB)
pp = torch.from_numpy(pts.view(np.uint8).reshape(n * LB, R * 48)).pin_memory(); pq = torch.from_numpy(poses).pin_memory()
hp = pp.numpy().view(pts.dtype).reshape(n * LB, R); hq = pq.numpy()
rows = []
for i in range(n):
    t0 = time.perf_counter()
    r = cc.addFirings(hp[i * LB:(i + 1) * LB], hq[i * LB:(i + 1) * LB])
    dt = 1e6 * (time.perf_counter() - t0)
    rows.append((i, dt, 1e3 * r.info.device_ms, int(r.info.n_events), int(r.info.n_clusters), int(r.info.n_cluster_points),
                 int(r.info.ground_to_gcol - r.info.ground_from_gcol), int(getattr(r.info, "n_unfinished_trees", 0))))
rows = rows[60:]
dev = np.array([x[2] for x in rows]); cp = np.array([x[5] for x in rows]); cl = np.array([x[4] for x in rows]); tr = np.array([x[7] for x in rows])
print("device us percentiles 10/50/90/99:", np.percentile(dev, [10, 50, 90, 99]).round(1))
slow = dev > np.median(dev) + 15
print("slow pushes:", int(slow.sum()), "of", len(rows))
print("fast: mean cluster points %.0f clusters %.1f trees %.0f | slow: cluster points %.0f clusters %.1f trees %.0f" % (
    cp[~slow].mean(), cl[~slow].mean(), tr[~slow].mean(), cp[slow].mean() if slow.any() else 0, cl[slow].mean() if slow.any() else 0, tr[slow].mean() if slow.any() else 0))
print("corr(device, cluster points) %.2f corr(device, clusters) %.2f corr(device, trees) %.2f" % (np.corrcoef(dev, cp)[0, 1], np.corrcoef(dev, cl)[0, 1], np.corrcoef(dev, tr)[0, 1]))
for x in rows[:40]:
    print("push %3d wall %6.1f dev %6.1f events %3d clusters %3d points %5d cols %3d trees %4d" % x)
cc.close()
