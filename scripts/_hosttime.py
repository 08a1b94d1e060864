import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from continuous_clustering_b200 import ContinuousClustering
from continuous_clustering_b200.presets import stream_configuration
base_pts, base_poses, sp = bench.make_rotations()
B, R = 2048, sp.rows
K = 30
pts, poses = bench.tile_stream(base_pts, base_poses, sp, 0, (K+3)*B)
d_pts = torch.from_numpy(pts.view(np.uint8).reshape(-1, R*48)).cuda(); d_poses = torch.from_numpy(poses).cuda()
cc = ContinuousClustering(device=0, max_firings_per_push=B); cc.setConfiguration(stream_configuration("velodyne64")); cc.reset(R); cc.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
def sub(s): cc.submitFiringsDevice(d_pts.data_ptr()+s*B*R*48, d_poses.data_ptr()+s*B*96, B, R)
for s in range(3): sub(s); cc.wait()
torch.cuda.synchronize()
ts=[]; tw=[]
t0=time.perf_counter()
sub(3)
for s in range(K):
    a=time.perf_counter()
    if s+1<K: sub(3+s+1)
    b=time.perf_counter()
    r=cc.wait()
    c=time.perf_counter()
    ts.append(b-a); tw.append(c-b)
torch.cuda.synchronize()
tot=time.perf_counter()-t0
print("per step us", 1e6*tot/K, "submit us", 1e6*np.median(ts), "wait us", 1e6*np.median(tw), "device_ms", r.info.device_ms)
# sync mode
t0=time.perf_counter()
# experiment: does the GPU make progress on push k+1 while the host sleeps?
def busy_sleep(us):
    t=time.perf_counter()
    while (time.perf_counter()-t)*1e6 < us: pass
for delay in (0, 200, 400):
    cc2 = ContinuousClustering(device=0, max_firings_per_push=B); cc2.setConfiguration(stream_configuration("velodyne64")); cc2.reset(R); cc2.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
    def sub2(s): cc2.submitFiringsDevice(d_pts.data_ptr()+s*B*R*48, d_poses.data_ptr()+s*B*96, B, R)
    for s in range(3): sub2(s); cc2.wait()
    torch.cuda.synchronize()
    w1=[];w2=[]
    for it in range(5):
        s=3+2*it
        sub2(s); sub2(s+1)
        busy_sleep(delay)
        a=time.perf_counter(); cc2.wait(); b=time.perf_counter(); cc2.wait(); c=time.perf_counter()
        w1.append((b-a)*1e6); w2.append((c-b)*1e6)
    print("delay",delay,"wait1 us",np.median(w1),"wait2 us",np.median(w2))
import ctypes
L = cc._L
L.cc_debug_event_query.argtypes=[ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
cc3 = ContinuousClustering(device=0, max_firings_per_push=B); cc3.setConfiguration(stream_configuration("velodyne64")); cc3.reset(R); cc3.setTransformRobotFrameFromSensorFrame(bench.IDENTITY)
def sub3(s): cc3.submitFiringsDevice(d_pts.data_ptr()+s*B*R*48, d_poses.data_ptr()+s*B*96, B, R)
for s in range(4): sub3(s); cc3.wait()
torch.cuda.synchronize()
for it in range(3):
    s=4+2*it
    t0=time.perf_counter()
    sub3(s); tA=(time.perf_counter()-t0)*1e6
    sub3(s+1); tB=(time.perf_counter()-t0)*1e6
    names={(0,0):"A.ev0",(0,1):"A.ev1",(0,2):"A.ready",(0,3):"A.done",(1,0):"B.ev0",(1,1):"B.ev1",(1,2):"B.ready",(1,3):"B.done"}
    tl={}
    while len(tl)<8:
        t=(time.perf_counter()-t0)*1e6
        for (sl,w),nm in names.items():
            if nm not in tl and L.cc_debug_event_query(cc3._h, sl, w)==1: tl[nm]=round(t)
        if t>5000: break
    cc3.wait(); cc3.wait()
    print("submitA",round(tA),"submitB",round(tB),tl)
