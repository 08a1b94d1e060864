"""Diagnostic (GPU box): how many firings of a push leave the regular (lite) insertion path, per stream geometry."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from continuous_clustering_b200 import ContinuousClustering, synth
from continuous_clustering_b200.presets import stream_configuration
IDENT = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
for spec, kw in (("vls128", {}), ("vls128", {"start_firing": 300}), ("os32_left", {}), ("velodyne64", {}), ("kitti64", {})):
    pts, poses, sp = synth.make_stream(spec, n_rotations=4.0, **kw)
    for B in (1024, 4096):
        cc = ContinuousClustering(max_firings_per_push=B)
        cc.setConfiguration(stream_configuration(spec)); cc.reset(sp.rows); cc.setTransformRobotFrameFromSensorFrame(IDENT)
        out = []
        for a in range(0, pts.shape[0] - B + 1, B):
            r = cc.addFirings(pts[a:a + B], poses[a:a + B])
            out.append((int(r.info.slow_insert_firings), round(r.info.device_ms, 3)))
        print(spec, kw, "B", B, "slow firings / device ms per push:", out, flush=True)
        cc.close()
