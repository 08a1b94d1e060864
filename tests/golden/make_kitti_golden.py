"""Generates tests/golden/kitti_golden.json from the reference build of the KITTI replay front-end
(oracle/_ref/libcc_eval_ref.so = excerpts of kitti_loader.cpp / kitti_demo.cpp compiled unmodified; `make -C oracle ref_eval`,
only possible where /root/reference exists). Run from the repository root: python tests/golden/make_kitti_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import test_kitti as tk  # noqa: E402

out = {}
for name, _, _ in tk.CASES:
    case = tk.make_case(name)
    ref = tk.reference(*case)
    laser = ref["laser_index"]
    xyzi = case[0]
    # rows found the way recoverLaserIndices counts them: azimuth wraps (capped at 64) + 1
    import numpy as np

    az = np.arctan2(xyzi[:, 1], xyzi[:, 0]).astype(np.float32).astype(np.float64)
    mono = np.where(az < 0, az + 2 * np.pi, az)
    wraps = int((np.diff(mono) < -0.7).sum())
    out[name] = {"sha256": tk.digest(ref), "n_points": int(xyzi.shape[0]), "rows_found": min(wraps, 64) + 1,
                 "cells_filled": int((ref["cell_point"] >= 0).sum())}
    print(name, out[name])
json.dump(out, open(os.path.join(HERE, "kitti_golden.json"), "w"), indent=1)
