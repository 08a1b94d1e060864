"""Command-line parity check (used by gpurun during development and by __graft_entry__.smoke):
python tests/run_parity.py [--emu] SPEC CHUNK ROTATIONS [KWARGS]"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import parity  # noqa: E402
import recorder  # noqa: E402
from continuous_clustering_b200 import _lib, synth  # noqa: E402
from continuous_clustering_b200.api import ContinuousClustering  # noqa: E402
from oracle import drvlib  # noqa: E402

IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]


def run(spec, chunk, rotations, kwargs=None, emu=False, oracle_lib=None, cfg_overrides=None, verbose=True,
        flag_period=0):
    kwargs = dict(kwargs or {})
    pts, poses, sp = synth.make_stream(spec, n_rotations=rotations, **kwargs)
    cfg = drvlib.stream_config(spec, **(cfg_overrides or {}))
    d = drvlib.Driver(oracle_lib or drvlib.ORACLE_LIB)
    d.configure(cfg, sp.rows)
    t0 = time.time()
    ref = parity.record(d, pts, poses)
    t_ref = time.time() - t0
    lib = _lib.load_library(os.path.join(HERE, "emu", "libcc_b200_emu_test.so")) if emu else None
    cc = ContinuousClustering(_library=lib, max_firings_per_push=max(4096, chunk))
    cc.setConfiguration(cfg)
    cc.reset(sp.rows)
    cc.setTransformRobotFrameFromSensorFrame(IDENTITY)
    cc.debug_flag_columns(flag_period)
    t0 = time.time()
    got = recorder.record(cc, pts, poses, chunk)
    t_got = time.time() - t0
    parity.compare(ref, got, name_a=d.name, name_b="emu" if emu else "cuda")
    tree_bad = int((ref["cluster_cells"]["tree_root_gcol"] != got["cluster_cells"]["tree_root_gcol"]).sum())
    if verbose:
        print(f"{spec} chunk={chunk} rot={rotations} {kwargs}: OK events={len(got['events'])} "
              f"clusters={len(got['clusters'])} exact_pushes={got['used_exact_path']} slow_insert_firings={got['slow_insert_firings']}/{pts.shape[0]} tree_root_mismatch={tree_bad} "
              f"oracle={t_ref:.2f}s impl={t_got:.2f}s launches={cc.total_launches}")
    cc.close()
    return ref, got


if __name__ == "__main__":
    a = sys.argv[1:]
    emu = False
    if a and a[0] == "--emu":
        emu = True
        a = a[1:]
    spec = a[0] if a else "tiny16"
    chunk = int(a[1]) if len(a) > 1 else 64
    rot = float(a[2]) if len(a) > 2 else 2.0
    kw = eval(a[3]) if len(a) > 3 else {}
    co = eval(a[4]) if len(a) > 4 else {}
    fp = int(a[5]) if len(a) > 5 else 0
    run(spec, chunk, rot, kw, emu=emu, cfg_overrides=co, flag_period=fp)
