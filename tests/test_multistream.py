"""N > 1: streams shard one per rank with no data-path collective. world_size-2 gloo run on CPU (emulation build):
every rank processes its own streams, digests are all-gathered, and must equal the single-process results."""
import os
import subprocess
import sys

import pytest

from continuous_clustering_b200 import multi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["CC_REPO"]); sys.path.insert(0, os.path.join(os.environ["CC_REPO"], "tests"))
import torch.distributed as dist
from continuous_clustering_b200 import _lib, multi, synth, ContinuousClustering
from oracle import drvlib
import parity, recorder
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = _lib.load_library(os.path.join(os.environ["CC_REPO"], "tests", "emu", "libcc_b200_emu_test.so"))
local = {}
for sid in multi.streams_of_rank(3, world, rank):
    pts, poses, sp = synth.make_stream("tiny16", n_rotations=1.5, seed=100 + sid, moving=bool(sid % 2))
    cc = ContinuousClustering(_library=lib)
    cc.setConfiguration(drvlib.stream_config("tiny16")); cc.reset(sp.rows)
    cc.setTransformRobotFrameFromSensorFrame([1,0,0,0,0,1,0,0,0,0,1,0])
    rec = recorder.record(cc, pts, poses, 128)
    local[sid] = multi.result_digest(rec["events"], [k for _, _, k in parity.cluster_multiset(rec)])
allr = multi.gather_digests(local)
if rank == 0:
    print("DIGESTS", sorted(allr.items()))
dist.barrier(); dist.destroy_process_group()
'''


def test_stream_assignment():
    assert multi.streams_of_rank(8, 8, 3) == [3]
    assert multi.streams_of_rank(3, 2, 0) == [0, 2] and multi.streams_of_rank(3, 2, 1) == [1]
    assert sorted(sum((multi.streams_of_rank(5, 4, r) for r in range(4)), [])) == list(range(5))
    with pytest.raises(ValueError):
        multi.streams_of_rank(2, 2, 2)


def run_world(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CC_REPO=REPO)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29400 + world + os.getpid() % 500), str(script)]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("DIGESTS")]
    assert line, out.stdout[-3000:]
    return line[0]


def test_two_ranks_equal_one_rank(tmp_path, emu_library, oracle_lib):
    assert run_world(2, tmp_path) == run_world(1, tmp_path)
