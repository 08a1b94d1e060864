"""The drop-in C++ class (facade/): the SAME recording driver source that drives the reference build
(oracle/cc_driver.cpp) is compiled against the facade; what it records must equal the oracle's recording.
CPU variant links the emulation library; the -m gpu variant links the CUDA library."""
import os
import subprocess

import numpy as np
import pytest

import parity
from continuous_clustering_b200 import synth
from oracle import drvlib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DRIVER = os.path.join(REPO, "build", "libcc_facade_driver_emu_test.so")


def run(lib, spec, kw, cfg_over, batch, pipelined=False):
    os.environ["CC_B200_BATCH"] = str(batch)
    os.environ["CC_B200_PIPELINE"] = "1" if pipelined else "0"
    pts, poses, sp = synth.make_stream(spec, **kw)
    cfg = drvlib.stream_config(spec, **cfg_over)
    d = drvlib.Driver(lib)
    d.configure(cfg, sp.rows)
    rec = parity.record(d, pts, poses, chunk=500)
    d.close()
    o = drvlib.Driver(drvlib.ORACLE_LIB)
    o.configure(cfg, sp.rows)
    want = parity.record(o, pts, poses)
    parity.compare(want, rec, name_a="oracle", name_b=d.name, check_tree_fields=True, check_published_tree_fields=True)
    assert np.array_equal(want["cluster_cells"]["tree_root_gcol"], rec["cluster_cells"]["tree_root_gcol"])


CASES = [("tiny16", dict(n_rotations=2.0, moving=True), {}, 64), ("tiny16", dict(n_rotations=1.5), {}, 1),
         ("tiny16", dict(n_rotations=2.0, dropout=0.1), dict(cluster_point_trees_every_nth_column=2), 100),
         ("velodyne64", dict(n_rotations=1.1), {}, 256)]


@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES)
def test_facade_on_emulation_library(emu_library, oracle_lib, spec, kw, cfg_over, batch):
    subprocess.run(["make", "-C", "facade", "emu"], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    run(EMU_DRIVER, spec, kw, cfg_over, batch)


@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES)
def test_pipelined_facade_on_emulation_library(emu_library, oracle_lib, spec, kw, cfg_over, batch):
    """Throughput mode of the facade (CC_B200_PIPELINE=1): batches are submitted asynchronously, callbacks arrive up to two
    batches late; what the driver records must not change."""
    subprocess.run(["make", "-C", "facade", "emu"], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    run(EMU_DRIVER, spec, kw, cfg_over, batch, pipelined=True)


def test_facade_throws_like_reference(emu_library, oracle_lib):
    subprocess.run(["make", "-C", "facade", "emu"], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    os.environ["CC_B200_BATCH"] = "32"
    os.environ["CC_B200_PIPELINE"] = "0"
    pts, poses, sp = synth.make_stream("tiny16", n_firings=200)
    cfg = drvlib.stream_config("tiny16")
    d = drvlib.Driver(EMU_DRIVER)
    d.configure(cfg, sp.rows, identity_robot_tf=False)
    assert d.add_firings(pts, poses, raise_on_error=False) == 1
    assert "Transform robot frame from sensor frame was not set yet" in d.error()
    d2 = drvlib.Driver(EMU_DRIVER)
    d2.configure(cfg, sp.rows)
    assert d2.add_firings(pts[:, :8], poses, raise_on_error=False) == 1
    assert "number of points in a firing has changed" in d2.error()


@pytest.mark.gpu
@pytest.mark.parametrize("spec,kw,cfg_over,batch", CASES + [("velodyne64", dict(n_rotations=2.2, moving=True), {}, 2048)])
def test_facade_on_cuda_library(cuda_library, oracle_lib, spec, kw, cfg_over, batch):
    if not os.path.exists(drvlib.FACADE_LIB):
        subprocess.run(["make", "-C", "facade"], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    run(drvlib.FACADE_LIB, spec, kw, cfg_over, batch)


@pytest.mark.gpu
@pytest.mark.parametrize("spec,kw,cfg_over,batch", [CASES[0], ("velodyne64", dict(n_rotations=3.2, moving=True), {}, 1024)])
def test_pipelined_facade_on_cuda_library(cuda_library, oracle_lib, spec, kw, cfg_over, batch):
    if not os.path.exists(drvlib.FACADE_LIB):
        subprocess.run(["make", "-C", "facade"], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    run(drvlib.FACADE_LIB, spec, kw, cfg_over, batch, pipelined=True)
