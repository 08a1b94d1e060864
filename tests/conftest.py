import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

EMU_LIB = os.path.join(HERE, "emu", "libcc_b200_emu_test.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run on the GPU box with -m gpu")


def _make(*targets):
    subprocess.run(["make", *targets], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU restatement (always buildable with g++)."""
    from oracle import drvlib

    if not drvlib.have_oracle():
        _make("-C", "oracle", "oracle")
    return drvlib.ORACLE_LIB


@pytest.fixture(scope="session")
def ref_lib():
    """The reference's own sources, compiled where /root/reference exists; prebuilt .so otherwise; else skip."""
    from oracle import drvlib

    if not drvlib.have_ref() and os.path.exists("/root/reference/src/clustering/continuous_clustering.cpp"):
        _make("-C", "oracle", "ref")
    if not drvlib.have_ref():
        pytest.skip("oracle/_ref/libcc_ref.so not available (needs /root/reference to build)")
    return drvlib.REF_LIB


@pytest.fixture(scope="session")
def emu_library():
    """CPU emulation build of the kernels + host code (tests only, see tests/emu/cuda_emu.h)."""
    from continuous_clustering_b200 import _lib

    if not os.path.exists(EMU_LIB):
        _make("emu")
    return _lib.load_library(EMU_LIB)


@pytest.fixture(scope="session")
def cuda_library():
    """The product library on a real GPU. Fails (does not skip, does not fall back) when it cannot run."""
    import torch

    from continuous_clustering_b200 import _lib

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return _lib.load_library()
