"""The C-ABI library: builds for sm_100a, loads, exports every symbol include/cc_b200.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from continuous_clustering_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "lib"], cwd=REPO, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _lib.load_library()


def declared_symbols():
    text = open(os.path.join(REPO, "include", "cc_b200.h")).read()
    return sorted(set(re.findall(r"CC_API\s+[\w\s\*]+?\b(cc_\w+)\s*\(", text)))


def test_header_symbols_are_exported(product_lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(product_lib, n), n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_library_is_sm100a_cuda_code():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out, out


def test_struct_sizes_match_header(product_lib, tmp_path):
    """sizeof of every struct of include/cc_b200.h as the C compiler lays it out == the ctypes / numpy mirrors."""
    import subprocess

    names = ["cc_config_t", "cc_batch_info_t", "cc_column_event_t", "cc_cluster_t", "cc_cluster_point_t", "cc_raw_point_t",
             "cc_cell_t", "cc_cloud_view_t", "cc_pack_request_t", "cc_column_fields_t"]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "cc_b200.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)], check=True)
    sizes = dict(zip(names, map(int, subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout.split())))
    assert C.sizeof(_lib.CcConfig) == sizes["cc_config_t"] == 33 * 4
    assert C.sizeof(_lib.CcBatchInfo) == sizes["cc_batch_info_t"]
    assert _lib.EVENT_DTYPE.itemsize == sizes["cc_column_event_t"] == 24
    assert _lib.CLUSTER_DTYPE.itemsize == sizes["cc_cluster_t"] == 64
    assert _lib.CLUSTER_POINT_DTYPE.itemsize == sizes["cc_cluster_point_t"] == 16
    assert _lib.CELL_DTYPE.itemsize == sizes["cc_cell_t"] == 128
    assert C.sizeof(_lib.CcCloudView) == sizes["cc_cloud_view_t"]
    assert C.sizeof(_lib.CcPackRequest) == sizes["cc_pack_request_t"]
    assert C.sizeof(_lib.CcColumnFields) == sizes["cc_column_fields_t"]
    assert sizes["cc_raw_point_t"] == 48
    cfg = _lib.CcConfig()
    product_lib.cc_config_default(C.byref(cfg))
    assert cfg.num_columns == 1700 and abs(cfg.max_distance - 0.7) < 1e-7 and cfg.max_steps_in_row == 20


def test_no_cpu_fallback(product_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert product_lib.cc_create(0, 0, C.byref(h)) == 2  # CC_ERR_CUDA: the path only exists on the device
    from continuous_clustering_b200 import ClusteringError, ContinuousClustering

    with pytest.raises(ClusteringError):
        ContinuousClustering()


def test_package_never_touches_the_oracle():
    """The product package neither imports nor includes nor dlopens anything under oracle/ (or the emulation build)."""
    pkg = os.path.join(REPO, "continuous_clustering_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(root, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
            assert not re.search(r"#include\s+[\"<][^\">]*oracle", text), f
            assert "libcc_oracle" not in text and "libcc_ref" not in text and "emu_test.so" not in text, f
