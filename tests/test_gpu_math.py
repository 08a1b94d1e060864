import pytest

from test_math import check_all


@pytest.mark.gpu
def test_device_math_matches_host_libm(cuda_library):
    """atan2f/asinf/atanf computed by the sm_100a build are bit-identical to the host libm."""
    check_all(cuda_library, 300_000)
