"""cc_math.cuh (the glibc-exact float transcendentals the kernels use) against the host libm, host build.
The exhaustive 2^32 sweep of asinf/atanf was run once while writing the functions; here a
few million structured + random inputs per function keep it pinned. tests/test_gpu_math.py repeats it on the
device build."""
import ctypes as C
import ctypes.util

import numpy as np

libm = C.CDLL(ctypes.util.find_library("m"))
for f in ("atan2f", "asinf", "atanf"):
    getattr(libm, f).restype = C.c_float
libm.atan2f.argtypes = [C.c_float, C.c_float]
libm.asinf.argtypes = [C.c_float]
libm.atanf.argtypes = [C.c_float]


def host(op, a, b=None):
    if op == 0:
        return np.array([libm.atan2f(float(x), float(y)) for x, y in zip(a, b)], dtype=np.float32)
    fn = libm.asinf if op == 1 else libm.atanf
    return np.array([fn(float(x)) for x in a], dtype=np.float32)


def samples(n, seed):
    rng = np.random.RandomState(seed)
    bits = rng.randint(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    lidar = (rng.uniform(-150, 150, size=n)).astype(np.float32)
    unit = rng.uniform(-1.001, 1.001, size=n).astype(np.float32)
    edge = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 0.975, 2.0**-27, 2.0**-29, 0.4375, 0.6875, 1.1875, 2.4375,
                     2.0**25, np.inf, -np.inf, np.nan, 1e-38, 1e-45, 3.4e38, 0.7, 120.0], dtype=np.float32)
    return bits, lidar, unit, edge


def same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


def run_selftest(lib, op, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b if b is not None else a, np.float32)
    out = np.zeros_like(a)
    rc = lib.cc_selftest_math(0, op, a.size, a.ctypes.data, b.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def check_all(lib, n):
    bits, lidar, unit, edge = samples(n, 99)
    for a, b in ((lidar, np.roll(lidar, 1)), (bits, np.roll(bits, 7)), (np.full(n, 0.7, np.float32), np.abs(lidar)),
                 (np.repeat(edge, edge.size), np.tile(edge, edge.size))):
        assert same(run_selftest(lib, 0, a, b), host(0, a, b))
    for a in (unit, bits, edge, lidar / 150.0):
        assert same(run_selftest(lib, 1, a), host(1, a))
        assert same(run_selftest(lib, 2, a * 50), host(2, a * 50))


def test_host_build_matches_libm(emu_library):
    check_all(emu_library, 200_000)
