"""sensor packets -> firings on the device (csrc/cc_packets.cuh, cc_ouster_* C ABI) against the CPU
restatement of OusterInput::onRawDataArrived (oracle/cc_packets_oracle.cpp; ouster_input.hpp:105-181). PARITY UNPINNED: the
ouster SDK the reference calls is not in the image, the restatement follows its published packet layout. Every field of
every RawPoint is compared bit for bit, plus the firing stamps and the firing-index bookkeeping across calls and resets."""
import ctypes as C
import os

import numpy as np
import pytest

from continuous_clustering_b200 import OusterInput
from continuous_clustering_b200.synth import RAW_POINT_DTYPE, make_ouster_packets, ouster_xyz_lut

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ["x", "y", "z", "firing_index", "intensity", "stamp", "globally_unique_point_index"]
H, W = 32, 1024


def oracle_decode(fmt, direction, offset, packets, stamps, after_reset, first_index):
    lib = C.CDLL(os.path.join(REPO, "oracle", "libcc_oracle.so"))
    vp = C.c_void_p
    lib.orc_ouster_decode.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_uint64, vp, vp]
    n = packets.shape[0]
    firings = np.zeros((n * 16, H), dtype=RAW_POINT_DTYPE)
    fst = np.zeros(n * 16, np.uint64)
    k = lib.orc_ouster_decode(C.addressof(fmt), direction.ctypes.data, offset.ctypes.data, n, packets.ctypes.data, packets.shape[1],
                              stamps.ctypes.data, int(after_reset), int(first_index), firings.ctypes.data, fst.ctypes.data)
    return firings[:k], fst[:k]


def same(a, b, what):
    assert a.shape == b.shape, f"{what}: {a.shape} vs {b.shape}"
    for f in FIELDS:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            x, y = x.view(np.uint32), y.view(np.uint32)
        assert np.array_equal(x, y), f"{what}: field {f}"


def run(library, side, n_packets, seed, chunks):
    direction, offset = ouster_xyz_lut(side)
    packets, stamps = make_ouster_packets(n_packets, rows=H, columns_per_frame=W, seed=seed, first_measurement_id=1000)
    dec = OusterInput(H, W, direction, offset, max_packets_per_call=max(chunks), _library=library)
    assert dec.packet_size == packets.shape[1] == 16 * (16 + 12 * H + 4)
    a, first, after_reset = 0, 0, True
    total = 0
    for c in chunks:
        pk, st = packets[a:a + c], stamps[a:a + c]
        got = dec.decode(pk, st)
        ref_f, ref_s = oracle_decode(dec.format, direction, offset, pk, st, after_reset, first)
        assert got["n_firings"] == ref_f.shape[0] and got["first_firing_index"] == first
        same(ref_f, dec.read_firings(got["n_firings"]), f"chunk at packet {a}")
        assert np.array_equal(ref_s, got["firing_stamps"])
        first += got["n_firings"]
        total += got["n_firings"]
        a += c
        after_reset = False
    # reset in the middle of the stream: the packet in flight is dropped, the firing index restarts
    dec.reset()
    got = dec.decode(packets[:3], stamps[:3])
    ref_f, ref_s = oracle_decode(dec.format, direction, offset, packets[:3], stamps[:3], True, 0)
    assert got["first_firing_index"] == 0 and got["n_firings"] == ref_f.shape[0] <= 32
    same(ref_f, dec.read_firings(got["n_firings"]), "after reset")
    dec.close()
    return total


def test_oracle_known_answers():
    """The restatement on a hand-made packet: one valid block with known range / signal words."""
    direction, offset = ouster_xyz_lut("left")
    packets, stamps = make_ouster_packets(2, rows=H, columns_per_frame=W, seed=1, first_measurement_id=5, p_invalid=0.0)
    col = packets[1].reshape(16, -1)[0]
    col[16:20] = np.array([2000 | (0xA << 28)], "<u4").view(np.uint8)  # ring 0: range 2000 mm, flag bits set
    col[16 + 6:16 + 8] = np.array([500], "<u2").view(np.uint8)  # signal 500 -> intensity 127
    col[16 + 12:16 + 16] = 0  # ring 1: no return
    from continuous_clustering_b200 import _lib

    fmt = _lib.CcOusterFormat(16, H, W, 0, 16, 4, 12, 8, 16 + 12 * H, 4, 0, 4, 0x000FFFFF, 0, 6, 2, 0, 0, 1)
    f, s = oracle_decode(fmt, direction, offset, packets, stamps, True, 0)
    assert f.shape[0] == 16 and s[0] == stamps[1]  # first packet dropped after the reset
    m_id = 5 + 16
    d = direction[m_id * H]
    assert f[0, 0]["x"] == np.float32(np.float32(2000.0) * d[0]) + d[0] and f[0, 0]["intensity"] == 127
    assert np.isnan(f[0, 1]["x"]) and f[0, 1]["intensity"] == 0 and f[0, 1]["firing_index"] == 0 and f[3, 0]["firing_index"] == 3


@pytest.mark.parametrize("side,n_packets,seed,chunks", [("left", 40, 1, [1, 7, 32]), ("right", 96, 2, [64, 32])])
def test_emulated_decoder_matches_oracle(emu_library, side, n_packets, seed, chunks):
    assert run(emu_library, side, n_packets, seed, chunks) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("side,n_packets,seed,chunks", [("left", 40, 1, [1, 7, 32]), ("right", 640, 2, [256, 256, 128])])
def test_cuda_decoder_matches_oracle(cuda_library, side, n_packets, seed, chunks):
    assert run(None, side, n_packets, seed, chunks) > 0


def test_no_packets_and_invalid_blocks(emu_library):
    """Degenerate inputs: an empty call, a call with only the packet that is dropped after a reset, packets whose
    measurement blocks are all invalid (status 0): no firings, firing index unchanged."""
    direction, offset = ouster_xyz_lut("left")
    packets, stamps = make_ouster_packets(6, rows=H, columns_per_frame=W, seed=3, p_invalid=1.0)
    dec = OusterInput(H, W, direction, offset, max_packets_per_call=8, _library=emu_library)
    assert dec.decode(packets[:0], stamps[:0])["n_firings"] == 0
    assert dec.decode(packets[:1], stamps[:1])["n_firings"] == 0  # dropped: first packet after the reset
    got = dec.decode(packets[1:], stamps[1:])
    ref_f, _ = oracle_decode(dec.format, direction, offset, packets[1:], stamps[1:], False, 0)
    assert got["n_firings"] == ref_f.shape[0] == 0 and got["first_firing_index"] == 0
    good, gstamps = make_ouster_packets(2, rows=H, columns_per_frame=W, seed=4, p_invalid=0.0)
    got = dec.decode(good, gstamps)
    assert got["n_firings"] == 32 and got["first_firing_index"] == 0
    assert dec.decode(good, gstamps)["first_firing_index"] == 32
    dec.close()


# ---- packets -> firings -> hot path, firings never leaving the device --------------------------------------------------
def scene_packets(n_packets, seed):
    """LEGACY packets whose ranges come from the synthetic OS-32 scene (the ranges of stream firing k go into measurement block
    k): decoded through the sensor's lookup table they are a coherent street scene again."""
    from continuous_clustering_b200 import synth

    pts, fposes, sp = synth.make_stream("os32_left", n_firings=n_packets * 16, seed=seed, start_firing=64)
    scene_packets.pose = fposes[0].copy()  # static sensor: odom_from_sensor = the mount roll, the same for every firing
    dist = np.sqrt(pts["x"].astype(np.float64) ** 2 + pts["y"].astype(np.float64) ** 2 + pts["z"].astype(np.float64) ** 2)
    mm = np.where(np.isnan(dist), 0, np.round(dist * 1000.0)).astype(np.uint32)
    # (encoder angle 2 pi (1 - col / W) and the lidar-to-sensor transform: the azimuth decreases with the measurement id, like
    # the stream's: a clockwise sensor)
    return make_ouster_packets(n_packets, rows=H, columns_per_frame=W, seed=seed, first_measurement_id=64, p_invalid=0.02, ranges_mm=mm)


def chain(library, oracle_lib, n_packets=160, per_call=32):
    import parity
    import recorder
    from continuous_clustering_b200 import ContinuousClustering
    from oracle import drvlib

    direction, offset = ouster_xyz_lut("left")
    packets, stamps = scene_packets(n_packets, 9)
    cfg = drvlib.stream_config("os32_left")
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float64)
    mount = scene_packets.pose
    dec = OusterInput(H, W, direction, offset, max_packets_per_call=per_call, _library=library)
    # oracle chain: restated decoder -> host RawPoints -> restated clustering
    ref_f, _ = oracle_decode(dec.format, direction, offset, packets, stamps, True, 0)
    d = drvlib.Driver(oracle_lib)
    d.configure(cfg, H)
    want = parity.record(d, np.ascontiguousarray(ref_f), np.tile(mount, (ref_f.shape[0], 1)))
    assert not want["reset_required"] and len(want["clusters"]) > 5 and int((want["ground_cells"]["ground_point_label"] == 54).sum()) > 5000
    # product chain: device decoder -> firings resident in device memory -> hot path (cc_push_firings_device)
    cc = ContinuousClustering(max_firings_per_push=per_call * 16, _library=library)
    cc.setConfiguration(cfg)
    cc.reset(H)
    cc.setTransformRobotFrameFromSensorFrame(ident)
    poses = np.tile(mount, (per_call * 16, 1))
    if library is None:  # the CUDA library: the poses have to be in device memory too
        import torch

        keep = torch.from_numpy(poses).cuda()
        d_poses = keep.data_ptr()
    else:  # the emulation's "device" memory is host memory
        d_poses = poses.ctypes.data

    class DecodeAndPush:
        """What tests/recorder.py drives instead of the clustering object: every addFirings call decodes the next packets
        and pushes the firings the decoder left on the device."""

        def __init__(self):
            self.call, self.firings = 0, 0

        def __getattr__(self, name):
            return getattr(cc, name)

        def addFirings(self, _points, _poses):
            a = self.call * per_call
            self.call += 1
            out = dec.decode(packets[a:a + per_call], stamps[a:a + per_call])
            self.firings += out["n_firings"]
            return cc.addFiringsDevice(out["d_firings"], d_poses, out["n_firings"], H)

    feeder = DecodeAndPush()
    n_calls = (n_packets + per_call - 1) // per_call
    got = recorder.record(feeder, np.zeros((n_calls, H), dtype=RAW_POINT_DTYPE), np.zeros((n_calls, 12)), 1)
    assert feeder.firings == ref_f.shape[0]
    parity.compare(want, got, name_a="oracle chain", name_b="product chain")
    dec.close()
    cc.close()
    return feeder.firings


def test_packets_feed_the_hot_path_emulation(emu_library, oracle_lib):
    assert chain(emu_library, oracle_lib) > 2000


@pytest.mark.gpu
def test_packets_feed_the_hot_path_cuda(cuda_library, oracle_lib):
    assert chain(None, oracle_lib) > 2000
