#!/usr/bin/env python
"""bench.py -- range-image columns/s of the per-column hot path on a synthetic LiDAR firing stream.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--spec velodyne64|vls128|os32_pair|kitti64]
                  [--moving] [--batch B] [--quick]

One STEP = one push of `--batch` consecutive firings (default 4096 = two sensor rotations of the 64 x 2048 synthetic
Velodyne-like stream, BASELINE.json configs[1]) through insertion, ground segmentation, association, finish detection
and ring recycling. The stream keeps going across steps (the range image is continuous).

  value     columns/s with the firings already resident in HBM when the timed region starts (cc_submit_firings_device /
            cc_wait, two pushes in flight), timed with CUDA events on the handle's stream over K back-to-back pushes.
            Inputs larger than L2: every push reads firings nothing has touched since L2 was flushed (256 MiB write)
            right before the timed region. `l2_flush_each_step` reports the same with a flush before every push.
  e2e       the same metric through the C ABI with page-locked HOST buffers in the reference's 48-byte RawPoint layout
            (cc_submit_firings / cc_wait: host -> device copy of the raw firings, kernels, device -> host copy of events,
            finished clusters, member lists and the ground labels of the new columns), wall clock.
  latency_mode   64-firing synchronous pushes with host buffers through the C ABI (one fused launch per push) and the
            per-call wall time of ContinuousClustering::addFiring on the drop-in C++ class (facade_latency), next to
            cpu_baseline.latency_us_*: the reference's single-threaded per-addFiring times on this box.
  facade    columns/s through the drop-in C++ class, one addFiring call per firing, callbacks registered.
  roofline  the dominant kernel of the step (largest share of device time, measured live with CUDA events around
            every launch, no exclusions): SURVEY 8d algorithmic bytes / its duration, against MEASURED_PEAKS.json.
  cpu_baseline  the reference's own CPU implementation (oracle/_ref/libcc_ref.so, built from the reference's
            sources) -- or the restatement in oracle/ when that build is absent -- timed on this box's host cores
            on a bounded sample of the same stream.

N > 1 (torchrun, one rank per GPU): every rank runs its own independent sensor stream (streams shard one per GPU;
there is no data-path collective), barrier + max-over-ranks timing, value = total columns / max time ("weak"); every
rank contributes its per-GPU push-latency histogram. `--impl reference` times only the CPU reference arm (rank 0
only; at N > 1 it runs N concurrent reference streams so that the ratio stays like for like).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

SPEC = "velodyne64"  # default workload (scripts/ import this)
IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
METRIC = "range-image columns/s, 64-ring stream"

SPEC_TEXT = {
    "velodyne64": "synthetic 64-ring 10 Hz Velodyne-like stream (64x2048 columns/rotation, ground plane + 150 boxes",
    "kitti64": "synthetic 64x2200 KITTI-replay-like stream (kitti_demo configuration, ground plane + 150 boxes",
    "vls128": "synthetic 128-ring 20 Hz VLS-128-like stream (128x1700 columns/rotation, 8 laser groups with azimuth offsets, "
              "ground plane + 150 boxes",
    "os32_left": "synthetic Ouster OS-32 stream, left sensor (32x1024, beam tables of the reference's calibration, rolled +30 deg",
    "os32_right": "synthetic Ouster OS-32 stream, right sensor (32x1024, beam tables of the reference's calibration, rolled -30 deg",
    "os32_pair": "synthetic Ouster OS-32 tilted-mount pair (32x1024 each, left sensor on even ranks / right on odd ranks",
}

# ALGORITHMIC bytes per range-image cell, attributed to the kernel that moves them (151 B/cell for the
# path: insert 37 r + 57 w, ground 21 r + 3 w, associate 21 r + 4 w, finish/label 4 r + 4 w). Staging arrays, the
# association view, list scratch and ring recycling are implementation traffic, reported separately.
KERNEL_BYTES_PER_CELL = {
    "k_prep": 12,            # x, y, z of the raw record
    "k_scatter": 25 + 57,    # the rest of the raw record + the 57 B of range-image fields
    "k_ground": 21 + 3,
    "k_probe": 21 + 4,       # every non-ignored point is probed once, by k_probe or by k_probe_heavy
    "k_probe_heavy": 0,
    "k_fin_label": 4 + 4,
    "k_push_fused": 151,     # the whole path in one launch (short pushes)
}
# implementation traffic per cell on top of that (what ncu's dram / L2 byte counters additionally see)
KERNEL_EXTRA_BYTES_PER_CELL = {
    "k_prep": 28 + 102,      # staged odom xyz, distance, azimuth, inclination, column (twice) + recycling of a retired cell
    "k_scan_check": 4, "k_insert_scan": 8, "k_gap_scan": 8,
    "k_ground": 16 + 4 + 8,  # association view (float4), mad, inclination-gap carry
    "k_commit_copy": 8, "k_commit_roots": 8, "k_commit_links": 16,
}


def load_peaks():
    p = os.path.join(HERE, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def rank_spec(spec_name: str, rank: int) -> str:
    if spec_name == "os32_pair":
        return "os32_left" if rank % 2 == 0 else "os32_right"
    return spec_name


def make_rotations(n_rot_unique=2, seed=1234, spec_name=None):
    """A few unique rotations of the static-sensor stream; longer streams tile them (a static scene repeats every
    rotation anyway) with fresh stamps / firing indices / unique point indices."""
    from continuous_clustering_b200 import synth

    # Sensors whose lasers fire at different azimuths (VLS-128: 8 groups over +-6.4 degrees, OS-32: 8 - 11 degrees) must not
    # START on the negative x axis: the reference drops a first firing that straddles it and asks for a reset (cpp:252-261,
    # SURVEY 8d "start just after the -x axis so the first firing does not straddle it"); 64 firings later no row does.
    name = spec_name or SPEC
    start = 64 if float(np.abs(synth.spec(name).azimuth_offsets_rad).max()) > 0 else 0
    pts, poses, sp = synth.make_stream(name, n_rotations=n_rot_unique, seed=seed, start_firing=start)
    return pts, poses, sp


def tile_stream(base_pts, base_poses, sp, start, n):
    """Firings [start, start + n) of the endless stream made by repeating the base rotations."""
    nb = base_pts.shape[0]
    idx = (np.arange(start, start + n)) % nb
    pts = base_pts[idx].copy()
    k = np.arange(start, start + n, dtype=np.uint64)
    t_rot_ns = 1e9 / sp.rotation_hz
    pts["stamp"] = (1_000_000_000 + k * (t_rot_ns / sp.num_columns)).astype(np.uint64)[:, None]
    pts["firing_index"] = k[:, None]
    pts["globally_unique_point_index"] = k[:, None] * np.uint64(sp.rows) + np.arange(sp.rows, dtype=np.uint64)[None, :]
    return pts, base_poses[idx].copy()


class Stream:
    """The endless firing stream of one rank: tiled static rotations, or (moving sensor, BASELINE.md section 2: 10 m/s,
    0.2 rad/s yaw) generated once for as many firings as the run needs -- poses never repeat."""

    def __init__(self, spec_name, seed, moving=False, total_firings=0):
        from continuous_clustering_b200 import synth

        self.moving = moving
        if moving:
            start = 64 if float(np.abs(synth.spec(spec_name).azimuth_offsets_rad).max()) > 0 else 0
            self.pts, self.poses, self.sp = synth.make_stream(spec_name, n_firings=total_firings, seed=seed, moving=True, start_firing=start)
        else:
            self.pts, self.poses, self.sp = make_rotations(seed=seed, spec_name=spec_name)

    def take(self, start, n):
        if self.moving:
            if start + n > self.pts.shape[0]:
                raise RuntimeError("moving stream exhausted: raise total_firings")
            return self.pts[start:start + n].copy(), self.poses[start:start + n].copy()
        return tile_stream(self.pts, self.poses, self.sp, start, n)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(args):
    motion = "moving sensor: 10 m/s, 0.2 rad/s yaw" if args.moving else "static sensor"
    return f"{SPEC_TEXT[args.spec]}, {motion}), {args.batch} firings per push"


def config_dict(args):
    """The same dict in both arms (the driver compares them)."""
    from continuous_clustering_b200 import synth

    sp = synth.spec(rank_spec(args.spec, 0))
    return {"workload": workload_name(args), "spec": args.spec, "moving": bool(args.moving), "batch_firings": args.batch,
            "rows": sp.rows, "columns_per_rotation": sp.num_columns, "rotation_hz": sp.rotation_hz}


def metric_name(args):
    return METRIC if args.spec in ("velodyne64", "kitti64") else f"range-image columns/s, {args.spec} stream"


# ---------------------------------------------------------------------------------------------------- CPU reference
def cpu_reference_run(stream, spec_name, columns_target, multi_threaded=True, time_budget_s=25.0, n_streams=1):
    """Times the reference's CPU implementation on this box's host cores (n_streams concurrent, independent sensor
    streams: one object each, like the reference's one-process-per-sensor deployment). Returns dict for `cpu_baseline`."""
    from oracle import drvlib

    kind = "reference" if drvlib.have_ref() else "port"
    lib = drvlib.REF_LIB if kind == "reference" else drvlib.ORACLE_LIB
    if kind == "port":
        multi_threaded = False  # the restatement is the deterministic single-threaded mode only
    sp = stream.sp
    cfg = drvlib.stream_config(spec_name, is_single_threaded=0 if multi_threaded else 1)
    rot = sp.num_columns
    drivers = []
    for _ in range(n_streams):
        d = drvlib.Driver(lib)
        d.configure(cfg, sp.rows)
        d.set_record(0)
        drivers.append(d)
    fed = 0
    for w in range(2):  # warm-up: two rotations
        pts, poses = stream.take(fed, rot)
        for d in drivers:
            d.prepare(pts, poses)
            d.run_prepared(0, rot, 3 * rot if multi_threaded else 0)
        fed += rot
    timed = 0
    t_total = 0.0
    t_wall0 = time.time()
    chunk = 16 * rot  # long runs, so that draining the reference's thread pipeline between runs does not matter
    while timed < columns_target and (time.time() - t_wall0) < time_budget_s:
        pts, poses = stream.take(fed, chunk)
        for d in drivers:
            d.prepare(pts, poses)  # shared_ptr construction outside the timed region
        secs = [0.0] * n_streams

        def work(i):
            secs[i] = drivers[i].run_prepared(0, chunk, 3 * rot if multi_threaded else 0)

        if n_streams == 1:
            work(0)
            t_total += secs[0]
        else:
            th = [threading.Thread(target=work, args=(i,)) for i in range(n_streams)]
            t0 = time.perf_counter()
            [t.start() for t in th]
            [t.join() for t in th]
            t_total += time.perf_counter() - t0
        fed += chunk
        timed += chunk
    for d in drivers:
        d.close()
    cores = (8 if multi_threaded else 1) * n_streams  # per stream: 4 stage threads + 3 publishers + producer (cpp:49-63)
    return {
        "value": n_streams * timed / t_total if t_total > 0 else 0.0,
        "unit": "columns/s",
        "cores": cores,
        "kind": kind,
        "sample": f"{n_streams} stream(s) x {timed} columns ({timed // rot} rotations of the {sp.rows}x{rot} synthetic stream) "
                  f"after 2 warm-up rotations, {'multi-threaded 5-stage pipeline' if multi_threaded else 'single-threaded mode'}, "
                  f"no-op callbacks, host has {os.cpu_count()} logical cores",
        "seconds": t_total,
    }


def cpu_reference_latency(stream, spec_name, n_firings=None):
    """per-addFiring wall time of the reference in its single-threaded mode (every stage a
    firing triggers, callbacks included, runs inside the call)."""
    from oracle import drvlib

    if not drvlib.have_ref():
        return None
    sp = stream.sp
    n = n_firings or 3 * sp.num_columns
    cfg = drvlib.stream_config(spec_name, is_single_threaded=1)
    d = drvlib.Driver(drvlib.REF_LIB)
    d.configure(cfg, sp.rows)
    d.set_record(0)
    pts, poses = stream.take(0, n)
    d.prepare(pts, poses)
    lat = d.run_prepared_latency(0, n)[sp.num_columns:]  # the first rotation fills the ring
    d.close()
    return {"latency_us_p50": float(np.median(lat)), "latency_us_p99": float(np.percentile(lat, 99)),
            "latency_us_mean": float(lat.mean()),
            "latency_sample": f"{len(lat)} addFiring calls, single-threaded mode (1 core), no-op callbacks"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    spec_name = rank_spec(args.spec, 0)
    stream = Stream(spec_name, 1234)
    steps, warm = args.steps, args.warmup
    n_streams = max(1, world)
    target = max(1, steps) * args.batch * 8
    try:
        res = cpu_reference_run(stream, spec_name, columns_target=target, multi_threaded=True, time_budget_s=60.0, n_streams=n_streams)
    except Exception as e:  # the reference's multi-threaded mode can throw its ring-overrun error (cpp:337-344)
        res = cpu_reference_run(stream, spec_name, columns_target=target, multi_threaded=False, time_budget_s=60.0, n_streams=n_streams)
        res["sample"] += f" (multi-threaded run failed: {str(e)[:80]})"
    cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    try:
        lat = cpu_reference_latency(stream, spec_name)
        if lat:
            cpu.update(lat)
    except Exception as e:
        cpu["latency_error"] = str(e)[:120]
    line = {
        "impl": "reference", "metric": metric_name(args), "value": res["value"], "unit": "columns/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * args.batch * n_streams / res["value"] if res["value"] else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64 (reference CPU arithmetic)",
        "data": "synthetic", "config": config_dict(args), "cpu_baseline": cpu,
        "e2e": {"value": res["value"], "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------- facade harness
def cabi_e2e(cfg_c, rows, pts, poses, batch, warm, steps, device, label_prefetch=True):
    """build/libcc_cabi_bench.so (facade/tools/cabi_bench.cpp): the C ABI driven by a C++ caller -- W synchronous warm-up
    pushes, then `steps` timed pushes with submit(k + 2); wait(k) over the given page-locked host buffers, results read on
    the host after every wait. Returns None when the harness has not been built."""
    path = os.path.join(HERE, "build", "libcc_cabi_bench.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.cb_e2e.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
    out = np.zeros(4, dtype=np.float64)
    marks = np.zeros(max(steps, 1), dtype=np.float64)
    slots = np.full((max(steps, 1), 5), -1.0, dtype=np.float32) if os.environ.get("CC_BENCH_SLOT_TIMES") else None
    err = C.create_string_buffer(256)
    tf = np.asarray(IDENTITY, dtype=np.float64)
    rc = lib.cb_e2e(C.addressof(cfg_c), rows, tf.ctypes.data, device, batch, warm, steps, pts.ctypes.data, poses.ctypes.data,
                    int(label_prefetch), out.ctypes.data, marks.ctypes.data, slots.ctypes.data if slots is not None else None, err)
    if rc != 0:
        raise RuntimeError("cabi_bench: " + err.value.decode(errors="replace"))
    if slots is not None:
        for i in range(min(steps, 6)):
            print(f"e2e push {i}: wait returned {marks[i]:.3f} | h2d {slots[i][0]:.3f}->{slots[i][1]:.3f} kernels {slots[i][2]:.3f}->{slots[i][3]:.3f} "
                  f"results {slots[i][4]:.3f}", file=sys.stderr)
    return {"seconds": float(out[0]), "d2h_bytes_per_step": float(out[1]), "exact_pushes": int(out[3]), "marks_ms": marks[:steps].tolist()}


def rows_around_the_path(device):
    """SURVEY 8f rows 2-4 (the steps before and after the hot path), each timed through its C-ABI call with host buffers in and
    its results on the host, beside the reference's own code (oracle/_ref/libcc_eval_ref.so, excerpts compiled unmodified) or the
    restatement (packet decode) on one host core. Synthetic inputs of the tests (tests/test_kitti.py, test_packets.py, test_evaluation.py)."""
    from continuous_clustering_b200 import KittiEvaluation, KittiReplay, OusterInput, synth

    out = {}

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    eval_ref = os.path.join(HERE, "oracle", "_ref", "libcc_eval_ref.so")
    ref = C.CDLL(eval_ref) if os.path.exists(eval_ref) else None
    # ---- f2: one KITTI frame -> 2200 pseudo firings resident on the device (+ their poses)
    xyzi, s0, s1, pstamps, poses, mid = synth.make_kitti_frame(seed=7, frame_index=3)
    kr = KittiReplay(device=device)
    kr.set_poses(pstamps, poses)
    t = timed(lambda: kr.frame(xyzi, s0, s1, mid, 0, 3), 20)
    kr.close()
    entry = {"unit": "frames/s", "points_per_frame": int(xyzi.shape[0]), "value": 1.0 / t, "ms_per_frame": 1e3 * t,
             "call": "cc_kitti_frame (host points in, 2200 x 64 RawPoint firings + poses left in device memory)"}
    if ref is not None:
        vp = C.c_void_p
        ref.ev_frame_to_firings.argtypes = [C.c_int, vp, C.c_uint64, C.c_uint64, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
        n = xyzi.shape[0]
        firings = np.zeros(2200 * 64 * 48, np.uint8)
        fposes, laser, cell, unc = np.zeros((2200, 12)), np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros((n, 3), np.float32)
        pc = np.ascontiguousarray(poses, dtype=np.float64)
        tc = timed(lambda: ref.ev_frame_to_firings(n, xyzi.ctypes.data, s0, s1, mid.ctypes.data, len(pstamps), pstamps.ctypes.data, pc.ctypes.data, 0, 3,
                                                   firings.ctypes.data, fposes.ctypes.data, laser.ctypes.data, cell.ctypes.data, unc.ctypes.data), 5, 1)
        entry["cpu_reference"] = {"value": 1.0 / tc, "ms_per_frame": 1e3 * tc, "cores": 1, "kind": "reference (kitti_loader.cpp / kitti_demo.cpp excerpts)"}
    out["kitti_frame_to_firings"] = entry
    # ---- f3: 256 Ouster lidar packets (16 blocks x 32 pixels) -> firings resident on the device
    direction, offset = synth.ouster_xyz_lut("left")
    packets, stamps = synth.make_ouster_packets(257, rows=32, columns_per_frame=1024, seed=1)
    dec = OusterInput(32, 1024, direction, offset, device=device, max_packets_per_call=256)
    dec.decode(packets[:1], stamps[:1])  # the packet in flight after a reset is dropped
    t = timed(lambda: dec.decode(packets[1:], stamps[1:]), 30)
    entry = {"unit": "packets/s", "value": 256 / t, "us_per_256_packets": 1e6 * t, "pixels_per_packet": 512,
             "call": "cc_ouster_decode (host packets in, RawPoint firings left in device memory, firing stamps on the host)"}
    orc_path = os.path.join(HERE, "oracle", "libcc_oracle.so")
    if os.path.exists(orc_path):
        orc = C.CDLL(orc_path)
        vp = C.c_void_p
        orc.orc_ouster_decode.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_uint64, vp, vp]
        f_out, s_out = np.zeros(256 * 16 * 32 * 48, np.uint8), np.zeros(256 * 16, np.uint64)
        pk, st = np.ascontiguousarray(packets[1:]), np.ascontiguousarray(stamps[1:])
        tc = timed(lambda: orc.orc_ouster_decode(C.addressof(dec.format), direction.ctypes.data, offset.ctypes.data, 256, pk.ctypes.data, pk.shape[1],
                                                 st.ctypes.data, 0, 0, f_out.ctypes.data, s_out.ctypes.data), 10, 1)
        entry["cpu_reference"] = {"value": 256 / tc, "us_per_256_packets": 1e6 * tc, "cores": 1, "kind": "port (oracle/cc_packets_oracle.cpp, parity unpinned)"}
    dec.close()
    out["ouster_packets_to_firings"] = entry
    # ---- f4: evaluation metrics of one frame
    rng = np.random.RandomState(4)
    n = 120000
    sem = rng.choice(np.array([60, 40, 44, 48, 49, 72, 0, 1, 10, 11, 30, 50, 51, 70, 71, 80, 81, 99], dtype=np.uint16), size=n).astype(np.uint16)
    ground = (rng.uniform(size=n) < 0.4).astype(np.uint8)
    gt = rng.randint(0, 300, size=n).astype(np.uint32)
    det = ((gt * 7 + (rng.uniform(size=n) < 0.3) * rng.randint(0, 5, size=n)) % 500).astype(np.uint32)
    ev = KittiEvaluation(device=device, max_points_per_frame=1 << 17)
    t = timed(lambda: ev.evaluate(sem, ground, gt, det), 30)
    ev.close()
    entry = {"unit": "frames/s", "points_per_frame": n, "value": 1.0 / t, "us_per_frame": 1e6 * t,
             "call": "cc_eval_frame (host label arrays in, six numbers out)"}
    if ref is not None:
        ref.ev_evaluate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        o6 = np.zeros(6)
        tc = timed(lambda: ref.ev_evaluate(n, sem.ctypes.data, ground.ctypes.data, gt.ctypes.data, det.ctypes.data, o6.ctypes.data), 5, 1)
        entry["cpu_reference"] = {"value": 1.0 / tc, "us_per_frame": 1e6 * tc, "cores": 1, "kind": "reference (kitti_evaluation.cpp excerpts)"}
    out["evaluation_metrics"] = entry
    return out


def facade_run(cfg_c, sp, pts, poses, batch, pipelined, callback_mode, warm, device, want_calls=False):
    """build/libcc_facade_bench.so (facade/tools/facade_bench.cpp): the stream through the drop-in C++ class, one
    addFiring call per firing."""
    path = os.path.join(HERE, "build", "libcc_facade_bench.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.fb_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_int, C.c_void_p, C.c_void_p, C.c_char_p]
    n = pts.shape[0]
    calls = np.zeros(n, dtype=np.float64) if want_calls else None
    result = np.zeros(8, dtype=np.float64)
    err = C.create_string_buffer(256)
    tf = np.asarray(IDENTITY, dtype=np.float64)
    pts = np.ascontiguousarray(pts)
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    rc = lib.fb_run(C.addressof(cfg_c), sp.rows, tf.ctypes.data, n, pts.ctypes.data, poses.ctypes.data, batch, int(pipelined),
                    callback_mode, warm, device, calls.ctypes.data if want_calls else None, result.ctypes.data, err)
    if rc != 0:
        raise RuntimeError("facade_bench: " + err.value.decode(errors="replace"))
    return {"seconds": float(result[0]), "column_callbacks": int(result[1]), "cluster_callbacks": int(result[2]),
            "cluster_points": int(result[3]), "packed_bytes": int(result[5]), "calls_us": calls}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="firings per push (= per step); 4096 = two rotations of 64x2048")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spec", default=SPEC, choices=sorted(SPEC_TEXT))
    ap.add_argument("--moving", action="store_true", help="moving sensor (10 m/s, 0.2 rad/s yaw) instead of the static pose")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident leg + kernel tables only (profiling runs)")
    ap.add_argument("--quick-e2e", action="store_true", help="device-resident and end-to-end legs only (A/B runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    from continuous_clustering_b200 import ContinuousClustering
    from continuous_clustering_b200.presets import stream_configuration

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200): there is no CPU path in continuous_clustering_b200")
    torch.cuda.set_device(local_rank)
    dist = None
    cpu_binding = None
    if world > 1 and not os.environ.get("CC_BENCH_NO_BIND"):
        # one rank per GPU on one host: give every rank's host thread (it polls for completion) its own share of the
        # cores the job may use, so that the ranks do not migrate over each other (round-1 verdict: e2e scaling)
        try:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // max(1, local_world))
            mine = cores[(local_rank * per) % len(cores):][:per]
            if mine:
                os.sched_setaffinity(0, set(mine))
                cpu_binding = {"cores": mine, "of": len(cores)}
        except Exception as e:
            cpu_binding = {"error": str(e)[:80]}
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W, B = args.steps, args.warmup, args.batch
    spec_name = rank_spec(args.spec, rank)
    # every rank = a different sensor stream; the moving variant needs every firing it will ever feed up front
    stream = Stream(spec_name, 1234 + rank, args.moving, total_firings=(W + K) * B + 64 * B if args.moving else 0)
    sp = stream.sp
    cfg = stream_configuration(spec_name)
    R = sp.rows
    B = min(B, 3 * sp.num_columns)  # the ring keeps 10 rotations: a push may span at most 3
    rec_bytes, pose_bytes = R * 48, 12 * 8
    full = rank == 0 and world == 1 and not args.quick and not args.moving and not args.quick_e2e

    def new_handle(batch=None):
        cc = ContinuousClustering(device=local_rank, max_firings_per_push=max(batch or B, 256))
        cc.setConfiguration(cfg)
        cc.reset(R)
        cc.setTransformRobotFrameFromSensorFrame(IDENTITY)
        return cc

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def l2_flush(tag):
        flush_buf.fill_(tag & 0xFF)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def to_device(pts, poses):
        n = pts.shape[0]
        return (torch.from_numpy(pts.view(np.uint8).reshape(n, R * 48)).cuda(), torch.from_numpy(poses).cuda())

    def pinned(pts, poses):
        n = pts.shape[0]
        pp = torch.from_numpy(pts.view(np.uint8).reshape(n, R * 48)).pin_memory()
        pq = torch.from_numpy(poses).pin_memory()
        return pp, pq, pp.numpy().view(pts.dtype).reshape(n, R), pq.numpy()

    # ------------------------------------------------------------------ device-resident leg ("value")
    def device_leg(B, K, W, sample_clocks, flush_each_step=False, src=None, reduce=True):
        """K timed pushes of B firings with the inputs already in HBM; returns a dict.
        flush_each_step=False: the pushes run back to back; every push reads input bytes nothing has touched since L2 was
        flushed right before the timed region (inputs larger than L2, streamed once), the stream's own state stays as
        warm as it is in steady state. flush_each_step=True: additionally a 256 MiB write before every push, its
        event-timed duration subtracted."""
        src = src or stream
        total = (W + K) * B
        pts, poses = src.take(0, total)
        d_pts, d_poses = to_device(pts, poses)
        cc = new_handle(B)
        cuda_stream = torch.cuda.ExternalStream(cc.stream)

        def submit_dev(step):
            cc.submitFiringsDevice(d_pts.data_ptr() + step * B * rec_bytes, d_poses.data_ptr() + step * B * pose_bytes, B, R)

        for s in range(W):
            cc.addFiringsDevice(d_pts.data_ptr() + s * B * rec_bytes, d_poses.data_ptr() + s * B * pose_bytes, B, R)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if reduce:
            barrier()
        else:
            torch.cuda.synchronize()
        if sampler:
            sampler.start()
        launches0 = cc.total_launches
        # Timed region: K pushes, two in flight (submit(k + 1); wait(k)) so that the host's result handling of push k
        # overlaps the kernels of push k + 1.
        ev_fa = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ev_fb = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ev_end = torch.cuda.Event(enable_timing=True)
        ev_begin = torch.cuda.Event(enable_timing=True)
        dev_ms = []
        exact_pushes = 0

        def flush_and_submit(s):
            if flush_each_step:
                with torch.cuda.stream(cuda_stream):
                    ev_fa[s].record(cuda_stream)
                    l2_flush(s)
                    ev_fb[s].record(cuda_stream)
            submit_dev(W + s)

        with torch.cuda.stream(cuda_stream):
            l2_flush(255)  # nothing of the inputs is cache resident when the timed region starts
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        with torch.cuda.stream(cuda_stream):
            ev_begin.record(cuda_stream)
        flush_and_submit(0)
        for s in range(K):
            if s + 1 < K:
                flush_and_submit(s + 1)
            res = cc.wait()
            dev_ms.append(res.info.device_ms)
            exact_pushes += int(res.info.used_exact_path)
        with torch.cuda.stream(cuda_stream):
            ev_end.record(cuda_stream)
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        launches = cc.total_launches - launches0
        flush_ms = sum(a.elapsed_time(b) for a, b in zip(ev_fa, ev_fb)) if flush_each_step else 0.0
        total_ms = ev_begin.elapsed_time(ev_end)
        clocks = sampler.stop() if sampler else None
        elapsed = (total_ms - flush_ms) / 1e3
        nranks = 1
        if dist is not None and reduce:
            t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed = float(t.item())
            nranks = world
        return {"cc": cc, "stream": cuda_stream, "value": nranks * K * B / elapsed, "elapsed": elapsed, "dev_ms": dev_ms,
                "launches": launches, "exact": exact_pushes, "clocks": clocks, "t_wall": t_wall, "fed": total}

    leg = device_leg(B, K, W, True)
    cc, cuda_stream = leg["cc"], leg["stream"]
    value, elapsed, dev_ms, launches, exact_pushes, clocks, t_wall = (leg["value"], leg["elapsed"], leg["dev_ms"], leg["launches"],
                                                                      leg["exact"], leg["clocks"], leg["t_wall"])
    total = leg["fed"]

    # ------------------------------------------------------------------ end-to-end leg through the public API
    # (runs right after the device-resident leg, before the other operating points: it is the headline and should not
    # depend on what the optional legs leave behind in the process)
    cce = new_handle()
    cce.set_label_prefetch(True)  # the ground labels of the new columns come back with every push's results
    K_e2e = min(K, 2) if args.quick else K
    h_pts, h_poses = stream.take(0, (W + K_e2e) * B)
    # page-locked host buffers (the contract's "pinned host memory"): cc_submit_firings copies them straight to the device
    pin_pts, pin_poses, h_pts, h_poses = pinned(h_pts, h_poses)
    d2h = 0
    for s in range(W):
        cce.addFirings(h_pts[s * B:(s + 1) * B], h_poses[s * B:(s + 1) * B])
    # (scripts/h2d_probe4.py: only the very first device access to freshly page-locked pages is slow, ~20 GB/s; CPU rewrites
    # of the buffer do not make it slow again.)
    # The page-locked staging memory is "DMA warm", as that of any long-running producer is: on this pool's hosts the
    # FIRST device access to freshly page-locked pages runs at ~35 GB/s, every later one at ~50 GB/s (scripts/h2d_probe2.py).
    # One untimed copy of the whole buffer to a scratch tensor touches every page; the data of the timed pushes has still
    # never been in HBM or in the GPU's L2 when its push starts.
    # (scripts/h2d_vs_kernels.py: after ONE warming pass the next pass still runs at ~42 GB/s, from the third on at the
    # link's 53-54 GB/s -- with or without pushes running beside it: three passes.)
    for _ in range(3):
        scratch, scratch_poses = pin_pts.cuda(), pin_poses.cuda()  # (the poses travel with every push too: 96 B per firing)
        torch.cuda.synchronize()
        del scratch, scratch_poses
    barrier()
    e2e_marks = []
    t0 = time.perf_counter()
    # two pushes in flight and a third one staged: the host->device copy of push k + 2 (input stream) overlaps the
    # kernels of pushes k and k + 1; its kernels are launched by the wait() that returns push k
    for s in range(W, min(W + 2, W + K_e2e)):
        cce.submitFirings(h_pts[s * B:(s + 1) * B], h_poses[s * B:(s + 1) * B])
    for s in range(W, W + K_e2e):
        if s + 2 < W + K_e2e:
            cce.submitFirings(h_pts[(s + 2) * B:(s + 3) * B], h_poses[(s + 2) * B:(s + 3) * B])
        res = cce.wait()
        e2e_marks.append(time.perf_counter() - t0)
        labels = cce.column_labels()  # [n_cols, rows, 4] u8: ground label, debug label, is_ignored, intensity
        assert labels.shape[0] == int(res.info.ground_to_gcol - res.info.ground_from_gcol)
        d2h += res.clusters.nbytes + res.cluster_points.nbytes + labels.nbytes + labels.shape[0] * 8 + 512
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_rank = K_e2e * B / e2e_s
    per_rank_e2e = None
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, e2e_rank)
        per_rank_e2e = gathered
    e2e_python = world * K_e2e * B / e2e_s
    cce.close()
    # The same loop written by a C++ caller of the C ABI (facade/tools/cabi_bench.cpp), same page-locked buffers: no
    # interpreter between the calls. This is the reported end-to-end number; the Python loop above is kept beside it
    # (its wait() returns are 10-20 % further apart on this host: interpreter jitter between submit and wait).
    e2e_c = None
    try:
        barrier()
        e2e_c = cabi_e2e(cfg.to_c(), R, h_pts, h_poses, B, W, K_e2e, local_rank)
    except Exception as e:
        print(f"cabi e2e leg failed: {e}", file=sys.stderr)
    e2e_value, e2e_how, e2e_marks_c = e2e_python, "python loop (ContinuousClustering.submitFirings / wait)", None
    if e2e_c is not None:
        c_s = e2e_c["seconds"]
        if dist is not None:
            t = torch.tensor([c_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            gathered = [None] * world
            dist.all_gather_object(gathered, K_e2e * B / c_s)
            per_rank_e2e = gathered
            c_s = float(t.item())
        e2e_value = world * K_e2e * B / c_s
        e2e_how = "C++ caller of the C ABI (cc_submit_firings / cc_wait, build/libcc_cabi_bench.so)"
        e2e_marks_c = [round(m, 3) for m in e2e_c["marks_ms"]]

    # per-push latency with ONE push of B firings in flight (the synchronous call), device-resident inputs
    sync_ms = []
    lat_pts, lat_poses = stream.take(total, 6 * B)
    d_lat, d_lat_poses = to_device(lat_pts, lat_poses)
    for r in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cc.addFiringsDevice(d_lat.data_ptr() + r * B * rec_bytes, d_lat_poses.data_ptr() + r * B * pose_bytes, B, R)
        sync_ms.append(1e3 * (time.perf_counter() - t0))
    total += 6 * B

    # device timeline of the kernels as they overlap in a normal push (globaltimer stamps, cc_debug_trace): the CUDA-event
    # table below serialises the launches and adds an event round trip to every kernel
    trace_table = None
    if rank == 0:
        tr_pts, tr_poses = stream.take(total, 4 * B)
        d_tr, d_tr_poses = to_device(tr_pts, tr_poses)
        cc.debug_trace(True)
        acc_tr = {}
        for r in range(4):
            with torch.cuda.stream(cuda_stream):
                l2_flush(r)
            cc.addFiringsDevice(d_tr.data_ptr() + r * B * rec_bytes, d_tr_poses.data_ptr() + r * B * pose_bytes, B, R)
            for name, a, z, longest, blocks in cc.get_trace():
                if name.startswith("k_"):
                    acc_tr.setdefault(name, []).append((z - a) / 1e3)
        cc.debug_trace(False)
        trace_table = {k: round(float(np.mean(v)), 2) for k, v in sorted(acc_tr.items(), key=lambda kv: -np.mean(kv[1]))}
        total += 4 * B

    # ------------------------------------------------------------------ per-kernel timing + roofline (rank 0)
    roofline = None
    kernel_table = None
    if rank == 0:
        cc.set_kernel_timing(True)
        acc = {}
        reps = 5
        extra, extra_poses = stream.take(total, reps * B)
        d_extra, d_extra_poses = to_device(extra, extra_poses)
        n_trees = []
        for r in range(reps):
            with torch.cuda.stream(cuda_stream):
                l2_flush(r)
            res = cc.addFiringsDevice(d_extra.data_ptr() + r * B * rec_bytes, d_extra_poses.data_ptr() + r * B * pose_bytes, B, R)
            n_trees.append(int(getattr(res.info, "n_unfinished_trees", 0)))
            for name, ms in cc.kernel_timings():
                a = acc.setdefault(name, [0.0, 0])
                a[0] += ms
                a[1] += 1
        cc.set_kernel_timing(False)
        tot = sum(v[0] for v in acc.values())
        kernel_table = {k: {"ms_per_step": v[0] / reps, "share": v[0] / tot} for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])}
        peak, peak_src = load_peaks()
        top = next(iter(kernel_table))  # the kernel with the largest share of the step, whatever it is

        def alg_bytes(k):
            if k == "k_fin_all":
                # list work, not per-cell: per unfinished tree 36 B read (list entry, union-find parent, finish azimuth,
                # root column, last column, point count) + 4 B written (compacted list); per new column 8 + 8 B
                return int(np.mean(n_trees) if n_trees else 0) * 40 + B * 16
            return KERNEL_BYTES_PER_CELL.get(k, 0) * B * R

        dur_s = kernel_table[top]["ms_per_step"] / 1e3
        achieved = alg_bytes(top) / dur_s / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(HERE, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
            with open(tp) as f:
                tj = json.load(f)
            if top in tj.get("kernels", {}):
                traffic = tj["kernels"][top]["dram_bytes"]
                traffic_src = tj.get("source")
        by_kernel = {}
        for k, v in kernel_table.items():
            ab = alg_bytes(k)
            by_kernel[k] = {"algorithmic_bytes_per_launch": ab, "achieved_gbs": ab / (v["ms_per_step"] / 1e3) / 1e9,
                            "frac": ab / (v["ms_per_step"] / 1e3) / 1e9 / peak,
                            "extra_implementation_bytes_per_launch": KERNEL_EXTRA_BYTES_PER_CELL.get(k, 0) * B * R}
        single_cta = {"k_fin_all", "k_scan_lite", "k_insert_scan"}
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes(top), "kernel_share_of_step": kernel_table[top]["share"],
                    "limiter": ("latency: single-CTA list phases separated by block barriers, not bandwidth" if top in single_cta
                                else "latency of dependent L2 round trips; the working set of a push is L2 resident"),
                    "by_kernel": by_kernel,
                    "path_bytes_per_cell": 151, "path_achieved_gbs": value / world * R * 151 / 1e9,
                    "path_frac": value / world * R * 151 / 1e9 / peak}
    cc.close()

    # ------------------------------------------------------------------ per-GPU push latency (every rank; BASELINE config 5)
    LB = 64

    def latency_leg_c(nl=200, warm=60):
        """The same through a C++ caller of the C ABI (facade/tools/cabi_bench.cpp cb_latency): no interpreter in the loop."""
        path = os.path.join(HERE, "build", "libcc_cabi_bench.so")
        if not os.path.exists(path):
            return None
        lib = C.CDLL(path)
        lib.cb_latency.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        lp, lq = stream.take(0, nl * LB)
        _keep, _keepq, hlp, hlq = pinned(lp, lq)
        call_us, dev_us, out = np.zeros(nl), np.zeros(nl), np.zeros(2)
        err = C.create_string_buffer(256)
        tf = np.asarray(IDENTITY, dtype=np.float64)
        cfg_c = cfg.to_c()
        rc = lib.cb_latency(C.addressof(cfg_c), R, tf.ctypes.data, local_rank, LB, nl, hlp.ctypes.data, hlq.ctypes.data, 1,
                            call_us.ctypes.data, dev_us.ctypes.data, out.ctypes.data, err)
        if rc != 0:
            raise RuntimeError("cabi_bench: " + err.value.decode(errors="replace"))
        return call_us[warm:], dev_us[warm:], int(out[1])

    def latency_leg(nl=200, warm=60):
        ccl = new_handle(LB)
        ccl.set_label_prefetch(True)
        lp, lq = stream.take(0, nl * LB)
        _keep, _keepq, hlp, hlq = pinned(lp, lq)
        lat, dev = [], []
        launches = 0
        for i in range(nl):
            t0 = time.perf_counter()
            r = ccl.addFirings(hlp[i * LB:(i + 1) * LB], hlq[i * LB:(i + 1) * LB])
            lat.append(1e6 * (time.perf_counter() - t0))
            dev.append(1e3 * r.info.device_ms)
            launches = int(r.info.gpu_launches)
        ccl.close()
        lat, dev = np.array(lat[warm:]), np.array(dev[warm:])
        python_loop = {"per_push_us_p50": float(np.median(lat)), "per_push_us_p99": float(np.percentile(lat, 99)),
                       "call": "ContinuousClustering.addFirings (Python mirror: ctypes + numpy views around the same C call)"}
        call = "cc_push_firings via ContinuousClustering.addFirings (page-locked host buffers in; events / clusters / member lists / labels back on the host)"
        try:
            c = latency_leg_c(nl, warm)
        except Exception as e:
            print(f"cabi latency leg failed: {e}", file=sys.stderr)
            c = None
        if c is not None:
            lat, dev, launches = c
            call = ("cc_push_firings from a C++ caller of the C ABI (build/libcc_cabi_bench.so; page-locked host buffers in; events / "
                    "clusters / member lists / labels read on the host before the clock stops)")
        hist, edges = np.histogram(lat, bins=[0, 60, 70, 80, 90, 100, 110, 120, 140, 160, 200, 300, 1e9])
        return {"batch_firings": LB, "call": call, "python_loop": python_loop, "launches_per_push": launches,
                "per_push_us_p50": float(np.median(lat)), "per_push_us_p99": float(np.percentile(lat, 99)),
                "per_push_us_mean": float(lat.mean()), "per_push_device_us_p50": float(np.median(dev)),
                "per_push_ms_p50": float(np.median(lat)) / 1e3, "per_push_ms_p99": float(np.percentile(lat, 99)) / 1e3,
                "columns_per_s": LB / (float(np.mean(lat)) / 1e6),
                "histogram_us": {"edges": [float(e) for e in edges[:-1]] + ["inf"], "counts": [int(c) for c in hist]}}

    latency_mode = None
    per_gpu_latency = None
    if not args.quick and not args.quick_e2e:
        barrier()
        mine = latency_leg()
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, {"rank": rank, "gpu": local_rank, "spec": spec_name,
                                              **{k: mine[k] for k in ("per_push_us_p50", "per_push_us_p99", "per_push_device_us_p50", "histogram_us")}})
            per_gpu_latency = gathered
        else:
            per_gpu_latency = [{"rank": 0, "gpu": local_rank, "spec": spec_name,
                                **{k: mine[k] for k in ("per_push_us_p50", "per_push_us_p99", "per_push_device_us_p50", "histogram_us")}}]
        latency_mode = mine

    # ------------------------------------------------------------------ other operating points (rank 0, single GPU)
    batch_sweep, flushed, facade, exact_path, rows_around, irregular = None, None, None, None, None, None
    if full:
        lg = device_leg(B, 10, 3, False, flush_each_step=True)
        flushed = {"columns_per_s": lg["value"], "ms_per_step": 1e3 * lg["elapsed"] / 10,
                   "per_push_device_ms_p50": float(np.median(lg["dev_ms"])),
                   "how": "256 MiB device write before every timed push, its event-timed duration subtracted"}
        lg["cc"].close()
        batch_sweep = {}
        for b2 in (256, 1024, 2048, 6144):
            if b2 == B or b2 > 3 * sp.num_columns:
                continue
            lg = device_leg(b2, 8, 3, False)
            batch_sweep[str(b2)] = {"columns_per_s": lg["value"], "ms_per_step": 1e3 * lg["elapsed"] / 8,
                                    "per_push_device_ms_p50": float(np.median(lg["dev_ms"])),
                                    "launches_per_push": lg["launches"] / 8}
            lg["cc"].close()
        # the exact column-sequential path: a closed wall around the sensor makes one cluster span a full rotation
        # (forced finish, cpp:909-919): speculative commits abort and the flagged columns go through k_careful
        try:
            from continuous_clustering_b200 import synth

            class WallStream:
                pass

            ws = WallStream()
            wp, wq, ws.sp = synth.make_stream(spec_name, n_rotations=3.0, seed=3, n_boxes=0, wall_radius=12.0)
            ws.take = lambda a, n: (wp[a:a + n].copy(), wq[a:a + n].copy())
            bw = 1024
            kw = wp.shape[0] // bw - 2
            lg = device_leg(bw, kw, 2, False, src=ws)
            exact_path = {"workload": "closed wall around the sensor (one cluster spans a rotation: forced finish)",
                          "batch_firings": bw, "steps": kw, "columns_per_s": lg["value"], "exact_path_pushes": lg["exact"],
                          "per_push_device_ms_p50": float(np.median(lg["dev_ms"])), "launches_per_push": lg["launches"] / kw}
            lg["cc"].close()
        except Exception as e:
            exact_path = {"error": str(e)[:200]}
        # a sensor that does not deliver exactly one firing per range-image column (5 % more firings per rotation than
        # columns: two firings share a column every ~20 firings, the collision rule cpp:188-208 fires): the insertion leaves
        # the grid-wide regular path at the first such firing of a push and the rest goes through the single-CTA scan
        try:
            from continuous_clustering_b200 import synth

            class IrregularStream:
                pass

            irs = IrregularStream()
            ib = 2048
            ip, iq, irs.sp = synth.make_stream(spec_name, n_firings=6 * ib, seed=5, az_step_scale=0.95)
            irs.take = lambda a, n: (ip[a:a + n].copy(), iq[a:a + n].copy())
            lg = device_leg(ib, 3, 3, False, src=irs)
            irregular = {"workload": "the same scene, sensor delivering 5 % more firings per rotation than the range image has columns "
                                     "(az_step_scale 0.95): column collisions every ~20 firings",
                         "batch_firings": ib, "steps": 3, "columns_per_s": lg["value"], "per_push_device_ms_p50": float(np.median(lg["dev_ms"])),
                         "exact_path_pushes": lg["exact"],
                         "note": "after the first irregular firing of a push the insertion runs in one CTA (k_insert_scan): the known slow case of the path"}
            lg["cc"].close()
        except Exception as e:
            irregular = {"error": str(e)[:200]}
        try:
            rows_around = rows_around_the_path(local_rank)
        except Exception as e:
            rows_around = {"error": str(e)[:200]}
        # the drop-in C++ class: one addFiring call per firing
        try:
            cfg_c = cfg.to_c()
            nf = 16 * 2048
            fp, fq = stream.take(0, nf + 2048)
            facade = {}
            for name, batch, pipelined, mode in (("sync_batch64_kitti_callbacks", 64, 0, 1), ("sync_batch64_no_callbacks", 64, 0, 0),
                                                  ("sync_batch64_packed_pointcloud2", 64, 0, 2),
                                                  ("pipelined_batch2048_kitti_callbacks", 2048, 1, 1),
                                                  ("pipelined_batch2048_packed_pointcloud2", 2048, 1, 2),
                                                  ("pipelined_batch2048_no_callbacks", 2048, 1, 0)):
                r = facade_run(cfg_c, sp, fp, fq, batch, pipelined, mode, 2048, local_rank, want_calls=(batch == 64))
                if r is None:
                    facade = {"unavailable": "build/libcc_facade_bench.so not built"}
                    break
                entry = {"columns_per_s": nf / r["seconds"], "column_callbacks": r["column_callbacks"],
                         "cluster_callbacks": r["cluster_callbacks"], "packed_bytes": r["packed_bytes"]}
                if r["calls_us"] is not None:
                    calls = r["calls_us"][2048:]
                    pushes = calls[batch - 1::batch]  # every batch-th call carries the device push and its callbacks
                    entry.update({"addFiring_us_p50": float(np.median(calls)), "addFiring_us_p99": float(np.percentile(calls, 99)),
                                  "push_call_us_p50": float(np.median(pushes)), "push_call_us_p99": float(np.percentile(pushes, 99))})
                facade[name] = entry
        except Exception as e:
            facade = {"error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_stream = stream if not args.moving else Stream(spec_name, 1234)
        try:
            cpu = cpu_reference_run(cpu_stream, spec_name, columns_target=10_000_000, multi_threaded=True, time_budget_s=20.0)
        except Exception as e:
            cpu = cpu_reference_run(cpu_stream, spec_name, columns_target=10_000_000, multi_threaded=False, time_budget_s=20.0)
            cpu["sample"] += f" (multi-threaded run failed: {str(e)[:80]})"
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            lat = cpu_reference_latency(cpu_stream, spec_name)
            if lat:
                cpu.update(lat)
        except Exception as e:
            cpu["latency_error"] = str(e)[:120]

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": "columns/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (+f64 rigid transforms, u32 union-find)", "data": "synthetic",
            "config": config_dict(args),
            "run": {"l2": f"inputs larger than L2: {(W + K) * B * (rec_bytes + pose_bytes) >> 20} MiB of firings streamed once, "
                          f"{B * (rec_bytes + pose_bytes) >> 20} MiB of never-touched input per step; L2 flushed (256 MiB write) once "
                          "before the timed region, timed pushes run back to back; l2_flush_each_step reports the same with a "
                          "flush before every push",
                    "pipelining": "two pushes in flight (cc_submit_firings_device / cc_wait); end-to-end leg: a third host push staged",
                    "cpu_binding": cpu_binding, "streams": f"{world} independent sensor stream(s), one per GPU, no data-path collective",
                    "exact_path_pushes": exact_pushes},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": B * (rec_bytes + pose_bytes),
                    "d2h_bytes_per_step": int(d2h // K_e2e),
                    "host_buffers": "page-locked, DMA-warm (three untimed device reads of the staging memory before the timed region, as the reused staging buffers of a long-running producer are; every timed push still reads bytes that have never been in HBM or L2)",
                    "caller": e2e_how, "wait_return_ms": e2e_marks_c, "python_loop": {"value": e2e_python, "wait_return_ms": [round(1e3 * m, 3) for m in e2e_marks]}, "h2d_gbs_per_rank": e2e_value / world * (rec_bytes + pose_bytes) / 1e9,
                    "per_rank_columns_per_s": per_rank_e2e},
            "latency": {"per_push_device_ms_p50": float(np.median(dev_ms)),
                        "per_push_sync_call_ms_p50": float(np.median(sync_ms)),
                        "note": "one push = batch_firings columns; every column of a push is charged the whole push",
                        "wall_s": t_wall},
            "roofline": roofline, "cpu_baseline": cpu, "kernels": kernel_table,
            "kernels_device_timeline_us": trace_table, "batch_sweep": batch_sweep, "latency_mode": latency_mode,
            "per_gpu_latency": per_gpu_latency, "facade": facade, "exact_path": exact_path, "irregular_stream": irregular, "rows_around_the_path": rows_around,
            "l2_flush_each_step": flushed,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
