#!/usr/bin/env python
"""bench.py -- range-image columns/s of the per-column hot path on a synthetic 64-ring / 10 Hz stream.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One STEP = one push of `--batch` consecutive firings (default 4096 = two sensor rotations of the 64 x 2048
synthetic Velodyne-like stream, BASELINE.json configs[1]) through insertion, ground segmentation, association,
finish detection and ring recycling. The stream keeps going across steps (the range image is continuous).

  value     columns/s with the firings already resident in HBM when the timed region starts (cc_submit_firings_device /
            cc_wait, two pushes in flight), timed with CUDA events on the handle's stream over K back-to-back pushes.
            Inputs larger than L2: every push reads firings nothing has touched since L2 was flushed right before the
            timed region. `l2_flush_each_step` reports the same with a 256 MiB write before every push.
  e2e       the same metric through the public API with HOST buffers (ContinuousClustering.submitFirings / wait:
            page-locked host -> device copy of the raw firings, kernels, device -> host copy of events, finished
            clusters, member lists and the ground labels of the new columns), wall clock.
  batch_sweep / latency_mode   other operating points: 1024 / 2048 / 6144 firings per push (device resident), and
            64-firing synchronous pushes with host buffers (per-push latency).
  roofline  the dominant kernel of the step (largest share of device time, measured live with CUDA events around
            every launch): algorithmic bytes it must move / its duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own CPU implementation (oracle/_ref/libcc_ref.so, built from the reference's
            sources) -- or the restatement in oracle/ when that build is absent -- timed on this box's host cores
            on a bounded sample of the same stream.

N > 1 (torchrun, one rank per GPU): every rank runs its own independent sensor stream (streams shard one per GPU;
there is no data-path collective), barrier + max-over-ranks timing, value = total columns / max time ("weak").
`--impl reference` times only the CPU reference arm (rank 0 only).
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

SPEC = "velodyne64"
IDENTITY = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
METRIC = "range-image columns/s, 64-ring stream"

# algorithmic bytes per range-image cell, per kernel (151 B/cell for the path)
KERNEL_BYTES_PER_CELL = {
    # read x,y,z (+ pose, amortised); write staged odom xyz, distance, azimuth, inclination, column (twice: per firing
    # and per row); + every field group reset for a recycled cell (one cell retired per cell inserted in steady state)
    "k_prep": 12 + 28 + (57 + 3 + 8 + 16 + 4 + 4 + 4 + 2 + 4),
    "k_insert_scan": 8 + 12 + 4,  # read column-in-rotation + distance; write resolved column + rotation; distance write-through
    "k_scatter": 37 + 57,  # SURVEY 8d insert: read raw record fields, write the 57 B of range-image fields
    "k_gap_scan": 4 + 4,
    "k_ground": 21 + 3 + 16 + 4,  # SURVEY 8d ground: 21 read + 3 written, + the association view (16) and mad (4)
    "k_probe": 21 + 4,  # SURVEY 8d associate
    "k_probe_heavy": 21 + 4,
    "k_commit_copy": 4 + 4,
    "k_commit_roots": 4 + 4,
    "k_fin_label": 4 + 4,  # SURVEY 8d finish/label
    "k_clear": 57 + 3 + 8 + 16 + 4 + 4 + 4 + 2 + 4,  # every field group reset for a recycled cell
}


def load_peaks():
    p = os.path.join(HERE, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_rotations(n_rot_unique=2, seed=1234):
    """A few unique rotations of the static-sensor stream; longer streams tile them (a static scene repeats every
    rotation anyway) with fresh stamps / firing indices / unique point indices."""
    from continuous_clustering_b200 import synth

    pts, poses, sp = synth.make_stream(SPEC, n_rotations=n_rot_unique, seed=seed)
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


def cpu_reference_run(base_pts, base_poses, sp, columns_target, multi_threaded=True, time_budget_s=25.0):
    """Times the reference's CPU implementation on this box's host cores. Returns dict for `cpu_baseline`."""
    from oracle import drvlib

    kind = "reference" if drvlib.have_ref() else "port"
    lib = drvlib.REF_LIB if kind == "reference" else drvlib.ORACLE_LIB
    if kind == "port":
        multi_threaded = False  # the restatement is the deterministic single-threaded mode only
    cfg = drvlib.stream_config(SPEC, is_single_threaded=0 if multi_threaded else 1)
    d = drvlib.Driver(lib)
    d.configure(cfg, sp.rows)
    d.set_record(0)
    rot = sp.num_columns
    fed = 0
    t_total = 0.0
    # warm-up: two rotations
    for w in range(2):
        pts, poses = tile_stream(base_pts, base_poses, sp, fed, rot)
        d.prepare(pts, poses)
        d.run_prepared(0, rot, 3 * rot if multi_threaded else 0)
        fed += rot
    timed = 0
    t_wall0 = time.time()
    chunk = 16 * rot  # long runs, so that draining the reference's thread pipeline between runs does not matter
    while timed < columns_target and (time.time() - t_wall0) < time_budget_s:
        pts, poses = tile_stream(base_pts, base_poses, sp, fed, chunk)
        d.prepare(pts, poses)  # shared_ptr construction outside the timed region
        t_total += d.run_prepared(0, chunk, 3 * rot if multi_threaded else 0)
        fed += chunk
        timed += chunk
    d.close()
    cores = 8 if multi_threaded else 1  # 4 stage threads + 3 publishers + producer (cpp:49-63)
    return {
        "value": timed / t_total if t_total > 0 else 0.0,
        "unit": "columns/s",
        "cores": cores,
        "kind": kind,
        "sample": f"{timed} columns ({timed // rot} rotations of the {sp.rows}x{rot} synthetic stream) after 2 warm-up "
                  f"rotations, {'multi-threaded 5-stage pipeline' if multi_threaded else 'single-threaded mode'}, "
                  f"no-op callbacks, host has {os.cpu_count()} logical cores",
        "seconds": t_total,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base_pts, base_poses, sp = make_rotations()
    steps, warm = args.steps, args.warmup
    try:
        res = cpu_reference_run(base_pts, base_poses, sp, columns_target=max(1, steps) * args.batch * 8,
                                multi_threaded=True, time_budget_s=60.0)
    except Exception as e:  # the reference's multi-threaded mode can throw its ring-overrun error (cpp:337-344)
        res = cpu_reference_run(base_pts, base_poses, sp, columns_target=max(1, steps) * args.batch * 8,
                                multi_threaded=False, time_budget_s=60.0)
        res["sample"] += f" (multi-threaded run failed: {str(e)[:80]})"
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "columns/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * args.batch / res["value"] if res["value"] else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64 (reference CPU arithmetic)",
        "data": "synthetic", "config": {"workload": workload_name(args), "batch_firings": args.batch},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_name(args):
    return (f"synthetic 64-ring 10 Hz Velodyne-like stream (64x2048 columns/rotation, ground plane + 150 boxes, "
            f"static sensor), {args.batch} firings per push")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="firings per push (= per step); 4096 = two rotations")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident leg + kernel tables only (profiling runs)")
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
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W, B = args.steps, args.warmup, args.batch
    base_pts, base_poses, sp = make_rotations(seed=1234 + rank)  # every rank = a different sensor stream
    cfg = stream_configuration(SPEC)
    R = sp.rows

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

    # ------------------------------------------------------------------ device-resident leg ("value")
    def device_leg(B, K, W, sample_clocks, flush_each_step=False):
        """K timed pushes of B firings with the inputs already in HBM; returns a dict.
        flush_each_step=False: the pushes run back to back; every push reads input bytes nothing has touched since L2 was
        flushed right before the timed region (inputs larger than L2, streamed once), the stream's own state stays as
        warm as it is in steady state. flush_each_step=True: additionally a 256 MiB write before every push, its
        event-timed duration subtracted."""
        total = (W + K) * B
        pts, poses = tile_stream(base_pts, base_poses, sp, 0, total)
        d_pts = torch.from_numpy(pts.view(np.uint8).reshape(total, R * 48)).cuda()
        d_poses = torch.from_numpy(poses).cuda()
        cc = new_handle(B)
        stream = torch.cuda.ExternalStream(cc.stream)

        def submit_dev(step):
            cc.submitFiringsDevice(d_pts.data_ptr() + step * B * rec_bytes, d_poses.data_ptr() + step * B * pose_bytes, B, R)

        for s in range(W):
            cc.addFiringsDevice(d_pts.data_ptr() + s * B * rec_bytes, d_poses.data_ptr() + s * B * pose_bytes, B, R)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        launches0 = cc.total_launches
        # Timed region: K pushes, two in flight (submit(k + 1); wait(k)) so that the host's result handling of push k
        # overlaps the kernels of push k + 1. Before every push L2 is flushed by a 256 MiB write on the same stream; the
        # flushes are bracketed by their own events and their device time is subtracted.
        ev_fa = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ev_fb = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ev_end = torch.cuda.Event(enable_timing=True)
        dev_ms = []
        exact_pushes = 0

        ev_begin = torch.cuda.Event(enable_timing=True)

        def flush_and_submit(s):
            if flush_each_step:
                with torch.cuda.stream(stream):
                    ev_fa[s].record(stream)
                    l2_flush(s)
                    ev_fb[s].record(stream)
            submit_dev(W + s)

        with torch.cuda.stream(stream):
            l2_flush(255)  # nothing of the inputs is cache resident when the timed region starts
            ev_begin.record(stream)
        torch.cuda.synchronize()
        t_wall0 = time.perf_counter()
        with torch.cuda.stream(stream):
            ev_begin.record(stream)
        flush_and_submit(0)
        for s in range(K):
            if s + 1 < K:
                flush_and_submit(s + 1)
            res = cc.wait()
            dev_ms.append(res.info.device_ms)
            exact_pushes += int(res.info.used_exact_path)
        with torch.cuda.stream(stream):
            ev_end.record(stream)
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        launches = cc.total_launches - launches0
        flush_ms = sum(a.elapsed_time(b) for a, b in zip(ev_fa, ev_fb)) if flush_each_step else 0.0
        total_ms = ev_begin.elapsed_time(ev_end)
        clocks = sampler.stop() if sampler else None
        elapsed = (total_ms - flush_ms) / 1e3
        if dist is not None:
            t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed = float(t.item())
        return {"cc": cc, "stream": stream, "value": world * K * B / elapsed, "elapsed": elapsed, "dev_ms": dev_ms,
                "launches": launches, "exact": exact_pushes, "clocks": clocks, "t_wall": t_wall, "fed": total}

    rec_bytes, pose_bytes = R * 48, 12 * 8
    leg = device_leg(B, K, W, True)
    cc, stream = leg["cc"], leg["stream"]
    value, elapsed, dev_ms, launches, exact_pushes, clocks, t_wall = (leg["value"], leg["elapsed"], leg["dev_ms"], leg["launches"],
                                                                      leg["exact"], leg["clocks"], leg["t_wall"])
    total = leg["fed"]

    # per-push latency with ONE push in flight (the synchronous call a latency-sensitive caller makes)
    sync_ms = []
    lat_pts, lat_poses = tile_stream(base_pts, base_poses, sp, total, 6 * B)
    d_lat = torch.from_numpy(lat_pts.view(np.uint8).reshape(6 * B, R * 48)).cuda()
    d_lat_poses = torch.from_numpy(lat_poses).cuda()
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
        tr_pts, tr_poses = tile_stream(base_pts, base_poses, sp, total, 4 * B)
        d_tr = torch.from_numpy(tr_pts.view(np.uint8).reshape(4 * B, R * 48)).cuda()
        d_tr_poses = torch.from_numpy(tr_poses).cuda()
        cc.debug_trace(True)
        acc_tr = {}
        for r in range(4):
            with torch.cuda.stream(stream):
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
        extra, extra_poses = tile_stream(base_pts, base_poses, sp, total, reps * B)
        d_extra = torch.from_numpy(extra.view(np.uint8).reshape(reps * B, R * 48)).cuda()
        d_extra_poses = torch.from_numpy(extra_poses).cuda()
        for r in range(reps):
            with torch.cuda.stream(stream):
                l2_flush(r)
            res = cc.addFiringsDevice(d_extra.data_ptr() + r * B * rec_bytes, d_extra_poses.data_ptr() + r * B * pose_bytes, B, R)
            ncols = int(res.info.ground_to_gcol - res.info.ground_from_gcol)
            for name, ms in cc.kernel_timings():
                a = acc.setdefault(name, [0.0, 0])
                a[0] += ms
                a[1] += 1
        cc.set_kernel_timing(False)
        tot = sum(v[0] for v in acc.values())
        kernel_table = {k: {"ms_per_step": v[0] / reps, "share": v[0] / tot} for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])}
        # the dominant kernel among those that stream the range image (k_fin_all works on the list of unfinished trees,
        # a few thousand entries: it has no per-cell traffic to put against a bandwidth roofline)
        top = next((k for k in kernel_table if k in KERNEL_BYTES_PER_CELL), next(iter(kernel_table)))
        peak, peak_src = load_peaks()
        bpc = KERNEL_BYTES_PER_CELL.get(top, 8)
        alg_bytes = bpc * B * R
        dur_s = kernel_table[top]["ms_per_step"] / 1e3
        achieved = alg_bytes / dur_s / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(HERE, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
            with open(tp) as f:
                tj = json.load(f)
            if top in tj.get("kernels", {}):
                traffic = tj["kernels"][top]["dram_bytes"]
                traffic_src = tj.get("source")
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_share_of_step": kernel_table[top]["share"],
                    "by_kernel": {k: {"algorithmic_bytes_per_launch": KERNEL_BYTES_PER_CELL[k] * B * R,
                                      "achieved_gbs": KERNEL_BYTES_PER_CELL[k] * B * R / (v["ms_per_step"] / 1e3) / 1e9,
                                      "frac": KERNEL_BYTES_PER_CELL[k] * B * R / (v["ms_per_step"] / 1e3) / 1e9 / peak}
                                  for k, v in kernel_table.items() if k in KERNEL_BYTES_PER_CELL},
                    "path_bytes_per_cell": 151, "path_achieved_gbs": value / world * R * 151 / 1e9,
                    "path_frac": value / world * R * 151 / 1e9 / peak}
    cc.close()

    # ------------------------------------------------------------------ other operating points (rank 0, single GPU)
    batch_sweep, latency_mode, flushed = None, None, None
    if rank == 0 and world == 1 and not args.quick:
        lg = device_leg(B, 10, 3, False, flush_each_step=True)
        flushed = {"columns_per_s": lg["value"], "ms_per_step": 1e3 * lg["elapsed"] / 10,
                   "per_push_device_ms_p50": float(np.median(lg["dev_ms"])),
                   "how": "256 MiB device write before every timed push, its event-timed duration subtracted"}
        lg["cc"].close()
        batch_sweep = {}
        for b2 in (1024, 2048, 6144):
            if b2 == B:
                continue
            lg = device_leg(b2, 8, 3, False)
            batch_sweep[str(b2)] = {"columns_per_s": lg["value"], "ms_per_step": 1e3 * lg["elapsed"] / 8,
                                    "per_push_device_ms_p50": float(np.median(lg["dev_ms"]))}
            lg["cc"].close()
        # latency mode: small synchronous pushes (one in flight), host buffers, results back on the host
        LB = 64
        ccl = new_handle(LB)
        ccl.set_label_prefetch(True)
        nl = 120
        lp, lq = tile_stream(base_pts, base_poses, sp, 0, nl * LB)
        pin_lp = torch.from_numpy(lp.view(np.uint8).reshape(nl * LB, R * 48)).pin_memory()
        pin_lq = torch.from_numpy(lq).pin_memory()
        hlp = pin_lp.numpy().view(lp.dtype).reshape(nl * LB, R)
        hlq = pin_lq.numpy()
        lat = []
        for i in range(nl):
            t0 = time.perf_counter()
            ccl.addFirings(hlp[i * LB:(i + 1) * LB], hlq[i * LB:(i + 1) * LB])
            lat.append(1e3 * (time.perf_counter() - t0))
        ccl.close()
        lat = np.array(lat[40:])
        latency_mode = {"batch_firings": LB, "call": "addFirings (host buffers in, events / clusters / labels back on the host)",
                        "per_push_ms_p50": float(np.median(lat)), "per_push_ms_p99": float(np.percentile(lat, 99)),
                        "columns_per_s": LB / (float(np.mean(lat)) / 1e3)}

    # ------------------------------------------------------------------ end-to-end leg through the public API
    cc = new_handle()
    cc.set_label_prefetch(True)  # the ground labels of the new columns come back with every push's results
    if args.quick:
        K_e2e = min(K, 2)
    else:
        K_e2e = K
    total = (W + K_e2e) * B
    h_pts, h_poses = tile_stream(base_pts, base_poses, sp, 0, total)
    # page-locked host buffers (the contract's "pinned host memory"): cc_push_firings copies them straight to the device
    pin_pts = torch.from_numpy(h_pts.view(np.uint8).reshape(total, R * 48)).pin_memory()
    pin_poses = torch.from_numpy(h_poses).pin_memory()
    h_pts = pin_pts.numpy().view(h_pts.dtype).reshape(total, R)
    h_poses = pin_poses.numpy()
    d2h = 0
    for s in range(W):
        cc.addFirings(h_pts[s * B:(s + 1) * B], h_poses[s * B:(s + 1) * B])
    barrier()
    t0 = time.perf_counter()
    # two pushes in flight and a third one staged: the host->device copy of push k + 2 (input stream) overlaps the
    # kernels of pushes k and k + 1; its kernels are launched by the wait() that returns push k
    for s in range(W, min(W + 2, W + K_e2e)):
        cc.submitFirings(h_pts[s * B:(s + 1) * B], h_poses[s * B:(s + 1) * B])
    for s in range(W, W + K_e2e):
        if s + 2 < W + K_e2e:
            cc.submitFirings(h_pts[(s + 2) * B:(s + 3) * B], h_poses[(s + 2) * B:(s + 3) * B])
        res = cc.wait()
        labels = cc.column_labels()  # [n_cols, rows, 4] u8: ground label, debug label, is_ignored, intensity
        assert labels.shape[0] == int(res.info.ground_to_gcol - res.info.ground_from_gcol)
        d2h += res.clusters.nbytes + res.cluster_points.nbytes + labels.nbytes + labels.shape[0] * 8 + 512
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * K_e2e * B / e2e_s
    cc.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_run(base_pts, base_poses, sp, columns_target=10_000_000, multi_threaded=True, time_budget_s=20.0)
        except Exception as e:
            cpu = cpu_reference_run(base_pts, base_poses, sp, columns_target=10_000_000, multi_threaded=False, time_budget_s=20.0)
            cpu["sample"] += f" (multi-threaded run failed: {str(e)[:80]})"
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (+f64 rigid transforms, u32 union-find)", "data": "synthetic",
            "config": {"workload": workload_name(args), "batch_firings": B, "rows": R, "columns_per_rotation": sp.num_columns,
                       "l2": f"inputs larger than L2: {(W + K) * B * (rec_bytes + pose_bytes) >> 20} MiB of firings streamed once, {B * (rec_bytes + pose_bytes) >> 20} MiB of never-touched input per step; L2 flushed (256 MiB write) once before the timed region; timed pushes run back to back. l2_flush_each_step reports the same with a flush before every push",
                       "pipelining": "two pushes in flight (cc_submit_firings_device / cc_wait); end-to-end leg: a third host push staged",
                       "streams": f"{world} independent sensor stream(s), one per GPU, no data-path collective",
                       "exact_path_pushes": exact_pushes},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": B * (rec_bytes + pose_bytes),
                    "d2h_bytes_per_step": int(d2h // K_e2e)},
            "latency": {"per_push_device_ms_p50": float(np.median(dev_ms)),
                        "per_push_sync_call_ms_p50": float(np.median(sync_ms)),
                        "note": "one push = batch_firings columns; every column of a push is charged the whole push",
                        "wall_s": t_wall},
            "roofline": roofline, "cpu_baseline": cpu, "kernels": kernel_table,
            "kernels_device_timeline_us": trace_table, "batch_sweep": batch_sweep, "latency_mode": latency_mode,
            "l2_flush_each_step": flushed,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
