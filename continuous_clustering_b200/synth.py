"""Seeded synthetic LiDAR firing streams.

There is no sensor data in this environment, so every parity test and benchmark runs on synthetic
Velodyne-/Ouster-like streams: a ground plane plus axis-aligned boxes, ray-cast per laser, emitted as
one firing (= one range-image column worth of RawPoints) at a time -- the same shape of input that
`ContinuousClustering::addFiring` receives from the reference's sensor adapters
(velodyne_input.hpp:46-91) or from kitti_demo's pseudo firings (kitti_demo.cpp:123-159).

Pure numpy; no dependency on the CUDA library or on anything under oracle/.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# Layout-identical to cc_raw_point_t / continuous_clustering::RawPoint (point_types.hpp:10-19): 48 bytes.
RAW_POINT_DTYPE = np.dtype(
    {
        "names": ["x", "y", "z", "firing_index", "intensity", "stamp", "globally_unique_point_index"],
        "formats": ["<f4", "<f4", "<f4", "<u8", "u1", "<u8", "<u8"],
        "offsets": [0, 4, 8, 16, 24, 32, 40],
        "itemsize": 48,
    }
)

# Ouster OS-32 beam tables of the reference's two Touareg sensors, in degrees, top beam first
# (sensor calibration DATA quoted from calibrations/touareg_os32_left.json / touareg_os32_right.json,
# keys "beam_altitude_angles" and "beam_azimuth_angles"; lidar_mode 1024x10).
_OS32_ALTITUDE_DEG = {
    "left": [46.09, 42.99, 39.96, 36.98, 34.02, 31.07, 28.14, 25.2, 22.29, 19.4, 16.51, 13.64, 10.78, 7.95, 5.11,
             2.3, -0.52, -3.35, -6.16, -8.99, -11.82, -14.66, -17.5, -20.36, -23.23, -26.11, -29.01, -31.93, -34.85,
             -37.79, -40.76, -43.75],
    "right": [43.54, 40.56, 37.6, 34.63, 31.7, 28.78, 25.88, 22.99, 20.11, 17.24, 14.4, 11.55, 8.72, 5.89, 3.08,
              0.26, -2.56, -5.37, -8.2, -11.04, -13.88, -16.74, -19.61, -22.5, -25.39, -28.3, -31.23, -34.16, -37.13,
              -40.12, -43.15, -46.19],
}
_OS32_AZIMUTH_DEG = {
    "left": [11.42, 10.93, 10.54, 10.19, 9.89, 9.63, 9.42, 9.2, 9.04, 8.89, 8.76, 8.66, 8.56, 8.5, 8.43, 8.39, 8.37,
             8.35, 8.36, 8.38, 8.42, 8.46, 8.52, 8.6, 8.7, 8.81, 8.95, 9.12, 9.32, 9.56, 9.84, 10.17],
    "right": [-10.11, -9.81, -9.53, -9.3, -9.1, -8.93, -8.8, -8.68, -8.58, -8.51, -8.45, -8.4, -8.38, -8.35, -8.36,
              -8.37, -8.4, -8.44, -8.49, -8.58, -8.66, -8.77, -8.88, -9.05, -9.2, -9.39, -9.63, -9.87, -10.17, -10.52,
              -10.93, -11.38],
}


@dataclasses.dataclass
class StreamSpec:
    """One of the BASELINE.json workload shapes."""

    name: str
    rows: int
    num_columns: int
    rotation_hz: float
    sensor_height: float
    inclinations_rad: np.ndarray  # row 0 = top laser (velodyne_input.hpp:55)
    azimuth_offsets_rad: np.ndarray  # per-row azimuth offset inside a firing
    mount_roll_rad: float = 0.0


def spec(name: str) -> StreamSpec:
    if name == "velodyne64":  # BASELINE configs 2 and 5: 64 rings, 2048 columns, 10 Hz
        inc = np.deg2rad(np.linspace(2.0, -24.8, 64))
        return StreamSpec(name, 64, 2048, 10.0, 1.73, inc, np.zeros(64))
    if name == "kitti64":  # config 1 stand-in: 64 x 2200 pseudo firings (kitti_demo.cpp:281)
        inc = np.deg2rad(np.linspace(2.0, -24.8, 64))
        return StreamSpec(name, 64, 2200, 10.0, 1.73, inc, np.zeros(64))
    if name == "vls128":  # config 3: 128 rings, 1700 columns (sensor_vls128_roof.launch:22), 20 Hz
        inc = np.deg2rad(np.linspace(15.0, -25.0, 128))
        # VLS-128 fires 8 laser groups with a few degrees of azimuth offset between them
        off = np.deg2rad(np.tile(np.array([-6.354, -4.548, -2.732, -0.911, 0.911, 2.732, 4.548, 6.354]), 16))
        return StreamSpec(name, 128, 1700, 20.0, 1.9, inc, off)
    if name in ("os32", "os32_left", "os32_right"):
        # config 4: Ouster OS-32, 1024 columns, 10 Hz. The mount angle is not in the reference repo; we roll
        # the sensor by +-30 degrees about x (chosen).
        side = "right" if name.endswith("right") else "left"
        roll = np.deg2rad(-30.0 if side == "right" else 30.0)
        return StreamSpec(
            name, 32, 1024, 10.0, 1.6, np.deg2rad(_OS32_ALTITUDE_DEG[side]), np.deg2rad(_OS32_AZIMUTH_DEG[side]), roll
        )
    if name == "tiny16":  # small case for fast CPU tests
        inc = np.deg2rad(np.linspace(2.0, -24.8, 16))
        return StreamSpec(name, 16, 256, 10.0, 1.73, inc, np.zeros(16))
    raise ValueError(f"unknown stream spec {name!r}")


def make_scene(seed: int, n_boxes: int = 150, extent: float = 60.0, min_dist: float = 5.0,
               height_range=(0.5, 3.0)):
    """150 axis-aligned boxes standing on the ground plane, none closer than 5 m to the origin."""
    rng = np.random.RandomState(seed)  # MT19937
    boxes = []
    while len(boxes) < n_boxes:
        cx, cy = rng.uniform(-extent, extent, 2)
        if np.hypot(cx, cy) < min_dist:
            continue
        hx, hy = rng.uniform(0.3, 2.5, 2)
        height = rng.uniform(*height_range)
        boxes.append((cx - hx, cx + hx, cy - hy, cy + hy, 0.0, height))
    return np.asarray(boxes, dtype=np.float64).reshape(-1, 6)  # z relative to the ground plane


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def make_stream(
    spec_name: str = "velodyne64",
    n_firings: int | None = None,
    n_rotations: float = 2.0,
    seed: int = 1234,
    moving: bool = False,
    start_firing: int = 0,
    max_range: float = 120.0,
    range_noise: float = 0.01,
    n_boxes: int = 150,
    dropout: float = 0.0,
    chunk: int = 128,
    extent: float = 60.0,
    min_box_dist: float = 5.0,
    box_height_range=(0.5, 3.0),
    wall_radius: float | None = None,
    wall_height: float = 2.5,
    az_jitter: float = 0.0,
    az_step_scale: float = 1.0,
):
    """Returns (points[n_firings, rows] of RAW_POINT_DTYPE, poses[n_firings, 12] float64, spec).

    Firing k has sensor-frame azimuth pi - ((k mod N) + 0.5) * 2pi/N (+ per-row offset): a clockwise
    sensor (sensor_is_clockwise=true) starting just after the negative x axis, so the first firing does
    not straddle it (cpp:252-261). Stamps are t0 + k * T_rot / N nanoseconds; intensity 100;
    globally_unique_point_index = k * rows + row. `moving` drives the sensor at 10 m/s with 0.2 rad/s yaw
    so the double-precision rigid transform (cpp:129-138) is exercised; otherwise the pose is identity.
    `dropout` randomly replaces that fraction of returns by NaN (missing returns). `wall_radius` adds a closed
    cylindrical wall around the sensor start position (a cluster that spans a full rotation: the reference's forced
    finish, cpp:909-919); `min_box_dist` / `box_height_range` / `extent` shape the box scene (tall, close boxes
    give steep inclinations: associations the reference refuses, cpp:654-659). `az_jitter` (in column widths) adds
    per-firing azimuth noise and `az_step_scale` stretches / shrinks the azimuth step per firing, so that firings land
    in the same column twice or skip columns: the cell-collision rule (cpp:188-208) and the "too far behind" cut.
    """
    sp = spec(spec_name)
    rows, ncols = sp.rows, sp.num_columns
    if n_firings is None:
        n_firings = int(round(n_rotations * ncols))
    boxes = make_scene(seed, n_boxes, extent, min_box_dist, box_height_range)
    rng = np.random.RandomState(seed + 1)
    t_rot_ns = 1e9 / sp.rotation_hz
    t0 = 1_000_000_000

    pts = np.zeros((n_firings, rows), dtype=RAW_POINT_DTYPE)
    poses = np.zeros((n_firings, 12), dtype=np.float64)
    mount = _rot_x(sp.mount_roll_rad)
    h = sp.sensor_height

    for c0 in range(0, n_firings, chunk):
        c1 = min(n_firings, c0 + chunk)
        k = np.arange(c0, c1) + start_firing
        f = k.shape[0]
        t_s = k * (t_rot_ns / ncols) * 1e-9
        # sensor pose in odom: R (f,3,3), t (f,3)
        if moving:
            yaw = 0.2 * t_s
            speed = 10.0
            # integrate a circular arc analytically
            tx = speed / 0.2 * np.sin(yaw)
            ty = speed / 0.2 * (1.0 - np.cos(yaw))
            rmat = np.stack([_rot_z(a) @ mount for a in yaw])
            tvec = np.stack([tx, ty, np.zeros(f)], axis=1)
        else:
            rmat = np.broadcast_to(mount, (f, 3, 3)).copy()
            tvec = np.zeros((f, 3))
        poses[c0:c1, 0:3] = rmat[:, 0, :]
        poses[c0:c1, 3] = tvec[:, 0]
        poses[c0:c1, 4:7] = rmat[:, 1, :]
        poses[c0:c1, 7] = tvec[:, 1]
        poses[c0:c1, 8:11] = rmat[:, 2, :]
        poses[c0:c1, 11] = tvec[:, 2]

        az = np.pi - (((k * az_step_scale) % ncols) + 0.5) * (2 * np.pi / ncols)  # (f,)
        if az_jitter > 0:
            az = az + rng.normal(0.0, az_jitter * 2 * np.pi / ncols, size=az.shape)
            az = (az + np.pi) % (2 * np.pi) - np.pi
        az = az[:, None] + sp.azimuth_offsets_rad[None, :]  # (f, rows)
        inc = sp.inclinations_rad[None, :]
        d_s = np.stack([np.cos(az) * np.cos(inc), np.sin(az) * np.cos(inc), np.broadcast_to(np.sin(inc), az.shape)], -1)
        d_w = np.einsum("fij,frj->fri", rmat, d_s)  # world direction
        o_w = tvec[:, None, :]  # sensor origin; ground plane is at z = -h in odom

        best = np.full((f, rows), max_range)
        # ground plane
        dz = d_w[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = (-h - o_w[..., 2]) / dz
        hit = (dz < -1e-9) & (tg > 0) & (tg < best)
        best = np.where(hit, tg, best)
        # boxes (slab test), z of boxes is relative to the ground plane
        lo = np.stack([boxes[:, 0], boxes[:, 2], boxes[:, 4] - h], 1)  # (b,3)
        hi = np.stack([boxes[:, 1], boxes[:, 3], boxes[:, 5] - h], 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d_w  # (f,r,3)
            ta = (lo[None, None] - o_w[:, :, None, :]) * inv[:, :, None, :]  # (f,r,b,3)
            tb = (hi[None, None] - o_w[:, :, None, :]) * inv[:, :, None, :]
        if boxes.shape[0]:
            tmin = np.nanmax(np.minimum(ta, tb), axis=-1)
            tmax = np.nanmin(np.maximum(ta, tb), axis=-1)
            ok = (tmax >= tmin) & (tmin > 0.5)
            tbox = np.where(ok, tmin, np.inf).min(axis=-1)
            best = np.minimum(best, tbox)
        if wall_radius is not None:
            # vertical cylinder x^2 + y^2 = r^2 around the odom origin, from the ground up to wall_height
            dx, dy = d_w[..., 0], d_w[..., 1]
            ox, oy = o_w[..., 0], o_w[..., 1]
            a = dx * dx + dy * dy
            b = 2 * (ox * dx + oy * dy)
            c = ox * ox + oy * oy - wall_radius**2
            with np.errstate(divide="ignore", invalid="ignore"):
                disc = b * b - 4 * a * c
                tw = (-b + np.sqrt(disc)) / (2 * a)
            zhit = o_w[..., 2] + tw * d_w[..., 2]
            okw = (disc > 0) & (tw > 0.5) & (zhit > -h) & (zhit < -h + wall_height)
            best = np.minimum(best, np.where(okw, tw, np.inf))

        valid = best < max_range
        rng_noise = rng.normal(0.0, range_noise, size=best.shape)
        rr = best + rng_noise
        if dropout > 0:
            valid &= rng.uniform(size=best.shape) >= dropout
        p = d_s * rr[..., None]
        p[~valid] = np.nan
        blk = pts[c0:c1]
        blk["x"] = p[..., 0].astype(np.float32)
        blk["y"] = p[..., 1].astype(np.float32)
        blk["z"] = p[..., 2].astype(np.float32)
        stamp = (t0 + k * (t_rot_ns / ncols)).astype(np.uint64)
        blk["stamp"] = stamp[:, None]
        blk["firing_index"] = k.astype(np.uint64)[:, None]
        blk["intensity"] = 100
        blk["globally_unique_point_index"] = (k.astype(np.uint64)[:, None] * np.uint64(rows)) + np.arange(
            rows, dtype=np.uint64
        )[None, :]
    return pts, poses, sp


def make_kitti_frame(seed: int = 7, frame_index: int = 3, n_poses: int = 8, dropout: float = 0.05, n_boxes: int = 150,
                     top_rows_empty: int = 0):
    """A synthetic frame in the layout of a SemanticKITTI velodyne .bin file, for the replay front-end (SURVEY 8f-2;
    there is no dataset in the image): the returns of one rotation of the 64-ring stream above, row after row (top laser
    first), NaN returns omitted, every row ordered by atan2 azimuth 0 -> pi -> -pi -> 0 (kitti_loader.cpp:49-54),
    ego-motion corrected to the pose at the middle of the rotation, intensity in [0, 1).
    Returns (xyzi [n, 4] float32, stamp_start, stamp_end, pose_stamps [n_poses] uint64, poses [n_poses, 12],
    frame_pose [12]) with poses = odom_from_velodyne of a sensor driving an arc, one per rotation."""
    sp = spec("velodyne64")
    n = sp.num_columns
    pts, fposes, _ = make_stream("velodyne64", n_firings=n, seed=seed, moving=True, start_firing=frame_index * n, dropout=dropout,
                                 n_boxes=n_boxes)
    rng = np.random.RandomState(seed + 77)
    t_rot = 1e9 / sp.rotation_hz
    t0 = 1_000_000_000
    stamp_start = int(t0 + frame_index * t_rot)
    stamp_end = int(stamp_start + t_rot)

    def pose_at(t_ns):
        t_s = t_ns * 1e-9
        yaw = 0.2 * t_s
        m = np.zeros(12)
        r = _rot_z(yaw)
        m[0:3], m[4:7], m[8:11] = r[0], r[1], r[2]
        m[3] = 10.0 / 0.2 * np.sin(yaw)
        m[7] = 10.0 / 0.2 * (1.0 - np.cos(yaw))
        return m

    first = frame_index - n_poses // 2
    pose_stamps = np.array([int(t0 + (first + i + 0.5) * t_rot) - t0 for i in range(n_poses)], dtype=np.int64)
    poses = np.stack([pose_at(s) for s in pose_stamps])
    pose_stamps = (pose_stamps + t0).astype(np.uint64)
    mid = pose_at(stamp_start - t0 + 0.5 * t_rot)
    # ego-motion correction: sensor frame at the firing -> odom -> sensor frame at the middle of the rotation
    rm = mid.reshape(3, 4)
    xyz = np.stack([pts["x"], pts["y"], pts["z"]], -1).astype(np.float64)  # (firing, row, 3)
    fp = fposes.reshape(-1, 3, 4)
    odom = np.einsum("fij,frj->fri", fp[:, :, :3], xyz) + fp[:, None, :, 3]
    corr = np.einsum("ji,frj->fri", rm[:, :3], odom - rm[:, 3])
    out = []
    for row in range(sp.rows):
        p = corr[:, row, :]
        p = p[~np.isnan(p[:, 0])]
        if row < top_rows_empty:
            continue
        az = np.arctan2(p[:, 1], p[:, 0])
        order = np.argsort(np.where(az < 0, az + 2 * np.pi, az), kind="stable")
        out.append(p[order])
    xyz = np.concatenate(out).astype(np.float32)
    inten = (rng.randint(0, 100, size=xyz.shape[0]) / 100.0).astype(np.float32)
    xyzi = np.concatenate([xyz, inten[:, None]], axis=1).astype(np.float32)
    return np.ascontiguousarray(xyzi), stamp_start, stamp_end, pose_stamps, poses, mid


def ouster_xyz_lut(side: str = "left", columns_per_frame: int = 1024, origin_to_beam_mm: float = 27.67,
                   lidar_to_sensor=(-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 36.18, 0, 0, 0, 1), range_unit: float = 0.001):
    """The sensor's lookup table the way OusterInput builds it (ouster_input.hpp:66-95): ouster::make_xyz_lut for the
    beam tables of calibrations/touareg_os32_<side>.json (published geometry: encoder angle 2 pi (1 - col / W), beam azimuth
    / altitude, beam origin offset, lidar-to-sensor transform, millimetres -> metres), cast to float and reordered so
    that the pixels of one measurement block are consecutive: ([W * H, 3] direction, [W * H, 3] offset).
    Test / bench scaffolding: a deployment takes the tables from the SDK."""
    alt = np.deg2rad(np.asarray(_OS32_ALTITUDE_DEG[side], dtype=np.float64))
    azi = -np.deg2rad(np.asarray(_OS32_AZIMUTH_DEG[side], dtype=np.float64))
    w, h = columns_per_frame, alt.shape[0]
    enc = 2.0 * np.pi * (1.0 - np.arange(w, dtype=np.float64) / w)  # (w,)
    e, a, t = enc[:, None], azi[None, :], alt[None, :]
    direction = np.stack([np.cos(e + a) * np.cos(t), np.sin(e + a) * np.cos(t), np.broadcast_to(np.sin(t), (w, h))], -1)
    offset = np.stack([(np.cos(e) - direction[..., 0]) * origin_to_beam_mm, (np.sin(e) - direction[..., 1]) * origin_to_beam_mm,
                       -direction[..., 2] * origin_to_beam_mm], -1)
    m = np.asarray(lidar_to_sensor, dtype=np.float64).reshape(4, 4)
    direction = direction @ m[:3, :3].T
    offset = offset @ m[:3, :3].T + m[:3, 3]
    direction = (direction * range_unit).astype(np.float32).reshape(w * h, 3)
    offset = (offset * range_unit).astype(np.float32).reshape(w * h, 3)
    return np.ascontiguousarray(direction), np.ascontiguousarray(offset)


def make_ouster_packets(n_packets: int, rows: int = 32, columns_per_frame: int = 1024, seed: int = 5, first_measurement_id: int = 0,
                        p_invalid: float = 0.03, p_no_return: float = 0.2, ranges_mm=None):
    """LEGACY-profile lidar packets (16 measurement blocks of 16 + 12 * rows + 4 bytes) with random ranges (20 bits, a
    share of zeros = no return), signal photons 0..3000 and a few invalid blocks; measurement ids run through the frame.
    Returns (packets [n_packets, packet_size] uint8, receive_stamps [n_packets] uint64)."""
    rng = np.random.RandomState(seed)
    col_size = 16 + 12 * rows + 4
    packets = np.zeros((n_packets, 16, col_size), dtype=np.uint8)
    m_id = (first_measurement_id + np.arange(n_packets * 16)) % columns_per_frame
    hdr = packets[:, :, :16].reshape(-1, 16)
    hdr[:, 0:8] = (np.arange(n_packets * 16, dtype=np.uint64) * 97656 + 1_000_000).view(np.uint8).reshape(-1, 8)
    hdr[:, 8:10] = m_id.astype("<u2").view(np.uint8).reshape(-1, 2)
    hdr[:, 10:12] = ((first_measurement_id + np.arange(n_packets * 16)) // columns_per_frame).astype("<u2").view(np.uint8).reshape(-1, 2)
    hdr[:, 12:16] = (m_id * (90112 // columns_per_frame)).astype("<u4").view(np.uint8).reshape(-1, 4)
    px = packets[:, :, 16:16 + 12 * rows].reshape(-1, rows, 12)
    if ranges_mm is not None:  # a coherent scene: [n_packets * 16, rows] millimetres, 0 = no return
        rng_mm = np.ascontiguousarray(ranges_mm, dtype="<u4").reshape(n_packets * 16, rows).copy()
    else:
        rng_mm = rng.randint(300, 120000, size=(n_packets * 16, rows)).astype("<u4")
        rng_mm[rng.uniform(size=rng_mm.shape) < p_no_return] = 0
    flags = rng.randint(0, 16, size=rng_mm.shape).astype("<u4") << 28  # the top bits of the range word are not range
    px[:, :, 0:4] = (rng_mm | flags).view(np.uint8).reshape(-1, rows, 4)
    px[:, :, 4:6] = rng.randint(0, 65536, size=rng_mm.shape).astype("<u2").view(np.uint8).reshape(-1, rows, 2)
    px[:, :, 6:8] = rng.randint(0, 3000, size=rng_mm.shape).astype("<u2").view(np.uint8).reshape(-1, rows, 2)
    px[:, :, 8:10] = rng.randint(0, 65536, size=rng_mm.shape).astype("<u2").view(np.uint8).reshape(-1, rows, 2)
    status = np.where(rng.uniform(size=n_packets * 16) < p_invalid, 0, 0xFFFFFFFF).astype("<u4")
    packets[:, :, 16 + 12 * rows:].reshape(-1, 4)[:] = status.view(np.uint8).reshape(-1, 4)
    stamps = (2_000_000_000 + np.arange(n_packets, dtype=np.uint64) * np.uint64(1_562_500)).astype(np.uint64)
    return np.ascontiguousarray(packets.reshape(n_packets, 16 * col_size)), stamps
