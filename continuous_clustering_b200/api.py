"""Host-side mirror of the reference's `continuous_clustering::ContinuousClustering` class
(include/continuous_clustering/clustering/continuous_clustering.hpp:197-290) on top of the CUDA C ABI.

Same method names, argument meaning and error behaviour as the reference class; every stage of the pipeline runs
in CUDA kernels on one B200 (continuous_clustering_b200/csrc). Firings are handed to the device in batches
(`addFirings`, or `addFiring` + `flush`); callbacks are delivered in the order of the reference's deterministic
single-threaded mode (thread_pool.hpp:31-35).
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import _lib
from .synth import RAW_POINT_DTYPE


@dataclasses.dataclass
class GeneralConfiguration:  # hpp:24-27
    is_single_threaded: bool = False


@dataclasses.dataclass
class ContinuousRangeImageConfiguration:  # hpp:29-34
    sensor_is_clockwise: bool = True
    num_columns: int = 1700
    supplement_inclination_angle_for_nan_cells: bool = True


@dataclasses.dataclass
class ContinuousGroundSegmentationConfiguration:  # hpp:36-66
    max_slope: float = 0.2
    first_ring_as_ground_max_allowed_z_diff: float = 0.4
    first_ring_as_ground_min_allowed_z_diff: float = -0.4
    last_ground_point_slope_higher_than: float = -0.1
    last_ground_point_distance_smaller_than: float = 5.0
    ground_because_close_to_last_certain_ground_max_z_diff: float = 0.4
    ground_because_close_to_last_certain_ground_max_dist_diff: float = 2.0
    obstacle_because_next_certain_obstacle_max_dist_diff: float = 0.3
    use_terrain: bool = False
    terrain_max_allowed_z_diff: float = 0.4
    height_ref_to_maximum_: float = 0.0
    height_ref_to_ground_: float = 0.0
    length_ref_to_front_end_: float = 0.0
    length_ref_to_rear_end_: float = 0.0
    width_ref_to_left_mirror_: float = 0.0
    width_ref_to_right_mirror_: float = 0.0
    fog_filtering_enabled: bool = False
    fog_filtering_intensity_below: int = 2
    fog_filtering_distance_below: float = 18.0
    fog_filtering_inclination_above: float = -0.06


@dataclasses.dataclass
class ContinuousClusteringConfiguration:  # hpp:68-79
    max_distance: float = 0.7
    max_steps_in_row: int = 20
    max_steps_in_column: int = 20
    stop_after_association_enabled: bool = True
    stop_after_association_min_steps: int = 1
    ignore_points_in_chessboard_pattern: bool = True
    ignore_points_with_too_big_inclination_angle_diff: bool = True
    use_last_point_for_cluster_stamp: bool = False
    cluster_point_trees_every_nth_column: int = 1


@dataclasses.dataclass
class Configuration:  # hpp:81-87
    general: GeneralConfiguration = dataclasses.field(default_factory=GeneralConfiguration)
    range_image: ContinuousRangeImageConfiguration = dataclasses.field(default_factory=ContinuousRangeImageConfiguration)
    ground_segmentation: ContinuousGroundSegmentationConfiguration = dataclasses.field(
        default_factory=ContinuousGroundSegmentationConfiguration)
    clustering: ContinuousClusteringConfiguration = dataclasses.field(default_factory=ContinuousClusteringConfiguration)

    def to_c(self) -> _lib.CcConfig:
        c = _lib.CcConfig()
        for group in (self.general, self.range_image, self.ground_segmentation, self.clustering):
            for f in dataclasses.fields(group):
                v = getattr(group, f.name)
                setattr(c, f.name, int(v) if isinstance(v, (bool, np.bool_)) else v)
        return c

    @staticmethod
    def from_c(c) -> "Configuration":
        cfg = Configuration()
        for group in (cfg.general, cfg.range_image, cfg.ground_segmentation, cfg.clustering):
            for f in dataclasses.fields(group):
                v = getattr(c, f.name)
                setattr(group, f.name, bool(v) if f.type in ("bool", bool) else v)
        return cfg


class ClusteringError(RuntimeError):
    """Raised where the reference throws std::runtime_error (cpp:90-91, 298-299, 337-344, 1072-1075)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{_lib.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class _EmptyViews(dict):
    def __missing__(self, key):
        for dt in (_lib.EVENT_DTYPE, _lib.CLUSTER_DTYPE, _lib.CLUSTER_POINT_DTYPE):
            if (dt.itemsize, dt.names) == key:
                self[key] = np.zeros(0, dtype=dt)
                return self[key]
        raise KeyError(key)


_EMPTY = _EmptyViews()


@dataclasses.dataclass
class BatchResult:
    """Results of one push. The arrays are zero-copy views of the handle's buffers: valid until the next push on the
    same object (copy them to keep them)."""

    info: _lib.CcBatchInfo
    events: np.ndarray  # EVENT_DTYPE, callback order
    clusters: np.ndarray  # CLUSTER_DTYPE, finish order (> 5 points; the cluster callback needs > 20)
    cluster_points: np.ndarray  # CLUSTER_POINT_DTYPE


class ContinuousClustering:
    """One sensor stream on one GPU. Mirrors hpp:197-251 (public methods and data members)."""

    def __init__(self, device: int = 0, max_firings_per_push: int = 4096, batch_firings: int = 256, _library=None):
        self._L = _library if _library is not None else _lib.load_library()
        h = C.c_void_p()
        rc = self._L.cc_create(device, max_firings_per_push, C.byref(h))
        if rc != 0:
            raise ClusteringError(rc, "cc_create failed: a CUDA device (B200, sm_100a) is required")
        self._h = h
        self.device = device
        self.max_firings_per_push = max_firings_per_push
        self.batch_firings = min(batch_firings, max_firings_per_push)
        self._pending_pts: list[np.ndarray] = []
        self._pending_poses: list[np.ndarray] = []
        self._column_cb = None
        self._cluster_cb = None
        self.last: BatchResult | None = None

    # ---- lifecycle ----
    def close(self):
        if getattr(self, "_h", None):
            self._L.cc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise ClusteringError(rc, self._L.cc_last_error(self._h).decode(errors="replace"))

    def setConfiguration(self, config):  # hpp:206
        c = config.to_c() if isinstance(config, Configuration) else config
        self._keep_cfg = c
        self._check(self._L.cc_set_config(self._h, C.addressof(c)))

    def reset(self, num_rows: int):  # hpp:205
        self._pending_pts.clear()
        self._pending_poses.clear()
        self._check(self._L.cc_reset(self._h, int(num_rows)))

    def resetRequired(self) -> bool:  # hpp:207
        return bool(self._L.cc_reset_required(self._h))

    def setTransformRobotFrameFromSensorFrame(self, tf):  # hpp:213
        m = np.ascontiguousarray(np.asarray(tf, dtype=np.float64).reshape(-1)[:12])
        self._check(self._L.cc_set_robot_from_sensor(self._h, m.ctypes.data))

    def hasTransformRobotFrameFromSensorFrame(self) -> bool:  # hpp:214
        return bool(self._L.cc_has_robot_from_sensor(self._h))

    def setFinishedColumnCallback(self, cb):  # hpp:217: cb(from_gcol, to_gcol, ground_points_only)
        self._column_cb = cb

    def setFinishedClusterCallback(self, cb):  # hpp:218: cb(points, stamp)
        self._cluster_cb = cb

    def recordJobQueueWorkload(self, num_jobs_sensor_input: int):  # hpp:221: there are no job queues on the device
        return None

    # public data members of the reference object (hpp:244-251)
    @property
    def num_rows_(self) -> int:
        return self._L.cc_num_rows(self._h)

    @property
    def num_columns_(self) -> int:
        return self._L.cc_num_columns(self._h)

    @property
    def ring_buffer_max_columns(self) -> int:
        return self._L.cc_ring_buffer_max_columns(self._h)

    @property
    def ring_buffer_start_global_column_index(self) -> int:
        return self.last.info.ring_start_gcol if self.last else -1

    @property
    def ring_buffer_end_global_column_index(self) -> int:
        return self.last.info.ring_end_gcol if self.last else -1

    # ---- the hot path ----
    def addFiring(self, firing, odom_from_sensor):  # hpp:210
        """One firing (array of num_rows RawPoints) + 3x4 pose. Buffered; pushed every `batch_firings` firings."""
        pts = np.ascontiguousarray(firing, dtype=RAW_POINT_DTYPE).reshape(-1)
        if pts.shape[0] != self.num_rows_:
            raise ClusteringError(3, "The number of points in a firing has changed. This is probably a bug!")
        self._pending_pts.append(pts)
        self._pending_poses.append(np.asarray(odom_from_sensor, dtype=np.float64).reshape(-1)[:12])
        if len(self._pending_pts) >= self.batch_firings:
            self.flush()

    def flush(self):
        if not self._pending_pts:
            return None
        pts = np.stack(self._pending_pts)
        poses = np.stack(self._pending_poses)
        self._pending_pts.clear()
        self._pending_poses.clear()
        return self.addFirings(pts, poses)

    def addFirings(self, points: np.ndarray, poses: np.ndarray) -> BatchResult:
        """A batch of consecutive firings: points[n, num_rows] (RAW_POINT_DTYPE), poses[n, 12]."""
        if not (points.flags.c_contiguous and points.dtype == RAW_POINT_DTYPE):
            points = np.ascontiguousarray(points, dtype=RAW_POINT_DTYPE)
        if not (poses.flags.c_contiguous and poses.dtype == np.float64):
            poses = np.ascontiguousarray(poses, dtype=np.float64)
        n, rows = points.shape
        out = None
        step = max(1, int(self._L.cc_max_firings_per_push(self._h)))
        if n <= step:  # the common case: one push
            self._check(self._L.cc_push_firings(self._h, n, rows, points.ctypes.data, poses.ctypes.data))
            out = self._collect()
            self._dispatch(out)
            return out
        for a in range(0, max(n, 1), step):  # larger batches are pushed piecewise; the LAST piece's result is returned
            b = min(n, a + step)
            self._check(self._L.cc_push_firings(self._h, b - a, rows, points[a:b].ctypes.data, poses[a:b].ctypes.data))
            out = self._collect()
            self._dispatch(out)
        return out

    def addFiringsDevice(self, d_points: int, d_poses: int, n: int, rows: int) -> BatchResult:
        """Same with device pointers (cc_push_firings_device): inputs already resident in HBM."""
        self._check(self._L.cc_push_firings_device(self._h, n, rows, d_points, d_poses))
        out = self._collect()
        self._dispatch(out)
        return out

    # ---- asynchronous pushes: keep the GPU busy with submit(k + 2); wait(k) (two pushes in flight; a third host push
    #      is staged: its input copy starts at once, its kernels when wait() makes room) ----
    def submitFirings(self, points: np.ndarray, poses: np.ndarray):
        points = np.ascontiguousarray(points, dtype=RAW_POINT_DTYPE)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        n, rows = points.shape
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((points, poses))  # inputs must stay alive until waited for
        self._check(self._L.cc_submit_firings(self._h, n, rows, points.ctypes.data, poses.ctypes.data))

    def submitFiringsDevice(self, d_points: int, d_poses: int, n: int, rows: int):
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append(None)
        self._check(self._L.cc_submit_firings_device(self._h, n, rows, d_points, d_poses))

    def wait(self) -> BatchResult:
        """Finishes the oldest submitted push and returns its results (callbacks are dispatched here)."""
        self._check(self._L.cc_wait(self._h))
        if getattr(self, "_inflight", None):
            self._inflight.pop(0)
        out = self._collect()
        self._dispatch(out)
        return out

    @property
    def pending(self) -> int:
        return int(self._L.cc_pending(self._h))

    def _collect(self) -> BatchResult:
        """Results of the last push as numpy views of the handle's own arrays (valid until the next push)."""
        info = _lib.CcBatchInfo()
        self._check(self._L.cc_get_batch_info(self._h, C.byref(info)))
        pe, pc, pp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._L.cc_get_result_views(self._h, C.byref(pe), C.byref(pc), C.byref(pp))

        cache = self.__dict__.setdefault("_views", {})

        def view(ptr, n, dtype):
            # the handle hands out a few long-lived buffers (three slots): one numpy view per buffer, sliced per push
            if n == 0 or not ptr.value:
                return _EMPTY[dtype.itemsize, dtype.names]
            ent = cache.get((ptr.value, dtype.itemsize))
            if ent is None or ent.shape[0] < n:
                cap = max(n, 1024)
                buf = (C.c_char * (cap * dtype.itemsize)).from_address(ptr.value)
                ent = cache[(ptr.value, dtype.itemsize)] = np.frombuffer(buf, dtype=dtype, count=cap)
                if len(cache) > 64:
                    cache.clear()
            return ent[:n]

        self.last = BatchResult(info, view(pe, info.n_events, _lib.EVENT_DTYPE), view(pc, info.n_clusters, _lib.CLUSTER_DTYPE),
                                view(pp, info.n_cluster_points, _lib.CLUSTER_POINT_DTYPE))
        return self.last

    def _dispatch(self, res: BatchResult):
        if self._column_cb is None and self._cluster_cb is None:
            return
        cells = None
        if self._cluster_cb is not None and len(res.clusters):
            big = res.clusters[res.clusters["num_points"] > 20]  # cpp:1023
            if len(big):
                lo, hi = int(big["min_gcol"].min()), int(big["max_gcol"].max())
                cells = (lo, self.read_columns(lo, hi))
        nxt = 0
        for e in res.events:
            while nxt < int(e["n_clusters_before"]):
                c = res.clusters[nxt]
                nxt += 1
                if self._cluster_cb is not None and c["num_points"] > 20:
                    p = res.cluster_points[int(c["point_offset"]) : int(c["point_offset"]) + int(c["num_points"])]
                    pts = cells[1][p["gcol"] - cells[0], p["row"]]
                    self._cluster_cb(pts, int(c["stamp"]))
            if self._column_cb is not None:
                self._column_cb(int(e["from_gcol"]), int(e["to_gcol"]), bool(e["ground_points_only"]))

    def read_columns(self, from_gcol: int, to_gcol: int, fields=None) -> np.ndarray:
        """Cells of columns [from, to] (inclusive) as a structured array [n_cols, num_rows]: what a caller of the
        reference reads from `range_image_` inside a column callback (ros_utils.cpp:56-63)."""
        names = list(fields) if fields is not None else list(_lib.COLUMN_FIELD_DTYPES)
        ncols = max(0, to_gcol - from_gcol + 1)
        rows = self.num_rows_
        dt = np.dtype([(n, _lib.COLUMN_FIELD_DTYPES[n][0], (3,)) if _lib.COLUMN_FIELD_DTYPES[n][1] == 3
                       else (n, _lib.COLUMN_FIELD_DTYPES[n][0]) for n in names])
        out = np.zeros((ncols, rows), dtype=dt)
        if ncols == 0:
            return out
        bufs = {n: np.zeros((ncols, rows) + ((3,) if _lib.COLUMN_FIELD_DTYPES[n][1] == 3 else ()),
                            dtype=_lib.COLUMN_FIELD_DTYPES[n][0]) for n in names}
        f = _lib.CcColumnFields()
        for n in names:
            setattr(f, n, bufs[n].ctypes.data)
        self._check(self._L.cc_read_columns(self._h, from_gcol, to_gcol, C.byref(f)))
        for n in names:
            out[n] = bufs[n]
        return out

    def export_columns(self, from_gcol: int, to_gcol: int) -> np.ndarray:
        """Cells of columns [from, to] (inclusive) as packed cc_cell_t records [n_cols, num_rows], gathered by one kernel
        straight into a page-locked buffer of the handle (cc_export_columns); a VIEW, valid until the next
        export_columns / read_columns call."""
        ncols = max(0, to_gcol - from_gcol + 1)
        rows = self.num_rows_
        if ncols == 0:
            return np.zeros((0, rows), dtype=_lib.CELL_DTYPE)
        ptr = C.c_void_p()
        self._check(self._L.cc_export_columns(self._h, from_gcol, to_gcol, C.byref(ptr)))
        buf = (C.c_char * (ncols * rows * _lib.CELL_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=_lib.CELL_DTYPE, count=ncols * rows).reshape(ncols, rows)

    def _cloud(self, view):
        n = int(view.data_size)
        data = np.frombuffer((C.c_char * n).from_address(view.data), dtype=np.uint8, count=n) if n else np.zeros(0, np.uint8)
        return dict(data=data, point_step=int(view.point_step), width=int(view.width), height=int(view.height),
                    stamp_ns=int(view.stamp_ns), n_fields=int(view.n_fields))

    def pack_columns_pointcloud2(self, from_gcol: int, to_gcol: int, ground_points_only: bool):
        """The sensor_msgs/PointCloud2 payload the ROS node publishes for a finished-column callback
        (ros_utils.cpp:34-77), packed on the device. `data` is a VIEW valid until the next pack_* call."""
        v = _lib.CcCloudView()
        self._check(self._L.cc_pack_columns_pointcloud2(self._h, from_gcol, to_gcol, int(ground_points_only), C.byref(v)))
        return self._cloud(v)

    def pack_cluster_pointcloud2(self, cluster_index: int):
        """Same for cluster `cluster_index` of the last finished push (ros_utils.cpp:11-32)."""
        v = _lib.CcCloudView()
        self._check(self._L.cc_pack_cluster_pointcloud2(self._h, int(cluster_index), C.byref(v)))
        return self._cloud(v)

    def pack_requests_pointcloud2(self, requests):
        """Every message of a push with one launch (cc_pack_requests_pointcloud2). requests: [(kind, from_gcol, to_gcol)]
        with kind 0 ground-stage columns, 1 clustered columns, or (2, cluster_index, cluster_index). Returns the list of
        clouds (payloads are views valid until the next pack_* call)."""
        n = len(requests)
        if n == 0:
            return []
        req = (_lib.CcPackRequest * n)()
        for i, (kind, a, b) in enumerate(requests):
            req[i].kind, req[i].cluster_index, req[i].from_gcol, req[i].to_gcol = kind, (a if kind == 2 else 0), a, b
        views = (_lib.CcCloudView * n)()
        self._check(self._L.cc_pack_requests_pointcloud2(self._h, n, C.addressof(req), C.addressof(views)))
        return [self._cloud(views[i]) for i in range(n)]

    def set_label_prefetch(self, enable: bool):
        """Bring the labels of every push's new columns back with its results (cc_set_label_prefetch)."""
        self._check(self._L.cc_set_label_prefetch(self._h, int(enable)))

    def column_labels(self) -> np.ndarray:
        """uint8 [n_cols, num_rows, 4] view: ground_point_label, debug label, is_ignored, intensity of the columns
        [ground_from_gcol, ground_to_gcol) of the last finished push (needs set_label_prefetch(True))."""
        ptr, n = C.c_void_p(), C.c_int(0)
        self._check(self._L.cc_get_column_labels(self._h, C.byref(ptr), C.byref(n)))
        rows = self.num_rows_
        if n.value == 0 or not ptr.value:
            return np.zeros((0, rows, 4), dtype=np.uint8)
        buf = (C.c_ubyte * (n.value * rows * 4)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(n.value, rows, 4)

    def debug_flag_columns(self, period: int):
        """Test hook (cc_debug_flag_columns): force every n-th column through the exact column-sequential path."""
        self._check(self._L.cc_debug_flag_columns(self._h, int(period)))

    def set_kernel_timing(self, enable: bool):
        self._check(self._L.cc_set_kernel_timing(self._h, int(enable)))

    def kernel_timings(self):
        """[(kernel name, milliseconds)] of the last push, launch order (needs set_kernel_timing(True))."""
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        n = C.c_int(0)
        self._check(self._L.cc_get_kernel_timings(self._h, names, 4096, ms, 64, C.byref(n)))
        parts = names.value.decode().split(";")[: n.value]
        return [(parts[i], float(ms[i])) for i in range(n.value)]

    def debug_trace(self, enable: bool):
        """Device-side timeline of the kernels as they overlap in a normal push (cc_debug_trace)."""
        self._check(self._L.cc_debug_trace(self._h, int(enable)))

    def get_trace(self):
        """[(kernel, first block entry ns, last block exit ns, longest block ns, blocks)] since the last call."""
        names = C.create_string_buffer(4096)
        out = (C.c_uint64 * (4 * 64))()
        n = C.c_int(0)
        self._check(self._L.cc_debug_get_trace(self._h, names, 4096, out, 64, C.byref(n)))
        parts = names.value.decode().split(";")
        return [(parts[i], int(out[4 * i]), int(out[4 * i + 1]), int(out[4 * i + 2]), int(out[4 * i + 3]))
                for i in range(n.value) if out[4 * i + 3]]

    @property
    def total_launches(self) -> int:
        return int(self._L.cc_total_launches(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.cc_stream(self._h) or 0)


class KittiEvaluation:
    """The per-frame evaluation metrics of the reference's KittiEvaluation (kitti_evaluation.cpp:44-146) on the device:
    ground-segmentation confusion counts and over- / under-segmentation entropies."""

    def __init__(self, device: int = 0, max_points_per_frame: int = 1 << 18, _library=None):
        self._L = _library or _lib.load_library()
        h = C.c_void_p()
        rc = self._L.cc_eval_create(device, max_points_per_frame, C.byref(h))
        if rc != 0:
            raise ClusteringError(rc, "cc_eval_create failed (no CUDA device: there is no CPU path)")
        self._h = h

    def evaluate(self, semantic_label, is_ground_point, euclidean_clustering_label, detection_label) -> dict:
        sem = np.ascontiguousarray(semantic_label, dtype=np.uint16)
        gr = np.ascontiguousarray(is_ground_point, dtype=np.uint8)
        gt = np.ascontiguousarray(euclidean_clustering_label, dtype=np.uint32)
        det = np.ascontiguousarray(detection_label, dtype=np.uint32)
        assert sem.shape == gr.shape == gt.shape == det.shape
        res = _lib.CcEvalResult()
        rc = self._L.cc_eval_frame(self._h, sem.size, sem.ctypes.data, gr.ctypes.data, gt.ctypes.data, det.ctypes.data, C.byref(res))
        if rc != 0:
            raise ClusteringError(rc, "cc_eval_frame failed")
        return {k: getattr(res, k) for k, _ in _lib.CcEvalResult._fields_}

    def close(self):
        if self._h:
            self._L.cc_eval_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KittiReplay:
    """The per-frame body of the reference's kitti_demo (kitti_demo.cpp:369-403) on the device: a
    SemanticKITTI frame becomes 2200 pseudo firings of 64 rows with their interpolated poses, resident in device memory
    (KittiLoader::recoverLaserIndices / undoEgoMotionCorrection / generateRangeImage / interpolate, kitti_loader.cpp:47-210,
    297-328; makePseudoFiringFromRangeImageColumn, kitti_demo.cpp:123-159)."""

    WIDTH, HEIGHT = 2200, 64

    def __init__(self, device: int = 0, max_points_per_frame: int = 1 << 18, _library=None):
        self._L = _library or _lib.load_library()
        h = C.c_void_p()
        rc = self._L.cc_kitti_create(device, max_points_per_frame, C.byref(h))
        if rc != 0:
            raise ClusteringError(rc, "cc_kitti_create failed (no CUDA device: there is no CPU path)")
        self._h = h
        self._n = 0

    def set_poses(self, stamps, poses12) -> None:
        """odom_from_velodyne of the sequence: stamps [n] uint64 ascending, poses [n][12] (3x4 row major)."""
        st = np.ascontiguousarray(stamps, dtype=np.uint64)
        ps = np.ascontiguousarray(poses12, dtype=np.float64).reshape(-1, 12)
        assert st.size == ps.shape[0]
        rc = self._L.cc_kitti_set_poses(self._h, st.size, st.ctypes.data, ps.ctypes.data)
        if rc != 0:
            raise ClusteringError(rc, "cc_kitti_set_poses failed")

    def frame(self, xyzi, stamp_start: int, stamp_end: int, frame_pose12, sequence_index: int, frame_index: int) -> dict:
        """Returns the device pointers of the frame's firings / poses (feed them to
        ContinuousClustering.push_device in chunks), the host copy of the poses and the reference's sanity values."""
        pts = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
        pose = np.ascontiguousarray(frame_pose12, dtype=np.float64).reshape(12)
        out = _lib.CcKittiFrame()
        rc = self._L.cc_kitti_frame(self._h, pts.shape[0], pts.ctypes.data, int(stamp_start), int(stamp_end), pose.ctypes.data,
                                    int(sequence_index), int(frame_index), C.byref(out))
        if rc != 0:
            raise ClusteringError(rc, f"cc_kitti_frame failed (longest row: {out.max_points_in_row} points)")
        self._n = pts.shape[0]
        poses = np.ctypeslib.as_array(C.cast(out.poses, C.POINTER(C.c_double)), shape=(out.n_firings, 12)).copy()
        return {"n_firings": out.n_firings, "rows_per_firing": out.rows_per_firing, "d_firings": out.d_firings,
                "d_poses": out.d_poses, "poses": poses, "rows_found": out.rows_found, "max_points_in_row": out.max_points_in_row}

    def read_debug(self) -> dict:
        """Intermediate results of the last frame (tests): laser indices, range-image cells, un-corrected points, firings."""
        laser = np.zeros(self._n, np.uint8)
        cells = np.zeros(self.WIDTH * self.HEIGHT, np.int32)
        unc = np.zeros((self._n, 3), np.float32)
        from .synth import RAW_POINT_DTYPE

        firings = np.zeros(self.WIDTH * self.HEIGHT, dtype=RAW_POINT_DTYPE)
        rc = self._L.cc_kitti_read_debug(self._h, laser.ctypes.data, cells.ctypes.data, unc.ctypes.data, firings.ctypes.data)
        if rc != 0:
            raise ClusteringError(rc, "cc_kitti_read_debug failed")
        return {"laser_index": laser, "cell_point": cells.reshape(self.HEIGHT, self.WIDTH), "uncorrected": unc,
                "firings": firings.reshape(self.WIDTH, self.HEIGHT)}

    def close(self):
        if self._h:
            self._L.cc_kitti_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OusterInput:
    """OusterInput::onRawDataArrived (ouster_input.hpp:105-181) for batches of lidar packets on the device:
    every valid measurement block becomes one firing of RawPoints in device memory."""

    def __init__(self, rows: int, columns_per_frame: int, direction, offset, device: int = 0, max_packets_per_call: int = 256,
                 fmt: "_lib.CcOusterFormat | None" = None, _library=None):
        self._L = _library or _lib.load_library()
        if fmt is None:
            fmt = _lib.CcOusterFormat()
            self._L.cc_ouster_format_legacy(rows, columns_per_frame, C.byref(fmt))
        self.format = fmt
        h = C.c_void_p()
        rc = self._L.cc_ouster_create(device, C.byref(fmt), max_packets_per_call, C.byref(h))
        if rc != 0:
            raise ClusteringError(rc, "cc_ouster_create failed (no CUDA device: there is no CPU path)")
        self._h = h
        d = np.ascontiguousarray(direction, dtype=np.float32)
        o = np.ascontiguousarray(offset, dtype=np.float32)
        assert d.shape == o.shape == (columns_per_frame * rows, 3)
        if self._L.cc_ouster_set_lut(h, d.ctypes.data, o.ctypes.data) != 0:
            raise ClusteringError(2, "cc_ouster_set_lut failed")
        self.packet_size = int(self._L.cc_ouster_packet_size(h))

    def reset(self):
        self._L.cc_ouster_reset(self._h)

    def decode(self, packets, receive_stamps) -> dict:
        pk = np.ascontiguousarray(packets, dtype=np.uint8).reshape(-1, self.packet_size)
        st = np.ascontiguousarray(receive_stamps, dtype=np.uint64)
        assert st.size == pk.shape[0]
        out = _lib.CcDecodedFirings()
        rc = self._L.cc_ouster_decode(self._h, pk.shape[0], pk.ctypes.data, st.ctypes.data, C.byref(out))
        if rc != 0:
            raise ClusteringError(rc, "cc_ouster_decode failed")
        stamps = (np.ctypeslib.as_array(C.cast(out.firing_stamps, C.POINTER(C.c_uint64)), shape=(out.n_firings,)).copy()
                  if out.n_firings else np.zeros(0, np.uint64))
        return {"n_firings": out.n_firings, "rows_per_firing": out.rows_per_firing, "d_firings": out.d_firings,
                "firing_stamps": stamps, "first_firing_index": out.first_firing_index}

    def read_firings(self, n_firings: int) -> np.ndarray:
        from .synth import RAW_POINT_DTYPE

        out = np.zeros((n_firings, self.format.pixels_per_column), dtype=RAW_POINT_DTYPE)
        if self._L.cc_ouster_read_firings(self._h, n_firings, out.ctypes.data) != 0:
            raise ClusteringError(2, "cc_ouster_read_firings failed")
        return out

    def close(self):
        if self._h:
            self._L.cc_ouster_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
