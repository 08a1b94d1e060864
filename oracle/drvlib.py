"""ctypes loader for the recording driver libraries (TEST INFRASTRUCTURE, see cc_driver.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (continuous_clustering_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

REF_LIB = os.path.join(HERE, "_ref", "libcc_ref.so")  # the reference's own sources (built where /root/reference exists)
# the same driver + the reference's own CALLER code (ros_utils.cpp / kitti_demo.cpp excerpts) against the facade
FACADE_CALLERS_LIB = os.path.join(HERE, "_ref", "libcc_facade_callers.so")  # facade + CUDA library
FACADE_CALLERS_EMU_LIB = os.path.join(HERE, "_ref", "libcc_facade_callers_emu_test.so")  # facade + CPU emulation (tests)
ORACLE_LIB = os.path.join(HERE, "libcc_oracle.so")  # the CPU restatement
FACADE_LIB = os.path.join(REPO, "build", "libcc_facade_driver.so")  # facade + CUDA library


class CcConfig(C.Structure):
    """cc_config_t (include/cc_b200.h), mirror of Configuration hpp:24-87."""

    _fields_ = [
        ("is_single_threaded", C.c_int32),
        ("sensor_is_clockwise", C.c_int32),
        ("num_columns", C.c_int32),
        ("supplement_inclination_angle_for_nan_cells", C.c_int32),
        ("max_slope", C.c_float),
        ("first_ring_as_ground_max_allowed_z_diff", C.c_float),
        ("first_ring_as_ground_min_allowed_z_diff", C.c_float),
        ("last_ground_point_slope_higher_than", C.c_float),
        ("last_ground_point_distance_smaller_than", C.c_float),
        ("ground_because_close_to_last_certain_ground_max_z_diff", C.c_float),
        ("ground_because_close_to_last_certain_ground_max_dist_diff", C.c_float),
        ("obstacle_because_next_certain_obstacle_max_dist_diff", C.c_float),
        ("use_terrain", C.c_int32),
        ("terrain_max_allowed_z_diff", C.c_float),
        ("height_ref_to_maximum_", C.c_float),
        ("height_ref_to_ground_", C.c_float),
        ("length_ref_to_front_end_", C.c_float),
        ("length_ref_to_rear_end_", C.c_float),
        ("width_ref_to_left_mirror_", C.c_float),
        ("width_ref_to_right_mirror_", C.c_float),
        ("fog_filtering_enabled", C.c_int32),
        ("fog_filtering_intensity_below", C.c_int32),
        ("fog_filtering_distance_below", C.c_float),
        ("fog_filtering_inclination_above", C.c_float),
        ("max_distance", C.c_float),
        ("max_steps_in_row", C.c_int32),
        ("max_steps_in_column", C.c_int32),
        ("stop_after_association_enabled", C.c_int32),
        ("stop_after_association_min_steps", C.c_int32),
        ("ignore_points_in_chessboard_pattern", C.c_int32),
        ("ignore_points_with_too_big_inclination_angle_diff", C.c_int32),
        ("use_last_point_for_cluster_stamp", C.c_int32),
        ("cluster_point_trees_every_nth_column", C.c_int32),
    ]


def default_config(**overrides) -> CcConfig:
    """Defaults of hpp:24-87 (NOT the dynamic_reconfigure defaults of cfg/ContinuousClustering.cfg)."""
    c = CcConfig()
    c.is_single_threaded = 0
    c.sensor_is_clockwise = 1
    c.num_columns = 1700
    c.supplement_inclination_angle_for_nan_cells = 1
    c.max_slope = 0.2
    c.first_ring_as_ground_max_allowed_z_diff = 0.4
    c.first_ring_as_ground_min_allowed_z_diff = -0.4
    c.last_ground_point_slope_higher_than = -0.1
    c.last_ground_point_distance_smaller_than = 5.0
    c.ground_because_close_to_last_certain_ground_max_z_diff = 0.4
    c.ground_because_close_to_last_certain_ground_max_dist_diff = 2.0
    c.obstacle_because_next_certain_obstacle_max_dist_diff = 0.3
    c.use_terrain = 0
    c.terrain_max_allowed_z_diff = 0.4
    c.fog_filtering_enabled = 0
    c.fog_filtering_intensity_below = 2
    c.fog_filtering_distance_below = 18.0
    c.fog_filtering_inclination_above = -0.06
    c.max_distance = 0.7
    c.max_steps_in_row = 20
    c.max_steps_in_column = 20
    c.stop_after_association_enabled = 1
    c.stop_after_association_min_steps = 1
    c.ignore_points_in_chessboard_pattern = 1
    c.ignore_points_with_too_big_inclination_angle_diff = 1
    c.use_last_point_for_cluster_stamp = 0
    c.cluster_point_trees_every_nth_column = 1
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def stream_config(spec_name: str, **overrides) -> CcConfig:
    """Per-workload configuration: the reference's struct defaults + the ego box of
    kitti_demo.cpp:286-291 + the launch-file overrides of the named sensor."""
    from continuous_clustering_b200 import synth

    sp = synth.spec(spec_name)
    kw = dict(
        is_single_threaded=1,
        num_columns=sp.num_columns,
        height_ref_to_maximum_=0.5,
        height_ref_to_ground_=-sp.sensor_height,
        length_ref_to_front_end_=3.0,
        length_ref_to_rear_end_=-3.0,
        width_ref_to_left_mirror_=1.5,
        width_ref_to_right_mirror_=-1.5,
    )
    if spec_name == "kitti64":  # kitti_demo.cpp:279-284
        kw.update(ignore_points_in_chessboard_pattern=0, max_distance=0.5)
    if spec_name.startswith("os32"):  # sensor_os32_left.launch:18-27
        kw.update(
            fog_filtering_intensity_below=3,
            fog_filtering_distance_below=5.0,
            fog_filtering_inclination_above=-0.17,
            ignore_points_in_chessboard_pattern=0,
            ignore_points_with_too_big_inclination_angle_diff=0,
        )
    kw.update(overrides)
    return default_config(**kw)


EVENT_DTYPE = np.dtype(
    [("from_gcol", "<i8"), ("to_gcol", "<i8"), ("ground_points_only", "<i4"), ("n_clusters_before", "<i4")]
)

CELL_DTYPE = np.dtype(
    [
        ("continuous_azimuth_angle", "<f8"),
        ("global_column_index", "<i8"),
        ("globally_unique_point_index", "<u8"),
        ("stamp", "<u8"),
        ("firing_index", "<u8"),
        ("id", "<u8"),
        ("tree_root_gcol", "<i8"),
        ("x", "<f4"),
        ("y", "<f4"),
        ("z", "<f4"),
        ("distance", "<f4"),
        ("azimuth_angle", "<f4"),
        ("inclination_angle", "<f4"),
        ("tree_root_row", "<i4"),
        ("number_of_visited_neighbors", "<i4"),
        ("intensity", "u1"),
        ("ground_point_label", "u1"),
        ("debug_ground_point_label", "u1"),
        ("is_ignored", "u1"),
        ("num_child_points", "<i4"),
        ("finished_at_continuous_azimuth_angle", "<f8"),
        ("tree_num_points", "<u4"),
        ("cluster_width", "<u4"),
        ("local_column_index", "<i4"),
        ("row_index", "<i4"),
    ]
)
assert CELL_DTYPE.itemsize == 120

CLUSTER_DTYPE = np.dtype(
    [("stamp", "<u8"), ("id", "<u8"), ("point_offset", "<i8"), ("num_points", "<i8"), ("event_index", "<i8")]
)
CLUSTER_POINT_DTYPE = np.dtype([("gcol", "<i8"), ("globally_unique_point_index", "<u8"), ("row", "<i4"), ("pad_", "<i4")])
CLOUD_DTYPE = np.dtype([("from_gcol", "<i8"), ("to_gcol", "<i8"), ("kind", "<i4"), ("width", "<u4"), ("height", "<u4"),
                        ("point_step", "<u4"), ("stamp_ns", "<u8"), ("data_offset", "<i8"), ("data_size", "<i8")])
assert CLOUD_DTYPE.itemsize == 56

# sensor_msgs/PointCloud2 point layout of the ROS node's messages (ros_utils.cpp:108-243): fields are packed without
# padding; the first 19 make up the ground-stage cloud (point_step 76), all 26 the clustered one (116)
POINTCLOUD2_FIELDS = [
    ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("firing_index", "<f8"), ("intensity", "u1"),
    ("globally_unique_point_index", "<f8"), ("time_sec", "<u4"), ("time_nsec", "<u4"), ("distance", "<f4"),
    ("azimuth_angle", "<f4"), ("inclination_angle", "<f4"), ("continuous_azimuth_angle", "<f8"),
    ("global_column_index", "<f8"), ("local_column_index", "<u2"), ("row_index", "<u2"), ("ground_point_label", "u1"),
    ("debug_ground_point_label", "u1"), ("height_over_ground", "<f4"), ("ignore_for_clustering", "u1"),
    ("finished_at_continuous_azimuth_angle", "<f8"), ("num_child_points", "<u2"), ("tree_root_row_index", "<u2"),
    ("tree_root_column_index", "<f8"), ("number_of_visited_neighbors", "<u4"), ("tree_id", "<f8"), ("id", "<f8"),
]


def pointcloud2_dtype(n_fields: int) -> np.dtype:
    return np.dtype(POINTCLOUD2_FIELDS[:n_fields])


assert pointcloud2_dtype(19).itemsize == 76 and pointcloud2_dtype(26).itemsize == 116


class Driver:
    """Thin wrapper around one drv_t of one driver library."""

    def __init__(self, lib_path: str):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(lib_path)
        self.lib = C.CDLL(lib_path)
        L = self.lib
        L.drv_impl_name.restype = C.c_char_p
        L.drv_create.restype = C.c_void_p
        L.drv_destroy.argtypes = [C.c_void_p]
        L.drv_last_error.restype = C.c_char_p
        L.drv_last_error.argtypes = [C.c_void_p]
        L.drv_configure.argtypes = [C.c_void_p, C.POINTER(CcConfig), C.c_int, C.c_void_p]
        L.drv_set_record.argtypes = [C.c_void_p, C.c_int]
        L.drv_add_firings.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.drv_reset_required.argtypes = [C.c_void_p]
        L.drv_num_rows.argtypes = [C.c_void_p]
        L.drv_ring_buffer_max_columns.argtypes = [C.c_void_p]
        L.drv_prepare.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.drv_run_prepared.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64]
        L.drv_run_prepared.restype = C.c_double
        L.drv_run_prepared_latency.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.drv_set_callbacks.argtypes = [C.c_void_p, C.c_int]
        for name in (
            "drv_num_events",
            "drv_num_ground_columns",
            "drv_num_cluster_columns",
            "drv_num_clusters",
            "drv_num_cluster_points",
        ):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int64
        L.drv_get_events.argtypes = [C.c_void_p, C.c_void_p]
        L.drv_get_ground_columns.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.drv_get_cluster_columns.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.drv_get_clusters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.drv_clear_records.argtypes = [C.c_void_p]
        L.drv_set_cloud_record.argtypes = [C.c_void_p, C.c_int]
        for name in ("drv_num_clouds", "drv_cloud_bytes"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int64
        L.drv_get_clouds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.drv_kitti_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.drv_kitti_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.drv_kitti_get.restype = C.c_int64
        self.h = L.drv_create()
        self.rows = 0
        self.name = L.drv_impl_name().decode()

    def close(self):
        if self.h:
            self.lib.drv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self) -> str:
        return self.lib.drv_last_error(self.h).decode(errors="replace")

    def configure(self, cfg: CcConfig, rows: int, robot_from_sensor=None, identity_robot_tf: bool = True):
        tf = None
        if robot_from_sensor is not None:
            tf = np.ascontiguousarray(robot_from_sensor, dtype=np.float64)
        elif identity_robot_tf:
            tf = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float64)
        rc = self.lib.drv_configure(self.h, C.byref(cfg), rows, tf.ctypes.data if tf is not None else None)
        if rc:
            raise RuntimeError(self.error())
        self.rows = rows

    def set_config(self, cfg: CcConfig):
        """setConfiguration without reset (cpp:66-81)."""
        self.lib.drv_set_config.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.drv_set_config(self.h, C.byref(cfg))

    def refusals(self):
        """(refused joins, refused links) counted by the restatement; (-1, -1) from the reference builds."""
        self.lib.drv_refusals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        a, b = C.c_int64(0), C.c_int64(0)
        self.lib.drv_refusals(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def set_record(self, level: int):
        self.lib.drv_set_record(self.h, level)

    def add_firings(self, pts: np.ndarray, poses: np.ndarray, raise_on_error: bool = True) -> int:
        pts = np.ascontiguousarray(pts)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        n, rows = pts.shape
        rc = self.lib.drv_add_firings(self.h, n, rows, pts.ctypes.data, poses.ctypes.data)
        if rc and raise_on_error:
            raise RuntimeError(self.error())
        return rc

    def reset_required(self) -> bool:
        return bool(self.lib.drv_reset_required(self.h))

    def prepare(self, pts, poses):
        pts = np.ascontiguousarray(pts)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        self._keep = (pts, poses)
        n, rows = pts.shape
        self.lib.drv_prepare(self.h, n, rows, pts.ctypes.data, poses.ctypes.data)

    def run_prepared(self, a: int, b: int, max_lag_columns: int = 0) -> float:
        s = self.lib.drv_run_prepared(self.h, a, b, max_lag_columns)
        if s < 0:
            raise RuntimeError(self.error())
        return s

    def run_prepared_latency(self, a: int, b: int) -> np.ndarray:
        """Wall-clock microseconds of every addFiring call for prepared firings [a, b)."""
        out = np.zeros(b - a, dtype=np.float64)
        if self.lib.drv_run_prepared_latency(self.h, a, b, out.ctypes.data):
            raise RuntimeError(self.error())
        return out

    def set_callbacks(self, on: bool):
        self.lib.drv_set_callbacks(self.h, int(on))

    # ---- recorded outputs -------------------------------------------------------------------------
    def events(self) -> np.ndarray:
        n = self.lib.drv_num_events(self.h)
        out = np.zeros(n, dtype=EVENT_DTYPE)
        if n:
            self.lib.drv_get_events(self.h, out.ctypes.data)
        return out

    def _columns(self, count_fn, get_fn):
        n = count_fn(self.h)
        cols = np.zeros(n, dtype=np.int64)
        cells = np.zeros((n, self.rows), dtype=CELL_DTYPE)
        if n:
            get_fn(self.h, cols.ctypes.data, cells.ctypes.data)
        return cols, cells

    def ground_columns(self):
        return self._columns(self.lib.drv_num_ground_columns, self.lib.drv_get_ground_columns)

    def cluster_columns(self):
        return self._columns(self.lib.drv_num_cluster_columns, self.lib.drv_get_cluster_columns)

    def clusters(self):
        n = self.lib.drv_num_clusters(self.h)
        m = self.lib.drv_num_cluster_points(self.h)
        cl = np.zeros(n, dtype=CLUSTER_DTYPE)
        pt = np.zeros(m, dtype=CLUSTER_POINT_DTYPE)
        if n:
            self.lib.drv_get_clusters(self.h, cl.ctypes.data, pt.ctypes.data)
        return cl, pt

    def clear_records(self):
        self.lib.drv_clear_records(self.h)

    # ---- the reference's own caller code (ros_utils.cpp / kitti_demo.cpp excerpts), where the library has it ----
    def has_caller_excerpts(self) -> bool:
        return bool(self.lib.drv_has_caller_excerpts())

    def set_cloud_record(self, on: bool):
        self.lib.drv_set_cloud_record(self.h, int(on))

    def clouds(self):
        """[(descriptor, structured array of the message's points [height, width])] in callback order."""
        n = self.lib.drv_num_clouds(self.h)
        desc = np.zeros(n, dtype=CLOUD_DTYPE)
        data = np.zeros(self.lib.drv_cloud_bytes(self.h), dtype=np.uint8)
        if n:
            self.lib.drv_get_clouds(self.h, desc.ctypes.data, data.ctypes.data)
        out = []
        for d in desc:
            dt = pointcloud2_dtype({76: 19, 116: 26, 69: 15, 37: 8}[int(d["point_step"])])
            raw = data[int(d["data_offset"]) : int(d["data_offset"]) + int(d["data_size"])]
            out.append((d, raw.view(dt).reshape(int(d["height"]), int(d["width"]))))
        return out

    def kitti_begin(self, sequence: int, points_per_frame):
        ppf = np.ascontiguousarray(points_per_frame, dtype=np.int32)
        self.lib.drv_kitti_begin(self.h, sequence, len(ppf), ppf.ctypes.data)

    def kitti_get(self, frame: int):
        n = self.lib.drv_kitti_get(self.h, frame, None, None)
        if n < 0:
            return None
        flags = np.zeros(n, dtype=np.uint8)
        labels = np.zeros(n, dtype=np.uint32)
        self.lib.drv_kitti_get(self.h, frame, flags.ctypes.data, labels.ctypes.data)
        return flags, labels


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_LIB)
