"""ctypes binding of the C ABI in include/cc_b200.h (continuous_clustering_b200/libcc_b200.so).

The library is CUDA-only (sm_100a). There is no CPU implementation of the path in this package: loading fails
loudly when the nvcc-built library is missing, and cc_create fails when there is no CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcc_b200.so")


class CcConfig(C.Structure):
    """cc_config_t: plain-C mirror of continuous_clustering::Configuration (hpp:24-87)."""

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


class CcBatchInfo(C.Structure):
    _fields_ = [
        ("ground_from_gcol", C.c_int64),
        ("ground_to_gcol", C.c_int64),
        ("first_unpublished_gcol", C.c_int64),
        ("ring_start_gcol", C.c_int64),
        ("ring_end_gcol", C.c_int64),
        ("cleared_from_gcol", C.c_int64),
        ("cleared_to_gcol", C.c_int64),
        ("n_events", C.c_int32),
        ("n_clusters", C.c_int32),
        ("n_cluster_points", C.c_int32),
        ("reset_required", C.c_int32),
        ("used_exact_path", C.c_int32),
        ("gpu_launches", C.c_int32),
        ("device_ms", C.c_float),
        ("slow_insert_firings", C.c_int32),
        ("n_unfinished_trees", C.c_int32),
        ("fused_launch", C.c_int32),
        ("visited_recounts", C.c_int32),
        ("pad_", C.c_int32),
    ]


class CcColumnFields(C.Structure):
    _fields_ = [
        (name, C.c_void_p)
        for name in (
            "xyz", "distance", "azimuth_angle", "inclination_angle", "continuous_azimuth_angle",
            "global_column_index", "stamp", "globally_unique_point_index", "firing_index", "intensity",
            "ground_point_label", "debug_ground_point_label", "is_ignored", "id", "tree_root_gcol", "tree_root_row",
            "finished_at_continuous_azimuth_angle", "tree_num_points", "cluster_width", "number_of_visited_neighbors",
            "first_parent_gcol", "first_parent_row", "belongs_to_finished_cluster",
        )
    ]


COLUMN_FIELD_DTYPES = {
    "xyz": ("<f4", 3), "distance": ("<f4", 1), "azimuth_angle": ("<f4", 1), "inclination_angle": ("<f4", 1),
    "continuous_azimuth_angle": ("<f8", 1), "global_column_index": ("<i8", 1), "stamp": ("<u8", 1),
    "globally_unique_point_index": ("<u8", 1), "firing_index": ("<u8", 1), "intensity": ("u1", 1),
    "ground_point_label": ("u1", 1), "debug_ground_point_label": ("u1", 1), "is_ignored": ("u1", 1),
    "id": ("<u8", 1), "tree_root_gcol": ("<i8", 1), "tree_root_row": ("<i4", 1),
    "finished_at_continuous_azimuth_angle": ("<f8", 1), "tree_num_points": ("<u4", 1), "cluster_width": ("<u4", 1),
    "number_of_visited_neighbors": ("<i4", 1), "first_parent_gcol": ("<i8", 1), "first_parent_row": ("<i4", 1),
    "belongs_to_finished_cluster": ("u1", 1),
}

CELL_DTYPE = np.dtype([  # cc_cell_t (include/cc_b200.h): one packed record per cell, cc_export_columns
    ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("distance", "<f4"), ("azimuth_angle", "<f4"), ("inclination_angle", "<f4"),
    ("continuous_azimuth_angle", "<f8"), ("global_column_index", "<i8"), ("stamp", "<u8"),
    ("globally_unique_point_index", "<u8"), ("firing_index", "<u8"), ("id", "<u8"),
    ("finished_at_continuous_azimuth_angle", "<f8"), ("tree_root_gcol", "<i8"), ("first_parent_gcol", "<i8"),
    ("tree_num_points", "<u4"), ("cluster_width", "<u4"), ("tree_root_row", "<i4"), ("first_parent_row", "<i4"),
    ("number_of_visited_neighbors", "<u2"), ("pad0_", "<u2"), ("intensity", "u1"), ("ground_point_label", "u1"),
    ("debug_ground_point_label", "u1"), ("is_ignored", "u1"), ("belongs_to_finished_cluster", "u1"), ("pad1_", "u1", (7,)),
])
assert CELL_DTYPE.itemsize == 128

EVENT_DTYPE = np.dtype(
    [("from_gcol", "<i8"), ("to_gcol", "<i8"), ("ground_points_only", "<i4"), ("n_clusters_before", "<i4")]
)
CLUSTER_DTYPE = np.dtype(
    [("id", "<u8"), ("stamp", "<u8"), ("min_stamp", "<u8"), ("max_stamp", "<u8"), ("finished_at_gcol", "<i8"),
     ("min_gcol", "<i8"), ("max_gcol", "<i8"), ("num_points", "<u4"), ("point_offset", "<u4")]
)
CLUSTER_POINT_DTYPE = np.dtype([("gcol", "<i8"), ("row", "<i4"), ("pad_", "<i4")])

STATUS_NAMES = {
    0: "CC_OK", 1: "CC_ERR_INVALID_ARGUMENT", 2: "CC_ERR_CUDA", 3: "CC_ERR_ROW_COUNT_CHANGED",
    4: "CC_ERR_NO_ROBOT_TRANSFORM", 5: "CC_ERR_COLUMN_NOT_CLEARED", 6: "CC_ERR_RING_START_DECREASED",
    7: "CC_ERR_NOT_RESET", 8: "CC_ERR_BATCH_TOO_LARGE", 9: "CC_ERR_INTERNAL",
}

# every symbol include/cc_b200.h declares
EXPORTED_SYMBOLS = [
    "cc_create", "cc_destroy", "cc_last_error", "cc_version", "cc_config_default", "cc_set_config", "cc_reset",
    "cc_reset_required", "cc_set_robot_from_sensor", "cc_has_robot_from_sensor", "cc_push_firings",
    "cc_push_firings_device", "cc_get_batch_info", "cc_get_column_events", "cc_get_clusters", "cc_get_cluster_points",
    "cc_read_columns", "cc_num_rows", "cc_num_columns", "cc_ring_buffer_max_columns", "cc_stream",
    "cc_total_launches", "cc_selftest_math", "cc_set_kernel_timing", "cc_get_kernel_timings",
    "cc_debug_flag_columns", "cc_get_result_views", "cc_submit_firings", "cc_submit_firings_device", "cc_wait", "cc_pending",
    "cc_max_firings_per_push", "cc_debug_event_query", "cc_set_label_prefetch", "cc_get_column_labels",
    "cc_debug_trace", "cc_debug_get_trace", "cc_debug_slot_base", "cc_debug_slot_times", "cc_export_columns",
    "cc_pack_columns_pointcloud2", "cc_pack_cluster_pointcloud2", "cc_pack_requests_pointcloud2",
    "cc_eval_create", "cc_eval_destroy", "cc_eval_frame",
    "cc_ouster_format_legacy", "cc_ouster_create", "cc_ouster_destroy", "cc_ouster_packet_size", "cc_ouster_set_lut",
    "cc_ouster_reset", "cc_ouster_decode", "cc_ouster_read_firings",
    "cc_kitti_create", "cc_kitti_destroy", "cc_kitti_set_poses", "cc_kitti_frame", "cc_kitti_read_debug",
]


class CcEvalResult(C.Structure):
    """cc_eval_result_t == EvaluationResultForFrame (kitti_evaluation.hpp:38-50)."""

    _fields_ = [("tp", C.c_double), ("fn", C.c_double), ("fp", C.c_double), ("tn", C.c_double),
                ("over_segmentation_entropy", C.c_double), ("under_segmentation_entropy", C.c_double)]


class CcKittiFrame(C.Structure):
    """cc_kitti_frame_t: the pseudo firings of one KITTI frame, resident on the device."""

    _fields_ = [("n_firings", C.c_int32), ("rows_per_firing", C.c_int32), ("d_firings", C.c_void_p), ("d_poses", C.c_void_p),
                ("poses", C.c_void_p), ("rows_found", C.c_int32), ("max_points_in_row", C.c_int32)]


class CcOusterFormat(C.Structure):
    """cc_ouster_format_t: the layout of a lidar packet (ouster::sensor::packet_format as data)."""

    _fields_ = [(n, C.c_int32) for n in ("columns_per_packet", "pixels_per_column", "columns_per_frame", "packet_header_size",
                                         "col_header_size", "col_footer_size", "pixel_bytes", "col_measurement_id_offset",
                                         "col_status_offset", "col_status_bytes", "range_offset", "range_bytes")] + [
        ("range_mask", C.c_uint32), ("range_shift", C.c_int32), ("signal_offset", C.c_int32), ("signal_bytes", C.c_int32),
        ("signal_mask", C.c_uint32), ("signal_shift", C.c_int32), ("offset_from_direction_table", C.c_int32)]


class CcDecodedFirings(C.Structure):
    _fields_ = [("n_firings", C.c_int32), ("rows_per_firing", C.c_int32), ("d_firings", C.c_void_p), ("firing_stamps", C.c_void_p),
                ("first_firing_index", C.c_uint64)]


class CcPackRequest(C.Structure):
    _fields_ = [("kind", C.c_int32), ("cluster_index", C.c_int32), ("from_gcol", C.c_int64), ("to_gcol", C.c_int64)]


class CcCloudView(C.Structure):
    """cc_cloud_view_t: a packed sensor_msgs/PointCloud2 payload in a page-locked buffer of the handle."""

    _fields_ = [("data", C.c_void_p), ("data_size", C.c_uint64), ("stamp_ns", C.c_uint64), ("point_step", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("n_fields", C.c_uint32)]


def bind(lib: C.CDLL) -> C.CDLL:
    """Declares the argument / result types of every entry point."""
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.cc_create.argtypes = [i32, i32, C.POINTER(vp)]
    lib.cc_destroy.argtypes = [vp]
    lib.cc_destroy.restype = None
    lib.cc_last_error.argtypes = [vp]
    lib.cc_last_error.restype = C.c_char_p
    lib.cc_version.restype = C.c_char_p
    lib.cc_config_default.argtypes = [C.POINTER(CcConfig)]
    lib.cc_config_default.restype = None
    lib.cc_set_config.argtypes = [vp, vp]
    lib.cc_reset.argtypes = [vp, i32]
    lib.cc_reset_required.argtypes = [vp]
    lib.cc_set_robot_from_sensor.argtypes = [vp, vp]
    lib.cc_has_robot_from_sensor.argtypes = [vp]
    lib.cc_push_firings.argtypes = [vp, i32, i32, vp, vp]
    lib.cc_push_firings_device.argtypes = [vp, i32, i32, vp, vp]
    lib.cc_submit_firings.argtypes = [vp, i32, i32, vp, vp]
    lib.cc_submit_firings_device.argtypes = [vp, i32, i32, vp, vp]
    lib.cc_wait.argtypes = [vp]
    lib.cc_pending.argtypes = [vp]
    lib.cc_max_firings_per_push.argtypes = [vp]
    lib.cc_debug_event_query.argtypes = [vp, i32, i32]
    lib.cc_set_label_prefetch.argtypes = [vp, i32]
    lib.cc_get_column_labels.argtypes = [vp, C.POINTER(vp), C.POINTER(i32)]
    lib.cc_get_batch_info.argtypes = [vp, C.POINTER(CcBatchInfo)]
    for name in ("cc_get_column_events", "cc_get_clusters", "cc_get_cluster_points"):
        getattr(lib, name).argtypes = [vp, vp, i32, C.POINTER(i32)]
    lib.cc_get_result_views.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.cc_read_columns.argtypes = [vp, i64, i64, C.POINTER(CcColumnFields)]
    lib.cc_export_columns.argtypes = [vp, i64, i64, C.POINTER(vp)]
    lib.cc_pack_columns_pointcloud2.argtypes = [vp, i64, i64, i32, C.POINTER(CcCloudView)]
    lib.cc_pack_cluster_pointcloud2.argtypes = [vp, i32, C.POINTER(CcCloudView)]
    lib.cc_pack_requests_pointcloud2.argtypes = [vp, i32, vp, vp]
    lib.cc_eval_create.argtypes = [i32, i32, C.POINTER(vp)]
    lib.cc_eval_destroy.argtypes = [vp]
    lib.cc_eval_destroy.restype = None
    lib.cc_eval_frame.argtypes = [vp, i32, vp, vp, vp, vp, C.POINTER(CcEvalResult)]
    lib.cc_ouster_format_legacy.argtypes = [i32, i32, C.POINTER(CcOusterFormat)]
    lib.cc_ouster_format_legacy.restype = None
    lib.cc_ouster_create.argtypes = [i32, C.POINTER(CcOusterFormat), i32, C.POINTER(vp)]
    lib.cc_ouster_destroy.argtypes = [vp]
    lib.cc_ouster_destroy.restype = None
    lib.cc_ouster_packet_size.argtypes = [vp]
    lib.cc_ouster_set_lut.argtypes = [vp, vp, vp]
    lib.cc_ouster_reset.argtypes = [vp]
    lib.cc_ouster_decode.argtypes = [vp, i32, vp, vp, C.POINTER(CcDecodedFirings)]
    lib.cc_ouster_read_firings.argtypes = [vp, i32, vp]
    lib.cc_kitti_create.argtypes = [i32, i32, C.POINTER(vp)]
    lib.cc_kitti_destroy.argtypes = [vp]
    lib.cc_kitti_destroy.restype = None
    lib.cc_kitti_set_poses.argtypes = [vp, i32, vp, vp]
    lib.cc_kitti_frame.argtypes = [vp, i32, vp, C.c_uint64, C.c_uint64, vp, i32, i32, C.POINTER(CcKittiFrame)]
    lib.cc_kitti_read_debug.argtypes = [vp, vp, vp, vp, vp]
    for name in ("cc_num_rows", "cc_num_columns", "cc_ring_buffer_max_columns"):
        getattr(lib, name).argtypes = [vp]
    lib.cc_stream.argtypes = [vp]
    lib.cc_stream.restype = vp
    lib.cc_total_launches.argtypes = [vp]
    lib.cc_total_launches.restype = C.c_uint64
    lib.cc_selftest_math.argtypes = [i32, i32, i32, vp, vp, vp]
    lib.cc_set_kernel_timing.argtypes = [vp, i32]
    lib.cc_debug_flag_columns.argtypes = [vp, i32]
    lib.cc_debug_trace.argtypes = [vp, i32]
    lib.cc_debug_slot_base.argtypes = [vp]
    lib.cc_debug_slot_times.argtypes = [vp, i32, vp]
    lib.cc_debug_get_trace.argtypes = [vp, vp, i32, vp, i32, C.POINTER(i32)]
    lib.cc_get_kernel_timings.argtypes = [vp, vp, i32, vp, i32, C.POINTER(i32)]
    return lib


_LIB = None


def load_library(path: str | None = None) -> C.CDLL:
    """Loads the CUDA library. Raises if it has not been built: there is no fallback."""
    global _LIB
    if path is None:
        if _LIB is not None:
            return _LIB
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make lib` (nvcc, sm_100a) or __graft_entry__.build(). "
                "continuous_clustering_b200 has no CPU implementation."
            )
        _LIB = bind(C.CDLL(LIB_PATH))
        return _LIB
    return bind(C.CDLL(path))
