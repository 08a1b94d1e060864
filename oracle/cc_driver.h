/*
 * cc_driver.h -- TEST INFRASTRUCTURE. A small C API ("driver") that feeds firings into a
 * continuous_clustering::ContinuousClustering object and records everything the object reports through
 * its callbacks and its public `range_image_` member.
 *
 * The same source (cc_driver.cpp) is compiled twice:
 *   oracle/_ref/libcc_ref.so       against the UNMODIFIED reference sources under /root/reference
 *                                   (oracle/Makefile, target `ref`; only possible where /root/reference exists)
 *   build/libcc_facade_driver.so   against the drop-in facade in facade/ (which calls the CUDA library)
 * and the restatement oracle/cc_oracle.cpp exports the same symbols directly. Tests load the libraries
 * with ctypes and compare the recorded outputs. Nothing here is on the product path.
 */
#ifndef CC_DRIVER_H
#define CC_DRIVER_H

#include <stdint.h>

#include "../include/cc_b200.h" /* cc_config_t, cc_raw_point_t, cc_column_event_t */

#ifdef __cplusplus
extern "C" {
#endif

#define DRV_API __attribute__((visibility("default")))

typedef struct drv drv_t;

/* Snapshot of one range-image cell taken inside a finished-column callback. 120 bytes. */
typedef struct drv_cell
{
    double continuous_azimuth_angle;
    int64_t global_column_index;
    uint64_t globally_unique_point_index;
    uint64_t stamp;
    uint64_t firing_index;
    uint64_t id;
    int64_t tree_root_gcol; /* global column of tree_root_, -1 if unassociated */
    float x, y, z;
    float distance;
    float azimuth_angle;
    float inclination_angle;
    int32_t tree_root_row;
    int32_t number_of_visited_neighbors;
    uint8_t intensity;
    uint8_t ground_point_label;
    uint8_t debug_ground_point_label;
    uint8_t is_ignored;
    int32_t num_child_points; /* child_points.size() */
    double finished_at_continuous_azimuth_angle;
    uint32_t tree_num_points;
    uint32_t cluster_width;
    int32_t local_column_index;
    int32_t row_index;
} drv_cell_t;

typedef struct drv_cluster
{
    uint64_t stamp;
    uint64_t id; /* id of the first point handed to the callback */
    int64_t point_offset;
    int64_t num_points;
    int64_t event_index; /* number of column events recorded before this cluster callback */
} drv_cluster_t;

typedef struct drv_cluster_point
{
    int64_t gcol;
    uint64_t globally_unique_point_index;
    int32_t row;
    int32_t pad_;
} drv_cluster_point_t;

/* record levels */
enum
{
    DRV_RECORD_NONE = 0,   /* callbacks registered but do nothing (timing runs) */
    DRV_RECORD_EVENTS = 1, /* column events + cluster summaries */
    DRV_RECORD_FULL = 2    /* + a drv_cell_t snapshot of every column reported by a callback */
};

DRV_API const char* drv_impl_name(void); /* "reference", "restatement" or "facade" */
DRV_API drv_t* drv_create(void);
DRV_API void drv_destroy(drv_t* d);
DRV_API const char* drv_last_error(drv_t* d);
/* setConfiguration + reset(num_rows) + setTransformRobotFrameFromSensorFrame (skipped if NULL) */
DRV_API int drv_configure(drv_t* d, const cc_config_t* cfg, int num_rows, const double* robot_from_sensor);
DRV_API int drv_set_config(drv_t* d, const cc_config_t* cfg);
DRV_API void drv_refusals(drv_t* d, int64_t* joins, int64_t* links);
DRV_API void drv_set_record(drv_t* d, int level);
/* addFiring for each of n firings; returns 0, or 1 if the object threw (message in drv_last_error) */
DRV_API int drv_add_firings(drv_t* d, int n, int rows, const cc_raw_point_t* pts, const double* poses);
DRV_API int drv_reset_required(drv_t* d);
DRV_API int drv_num_rows(drv_t* d);
DRV_API int drv_ring_buffer_max_columns(drv_t* d);

/* timing helpers: build the shared_ptr firings outside the timed region, then feed [from, to).
 * In multi-threaded mode the producer is throttled so that fed - finished <= max_lag_columns
 * (the reference throws its ring-overrun error otherwise, cpp:337-344); the call returns when the
 * pipeline has drained. Returns seconds of wall time, negative on error. */
DRV_API int drv_prepare(drv_t* d, int n, int rows, const cc_raw_point_t* pts, const double* poses);
DRV_API double drv_run_prepared(drv_t* d, int from, int to, int64_t max_lag_columns);

/* Per-call latency (per-addFiring p50 / p99): feeds prepared firings [from, to) one addFiring
 * call at a time and stores the wall-clock duration of every call in out_us[to - from] (microseconds). In the
 * reference's single-threaded mode the call returns when everything the firing triggered -- segmentation, association,
 * finish detection, callbacks -- is done; on the facade most calls only buffer and every batch-th call carries the whole
 * device push and its callbacks. Returns 0, or 1 if the object threw. */
DRV_API int drv_run_prepared_latency(drv_t* d, int from, int to, double* out_us);
/* callbacks != 0 (default): the recording callbacks are registered; 0: none are (the object skips whatever it only does
 * for callbacks, e.g. the facade does not fill range_image_) */
DRV_API void drv_set_callbacks(drv_t* d, int callbacks);

/* recorded outputs */
DRV_API int64_t drv_num_events(drv_t* d);
DRV_API void drv_get_events(drv_t* d, cc_column_event_t* out);
/* snapshots, in callback order: ground_only=1 columns and ground_only=0 columns separately */
DRV_API int64_t drv_num_ground_columns(drv_t* d);
DRV_API void drv_get_ground_columns(drv_t* d, int64_t* gcols, drv_cell_t* cells /* n*rows */);
DRV_API int64_t drv_num_cluster_columns(drv_t* d);
DRV_API void drv_get_cluster_columns(drv_t* d, int64_t* gcols, drv_cell_t* cells /* n*rows */);
DRV_API int64_t drv_num_clusters(drv_t* d);
DRV_API int64_t drv_num_cluster_points(drv_t* d);
DRV_API void drv_get_clusters(drv_t* d, drv_cluster_t* clusters, drv_cluster_point_t* points);
DRV_API void drv_clear_records(drv_t* d);

/* ---- the reference's own CALLER code run against the object (only in libraries built with
 * -DDRV_WITH_CALLER_EXCERPTS, i.e. where /root/reference was present at build time: oracle/_ref/*.so).
 * The excerpts (oracle/extract_caller_excerpts.py) are ros_utils.cpp's columnToPointCloud / clusterToPointCloud /
 * addPointToMessage and kitti_demo.cpp's column callback body, compiled unmodified. ---- */
typedef struct drv_cloud
{
    int64_t from_gcol, to_gcol; /* column range of a column cloud; cluster clouds: -1, -1 */
    int32_t kind;               /* 0 ground-stage columns, 1 clustered columns, 2 finished cluster */
    uint32_t width, height, point_step;
    uint64_t stamp_ns;          /* msg->header.stamp */
    int64_t data_offset, data_size; /* into the byte blob returned by drv_get_clouds */
} drv_cloud_t;

DRV_API int drv_has_caller_excerpts(void);
/* on != 0: every column / cluster callback also builds the PointCloud2 the ROS node would publish
 * (continuous_clustering_node.cpp:166-178) and keeps its bytes */
DRV_API void drv_set_cloud_record(drv_t* d, int on);
DRV_API int64_t drv_num_clouds(drv_t* d);
DRV_API int64_t drv_cloud_bytes(drv_t* d);
DRV_API void drv_get_clouds(drv_t* d, drv_cloud_t* clouds, uint8_t* data);
/* kitti_demo.cpp:173-224 run from the clustered-column callback: frames of `points_per_frame[f]` points each
 * (guid = seq << 48 | frame << 32 | index); results per point: bit 0 has_corresponding_point, bit 1 is_ground_point */
DRV_API void drv_kitti_begin(drv_t* d, int sequence, int n_frames, const int32_t* points_per_frame);
DRV_API int64_t drv_kitti_get(drv_t* d, int frame, uint8_t* flags, uint32_t* detection_label);

#ifdef __cplusplus
}
#endif
#endif
