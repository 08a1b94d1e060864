/*
 * cc_b200.h -- C ABI of the B200-native continuous-clustering hot path.
 *
 * One handle == one sensor stream == one `continuous_clustering::ContinuousClustering` object of
 * the reference (include/continuous_clustering/clustering/continuous_clustering.hpp:197-290).
 * The reference has no FFI; the only caller of this ABI is the header-compatible C++ facade in
 * facade/ (and, for tests/bench, Python ctypes).  Each entry point cites the reference interface it
 * replaces ("hpp" = continuous_clustering.hpp, "cpp" = src/clustering/continuous_clustering.cpp).
 *
 * Conventions
 *  - plain pointers + sizes, no C++/torch types; every function returns a cc_status_t (0 = ok) unless
 *    stated otherwise; cc_last_error() gives the message the facade re-throws as std::runtime_error.
 *  - a handle may be used from one host thread at a time; handles are independent (one per GPU/stream).
 *  - poses are 3x4 row-major doubles [R|t] (odom_from_sensor), 12 per firing.
 *  - global column index ("gcol") = the reference's global_column_index (cpp:152-153);
 *    ring cell = (gcol % (10*num_columns)) * num_rows + row   (cpp:17, 178-181).
 *  - all compute runs in CUDA kernels on the handle's device; there is no CPU fallback.
 */
#ifndef CC_B200_H
#define CC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define CC_API
#else
#define CC_API __attribute__((visibility("default")))
#endif

typedef struct cc_handle cc_handle_t;

typedef enum cc_status
{
    CC_OK = 0,
    CC_ERR_INVALID_ARGUMENT = 1,
    CC_ERR_CUDA = 2,
    CC_ERR_ROW_COUNT_CHANGED = 3,   /* cpp:90-91   "The number of points in a firing has changed..."      */
    CC_ERR_NO_ROBOT_TRANSFORM = 4,  /* cpp:298-299 "Transform robot frame from sensor frame was not set yet!" */
    CC_ERR_COLUMN_NOT_CLEARED = 5,  /* cpp:337-344 ring buffer overrun                                     */
    CC_ERR_RING_START_DECREASED = 6,/* cpp:1072-1075                                                      */
    CC_ERR_NOT_RESET = 7,           /* push before cc_reset()                                              */
    CC_ERR_BATCH_TOO_LARGE = 8,     /* more firings/columns in one push than the handle was sized for      */
    CC_ERR_INTERNAL = 9
} cc_status_t;

/* Layout-identical to continuous_clustering::RawPoint (point_types.hpp:10-19), 48 bytes, so the facade
 * can hand `firing->points.data()` straight to cc_push_firings. */
typedef struct cc_raw_point
{
    float x, y, z;
    uint32_t pad0_;
    uint64_t firing_index;
    uint8_t intensity;
    uint8_t pad1_[7];
    uint64_t stamp;
    uint64_t globally_unique_point_index;
} cc_raw_point_t;

/* Plain-C mirror of continuous_clustering::Configuration (hpp:24-87); same names, same defaults
 * (cc_config_default). Booleans are int32. */
typedef struct cc_config
{
    /* GeneralConfiguration hpp:24-27. Accepted for API parity; the device pipeline always has the
     * deterministic ordering of the reference's single-threaded mode. */
    int32_t is_single_threaded;
    /* ContinuousRangeImageConfiguration hpp:29-34 */
    int32_t sensor_is_clockwise;
    int32_t num_columns;
    int32_t supplement_inclination_angle_for_nan_cells;
    /* ContinuousGroundSegmentationConfiguration hpp:36-66 */
    float max_slope;
    float first_ring_as_ground_max_allowed_z_diff;
    float first_ring_as_ground_min_allowed_z_diff;
    float last_ground_point_slope_higher_than;
    float last_ground_point_distance_smaller_than;
    float ground_because_close_to_last_certain_ground_max_z_diff;
    float ground_because_close_to_last_certain_ground_max_dist_diff;
    float obstacle_because_next_certain_obstacle_max_dist_diff;
    int32_t use_terrain;
    float terrain_max_allowed_z_diff;
    float height_ref_to_maximum_;
    float height_ref_to_ground_;
    float length_ref_to_front_end_;
    float length_ref_to_rear_end_;
    float width_ref_to_left_mirror_;
    float width_ref_to_right_mirror_;
    int32_t fog_filtering_enabled;
    int32_t fog_filtering_intensity_below; /* uint8 in the reference */
    float fog_filtering_distance_below;
    float fog_filtering_inclination_above;
    /* ContinuousClusteringConfiguration hpp:68-79 */
    float max_distance;
    int32_t max_steps_in_row;
    int32_t max_steps_in_column;
    int32_t stop_after_association_enabled;
    int32_t stop_after_association_min_steps;
    int32_t ignore_points_in_chessboard_pattern;
    int32_t ignore_points_with_too_big_inclination_angle_diff;
    int32_t use_last_point_for_cluster_stamp;
    int32_t cluster_point_trees_every_nth_column;
} cc_config_t;

/* One finished-column callback of the reference: finished_column_callback_(from, to, ground_only)
 * (cpp:618-620 with ground_only=1; cpp:1087-1089 with ground_only=0; `to < from` is an empty range and is
 * reported exactly as the reference reports it). Events come in the order the reference's single-threaded
 * mode invokes them. `n_clusters_before` = how many entries of the batch's cluster list precede this
 * event in callback order (clusters finished by column c are delivered before c's ground_only=0 event). */
typedef struct cc_column_event
{
    int64_t from_gcol;
    int64_t to_gcol;
    int32_t ground_points_only;
    int32_t n_clusters_before;
} cc_column_event_t;

/* One finished cluster with more than 5 points (cpp:936-940). Points are cell references into the ring;
 * `stamp` is what the reference passes to finished_cluster_callback_ (cpp:1025-1028); the reference only
 * invokes that callback when num_points > 20 (cpp:1023). */
typedef struct cc_cluster
{
    uint64_t id;            /* Point::id of every member, cpp:939, 1005 (ids are a permutation of the reference's) */
    uint64_t stamp;         /* cpp:1025-1028 */
    uint64_t min_stamp;
    uint64_t max_stamp;
    int64_t finished_at_gcol; /* column whose tree-combination pass finished the cluster (cpp:837) */
    int64_t min_gcol;
    int64_t max_gcol;
    uint32_t num_points;
    uint32_t point_offset;  /* first entry in the batch's cluster-point list */
} cc_cluster_t;

/* Member of a finished cluster: which ring cell. */
typedef struct cc_cluster_point
{
    int64_t gcol;
    int32_t row;
    int32_t pad_;
} cc_cluster_point_t;

/* Summary of the last cc_push_firings call. */
typedef struct cc_batch_info
{
    int64_t ground_from_gcol; /* columns [from, to) went through ground segmentation + association in this batch */
    int64_t ground_to_gcol;
    int64_t first_unpublished_gcol;  /* sc_first_unpublished_global_column_index after the batch (hpp:270) */
    int64_t ring_start_gcol;         /* ring_buffer_start_global_column_index (hpp:250) */
    int64_t ring_end_gcol;           /* ring_buffer_end_global_column_index (hpp:251) */
    int64_t cleared_from_gcol;       /* columns [from, to) were recycled by clearColumns (cpp:1091) in this batch */
    int64_t cleared_to_gcol;
    int32_t n_events;
    int32_t n_clusters;
    int32_t n_cluster_points;
    int32_t reset_required;          /* cpp:83-86 */
    int32_t used_exact_path;         /* 1 if the batch needed the column-sequential exact kernels */
    int32_t gpu_launches;            /* kernels launched for this batch */
    float device_ms;                 /* CUDA-event time of the batch's kernels on the handle's stream */
    int32_t slow_insert_firings;     /* firings that went through the per-firing insertion path (collisions) */
    int32_t n_unfinished_trees;      /* sc_unfinished_point_trees_.size() after the batch (hpp:273)            */
    int32_t fused_launch;            /* 1 if the batch ran as ONE fused kernel launch (short pushes) */
    int32_t visited_recounts;        /* points whose number_of_visited_neighbors was recounted (walk cut at the first
                                        unpublished column, cpp:762-763) */
    int32_t pad_;
} cc_batch_info_t;

/* Field selector + destination pointers for cc_read_columns. Each non-NULL pointer receives
 * n_cols*num_rows values in the reference's column-major cell order (column-by-column, row 0 = top
 * laser). Cleared / never-written cells hold the reference's cleared values (cpp:1110-1142). */
typedef struct cc_column_fields
{
    float* xyz;                        /* 3 floats per cell, Point::xyz                        */
    float* distance;                   /* Point::distance                                      */
    float* azimuth_angle;              /* Point::azimuth_angle                                 */
    float* inclination_angle;          /* Point::inclination_angle (incl. NaN-cell supplement) */
    double* continuous_azimuth_angle;  /* Point::continuous_azimuth_angle                      */
    int64_t* global_column_index;      /* Point::global_column_index                           */
    uint64_t* stamp;                   /* Point::stamp                                         */
    uint64_t* globally_unique_point_index;
    uint64_t* firing_index;
    uint8_t* intensity;
    uint8_t* ground_point_label;       /* Point::ground_point_label                            */
    uint8_t* debug_ground_point_label; /* Point::debug_ground_point_label                      */
    uint8_t* is_ignored;               /* Point::is_ignored                                    */
    uint64_t* id;                      /* Point::id (0 = not part of a published cluster)      */
    int64_t* tree_root_gcol;           /* global column of Point::tree_root_ (-1 = none)       */
    int32_t* tree_root_row;            /* Point::tree_root_.row_index                          */
    /* clustering bookkeeping the ROS node publishes (ros_utils.cpp:287-298) */
    double* finished_at_continuous_azimuth_angle; /* tree roots: Point::finished_at_continuous_azimuth_angle, else 0 */
    uint32_t* tree_num_points;         /* tree roots: Point::tree_num_points, else 0           */
    uint32_t* cluster_width;           /* tree roots: Point::cluster_width, else 0             */
    int32_t* number_of_visited_neighbors; /* Point::number_of_visited_neighbors                */
    int64_t* first_parent_gcol;        /* the point whose child_points holds this one (cpp:663): global column, -1 none */
    int32_t* first_parent_row;
    uint8_t* belongs_to_finished_cluster; /* tree roots: Point::belongs_to_finished_cluster    */
} cc_column_fields_t;

/* Every field of one range-image cell (`Point`, hpp:126-161) as the device packs it: 128 bytes, one record per cell in
 * the reference's cell order (column by column, row 0 = top laser). Cleared / never-written cells hold the reference's
 * cleared values (cpp:1110-1142). child_points are represented by their inverse: first_parent_* names the point whose
 * list holds this point (children are at most max_steps_in_row columns ahead of their parent). */
typedef struct cc_cell
{
    float x, y, z, distance;
    float azimuth_angle, inclination_angle;
    double continuous_azimuth_angle;
    int64_t global_column_index;
    uint64_t stamp, globally_unique_point_index, firing_index;
    uint64_t id;
    double finished_at_continuous_azimuth_angle;
    int64_t tree_root_gcol;
    int64_t first_parent_gcol;
    uint32_t tree_num_points, cluster_width;
    int32_t tree_root_row, first_parent_row;
    uint16_t number_of_visited_neighbors, pad0_;
    uint8_t intensity, ground_point_label, debug_ground_point_label, is_ignored;
    uint8_t belongs_to_finished_cluster, pad1_[7];
} cc_cell_t;

/* ---- lifecycle --------------------------------------------------------------------------------- */

/* ContinuousClustering::ContinuousClustering() hpp:201. `max_firings_per_push` sizes the staging
 * buffers (0 = default 4096). */
CC_API cc_status_t cc_create(int device_ordinal, int max_firings_per_push, cc_handle_t** out);
CC_API void cc_destroy(cc_handle_t* h);
CC_API const char* cc_last_error(const cc_handle_t* h);
CC_API const char* cc_version(void);

/* Defaults of hpp:24-87. */
CC_API void cc_config_default(cc_config_t* cfg);
/* ContinuousClustering::setConfiguration hpp:206, cpp:66-81 (sets reset_required when
 * is_single_threaded / sensor_is_clockwise / num_columns change). */
CC_API cc_status_t cc_set_config(cc_handle_t* h, const cc_config_t* cfg);
/* ContinuousClustering::reset hpp:205, cpp:11-64. */
CC_API cc_status_t cc_reset(cc_handle_t* h, int num_rows);
/* ContinuousClustering::resetRequired hpp:207, cpp:83-86. Returns 0/1. */
CC_API int cc_reset_required(const cc_handle_t* h);
/* setTransformRobotFrameFromSensorFrame / hasTransformRobotFrameFromSensorFrame hpp:213-214, cpp:626-636. */
CC_API cc_status_t cc_set_robot_from_sensor(cc_handle_t* h, const double robot_from_sensor[12]);
CC_API int cc_has_robot_from_sensor(const cc_handle_t* h);

/* ---- the hot path ------------------------------------------------------------------------------- */

/* ContinuousClustering::addFiring hpp:210, cpp:88-93, for `n_firings` consecutive firings: runs
 * insertFiringIntoRangeImage (cpp:105-292), performGroundPointSegmentationForColumn (cpp:294-624),
 * associatePointsInColumn (cpp:773-835), findFinishedTreesAndAssignSameId (cpp:837-974), the id /
 * bookkeeping half of collectPointsForCusterAndPublish (cpp:976-1092) and clearColumns (cpp:1094-1145)
 * for every column the firings complete, on the device. `points` = n_firings*rows_per_firing host records,
 * `poses` = n_firings*12 host doubles (page-locked buffers are copied to the device without staging). Returns after
 * the results are on the host.
 * rows_per_firing != num_rows -> CC_ERR_ROW_COUNT_CHANGED (cpp:90-91). */
CC_API cc_status_t cc_push_firings(cc_handle_t* h, int n_firings, int rows_per_firing,
                                   const cc_raw_point_t* points, const double* poses);

/* Same, with inputs already resident in device memory (device pointers on the handle's device). */
CC_API cc_status_t cc_push_firings_device(cc_handle_t* h, int n_firings, int rows_per_firing,
                                          const cc_raw_point_t* d_points, const double* d_poses);

/* Asynchronous variant: cc_submit_* enqueues the host->device copy (on its own input stream) and every kernel of
 * the push and returns at once; cc_wait() blocks until the OLDEST submitted push is finished and makes its results
 * current (cc_get_batch_info & co). At most two pushes are in flight. A third HOST push may be submitted while two
 * are in flight: it is STAGED -- its input copy starts at once (into the third of three device input buffers), its
 * kernels are launched by the first cc_submit_* / cc_wait call after the cc_wait() that made room (not by that cc_wait
 * itself, so that the caller can still read the columns the finished push reported: the staged push's first kernel
 * recycles ring columns). A caller keeps the GPU and the PCIe link busy with `submit(k+2); wait(k)`; with device
 * inputs `submit(k+1); wait(k)`. Inputs must stay valid until the push has been waited for. cc_push_firings* ==
 * submit + wait. If a push in flight cannot be committed speculatively the pushes behind it
 * skip themselves on the device and cc_wait() transparently re-runs them after finishing it. Result views
 * (cc_get_result_views, cc_get_column_labels) stay valid until the next cc_wait() on the handle. */
CC_API cc_status_t cc_submit_firings(cc_handle_t* h, int n_firings, int rows_per_firing,
                                     const cc_raw_point_t* points, const double* poses);
CC_API cc_status_t cc_submit_firings_device(cc_handle_t* h, int n_firings, int rows_per_firing,
                                            const cc_raw_point_t* d_points, const double* d_poses);
CC_API cc_status_t cc_wait(cc_handle_t* h);
CC_API int cc_pending(const cc_handle_t* h); /* pushes in flight + staged (0..3) */
/* Largest n_firings one push accepts: min(max_firings_per_push, 3 * num_columns). */
CC_API int cc_max_firings_per_push(const cc_handle_t* h);

/* Results of the last push. */
CC_API cc_status_t cc_get_batch_info(const cc_handle_t* h, cc_batch_info_t* out);
/* Copies min(cap, n) entries; returns the number copied through *n_out. */
CC_API cc_status_t cc_get_column_events(const cc_handle_t* h, cc_column_event_t* out, int cap, int* n_out);
CC_API cc_status_t cc_get_clusters(const cc_handle_t* h, cc_cluster_t* out, int cap, int* n_out);
CC_API cc_status_t cc_get_cluster_points(const cc_handle_t* h, cc_cluster_point_t* out, int cap, int* n_out);

/* Zero-copy variant: pointers to the handle's own result arrays (n_events / n_clusters / n_cluster_points entries,
 * see cc_get_batch_info), valid until the next cc_wait / synchronous push or reset on this handle. */
CC_API cc_status_t cc_get_result_views(const cc_handle_t* h, const cc_column_event_t** events,
                                       const cc_cluster_t** clusters, const cc_cluster_point_t** points);

/* Optional: bring the per-cell labels of every push's new columns (columns [ground_from_gcol, ground_to_gcol) of
 * cc_batch_info_t) back together with the other results -- 4 bytes per cell in cell order: ground_point_label,
 * debug_ground_point_label, is_ignored, intensity. cc_get_column_labels returns a pointer into the handle's
 * page-locked buffer, valid until the next cc_wait / push. Saves a cc_read_columns round trip per push. */
CC_API cc_status_t cc_set_label_prefetch(cc_handle_t* h, int enable);
CC_API cc_status_t cc_get_column_labels(const cc_handle_t* h, const uint8_t** labels, int* n_cols);

/* Reads cells of columns [from_gcol, to_gcol] (inclusive, like the callback ranges) from the device
 * ring -- what a caller reads from `range_image_` inside a column callback (ros_utils.cpp:56-63,
 * kitti_demo.cpp:183-216). Only valid for columns still inside the ring. */
CC_API cc_status_t cc_read_columns(cc_handle_t* h, int64_t from_gcol, int64_t to_gcol,
                                   const cc_column_fields_t* fields);

/* The same cells as packed records, gathered by ONE kernel that writes straight into a page-locked buffer of the
 * handle (the publish side on the device): *cells points at (to - from + 1) * num_rows records, valid until
 * the next cc_export_columns / cc_read_columns call on the handle. This is what the facade fills `range_image_` from. */
CC_API cc_status_t cc_export_columns(cc_handle_t* h, int64_t from_gcol, int64_t to_gcol, const cc_cell_t** cells);

/* The sensor_msgs/PointCloud2 payloads the ROS node publishes, packed ON THE DEVICE byte for byte as the reference's
 * message conversion builds them (src/ros/ros_utils.cpp:11-77 columnToPointCloud / clusterToPointCloud, point layout
 * ros_utils.cpp:108-243, field values ros_utils.cpp:245-298): fields without padding, point_step 76 for the
 * ground-segmentation stage (ground_points_only callbacks) and 116 with the clustering fields. Column messages are
 * row-major images (height = num_rows, width = columns; point of (row, column) at index row * width + column) whose
 * header stamp is the smallest non-zero point stamp (0 if none); a cluster message is height 1 with the cluster's
 * stamp (cpp:1025-1028). `data` points into a page-locked buffer of the handle, valid until the next cc_pack_* call.
 * cc_pack_cluster_pointcloud2 takes the index of a cluster of the LAST finished push (cc_get_clusters order) and is
 * valid until the next push is submitted. */
typedef struct cc_cloud_view
{
    const uint8_t* data;
    uint64_t data_size;
    uint64_t stamp_ns;
    uint32_t point_step, width, height, n_fields;
} cc_cloud_view_t;
CC_API cc_status_t cc_pack_columns_pointcloud2(cc_handle_t* h, int64_t from_gcol, int64_t to_gcol, int ground_points_only,
                                               cc_cloud_view_t* out);
CC_API cc_status_t cc_pack_cluster_pointcloud2(cc_handle_t* h, int cluster_index, cc_cloud_view_t* out);
/* Every message of a push with ONE launch: request i is a range of columns (kind 0: ground_points_only callback, kind 1:
 * clustered columns) or cluster `cluster_index` of the last finished push (kind 2); out[i] receives its view (an empty
 * column range gives an empty view, like columnToPointCloud's nullptr). All payloads live in one page-locked buffer,
 * valid until the next cc_pack_* call. */
typedef struct cc_pack_request
{
    int32_t kind;
    int32_t cluster_index;
    int64_t from_gcol, to_gcol;
} cc_pack_request_t;
CC_API cc_status_t cc_pack_requests_pointcloud2(cc_handle_t* h, int n, const cc_pack_request_t* requests, cc_cloud_view_t* out);

/* Public data members of the reference object (hpp:244-251). */
CC_API int cc_num_rows(const cc_handle_t* h);
CC_API int cc_num_columns(const cc_handle_t* h);
CC_API int cc_ring_buffer_max_columns(const cc_handle_t* h);

/* The CUDA stream (cudaStream_t) the handle launches on, for callers that time with CUDA events. */
CC_API void* cc_stream(const cc_handle_t* h);
/* Total kernels launched by this handle so far. */
CC_API uint64_t cc_total_launches(const cc_handle_t* h);

/* Optional per-kernel timing (bench.py's roofline leg): when enabled, every kernel launch of the next pushes is
 * bracketed by CUDA events on the handle's stream; cc_get_kernel_timings returns, for the LAST push, the kernel
 * names (';'-separated, launch order) and their durations in milliseconds. */
CC_API cc_status_t cc_set_kernel_timing(cc_handle_t* h, int enable);
CC_API cc_status_t cc_get_kernel_timings(cc_handle_t* h, char* names, int names_cap, float* ms, int cap, int* n_out);

/* Test hook: treat every `period`-th column as if the association probe had flagged it, which routes the push
 * through the column-sequential exact kernels. Results must not change. 0 = off. */
CC_API cc_status_t cc_debug_flag_columns(cc_handle_t* h, int period);

/* Debug hook: device-side timeline. While enabled, thread 0 of every block of every kernel stamps its entry and exit
 * with the GPU's global nanosecond timer; cc_debug_get_trace reduces what the pushes since the last call left behind
 * to four values per kernel (first block entry, last block exit, longest single block, blocks seen; nanoseconds) and
 * the ';'-separated kernel names, then clears the buffer. Unlike cc_set_kernel_timing this does not serialise the
 * launches: it shows the kernels as they overlap in a normal push. */
CC_API cc_status_t cc_debug_trace(cc_handle_t* h, int enable);
CC_API cc_status_t cc_debug_get_trace(cc_handle_t* h, char* names, int names_cap, uint64_t* out, int cap_kernels, int* n_out);

/* Debug hook: host-visible milestones of the push that last used in-flight slot 0/1, in milliseconds since
 * cc_debug_slot_base(): input copy start, input copy end, first kernel, last kernel, results on the host. */
CC_API cc_status_t cc_debug_slot_base(cc_handle_t* h);
CC_API cc_status_t cc_debug_slot_times(cc_handle_t* h, int slot, float out_ms[5]);

/* Debug hook: 1 if the event `which` (0 start, 1 end of kernels, 2 state snapshot ready, 3 results on the host) of
 * in-flight slot 0/1 has completed. */
CC_API int cc_debug_event_query(cc_handle_t* h, int slot, int which);

/* ---- evaluation metrics ---------------------------------------------------------------
 * The per-frame metrics of the reference's KittiEvaluation (src/evaluation/kitti_evaluation.cpp:44-146) on the device:
 * ground-segmentation confusion counts against the SemanticKITTI ground classes (evaluateGroundPoints, cpp:44-84) and
 * the over- / under-segmentation entropies between ground-truth clusters and detections (evaluateClusters, cpp:86-146).
 * Per point: semantic_label (KittiPoint::semantic_label), is_ground_point, euclidean_clustering_label (0 = none) and
 * detection_label (Point::id, 0 = none) -- the members kitti_demo.cpp:211-213 fills. Host arrays in, six doubles out. */
typedef struct cc_eval cc_eval_t;
typedef struct cc_eval_result
{
    double tp, fn, fp, tn;                 /* EvaluationResultForFrame, kitti_evaluation.hpp:38-50 */
    double over_segmentation_entropy;
    double under_segmentation_entropy;
} cc_eval_result_t;
CC_API cc_status_t cc_eval_create(int device_ordinal, int max_points_per_frame, cc_eval_t** out);
CC_API void cc_eval_destroy(cc_eval_t* e);
CC_API cc_status_t cc_eval_frame(cc_eval_t* e, int n_points, const uint16_t* semantic_label, const uint8_t* is_ground_point,
                                 const uint32_t* euclidean_clustering_label, const uint32_t* detection_label,
                                 cc_eval_result_t* out);

/* ---- KITTI replay front-end -------------------------------------------------------------
 * The per-frame body of the reference's kitti_demo (src/tools/kitti_demo.cpp:369-403) on the device: one SemanticKITTI
 * frame (n points x {x, y, z, intensity} floats in file order) -> KittiLoader::recoverLaserIndices
 * (src/evaluation/kitti_loader.cpp:47-99) -> undoEgoMotionCorrection (kitti_loader.cpp:176-210) -> generateRangeImage
 * (kitti_loader.cpp:101-174) -> one pseudo firing per range-image column (makePseudoFiringFromRangeImageColumn,
 * kitti_demo.cpp:123-159) with its interpolated pose (KittiLoader::interpolate, kitti_loader.cpp:297-328). The 2200 x 64
 * RawPoint records stay in device memory, ready for cc_submit_firings_device / cc_push_firings_device (in chunks of at most
 * cc_max_firings_per_push firings: d_firings + k * 64, d_poses + k * 12). */
typedef struct cc_kitti cc_kitti_t;
typedef struct cc_kitti_frame
{
    int32_t n_firings;               /* KittiLoader::RANGE_IMAGE_WIDTH = 2200 */
    int32_t rows_per_firing;         /* KittiLoader::RANGE_IMAGE_HEIGHT = 64 */
    const cc_raw_point_t* d_firings; /* device: [n_firings][rows_per_firing] */
    const double* d_poses;           /* device: [n_firings][12] odom_from_velodyne at the firing's stamp */
    const double* poses;             /* the same on the host (page-locked), valid until the next cc_kitti_frame */
    int32_t rows_found;              /* laser_index + 1 of recoverLaserIndices: != 64 is the reference's "Wrong number of rows found" */
    int32_t max_points_in_row;       /* > 2200 is the reference's "More points in a single row than expected" (status != CC_OK) */
} cc_kitti_frame_t;
CC_API cc_status_t cc_kitti_create(int device_ordinal, int max_points_per_frame, cc_kitti_t** out);
CC_API void cc_kitti_destroy(cc_kitti_t* k);
/* The sequence's odom_from_velodyne transforms (3x4 row major) and their stamps, ascending (kitti_demo.cpp:331-341). */
CC_API cc_status_t cc_kitti_set_poses(cc_kitti_t* k, int n_poses, const uint64_t* stamps, const double* poses12);
/* frame_pose12: odom_from_velodyne at the middle of this frame's rotation (transforms_odom_from_velodyne[frame_index]). */
CC_API cc_status_t cc_kitti_frame(cc_kitti_t* k, int n_points, const float* xyzi, uint64_t stamp_start, uint64_t stamp_end,
                                  const double* frame_pose12, int sequence_index, int frame_index, cc_kitti_frame_t* out);
/* Intermediate results of the last frame for tests (any pointer may be NULL): laser_index[n], the point held by every
 * range-image cell [64 * 2200] (row major, -1 = empty), the un-corrected coordinates [n][3], the firings [2200][64]. */
CC_API cc_status_t cc_kitti_read_debug(cc_kitti_t* k, uint8_t* laser_index, int32_t* cell_point, float* uncorrected_xyz,
                                       cc_raw_point_t* firings);

/* ---- sensor packets -> firings -----------------------------------------------------------
 * OusterInput::onRawDataArrived (include/continuous_clustering/ros/ouster_input.hpp:105-181) for a batch of lidar UDP
 * packets on the device: every VALID measurement block of every packet becomes one firing of pixels_per_column RawPoints
 * (x, y, z from the sensor's lookup table, NaN for range 0; intensity = min(1, signal / 1000) * 255; firing_index counts
 * the published firings; stamp = the packet's receive time), resident in device memory for cc_submit_firings_device.
 * The packet layout is data: fill cc_ouster_format_t from ouster::sensor::packet_format (col_field offsets / masks of
 * ChanField::RANGE and ChanField::SIGNAL) or use cc_ouster_format_legacy. The lookup tables are the caller's
 * ouster::make_xyz_lut(info) cast to float and reordered column-major exactly like ouster_input.hpp:72-95:
 * [columns_per_frame * pixels_per_column][3]. The ouster SDK is not part of the reference tree (dependencies.repos:10-13):
 * parity for this row is pinned only to the restatement in oracle/cc_packets_oracle.cpp.
 * VelodyneInput (velodyne_input.hpp:46-91) delegates the whole decode to velodyne_rawdata::RawData::unpack of a
 * DataContainerBase variant that is neither in the reference tree nor published upstream (addPoint with a time argument +
 * newLine): not restated, see DESIGN.md. */
typedef struct cc_ouster_format
{
    int32_t columns_per_packet, pixels_per_column, columns_per_frame;
    int32_t packet_header_size, col_header_size, col_footer_size, pixel_bytes;
    int32_t col_measurement_id_offset; /* u16, from the start of the measurement block */
    int32_t col_status_offset;         /* from the start of the measurement block; bit 0 = valid (:120-124) */
    int32_t col_status_bytes;          /* 2 or 4 */
    int32_t range_offset, range_bytes; /* within the pixel; value = (little-endian word & mask) >> shift */
    uint32_t range_mask;               /* 0 = no mask */
    int32_t range_shift;
    int32_t signal_offset, signal_bytes;
    uint32_t signal_mask;
    int32_t signal_shift;
    int32_t offset_from_direction_table; /* 1 = like ouster_input.hpp:134, which cuts the offset block out of lut_direction [sic] */
} cc_ouster_format_t;
typedef struct cc_decoded_firings
{
    int32_t n_firings;
    int32_t rows_per_firing;
    const cc_raw_point_t* d_firings; /* device: [n_firings][rows_per_firing] */
    const uint64_t* firing_stamps;   /* host (page-locked): RawPoints::stamp of every firing (sensor_input.hpp:31) */
    uint64_t first_firing_index;
} cc_decoded_firings_t;
typedef struct cc_ouster cc_ouster_t;
CC_API void cc_ouster_format_legacy(int pixels_per_column, int columns_per_frame, cc_ouster_format_t* out);
CC_API cc_status_t cc_ouster_create(int device_ordinal, const cc_ouster_format_t* format, int max_packets_per_call, cc_ouster_t** out);
CC_API void cc_ouster_destroy(cc_ouster_t* o);
CC_API int cc_ouster_packet_size(const cc_ouster_t* o);
CC_API cc_status_t cc_ouster_set_lut(cc_ouster_t* o, const float* direction, const float* offset);
/* SensorInput::reset + OusterInput::reset (sensor_input.hpp:15-19, ouster_input.hpp:97-101): firing index 0, the next
 * packet is discarded. A new decoder starts in this state. */
CC_API cc_status_t cc_ouster_reset(cc_ouster_t* o);
CC_API cc_status_t cc_ouster_decode(cc_ouster_t* o, int n_packets, const uint8_t* packets, const uint64_t* receive_stamps,
                                    cc_decoded_firings_t* out);
CC_API cc_status_t cc_ouster_read_firings(cc_ouster_t* o, int n_firings, cc_raw_point_t* firings); /* tests */

/* ---- device math self-test (used by tests: bit-equality with host libm) ---------------- */
/* Evaluates the device re-implementations on n host inputs: out[i] = atan2f(a[i], b[i]) (op 0),
 * asinf(a[i]) (op 1). */
CC_API cc_status_t cc_selftest_math(int device_ordinal, int op, int n, const float* a, const float* b, float* out);

#ifdef __cplusplus
}
#endif
#endif /* CC_B200_H */
