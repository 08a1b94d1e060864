// cc_driver.cpp -- TEST INFRASTRUCTURE (see cc_driver.h). Drives a
// continuous_clustering::ContinuousClustering object purely through its public class API
// (hpp:197-251), exactly like kitti_demo.cpp:276-313,403 and continuous_clustering_node.cpp:34-38,163
// do, and records what the callbacks report. Compiled against the reference (oracle/_ref) or the
// facade (facade/) -- the source is identical, which is the drop-in proof.
#include "cc_driver.h"

#include <continuous_clustering/clustering/continuous_clustering.hpp>

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#ifndef DRV_IMPL_NAME
#define DRV_IMPL_NAME "unknown"
#endif

#ifdef DRV_WITH_CALLER_EXCERPTS
// The reference's own caller code, cut out of /root/reference at build time (oracle/extract_caller_excerpts.py ->
// oracle/_ref/*.inc, never committed) and compiled UNMODIFIED against whichever class header this driver is built
// with: against the facade's headers this is the compile-time half of the drop-in proof (ros_utils.cpp:245-298 reads
// every Point field the ROS node publishes, kitti_demo.cpp:173-224 is the evaluation callback), and running it on
// both objects is the oracle for the published PointCloud2 bytes.
#include <map>
#include <optional>

#include <sensor_msgs/PointCloud2.h>
#include <sensor_msgs/point_cloud2_iterator.h>

namespace continuous_clustering
{
#include "ros_utils_types.inc"
PointCloud2Iterators prepareMessageAndCreateIterators(sensor_msgs::PointCloud2& msg, ProcessingStage fill_fields_up_to_stage);
void addPointToMessage(PointCloud2Iterators& container, int data_index_message, const Point& point, int num_rows,
                       ProcessingStage fill_fields_up_to_stage);
#include "ros_utils_clouds.inc"
#include "ros_utils_fields.inc"

struct KittiSegmentationEvaluationPoint // the three members the callback writes (kitti_evaluation.hpp:18-36)
{
    bool has_corresponding_point_in_detection_point_cloud{false};
    bool is_ground_point{false};
    uint32_t detection_label{0};
};

struct KittiDemoCallback // the members of class KittiDemo the callback body touches (kitti_demo.cpp:444-448)
{
    std::map<std::pair<uint16_t, uint16_t>, std::vector<KittiSegmentationEvaluationPoint>> map_frame_to_point_cloud;
    int current_sequence_index{0};
    int previous_frame_index{0};
    std::map<int, std::vector<KittiSegmentationEvaluationPoint>> evaluated; // by frame
    void evaluatePreviousFrame() // kitti_demo.cpp:160-171 without the evaluation itself: keeps the labelled frame
    {
        auto it = map_frame_to_point_cloud.find({static_cast<uint16_t>(current_sequence_index), static_cast<uint16_t>(previous_frame_index)});
        if (it != map_frame_to_point_cloud.end())
        {
            evaluated[previous_frame_index] = std::move(it->second);
            map_frame_to_point_cloud.erase(it);
        }
        previous_frame_index++;
    }
#include "kitti_demo_callback.inc"
};
} // namespace continuous_clustering
#endif

using namespace continuous_clustering;

struct drv
{
    ContinuousClustering cc;
    Configuration config;
    int num_rows{0};
    int record{DRV_RECORD_FULL};
    std::string error;

    std::mutex mutex; // callbacks may arrive on worker threads in multi-threaded mode
    std::vector<cc_column_event_t> events;
    std::vector<int64_t> ground_cols, cluster_cols;
    std::vector<drv_cell_t> ground_cells, cluster_cells;
    std::vector<drv_cluster_t> clusters;
    std::vector<drv_cluster_point_t> cluster_points;
    std::atomic<int64_t> last_clustered_column{-1};
    std::atomic<int64_t> num_column_callbacks{0};

    std::vector<RawPoints::Ptr> prepared;
    std::vector<Eigen::Isometry3d> prepared_poses;

    bool record_clouds{false};
    std::vector<drv_cloud_t> clouds;
    std::vector<uint8_t> cloud_data;
#ifdef DRV_WITH_CALLER_EXCERPTS
    bool kitti{false};
    KittiDemoCallback kitti_cb;
#endif
};

#ifdef DRV_WITH_CALLER_EXCERPTS
static void keepCloud(drv* d, const sensor_msgs::PointCloud2Ptr& msg, int kind, int64_t from, int64_t to)
{
    if (!msg)
        return;
    drv_cloud_t c;
    std::memset(&c, 0, sizeof(c));
    c.from_gcol = from;
    c.to_gcol = to;
    c.kind = kind;
    c.width = msg->width;
    c.height = msg->height;
    c.point_step = msg->point_step;
    c.stamp_ns = msg->header.stamp.toNSec();
    c.data_offset = static_cast<int64_t>(d->cloud_data.size());
    c.data_size = static_cast<int64_t>(msg->data.size());
    d->cloud_data.insert(d->cloud_data.end(), msg->data.begin(), msg->data.end());
    d->clouds.push_back(c);
}
#endif

static Eigen::Isometry3d poseFrom12(const double* m)
{
    Eigen::Isometry3d t = Eigen::Isometry3d::Identity();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++)
            t(i, j) = m[i * 4 + j];
    return t;
}

static void snapshotCell(const ContinuousClustering& cc, const Point& p, drv_cell_t& c)
{
    std::memset(&c, 0, sizeof(c));
    c.continuous_azimuth_angle = p.continuous_azimuth_angle;
    c.global_column_index = p.global_column_index;
    c.globally_unique_point_index = p.globally_unique_point_index;
    c.stamp = p.stamp;
    c.firing_index = p.firing_index;
    c.id = p.id;
    c.x = p.xyz.x;
    c.y = p.xyz.y;
    c.z = p.xyz.z;
    c.distance = p.distance;
    c.azimuth_angle = p.azimuth_angle;
    c.inclination_angle = p.inclination_angle;
    c.number_of_visited_neighbors = p.number_of_visited_neighbors;
    c.intensity = p.intensity;
    c.ground_point_label = p.ground_point_label;
    c.debug_ground_point_label = p.debug_ground_point_label;
    c.is_ignored = p.is_ignored ? 1 : 0;
    c.num_child_points = static_cast<int32_t>(p.child_points.size());
    c.finished_at_continuous_azimuth_angle = p.finished_at_continuous_azimuth_angle;
    c.tree_num_points = p.tree_num_points;
    c.cluster_width = p.cluster_width;
    c.local_column_index = p.local_column_index;
    c.row_index = p.row_index;
    c.tree_root_row = p.tree_root_.row_index;
    c.tree_root_gcol = -1;
    if (p.tree_root_.column_index >= 0)
    {
        // tree_root_ holds the LOCAL ring column (cpp:661, 814); report the root's global column
        const Point& root = cc.range_image_[p.tree_root_.column_index * cc.num_rows_ + p.tree_root_.row_index];
        c.tree_root_gcol = root.global_column_index;
    }
}

static void onColumns(drv* d, int64_t from, int64_t to, bool ground_only)
{
    d->num_column_callbacks++;
    if (!ground_only)
        d->last_clustered_column = to;
#ifdef DRV_WITH_CALLER_EXCERPTS
    if (d->record_clouds || d->kitti)
    {
        std::lock_guard<std::mutex> lock(d->mutex);
        if (d->record_clouds) // continuous_clustering_node.cpp:171-178
            keepCloud(d, columnToPointCloud(d->cc, from, to, "odom", ground_only ? GROUND_POINT_SEGMENTATION : CONTINUOUS_CLUSTERING),
                      ground_only ? 0 : 1, from, to);
        if (d->kitti && !ground_only) // kitti_demo.cpp:297-306
            d->kitti_cb.addColumnAndEvaluateFrameIfCompleted(d->cc, from, to);
    }
#endif
    if (d->record == DRV_RECORD_NONE)
        return;
    std::lock_guard<std::mutex> lock(d->mutex);
    cc_column_event_t ev;
    ev.from_gcol = from;
    ev.to_gcol = to;
    ev.ground_points_only = ground_only ? 1 : 0;
    ev.n_clusters_before = static_cast<int32_t>(d->clusters.size());
    d->events.push_back(ev);
    if (d->record < DRV_RECORD_FULL)
        return;
    const ContinuousClustering& cc = d->cc;
    for (int64_t g = from; g <= to; g++)
    {
        // same indexing as ros_utils.cpp:56-63 / kitti_demo.cpp:183-193
        int local = static_cast<int>(g % cc.ring_buffer_max_columns);
        std::vector<drv_cell_t>& cells = ground_only ? d->ground_cells : d->cluster_cells;
        (ground_only ? d->ground_cols : d->cluster_cols).push_back(g);
        size_t base = cells.size();
        cells.resize(base + cc.num_rows_);
        for (int r = 0; r < cc.num_rows_; r++)
            snapshotCell(cc, cc.range_image_[local * cc.num_rows_ + r], cells[base + r]);
    }
}

static void onCluster(drv* d, const std::vector<Point>& points, uint64_t stamp)
{
#ifdef DRV_WITH_CALLER_EXCERPTS
    if (d->record_clouds) // continuous_clustering_node.cpp:166-169
    {
        std::lock_guard<std::mutex> lock(d->mutex);
        keepCloud(d, clusterToPointCloud(points, d->cc.num_rows_, stamp, "odom"), 2, -1, -1);
    }
#endif
    if (d->record == DRV_RECORD_NONE)
        return;
    std::lock_guard<std::mutex> lock(d->mutex);
    drv_cluster_t c;
    c.stamp = stamp;
    c.id = points.empty() ? 0 : points.front().id;
    c.point_offset = static_cast<int64_t>(d->cluster_points.size());
    c.num_points = static_cast<int64_t>(points.size());
    c.event_index = static_cast<int64_t>(d->events.size());
    d->clusters.push_back(c);
    for (const Point& p : points)
    {
        drv_cluster_point_t cp;
        cp.gcol = p.global_column_index;
        cp.globally_unique_point_index = p.globally_unique_point_index;
        cp.row = p.row_index;
        cp.pad_ = 0;
        d->cluster_points.push_back(cp);
    }
}

static void toConfiguration(const cc_config_t& c, Configuration& o)
{
    o.general.is_single_threaded = c.is_single_threaded != 0;
    o.range_image.sensor_is_clockwise = c.sensor_is_clockwise != 0;
    o.range_image.num_columns = c.num_columns;
    o.range_image.supplement_inclination_angle_for_nan_cells = c.supplement_inclination_angle_for_nan_cells != 0;
    auto& g = o.ground_segmentation;
    g.max_slope = c.max_slope;
    g.first_ring_as_ground_max_allowed_z_diff = c.first_ring_as_ground_max_allowed_z_diff;
    g.first_ring_as_ground_min_allowed_z_diff = c.first_ring_as_ground_min_allowed_z_diff;
    g.last_ground_point_slope_higher_than = c.last_ground_point_slope_higher_than;
    g.last_ground_point_distance_smaller_than = c.last_ground_point_distance_smaller_than;
    g.ground_because_close_to_last_certain_ground_max_z_diff = c.ground_because_close_to_last_certain_ground_max_z_diff;
    g.ground_because_close_to_last_certain_ground_max_dist_diff =
        c.ground_because_close_to_last_certain_ground_max_dist_diff;
    g.obstacle_because_next_certain_obstacle_max_dist_diff = c.obstacle_because_next_certain_obstacle_max_dist_diff;
    g.use_terrain = c.use_terrain != 0;
    g.terrain_max_allowed_z_diff = c.terrain_max_allowed_z_diff;
    g.height_ref_to_maximum_ = c.height_ref_to_maximum_;
    g.height_ref_to_ground_ = c.height_ref_to_ground_;
    g.length_ref_to_front_end_ = c.length_ref_to_front_end_;
    g.length_ref_to_rear_end_ = c.length_ref_to_rear_end_;
    g.width_ref_to_left_mirror_ = c.width_ref_to_left_mirror_;
    g.width_ref_to_right_mirror_ = c.width_ref_to_right_mirror_;
    g.fog_filtering_enabled = c.fog_filtering_enabled != 0;
    g.fog_filtering_intensity_below = static_cast<uint8_t>(c.fog_filtering_intensity_below);
    g.fog_filtering_distance_below = c.fog_filtering_distance_below;
    g.fog_filtering_inclination_above = c.fog_filtering_inclination_above;
    auto& k = o.clustering;
    k.max_distance = c.max_distance;
    k.max_steps_in_row = c.max_steps_in_row;
    k.max_steps_in_column = c.max_steps_in_column;
    k.stop_after_association_enabled = c.stop_after_association_enabled != 0;
    k.stop_after_association_min_steps = c.stop_after_association_min_steps;
    k.ignore_points_in_chessboard_pattern = c.ignore_points_in_chessboard_pattern != 0;
    k.ignore_points_with_too_big_inclination_angle_diff = c.ignore_points_with_too_big_inclination_angle_diff != 0;
    k.use_last_point_for_cluster_stamp = c.use_last_point_for_cluster_stamp != 0;
    k.cluster_point_trees_every_nth_column = c.cluster_point_trees_every_nth_column;
}

static RawPoints::Ptr makeFiring(int rows, const cc_raw_point_t* pts)
{
    static_assert(sizeof(RawPoint) == sizeof(cc_raw_point_t), "RawPoint layout");
    RawPoints::Ptr firing(new RawPoints);
    firing->stamp = rows > 0 ? pts[0].stamp : 0;
    firing->points.resize(rows);
    for (int r = 0; r < rows; r++)
    {
        RawPoint& p = firing->points[r];
        p.x = pts[r].x;
        p.y = pts[r].y;
        p.z = pts[r].z;
        p.firing_index = pts[r].firing_index;
        p.intensity = pts[r].intensity;
        p.stamp = pts[r].stamp;
        p.globally_unique_point_index = pts[r].globally_unique_point_index;
    }
    return firing;
}

extern "C" {

const char* drv_impl_name(void)
{
    return DRV_IMPL_NAME;
}

drv_t* drv_create(void)
{
    drv* d = new drv();
    d->cc.setFinishedColumnCallback([d](int64_t from, int64_t to, bool ground_only)
                                    { onColumns(d, from, to, ground_only); });
    d->cc.setFinishedClusterCallback([d](const std::vector<Point>& points, uint64_t stamp)
                                     { onCluster(d, points, stamp); });
    return d;
}

void drv_destroy(drv_t* d)
{
    delete d;
}

const char* drv_last_error(drv_t* d)
{
    return d->error.c_str();
}

int drv_configure(drv_t* d, const cc_config_t* cfg, int num_rows, const double* robot_from_sensor)
{
    try
    {
        toConfiguration(*cfg, d->config);
        d->cc.setConfiguration(d->config);
        d->cc.reset(num_rows);
        d->num_rows = num_rows;
        if (robot_from_sensor)
            d->cc.setTransformRobotFrameFromSensorFrame(poseFrom12(robot_from_sensor));
    }
    catch (const std::exception& e)
    {
        d->error = e.what();
        return 1;
    }
    return 0;
}

// setConfiguration without reset (cpp:66-81), as a caller may do mid-stream (node.cpp:234)
int drv_set_config(drv_t* d, const cc_config_t* cfg)
{
    toConfiguration(*cfg, d->config);
    d->cc.setConfiguration(d->config);
    return 0;
}

// the reference does not count the associations it refuses: only the restatement (cc_oracle.cpp) reports them
void drv_refusals(drv_t*, int64_t* joins, int64_t* links)
{
    *joins = -1;
    *links = -1;
}

void drv_set_record(drv_t* d, int level)
{
    d->record = level;
}

int drv_add_firings(drv_t* d, int n, int rows, const cc_raw_point_t* pts, const double* poses)
{
    try
    {
        for (int k = 0; k < n; k++)
            d->cc.addFiring(makeFiring(rows, pts + static_cast<size_t>(k) * rows), poseFrom12(poses + 12 * k));
#ifdef CC_B200_FACADE
        d->cc.flush(); // the facade batches firings; the reference delivers everything inside addFiring
        d->cc.drain(); // and, in its throughput mode, delivers a batch's callbacks up to two batches late
#endif
    }
    catch (const std::exception& e)
    {
        d->error = e.what();
        return 1;
    }
    return 0;
}

int drv_reset_required(drv_t* d)
{
    return d->cc.resetRequired() ? 1 : 0;
}

int drv_num_rows(drv_t* d)
{
    return d->cc.num_rows_;
}

int drv_ring_buffer_max_columns(drv_t* d)
{
    return d->cc.ring_buffer_max_columns;
}

int drv_prepare(drv_t* d, int n, int rows, const cc_raw_point_t* pts, const double* poses)
{
    d->prepared.clear();
    d->prepared_poses.clear();
    d->prepared.reserve(n);
    d->prepared_poses.reserve(n);
    for (int k = 0; k < n; k++)
    {
        d->prepared.push_back(makeFiring(rows, pts + static_cast<size_t>(k) * rows));
        d->prepared_poses.push_back(poseFrom12(poses + 12 * k));
    }
    return 0;
}

double drv_run_prepared(drv_t* d, int from, int to, int64_t max_lag_columns)
{
    if (from < 0 || to > static_cast<int>(d->prepared.size()) || from > to)
    {
        d->error = "drv_run_prepared: bad range";
        return -1.;
    }
    const bool throttle = !d->config.general.is_single_threaded && max_lag_columns > 0;
    auto t0 = std::chrono::steady_clock::now();
    try
    {
        int64_t fed_base = d->last_clustered_column.load();
        for (int k = from; k < to; k++)
        {
            d->cc.addFiring(d->prepared[k], d->prepared_poses[k]);
            if (throttle && (k & 15) == 0)
            {
                // fed - finished <= max_lag_columns  (each firing advances the front by ~1 column)
                while ((fed_base + (k - from)) - d->last_clustered_column.load() > max_lag_columns)
                    std::this_thread::yield();
            }
        }
#ifdef CC_B200_FACADE
        d->cc.flush();
        d->cc.drain();
#endif
        if (!d->config.general.is_single_threaded)
        {
            // drain: wait until no callback has arrived for a while
            int64_t seen = -1;
            int idle = 0;
            while (idle < 20)
            {
                int64_t now = d->num_column_callbacks.load();
                idle = (now == seen) ? idle + 1 : 0;
                seen = now;
                std::this_thread::sleep_for(std::chrono::microseconds(500));
            }
        }
    }
    catch (const std::exception& e)
    {
        d->error = e.what();
        return -1.;
    }
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    if (!d->config.general.is_single_threaded)
        s -= 20 * 500e-6; // the idle detection window is not pipeline work
    return s;
}

int drv_run_prepared_latency(drv_t* d, int from, int to, double* out_us)
{
    if (from < 0 || to > static_cast<int>(d->prepared.size()) || from > to)
    {
        d->error = "drv_run_prepared_latency: bad range";
        return 1;
    }
    try
    {
        for (int k = from; k < to; k++)
        {
            const auto t0 = std::chrono::steady_clock::now();
            d->cc.addFiring(d->prepared[k], d->prepared_poses[k]);
            out_us[k - from] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        }
#ifdef CC_B200_FACADE
        d->cc.flush();
        d->cc.drain();
#endif
    }
    catch (const std::exception& e)
    {
        d->error = e.what();
        return 1;
    }
    return 0;
}

void drv_set_callbacks(drv_t* d, int callbacks)
{
    if (callbacks)
    {
        d->cc.setFinishedColumnCallback([d](int64_t from, int64_t to, bool ground_only) { onColumns(d, from, to, ground_only); });
        d->cc.setFinishedClusterCallback([d](const std::vector<Point>& points, uint64_t stamp) { onCluster(d, points, stamp); });
    }
    else
    {
        d->cc.setFinishedColumnCallback(nullptr);
        d->cc.setFinishedClusterCallback(nullptr);
    }
}

int64_t drv_num_events(drv_t* d)
{
    return static_cast<int64_t>(d->events.size());
}

void drv_get_events(drv_t* d, cc_column_event_t* out)
{
    std::memcpy(out, d->events.data(), d->events.size() * sizeof(cc_column_event_t));
}

int64_t drv_num_ground_columns(drv_t* d)
{
    return static_cast<int64_t>(d->ground_cols.size());
}

void drv_get_ground_columns(drv_t* d, int64_t* gcols, drv_cell_t* cells)
{
    std::memcpy(gcols, d->ground_cols.data(), d->ground_cols.size() * sizeof(int64_t));
    std::memcpy(cells, d->ground_cells.data(), d->ground_cells.size() * sizeof(drv_cell_t));
}

int64_t drv_num_cluster_columns(drv_t* d)
{
    return static_cast<int64_t>(d->cluster_cols.size());
}

void drv_get_cluster_columns(drv_t* d, int64_t* gcols, drv_cell_t* cells)
{
    std::memcpy(gcols, d->cluster_cols.data(), d->cluster_cols.size() * sizeof(int64_t));
    std::memcpy(cells, d->cluster_cells.data(), d->cluster_cells.size() * sizeof(drv_cell_t));
}

int64_t drv_num_clusters(drv_t* d)
{
    return static_cast<int64_t>(d->clusters.size());
}

int64_t drv_num_cluster_points(drv_t* d)
{
    return static_cast<int64_t>(d->cluster_points.size());
}

void drv_get_clusters(drv_t* d, drv_cluster_t* clusters, drv_cluster_point_t* points)
{
    std::memcpy(clusters, d->clusters.data(), d->clusters.size() * sizeof(drv_cluster_t));
    std::memcpy(points, d->cluster_points.data(), d->cluster_points.size() * sizeof(drv_cluster_point_t));
}

void drv_clear_records(drv_t* d)
{
    std::lock_guard<std::mutex> lock(d->mutex);
    d->events.clear();
    d->ground_cols.clear();
    d->cluster_cols.clear();
    d->ground_cells.clear();
    d->cluster_cells.clear();
    d->clusters.clear();
    d->cluster_points.clear();
    d->clouds.clear();
    d->cloud_data.clear();
}

int drv_has_caller_excerpts(void)
{
#ifdef DRV_WITH_CALLER_EXCERPTS
    return 1;
#else
    return 0;
#endif
}

void drv_set_cloud_record(drv_t* d, int on)
{
    d->record_clouds = on != 0;
}

int64_t drv_num_clouds(drv_t* d)
{
    return static_cast<int64_t>(d->clouds.size());
}

int64_t drv_cloud_bytes(drv_t* d)
{
    return static_cast<int64_t>(d->cloud_data.size());
}

void drv_get_clouds(drv_t* d, drv_cloud_t* clouds, uint8_t* data)
{
    std::memcpy(clouds, d->clouds.data(), d->clouds.size() * sizeof(drv_cloud_t));
    std::memcpy(data, d->cloud_data.data(), d->cloud_data.size());
}

void drv_kitti_begin(drv_t* d, int sequence, int n_frames, const int32_t* points_per_frame)
{
#ifdef DRV_WITH_CALLER_EXCERPTS
    d->kitti = true;
    d->kitti_cb = KittiDemoCallback();
    d->kitti_cb.current_sequence_index = sequence; // kitti_demo.cpp:316-317
    d->kitti_cb.previous_frame_index = 0;
    for (int f = 0; f < n_frames; f++) // kitti_demo.cpp:361-365
        d->kitti_cb.map_frame_to_point_cloud.insert(
            {{static_cast<uint16_t>(sequence), static_cast<uint16_t>(f)},
             std::vector<KittiSegmentationEvaluationPoint>(static_cast<size_t>(points_per_frame[f]))});
#else
    (void)d, (void)sequence, (void)n_frames, (void)points_per_frame;
#endif
}

int64_t drv_kitti_get(drv_t* d, int frame, uint8_t* flags, uint32_t* detection_label)
{
#ifdef DRV_WITH_CALLER_EXCERPTS
    // a frame the callback has not closed yet is still in the map (kitti_demo evaluates the last frame at the end)
    const std::vector<KittiSegmentationEvaluationPoint>* v = nullptr;
    auto e = d->kitti_cb.evaluated.find(frame);
    if (e != d->kitti_cb.evaluated.end())
        v = &e->second;
    else
    {
        auto it = d->kitti_cb.map_frame_to_point_cloud.find(
            {static_cast<uint16_t>(d->kitti_cb.current_sequence_index), static_cast<uint16_t>(frame)});
        if (it != d->kitti_cb.map_frame_to_point_cloud.end())
            v = &it->second;
    }
    if (!v)
        return -1;
    if (flags && detection_label)
        for (size_t i = 0; i < v->size(); i++)
        {
            flags[i] = ((*v)[i].has_corresponding_point_in_detection_point_cloud ? 1 : 0) | ((*v)[i].is_ground_point ? 2 : 0);
            detection_label[i] = (*v)[i].detection_label;
        }
    return static_cast<int64_t>(v->size());
#else
    (void)d, (void)frame, (void)flags, (void)detection_label;
    return -1;
#endif
}

} // extern "C"
