// facade_bench.cpp -- measurement harness for the drop-in class (bench.py's facade legs): feeds a pre-built stream of
// firings into continuous_clustering::ContinuousClustering one addFiring call at a time, like the ROS node and
// kitti_demo do (continuous_clustering_node.cpp:163, kitti_demo.cpp:403), with callbacks that consume the results the
// way those callers do, and reports the wall time of every call. Product-side tool: links only the facade library.
#include <continuous_clustering/clustering/continuous_clustering.hpp>

#include <chrono>
#include <cstring>
#include <string>

#include "../../include/cc_b200.h"

using namespace continuous_clustering;

namespace
{
void toConfiguration(const cc_config_t& c, Configuration& o)
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
    g.ground_because_close_to_last_certain_ground_max_dist_diff = c.ground_because_close_to_last_certain_ground_max_dist_diff;
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

Eigen::Isometry3d poseFrom12(const double* m)
{
    Eigen::Isometry3d t = Eigen::Isometry3d::Identity();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++)
            t(i, j) = m[i * 4 + j];
    return t;
}
} // namespace

extern "C" {

// callback_mode 0: no callbacks registered; 1: the evaluation callback of kitti_demo.cpp:173-224 (reads
// globally_unique_point_index, ground_point_label and id of every cell of the clustered columns from range_image_) plus
// a cluster callback that walks the points; 2: the node's publishers on the packed path (packColumnsPointCloud2 for
// every ground / clustered column range + packed finished clusters, range_image_ not filled).
// call_us[n] (or null): wall time of every addFiring call in microseconds. result[8]: seconds for firings
// [warm_firings, n) incl. the final flush; column callbacks; cluster callbacks; points seen by cluster callbacks;
// checksum of what the callbacks read; bytes of packed messages; 0; 0. Returns 0, or 1 with the message in err[256].
__attribute__((visibility("default"))) int fb_run(const cc_config_t* cfg, int rows, const double* robot_from_sensor, int n,
                                                  const cc_raw_point_t* pts, const double* poses, int batch, int pipelined,
                                                  int callback_mode, int warm_firings, int device, double* call_us,
                                                  double* result, char* err)
{
    try
    {
        std::vector<RawPoints::Ptr> firings(n);
        std::vector<Eigen::Isometry3d> tfs(n);
        for (int k = 0; k < n; k++) // built before the clock starts, like the reference arm's drv_prepare
        {
            RawPoints::Ptr f(new RawPoints);
            f->points.resize(rows);
            std::memcpy(f->points.data(), pts + static_cast<size_t>(k) * rows, static_cast<size_t>(rows) * sizeof(RawPoint));
            f->stamp = rows > 0 ? f->points[0].stamp : 0;
            firings[k] = f;
            tfs[k] = poseFrom12(poses + 12 * k);
        }
        ContinuousClustering cc;
        cc.setDevice(device);
        Configuration config;
        toConfiguration(*cfg, config);
        cc.setConfiguration(config);
        cc.reset(rows);
        cc.setBatchSize(batch);
        cc.setPipelined(pipelined != 0);
        cc.setMaxBatchLatency(-1); // full batches only: the harness controls the batch size
        cc.setTransformRobotFrameFromSensorFrame(poseFrom12(robot_from_sensor));
        uint64_t checksum = 0, n_col_cb = 0, n_cluster_cb = 0, n_cluster_points = 0, packed_bytes = 0;
        if (callback_mode == 1)
        {
            cc.setFinishedColumnCallback(
                [&](int64_t from, int64_t to, bool ground_only)
                {
                    n_col_cb++;
                    if (ground_only)
                        return;
                    for (int64_t g = from; g <= to; g++)
                    {
                        const int local = static_cast<int>(g % cc.ring_buffer_max_columns);
                        for (int r = 0; r < cc.num_rows_; r++)
                        {
                            const Point& p = cc.range_image_[static_cast<size_t>(local) * cc.num_rows_ + r];
                            if (p.globally_unique_point_index != static_cast<uint64_t>(-1))
                                checksum += (p.ground_point_label == GP_GROUND ? 1 : 0) + p.id + (p.globally_unique_point_index & 0xff);
                        }
                    }
                });
            cc.setFinishedClusterCallback(
                [&](const std::vector<Point>& points, uint64_t stamp)
                {
                    n_cluster_cb++;
                    n_cluster_points += points.size();
                    checksum += stamp & 0xffff;
                });
        }
        else if (callback_mode == 2)
        {
            cc.setMaterialiseRangeImage(false);
            cc.setFinishedColumnCallback(
                [&](int64_t from, int64_t to, bool ground_only)
                {
                    n_col_cb++;
                    const ContinuousClustering::PackedPointCloud2 m = cc.packColumnsPointCloud2(from, to, ground_only);
                    packed_bytes += m.size;
                    if (m.size)
                        checksum += m.data[m.size - 1] + m.stamp_ns % 1000;
                });
            cc.setFinishedClusterPackedCallback(
                [&](const ContinuousClustering::PackedPointCloud2& m)
                {
                    n_cluster_cb++;
                    n_cluster_points += m.width;
                    packed_bytes += m.size;
                });
        }
        std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
        for (int k = 0; k < n; k++)
        {
            if (k == warm_firings)
            {
                cc.flush();
                cc.drain();
                t_start = std::chrono::steady_clock::now();
            }
            const auto t0 = std::chrono::steady_clock::now();
            cc.addFiring(firings[k], tfs[k]);
            if (call_us)
                call_us[k] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        }
        cc.flush();
        cc.drain();
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        result[0] = seconds;
        result[1] = static_cast<double>(n_col_cb);
        result[2] = static_cast<double>(n_cluster_cb);
        result[3] = static_cast<double>(n_cluster_points);
        result[4] = static_cast<double>(checksum % 1000000007ull);
        result[5] = static_cast<double>(packed_bytes);
        result[6] = result[7] = 0;
        return 0;
    }
    catch (const std::exception& e)
    {
        if (err)
        {
            std::strncpy(err, e.what(), 255);
            err[255] = 0;
        }
        return 1;
    }
}

} // extern "C"
