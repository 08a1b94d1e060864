// Drop-in replacement for the reference's clustering/continuous_clustering.hpp (hpp:1-293): same namespace, same
// configuration structs, same `Point`, same public methods and public data members of `ContinuousClustering`, so
// that continuous_clustering_node.cpp, ros_utils.cpp and kitti_demo.cpp compile and link against this library
// instead of the CPU implementation. Everything behind addFiring runs in CUDA kernels on a B200 through the C ABI
// of include/cc_b200.h; this class only batches firings, converts results and invokes the callbacks.
//
// Behavioural notes for integrators:
//  * firings are handed to the GPU in batches of `setBatchSize()` firings (default 64, env CC_B200_BATCH); callbacks
//    are delivered from inside the addFiring()/flush() call that completes a batch, on the caller's thread, in the
//    order of the reference's single-threaded mode. flush() pushes a partial batch. setPipelined(true) trades
//    callback latency (up to two batches) for throughput.
//  * range_image_ holds host copies of exactly the columns reported by finished-column callbacks (that is what the
//    reference's consumers read, ros_utils.cpp:56-63, kitti_demo.cpp:183-216).
#ifndef CONTINUOUS_CLUSTERING_CONTINUOUS_CLUSTERING_HPP
#define CONTINUOUS_CLUSTERING_CONTINUOUS_CLUSTERING_HPP

#include <cstdint>
#include <functional>
#include <list>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include <Eigen/Geometry>

#include <continuous_clustering/clustering/general.hpp>
#include <continuous_clustering/clustering/point_types.hpp>

struct cc_handle;

namespace continuous_clustering
{

enum
{
    GP_UNKNOWN = WHITE,
    GP_GROUND = GREEN,
    GP_OBSTACLE = RED,
    GP_EGO_VEHICLE = MAGENTA,
    GP_FOG = LIGHTGRAY,
};

struct GeneralConfiguration
{
    bool is_single_threaded{false};
};

struct ContinuousRangeImageConfiguration
{
    bool sensor_is_clockwise{true};
    int num_columns{1700};
    bool supplement_inclination_angle_for_nan_cells{true};
};

struct ContinuousGroundSegmentationConfiguration
{
    float max_slope{0.2};
    float first_ring_as_ground_max_allowed_z_diff{0.4};
    float first_ring_as_ground_min_allowed_z_diff{-0.4};
    float last_ground_point_slope_higher_than{-0.1};
    float last_ground_point_distance_smaller_than{5.};
    float ground_because_close_to_last_certain_ground_max_z_diff{0.4};
    float ground_because_close_to_last_certain_ground_max_dist_diff{2.0};
    float obstacle_because_next_certain_obstacle_max_dist_diff{0.3};
    bool use_terrain{false};
    float terrain_max_allowed_z_diff{0.4};
    float height_ref_to_maximum_{}, height_ref_to_ground_{};
    float length_ref_to_front_end_{}, length_ref_to_rear_end_{};
    float width_ref_to_left_mirror_{}, width_ref_to_right_mirror_{};
    bool fog_filtering_enabled{false};
    uint8_t fog_filtering_intensity_below{2};
    float fog_filtering_distance_below{18};
    float fog_filtering_inclination_above{-0.06};
};

struct ContinuousClusteringConfiguration
{
    float max_distance{0.7};
    int max_steps_in_row{20};
    int max_steps_in_column{20};
    bool stop_after_association_enabled{true};
    int stop_after_association_min_steps{1};
    bool ignore_points_in_chessboard_pattern{true};
    bool ignore_points_with_too_big_inclination_angle_diff{true};
    bool use_last_point_for_cluster_stamp{false};
    int cluster_point_trees_every_nth_column{1};
};

struct Configuration
{
    GeneralConfiguration general{};
    ContinuousRangeImageConfiguration range_image{};
    ContinuousGroundSegmentationConfiguration ground_segmentation{};
    ContinuousClusteringConfiguration clustering{};
};

class RangeImageIndex
{
  public:
    RangeImageIndex(uint16_t row, int64_t column) : column_index(column), row_index(row) {}
    bool operator==(const RangeImageIndex& o) const { return row_index == o.row_index && column_index == o.column_index; }
    bool operator!=(const RangeImageIndex& o) const { return !(*this == o); }
    bool operator<(const RangeImageIndex& o) const
    {
        return row_index < o.row_index || (row_index == o.row_index && column_index < o.column_index);
    }
    int64_t column_index{0};
    uint16_t row_index{0};
};

// One range-image cell as callers see it (hpp:126-161). The device keeps these fields as separate arrays; host
// copies are materialised for the columns / clusters handed to callbacks. child_points / associated_trees exist for
// source compatibility and stay empty (the device keeps trees as parent pointers + a union-find).
struct Point
{
    Point3D xyz{std::nanf(""), std::nanf(""), std::nanf("")};
    uint64_t firing_index{0};
    uint8_t intensity{0};
    float distance{std::nanf("")};
    float azimuth_angle{std::nanf("")};
    float inclination_angle{std::nanf("")};
    double continuous_azimuth_angle{std::nan("")};
    int64_t global_column_index{-1};
    int local_column_index{-1};
    int row_index{-1};
    uint64_t stamp{0};
    uint64_t globally_unique_point_index{static_cast<uint64_t>(-1)};

    uint8_t ground_point_label{0};
    float height_over_ground{std::nanf("")};
    uint8_t debug_ground_point_label{WHITE};

    bool is_ignored{false};
    double finished_at_continuous_azimuth_angle{0.f};
    std::list<RangeImageIndex> child_points{};
    std::set<RangeImageIndex> associated_trees{};
    RangeImageIndex tree_root_{0, -1};
    uint32_t tree_num_points{0};
    uint32_t cluster_width{0};
    uint64_t tree_id{0};
    uint64_t id{0};
    double visited_at_continuous_azimuth_angle{-1.};
    bool belongs_to_finished_cluster{false};
    int number_of_visited_neighbors{0};
};

class ContinuousClustering
{
  public:
    ContinuousClustering();
    ~ContinuousClustering();
    ContinuousClustering(const ContinuousClustering&) = delete;
    ContinuousClustering& operator=(const ContinuousClustering&) = delete;

    // general (hpp:205-207)
    void reset(int num_rows);
    void setConfiguration(const Configuration& config);
    bool resetRequired() const;

    // range image generation (hpp:210)
    void addFiring(const RawPoints::ConstPtr& firing, const Eigen::Isometry3d& odom_from_sensor);

    // ground point segmentation (hpp:213-214)
    void setTransformRobotFrameFromSensorFrame(const Eigen::Isometry3d& tf);
    bool hasTransformRobotFrameFromSensorFrame();

    // continuous clustering (hpp:217-218)
    void setFinishedColumnCallback(std::function<void(int64_t, int64_t, bool)> cb);
    void setFinishedClusterCallback(std::function<void(const std::vector<Point>&, uint64_t)> cb);

    // debugging (hpp:221): there are no job queues on the device; kept so callers link
    void recordJobQueueWorkload(size_t num_jobs_sensor_input);

    // ---- extensions of the B200 facade ----
    void flush();                    // push the firings buffered so far and deliver their callbacks
    void setBatchSize(int firings);  // firings per device push (1 = a push per addFiring call)
    // A partial batch is pushed by the addFiring call that finds its oldest firing waiting longer than this (default
    // 2000 us, env CC_B200_MAX_WAIT_US; negative = only full batches), by setTransformRobotFrameFromSensorFrame /
    // setConfiguration / reset (order relative to addFiring is the reference's) and by the destructor.
    void setMaxBatchLatency(int64_t microseconds);
    void setDevice(int ordinal);     // CUDA device of this stream; call before the first reset()
    // Throughput mode (off by default, env CC_B200_PIPELINE=1): a full batch is submitted asynchronously and its
    // callbacks are delivered by a later addFiring()/flush() call -- up to two batches late, still in order and on the
    // caller's thread -- so that the host (buffering, callbacks) and the device overlap. drain() delivers whatever is
    // still outstanding (call it at the end of a replay; reset() discards it like the reference's reset does).
    void setPipelined(bool on);
    void drain();
    // Publish side on the device: the sensor_msgs/PointCloud2 payloads the ROS node builds with
    // columnToPointCloud / clusterToPointCloud (ros_utils.cpp:11-77), packed by a kernel byte for byte (point layout
    // ros_utils.cpp:108-243; 76-byte points for ground_points_only, else 116) into page-locked memory. A node replaces
    // `msg = columnToPointCloud(clustering_, from, to, frame, stage)` by copying `data` into msg->data.
    // packColumnsPointCloud2 may be called from inside the finished-column callback; the view is valid until the next call.
    struct PackedPointCloud2
    {
        const uint8_t* data{nullptr};
        size_t size{0};
        uint32_t point_step{0}, width{0}, height{0}, n_fields{0};
        uint64_t stamp_ns{0}; // header stamp: smallest non-zero point stamp / the cluster stamp
    };
    PackedPointCloud2 packColumnsPointCloud2(int64_t from, int64_t to, bool ground_points_only);
    // finished clusters (> 20 points, cpp:1023) as packed messages instead of std::vector<Point>
    void setFinishedClusterPackedCallback(std::function<void(const PackedPointCloud2&)> cb);
    // false: range_image_ is not filled before the callbacks (for callers that only use the packed messages)
    void setMaterialiseRangeImage(bool on);

  public:
    // range image (implemented as ring buffer) -- public data members read by callers (hpp:244-251)
    int ring_buffer_max_columns{0};
    int num_columns_{};
    int num_rows_{-1};
    std::vector<Point> range_image_{0};
    int64_t ring_buffer_start_global_column_index{};
    int64_t ring_buffer_end_global_column_index{};

  private:
    void ensureHandle();
    void deliver();
    void materialise(int64_t from, int64_t to);
    [[noreturn]] void fail(int status);

    cc_handle* handle_{nullptr};
    int device_{0};
    int batch_size_{64};
    Configuration config_;
    bool config_dirty_{true};
    // firings buffered for the next push: cc_raw_point_t records + poses. Three buffers rotate in pipelined mode (a
    // buffer handed to cc_submit_firings must stay valid until its push has been waited for; at most two pushes are
    // outstanding when the next buffer is taken)
    std::vector<unsigned char> points_buf_[3];
    std::vector<double> poses_buf_[3];
    int cur_buf_{0};
    int pending_{0};
    int64_t max_wait_us_{2000};
    int64_t first_pending_us_{0};
    bool pipelined_{false};
    std::function<void(int64_t, int64_t, bool)> finished_column_callback_;
    std::function<void(const std::vector<Point>&, uint64_t)> finished_cluster_callback_;
    std::function<void(const PackedPointCloud2&)> finished_cluster_packed_callback_;
    bool materialise_{true};
    // packed messages of the push being delivered, built by ONE device launch before its callbacks run
    struct Prepacked
    {
        int64_t from, to;
        int kind;
        PackedPointCloud2 msg;
    };
    std::vector<Prepacked> prepacked_;
    void prepack(const void* events, int n_events, const void* clusters, int n_clusters);
    std::vector<Point> cluster_buffer_;
};

} // namespace continuous_clustering
#endif
