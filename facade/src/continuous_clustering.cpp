// B200 facade: the reference's ContinuousClustering class API (hpp:197-251) on top of the CUDA C ABI.
// Replaces src/clustering/continuous_clustering.cpp of the reference at link time. No pipeline stage is computed here.
#include <continuous_clustering/clustering/continuous_clustering.hpp>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../../include/cc_b200.h"

namespace continuous_clustering
{

static_assert(sizeof(RawPoint) == sizeof(cc_raw_point_t), "RawPoint must stay layout-identical to cc_raw_point_t");

static void toC(const Configuration& c, cc_config_t& o)
{
    cc_config_default(&o);
    o.is_single_threaded = c.general.is_single_threaded;
    o.sensor_is_clockwise = c.range_image.sensor_is_clockwise;
    o.num_columns = c.range_image.num_columns;
    o.supplement_inclination_angle_for_nan_cells = c.range_image.supplement_inclination_angle_for_nan_cells;
    const auto& g = c.ground_segmentation;
    o.max_slope = g.max_slope;
    o.first_ring_as_ground_max_allowed_z_diff = g.first_ring_as_ground_max_allowed_z_diff;
    o.first_ring_as_ground_min_allowed_z_diff = g.first_ring_as_ground_min_allowed_z_diff;
    o.last_ground_point_slope_higher_than = g.last_ground_point_slope_higher_than;
    o.last_ground_point_distance_smaller_than = g.last_ground_point_distance_smaller_than;
    o.ground_because_close_to_last_certain_ground_max_z_diff = g.ground_because_close_to_last_certain_ground_max_z_diff;
    o.ground_because_close_to_last_certain_ground_max_dist_diff = g.ground_because_close_to_last_certain_ground_max_dist_diff;
    o.obstacle_because_next_certain_obstacle_max_dist_diff = g.obstacle_because_next_certain_obstacle_max_dist_diff;
    o.use_terrain = g.use_terrain;
    o.terrain_max_allowed_z_diff = g.terrain_max_allowed_z_diff;
    o.height_ref_to_maximum_ = g.height_ref_to_maximum_;
    o.height_ref_to_ground_ = g.height_ref_to_ground_;
    o.length_ref_to_front_end_ = g.length_ref_to_front_end_;
    o.length_ref_to_rear_end_ = g.length_ref_to_rear_end_;
    o.width_ref_to_left_mirror_ = g.width_ref_to_left_mirror_;
    o.width_ref_to_right_mirror_ = g.width_ref_to_right_mirror_;
    o.fog_filtering_enabled = g.fog_filtering_enabled;
    o.fog_filtering_intensity_below = g.fog_filtering_intensity_below;
    o.fog_filtering_distance_below = g.fog_filtering_distance_below;
    o.fog_filtering_inclination_above = g.fog_filtering_inclination_above;
    const auto& k = c.clustering;
    o.max_distance = k.max_distance;
    o.max_steps_in_row = k.max_steps_in_row;
    o.max_steps_in_column = k.max_steps_in_column;
    o.stop_after_association_enabled = k.stop_after_association_enabled;
    o.stop_after_association_min_steps = k.stop_after_association_min_steps;
    o.ignore_points_in_chessboard_pattern = k.ignore_points_in_chessboard_pattern;
    o.ignore_points_with_too_big_inclination_angle_diff = k.ignore_points_with_too_big_inclination_angle_diff;
    o.use_last_point_for_cluster_stamp = k.use_last_point_for_cluster_stamp;
    o.cluster_point_trees_every_nth_column = k.cluster_point_trees_every_nth_column;
}

static void pose12(const Eigen::Isometry3d& t, double* m)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++)
            m[i * 4 + j] = t(i, j);
}

ContinuousClustering::ContinuousClustering()
{
    if (const char* b = std::getenv("CC_B200_BATCH"))
        batch_size_ = std::max(1, std::atoi(b));
    if (const char* d = std::getenv("CC_B200_DEVICE"))
        device_ = std::atoi(d);
    if (const char* q = std::getenv("CC_B200_PIPELINE"))
        pipelined_ = std::atoi(q) != 0;
    if (const char* w = std::getenv("CC_B200_MAX_WAIT_US"))
        max_wait_us_ = std::atoll(w);
}

static int64_t nowMicros()
{
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

ContinuousClustering::~ContinuousClustering()
{
    // an unchanged caller (kitti_demo, the ROS node) never calls flush(): whatever is still buffered or in flight is
    // processed and delivered here, like the reference's worker threads finish their queues before the object dies
    try
    {
        flush();
        drain();
    }
    catch (...)
    {
    }
    if (handle_)
        cc_destroy(handle_);
}

void ContinuousClustering::fail(int status)
{
    // the reference reports these conditions as std::runtime_error (cpp:90-91, 298-299, 337-344, 1072-1075)
    std::string msg = handle_ ? cc_last_error(handle_) : "no CUDA device";
    if (msg.empty())
        msg = "continuous_clustering_b200 error " + std::to_string(status);
    throw std::runtime_error(msg);
}

void ContinuousClustering::ensureHandle()
{
    if (handle_)
        return;
    int s = cc_create(device_, std::max(4096, batch_size_), &handle_);
    if (s != CC_OK)
    {
        handle_ = nullptr;
        throw std::runtime_error("continuous_clustering_b200: no usable CUDA device (the hot path has no CPU fallback)");
    }
}

void ContinuousClustering::setDevice(int ordinal)
{
    device_ = ordinal;
}

void ContinuousClustering::setBatchSize(int firings)
{
    flush();
    batch_size_ = std::max(1, std::min(firings, 4096));
}

void ContinuousClustering::setPipelined(bool on)
{
    flush();
    drain();
    pipelined_ = on;
}

void ContinuousClustering::drain()
{
    while (handle_ && cc_pending(handle_) > 0)
    {
        int s = cc_wait(handle_);
        if (s != CC_OK)
            fail(s);
        deliver();
    }
}

void ContinuousClustering::setConfiguration(const Configuration& config)
{
    ensureHandle();
    flush(); // firings already handed over were processed with the previous parameters
    drain();
    config_ = config;
    cc_config_t c;
    toC(config, c);
    int s = cc_set_config(handle_, &c);
    if (s != CC_OK)
        fail(s);
}

bool ContinuousClustering::resetRequired() const
{
    return handle_ && cc_reset_required(handle_);
}

void ContinuousClustering::reset(int num_rows)
{
    ensureHandle();
    // firings handed over before reset() were already processed by the reference at this point: push them first (an
    // error they raise would have escaped an earlier addFiring call there; here it is dropped with the state)
    try
    {
        flush();
        drain();
    }
    catch (...)
    {
    }
    pending_ = 0;
    for (int b = 0; b < 3; b++)
    {
        points_buf_[b].clear();
        poses_buf_[b].clear();
    }
    int s = cc_reset(handle_, num_rows); // pushes still in flight are discarded with the rest of the state
    if (s != CC_OK)
        fail(s);
    num_rows_ = cc_num_rows(handle_);
    num_columns_ = cc_num_columns(handle_);
    ring_buffer_max_columns = cc_ring_buffer_max_columns(handle_);
    range_image_.assign(static_cast<size_t>(ring_buffer_max_columns) * num_rows_, Point{});
    ring_buffer_start_global_column_index = -1;
    ring_buffer_end_global_column_index = -1;
}

void ContinuousClustering::setTransformRobotFrameFromSensorFrame(const Eigen::Isometry3d& tf)
{
    ensureHandle();
    // order relative to addFiring is kept: firings buffered so far are segmented with the transform they were handed
    // over under (or raise "Transform ... not set yet" like the reference, cpp:298-299)
    flush();
    drain();
    double m[12];
    pose12(tf, m);
    cc_set_robot_from_sensor(handle_, m);
}

bool ContinuousClustering::hasTransformRobotFrameFromSensorFrame()
{
    return handle_ && cc_has_robot_from_sensor(handle_);
}

void ContinuousClustering::setFinishedColumnCallback(std::function<void(int64_t, int64_t, bool)> cb)
{
    finished_column_callback_ = std::move(cb);
}

void ContinuousClustering::setFinishedClusterCallback(std::function<void(const std::vector<Point>&, uint64_t)> cb)
{
    finished_cluster_callback_ = std::move(cb);
}

void ContinuousClustering::recordJobQueueWorkload(size_t) {}

void ContinuousClustering::setFinishedClusterPackedCallback(std::function<void(const PackedPointCloud2&)> cb)
{
    finished_cluster_packed_callback_ = std::move(cb);
}

void ContinuousClustering::setMaterialiseRangeImage(bool on)
{
    materialise_ = on;
}

static ContinuousClustering::PackedPointCloud2 toPacked(const cc_cloud_view_t& v)
{
    ContinuousClustering::PackedPointCloud2 m;
    m.data = v.data;
    m.size = static_cast<size_t>(v.data_size);
    m.point_step = v.point_step;
    m.width = v.width;
    m.height = v.height;
    m.n_fields = v.n_fields;
    m.stamp_ns = v.stamp_ns;
    return m;
}

// Every message the callbacks of this push can ask for, with one device launch: one per finished-column event and one
// per finished cluster handed to the packed cluster callback
void ContinuousClustering::prepack(const void* events_, int n_events, const void* clusters_, int n_clusters)
{
    const cc_column_event_t* events = static_cast<const cc_column_event_t*>(events_);
    const cc_cluster_t* clusters = static_cast<const cc_cluster_t*>(clusters_);
    std::vector<cc_pack_request_t> req;
    req.reserve(static_cast<size_t>(n_events) + n_clusters);
    if (finished_column_callback_)
        for (int i = 0; i < n_events; i++)
        {
            cc_pack_request_t r{};
            r.kind = events[i].ground_points_only ? 0 : 1;
            r.from_gcol = events[i].from_gcol;
            r.to_gcol = events[i].to_gcol;
            req.push_back(r);
        }
    if (finished_cluster_packed_callback_)
        for (int i = 0; i < n_clusters; i++)
            if (clusters[i].num_points > 20) // cpp:1023
            {
                cc_pack_request_t r{};
                r.kind = 2;
                r.cluster_index = i;
                r.from_gcol = i; // key of the lookup below
                r.to_gcol = i;
                req.push_back(r);
            }
    prepacked_.clear();
    if (req.empty())
        return;
    std::vector<cc_cloud_view_t> views(req.size());
    int s = cc_pack_requests_pointcloud2(handle_, static_cast<int>(req.size()), req.data(), views.data());
    if (s != CC_OK)
        fail(s);
    prepacked_.resize(req.size());
    for (size_t i = 0; i < req.size(); i++)
    {
        prepacked_[i].from = req[i].from_gcol;
        prepacked_[i].to = req[i].to_gcol;
        prepacked_[i].kind = req[i].kind;
        prepacked_[i].msg = toPacked(views[i]);
    }
}

ContinuousClustering::PackedPointCloud2 ContinuousClustering::packColumnsPointCloud2(int64_t from, int64_t to, bool ground_points_only)
{
    for (const Prepacked& q : prepacked_) // inside a callback: the message was packed with the rest of the push
        if (q.kind == (ground_points_only ? 0 : 1) && q.from == from && q.to == to)
            return q.msg;
    cc_cloud_view_t v;
    int s = cc_pack_columns_pointcloud2(handle_, from, to, ground_points_only ? 1 : 0, &v);
    if (s != CC_OK)
        fail(s);
    return toPacked(v);
}

void ContinuousClustering::addFiring(const RawPoints::ConstPtr& firing, const Eigen::Isometry3d& odom_from_sensor)
{
    if (num_rows_ != static_cast<int>(firing->points.size())) // cpp:90-91
        throw std::runtime_error("The number of points in a firing has changed. This is probably a bug!");
    const size_t bytes = firing->points.size() * sizeof(RawPoint);
    std::vector<unsigned char>& pts = points_buf_[cur_buf_];
    std::vector<double>& poses = poses_buf_[cur_buf_];
    const size_t off = pts.size();
    pts.resize(off + bytes);
    std::memcpy(pts.data() + off, firing->points.data(), bytes);
    poses.resize(poses.size() + 12);
    pose12(odom_from_sensor, poses.data() + poses.size() - 12);
    if (pending_ == 0)
        first_pending_us_ = nowMicros();
    // a batch goes to the device when it is full or when its oldest firing has waited max_wait_us_ (firings arrive
    // continuously, ~50 us apart at 64x2048 @ 10 Hz, so checking on arrival bounds the added latency)
    if (++pending_ >= batch_size_ || (max_wait_us_ >= 0 && nowMicros() - first_pending_us_ >= max_wait_us_))
        flush();
}

void ContinuousClustering::setMaxBatchLatency(int64_t microseconds)
{
    max_wait_us_ = microseconds;
}

void ContinuousClustering::flush()
{
    if (!handle_ || pending_ == 0)
        return;
    const int n = pending_;
    pending_ = 0;
    // whatever way this function is left (fail() throws), the buffer that was handed over is not reused with stale
    // records behind it: in pipelined mode the next buffer is taken, and the buffer taken next starts out empty
    struct Rotate
    {
        ContinuousClustering* self;
        ~Rotate()
        {
            if (self->pipelined_)
                self->cur_buf_ = (self->cur_buf_ + 1) % 3;
            self->points_buf_[self->cur_buf_].clear();
            self->poses_buf_[self->cur_buf_].clear();
        }
    } rotate{this};
    std::vector<unsigned char>& pts = points_buf_[cur_buf_];
    std::vector<double>& poses = poses_buf_[cur_buf_];
    const int step = std::max(1, cc_max_firings_per_push(handle_));
    const size_t rec = static_cast<size_t>(num_rows_) * sizeof(cc_raw_point_t);
    for (int a = 0; a < n; a += step)
    {
        const int m = std::min(step, n - a);
        const cc_raw_point_t* src = reinterpret_cast<const cc_raw_point_t*>(pts.data() + static_cast<size_t>(a) * rec);
        const double* src_poses = poses.data() + static_cast<size_t>(a) * 12;
        if (!pipelined_)
        {
            int s = cc_push_firings(handle_, m, num_rows_, src, src_poses);
            if (s != CC_OK)
                fail(s);
            deliver();
            continue;
        }
        // throughput mode: make room (at most two pushes in flight and one staged), hand the push over, and deliver what
        // has finished in the meantime; at most two pushes stay outstanding, so the buffer taken next is free again
        while (cc_pending(handle_) > 2)
        {
            int s = cc_wait(handle_);
            if (s != CC_OK)
                fail(s);
            deliver();
        }
        int s = cc_submit_firings(handle_, m, num_rows_, src, src_poses);
        if (s != CC_OK)
            fail(s);
        while (cc_pending(handle_) > 2)
        {
            s = cc_wait(handle_);
            if (s != CC_OK)
                fail(s);
            deliver();
        }
    }
}

// host copies of columns [from, to] into range_image_ (what the reference's consumers index, ros_utils.cpp:56-63): one
// device-side gather of packed cell records (cc_export_columns), then a single pass that fills every `Point` member
// the callers read (ros_utils.cpp:245-298, kitti_demo.cpp:196-216)
void ContinuousClustering::materialise(int64_t from, int64_t to)
{
    if (to < from)
        return;
    const size_t R = static_cast<size_t>(num_rows_);
    // children sit at most max_steps_in_row columns ahead of their parent (cpp:704-705): the child lists of [from, to]
    // need the parent pointers of that many more columns (as far as they have been associated yet)
    const int64_t ahead = std::max(0, config_.clustering.max_steps_in_row);
    const int64_t last = std::max(to, std::min(to + ahead, ring_buffer_end_global_column_index));
    const cc_cell_t* cells = nullptr;
    int s = cc_export_columns(handle_, from, last, &cells);
    if (s != CC_OK)
        fail(s);
    const size_t ncols = static_cast<size_t>(to - from + 1), ncols_ext = static_cast<size_t>(last - from + 1);
    for (size_t c = 0; c < ncols; c++)
    {
        const int local = static_cast<int>((from + static_cast<int64_t>(c)) % ring_buffer_max_columns);
        for (size_t r = 0; r < R; r++)
        {
            const cc_cell_t& q = cells[c * R + r];
            Point& p = range_image_[static_cast<size_t>(local) * R + r];
            p.xyz = Point3D(q.x, q.y, q.z);
            p.firing_index = q.firing_index;
            p.intensity = q.intensity;
            p.distance = q.distance;
            p.azimuth_angle = q.azimuth_angle;
            p.inclination_angle = q.inclination_angle;
            p.continuous_azimuth_angle = q.continuous_azimuth_angle;
            p.global_column_index = q.global_column_index;
            p.local_column_index = q.global_column_index >= 0 ? local : -1;
            p.row_index = std::isnan(q.distance) ? -1 : static_cast<int>(r); // set by insertion only (cpp:235)
            p.stamp = q.stamp;
            p.globally_unique_point_index = q.globally_unique_point_index;
            p.ground_point_label = q.ground_point_label;
            p.debug_ground_point_label = q.debug_ground_point_label;
            p.is_ignored = q.is_ignored != 0;
            p.id = q.id;
            p.finished_at_continuous_azimuth_angle = q.finished_at_continuous_azimuth_angle;
            p.tree_num_points = q.tree_num_points;
            p.cluster_width = q.cluster_width;
            p.number_of_visited_neighbors = q.number_of_visited_neighbors;
            p.belongs_to_finished_cluster = q.belongs_to_finished_cluster != 0;
            p.child_points.clear();
            if (q.tree_root_gcol >= 0)
            {
                p.tree_root_ = RangeImageIndex(static_cast<uint16_t>(q.tree_root_row),
                                               static_cast<int64_t>(q.tree_root_gcol % ring_buffer_max_columns));
                p.tree_id = static_cast<uint64_t>(q.tree_root_gcol) * R + static_cast<uint64_t>(q.tree_root_row);
            }
            else
            {
                p.tree_root_ = RangeImageIndex(0, -1);
                p.tree_id = 0;
            }
        }
    }
    // child_points (cpp:663): every associated point names the point whose list holds it; appended in association order
    // (column by column, rows top to bottom), which is the order the reference's lists have
    for (size_t c = 0; c < ncols_ext; c++)
        for (size_t r = 0; r < R; r++)
        {
            const cc_cell_t& q = cells[c * R + r];
            if (q.first_parent_gcol < from || q.first_parent_gcol > to)
                continue;
            const int parent_local = static_cast<int>(q.first_parent_gcol % ring_buffer_max_columns);
            const int child_local = static_cast<int>((from + static_cast<int64_t>(c)) % ring_buffer_max_columns);
            range_image_[static_cast<size_t>(parent_local) * R + static_cast<size_t>(q.first_parent_row)].child_points.emplace_back(
                static_cast<uint16_t>(r), static_cast<int64_t>(child_local));
        }
}

// callbacks of the last push, in the order of the reference's single-threaded mode
void ContinuousClustering::deliver()
{
    cc_batch_info_t info;
    cc_get_batch_info(handle_, &info);
    ring_buffer_start_global_column_index = info.ring_start_gcol;
    ring_buffer_end_global_column_index = info.ring_end_gcol;
    if (!finished_column_callback_ && !finished_cluster_callback_ && !finished_cluster_packed_callback_)
        return;
    std::vector<cc_column_event_t> events(info.n_events);
    std::vector<cc_cluster_t> clusters(info.n_clusters);
    std::vector<cc_cluster_point_t> points(info.n_cluster_points);
    int n = 0;
    cc_get_column_events(handle_, events.data(), info.n_events, &n);
    cc_get_clusters(handle_, clusters.data(), info.n_clusters, &n);
    cc_get_cluster_points(handle_, points.data(), info.n_cluster_points, &n);
    // every column a callback of this push can look at is read once: the ranges of the events (ground columns at the front
    // of the stream, clustered columns up to a rotation behind them) and the spans of the clusters handed out, merged where
    // they touch -- not their hull: the columns in between were delivered by earlier pushes or are not due yet
    if (materialise_ || finished_cluster_callback_)
    {
        std::vector<std::pair<int64_t, int64_t>> spans;
        for (const auto& e : events)
            if (e.to_gcol >= e.from_gcol && e.to_gcol >= 0)
                spans.emplace_back(e.from_gcol, e.to_gcol);
        if (finished_cluster_callback_)
            for (const auto& c : clusters)
                if (c.num_points > 20 && c.max_gcol >= c.min_gcol)
                    spans.emplace_back(c.min_gcol, c.max_gcol);
        std::sort(spans.begin(), spans.end());
        // a gap shorter than this many columns is cheaper to read than another device round trip (~25 us)
        const int64_t min_gap = 16;
        size_t i = 0;
        while (i < spans.size())
        {
            int64_t lo = spans[i].first, hi = spans[i].second;
            size_t j = i + 1;
            while (j < spans.size() && spans[j].first <= hi + min_gap)
            {
                hi = std::max(hi, spans[j].second);
                j++;
            }
            materialise(std::max<int64_t>(lo, 0), hi);
            i = j;
        }
    }
    prepacked_.clear();
    if (!materialise_ || finished_cluster_packed_callback_)
        prepack(events.data(), static_cast<int>(events.size()), clusters.data(), static_cast<int>(clusters.size()));
    size_t next = 0;
    for (const auto& e : events)
    {
        while (next < static_cast<size_t>(e.n_clusters_before) && next < clusters.size())
        {
            const cc_cluster_t& c = clusters[next++];
            if (c.num_points > 20 && finished_cluster_packed_callback_) // cpp:1023
                for (const Prepacked& q : prepacked_)
                    if (q.kind == 2 && q.from == static_cast<int64_t>(next - 1))
                    {
                        finished_cluster_packed_callback_(q.msg);
                        break;
                    }
            if (c.num_points > 20 && finished_cluster_callback_) // cpp:1023
            {
                cluster_buffer_.clear();
                for (uint32_t i = 0; i < c.num_points; i++)
                {
                    const cc_cluster_point_t& cp = points[c.point_offset + i];
                    const int local = static_cast<int>(cp.gcol % ring_buffer_max_columns);
                    cluster_buffer_.push_back(range_image_[static_cast<size_t>(local) * num_rows_ + cp.row]);
                }
                finished_cluster_callback_(cluster_buffer_, c.stamp);
            }
        }
        if (finished_column_callback_)
            finished_column_callback_(e.from_gcol, e.to_gcol, e.ground_points_only != 0);
    }
}

} // namespace continuous_clustering
