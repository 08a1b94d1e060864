// cc_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement of the reference's per-column hot path.
//
// This is the parity oracle of the B200 implementation. It restates, in plain sequential C++ over flat
// arrays, what UniBwTAS/continuous_clustering's `ContinuousClustering` does in its deterministic
// single-threaded mode (every stage runs synchronously and nested inside addFiring, thread_pool.hpp:31-35,
// 60-64). "cpp:" = /root/reference/src/clustering/continuous_clustering.cpp, "hpp:" =
// include/continuous_clustering/clustering/continuous_clustering.hpp.
//
//   insert_firing()        <- insertFiringIntoRangeImage                 cpp:105-292
//   segment_column()       <- performGroundPointSegmentationForColumn    cpp:294-624
//   associate_column()     <- associatePointsInColumn/traverseFieldOfView/
//                             associatePointToPointTree/...TreeToPointTree cpp:638-835
//   finish_pass()          <- findFinishedTreesAndAssignSameId            cpp:837-974
//   publish()              <- collectPointsForCusterAndPublish            cpp:976-1092
//   clear_columns()        <- clearColumns                                cpp:1094-1145
//
// Differences in REPRESENTATION (not in results): links between point trees are kept in a union-find over
// tree roots instead of per-root std::set adjacency + BFS (cpp:851-907) -- in single-threaded mode a link is
// only ever created between two unfinished trees (cpp:685-695) and a finish pass always finishes a whole
// connected component (cpp:922-934), so "connected component of the link graph" and "union-find class" are
// the same sets; child lists (cpp:663) are kept as one intrusive singly linked list per tree. Double
// precision rigid transforms use the evaluation order of oracle/eigen_standin (Eigen is not installed).
//
// PINNING: the reference ships no tests or golden vectors. This restatement is pinned
// against the reference's OWN sources compiled into oracle/_ref/libcc_ref.so (oracle/Makefile target `ref`)
// by tests/test_oracle.py, and against the fixtures in tests/golden generated from that build.
//
// Nothing under continuous_clustering_b200/ may include, link or load this file. It exports the recording
// driver API of cc_driver.h so tests can swap it for the reference build or the facade.
#include "cc_driver.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{

// label values = PointCloudColors entries used by the reference (general.hpp:208-357, hpp:15-22)
enum : uint8_t
{
    C_DARKRED = 32,
    C_GRAY = 53,
    C_GREEN = 54,
    C_LIGHTGRAY = 71,
    C_MAGENTA = 85,
    C_ORANGE = 105,
    C_RED = 119,
    C_VIOLET = 141,
    C_WHITE = 143,
    C_YELLOW = 145,
    C_YELLOWGREEN = 146,
    GP_UNKNOWN = C_WHITE,
    GP_GROUND = C_GREEN,
    GP_OBSTACLE = C_RED,
    GP_EGO_VEHICLE = C_MAGENTA,
    GP_FOG = C_LIGHTGRAY
};

const float NANF = std::numeric_limits<float>::quiet_NaN();
const uint32_t NONE = 0xffffffffu;

struct Iso // 3x4 [R|t], double; evaluation orders as in oracle/eigen_standin/Eigen/Geometry
{
    double m[3][4];
    static Iso from12(const double* p)
    {
        Iso t;
        std::memcpy(t.m, p, sizeof(t.m));
        return t;
    }
    void apply(double x, double y, double z, double out[3]) const
    {
        for (int i = 0; i < 3; i++)
            out[i] = ((m[i][0] * x + m[i][1] * y) + m[i][2] * z) + m[i][3];
    }
    Iso inverse() const
    {
        Iso r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                r.m[i][j] = m[j][i];
        for (int i = 0; i < 3; i++)
            r.m[i][3] = -(r.m[i][0] * m[0][3] + (r.m[i][1] * m[1][3] + r.m[i][2] * m[2][3]));
        return r;
    }
    Iso mul(const Iso& o) const
    {
        Iso r;
        for (int i = 0; i < 3; i++)
        {
            for (int j = 0; j < 3; j++)
                r.m[i][j] = m[i][0] * o.m[0][j] + (m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j]);
            r.m[i][3] = (m[i][0] * o.m[0][3] + (m[i][1] * o.m[1][3] + m[i][2] * o.m[2][3])) + m[i][3];
        }
        return r;
    }
};

struct Cell // mirror of `Point` hpp:126-161, flat
{
    float x, y, z;
    float distance, azimuth, inclination;
    double cont_az;
    int64_t gcol;
    int local_col, row;
    uint64_t firing_index, stamp, guid;
    uint8_t intensity, label, dbg, ignored;
    // clustering
    double finished_at;
    int64_t root_col; // LOCAL ring column of the tree root, -1 = unassociated (cpp:661, 1134-1135)
    int root_row;
    uint32_t tree_num_points, cluster_width;
    uint64_t tree_id, id;
    uint8_t finished; // belongs_to_finished_cluster
    int visited;      // number_of_visited_neighbors
    int num_children; // child_points.size() (cpp:663)
    uint32_t link_parent; // union-find over tree roots (replaces associated_trees)
    uint32_t next_in_tree, tree_tail; // intrusive member list of a tree (replaces child_points)
    uint64_t pass_id;   // per-pass scratch (replaces visited_at_continuous_azimuth_angle)
    uint32_t pass_slot; // per-pass scratch: aggregate slot of a component representative
};

} // namespace

struct drv
{
    // ---- configuration / lifecycle (cpp:11-86) ----
    cc_config_t cfg{};
    cc_config_t pending_cfg{};
    int R{-1}, N{0}, ringcols{0};
    float width{0};
    float max_distance_squared{0.7f * 0.7f};
    bool reset_required{false};
    bool has_robot_tf{false};
    Iso robot_from_sensor{};

    std::vector<Cell> ring;
    int64_t ring_start{-1}, ring_end{-1};
    int64_t prev_rearmost{0}, prev_foremost{-1}, first_unfinished{-1};
    double sensor_pos_d[3]{0, 0, 0};
    float sensor_pos[3]{0, 0, 0};
    int64_t first_unpublished{-1};
    std::vector<uint32_t> unfinished_trees; // sc_unfinished_point_trees_ (cell indices of roots, list order)
    uint64_t cluster_counter{1};
    std::vector<float> incl_gap; // sc_inclination_angles_between_lasers_

    // ---- recording (same outputs as cc_driver.cpp) ----
    int record{DRV_RECORD_FULL};
    std::string error;
    std::vector<cc_column_event_t> events;
    std::vector<int64_t> ground_cols, cluster_cols;
    std::vector<drv_cell_t> ground_cells, cluster_cells;
    std::vector<drv_cluster_t> clusters;
    std::vector<drv_cluster_point_t> cluster_points;
    std::vector<cc_raw_point_t> prepared;
    std::vector<double> prepared_poses;
    int prepared_rows{0};

    Cell& at(int64_t local_col, int row) { return ring[local_col * R + row]; }

    // clearColumns cpp:1094-1145
    void clear_columns(int64_t from, int64_t to)
    {
        for (int64_t g = from; g <= to; g++)
        {
            int local = static_cast<int>(g % ringcols);
            for (int r = 0; r < R; r++)
            {
                Cell& c = at(local, r);
                c.x = c.y = c.z = NANF;
                c.distance = c.azimuth = c.inclination = NANF;
                c.cont_az = std::numeric_limits<double>::quiet_NaN();
                c.gcol = -1;
                c.local_col = -1;
                c.row = -1;
                c.intensity = 0;
                c.stamp = 0;
                c.guid = static_cast<uint64_t>(-1);
                c.label = GP_UNKNOWN;
                c.dbg = C_WHITE;
                c.ignored = 0;
                c.finished_at = 0.;
                c.root_row = 0;
                c.root_col = -1;
                c.tree_num_points = 0;
                c.cluster_width = 0;
                c.tree_id = 0;
                c.id = 0;
                c.finished = 0;
                c.visited = 0;
                c.num_children = 0;
                c.link_parent = NONE;
                c.next_in_tree = NONE;
                c.tree_tail = NONE;
                c.pass_id = 0;
                c.pass_slot = 0;
                // note: firing_index is not cleared by the reference (cpp:1107-1142)
            }
        }
    }

    // reset cpp:11-64
    void reset(int num_rows)
    {
        N = cfg.num_columns;
        const bool same_rows = (num_rows == R);
        R = num_rows;
        width = static_cast<float>(2 * M_PI) / static_cast<float>(N);
        ringcols = N * 10;
        const size_t old = ring.size();
        ring.resize(static_cast<size_t>(ringcols) * R);
        for (size_t i = old; i < ring.size(); i++)
            ring[i].firing_index = 0;
        clear_columns(0, ringcols - 1);
        ring_start = ring_end = -1;
        prev_rearmost = 0;
        prev_foremost = -1;
        first_unfinished = -1;
        reset_required = false;
        has_robot_tf = false;
        first_unpublished = -1;
        unfinished_trees.clear();
        cluster_counter = 1;
        // std::vector::resize(n, NaN) only fills NEW elements (cpp:46): values survive a reset with the same row count
        (void)same_rows;
        incl_gap.resize(R, NANF);
    }

    int64_t refused_joins{0}, refused_links{0}; // test statistics: associations the reference refuses (cpp:654-659, 688-690)

    // setConfiguration cpp:66-81
    void set_config(const cc_config_t& c)
    {
        if ((cfg.is_single_threaded != 0) != (c.is_single_threaded != 0))
            reset_required = true;
        if ((cfg.sensor_is_clockwise != 0) != (c.sensor_is_clockwise != 0))
            reset_required = true;
        if (cfg.num_columns != c.num_columns)
            reset_required = true;
        cfg = c;
        max_distance_squared = cfg.max_distance * cfg.max_distance;
    }

    // ---- callbacks -> records ----
    void snapshot(const Cell& p, drv_cell_t& c)
    {
        std::memset(&c, 0, sizeof(c));
        c.continuous_azimuth_angle = p.cont_az;
        c.global_column_index = p.gcol;
        c.globally_unique_point_index = p.guid;
        c.stamp = p.stamp;
        c.firing_index = p.firing_index;
        c.id = p.id;
        c.x = p.x;
        c.y = p.y;
        c.z = p.z;
        c.distance = p.distance;
        c.azimuth_angle = p.azimuth;
        c.inclination_angle = p.inclination;
        c.number_of_visited_neighbors = p.visited;
        c.intensity = p.intensity;
        c.ground_point_label = p.label;
        c.debug_ground_point_label = p.dbg;
        c.is_ignored = p.ignored;
        c.num_child_points = p.num_children;
        c.finished_at_continuous_azimuth_angle = p.finished_at;
        c.tree_num_points = p.tree_num_points;
        c.cluster_width = p.cluster_width;
        c.local_column_index = p.local_col;
        c.row_index = p.row;
        c.tree_root_row = p.root_row;
        c.tree_root_gcol = -1;
        if (p.root_col >= 0)
            c.tree_root_gcol = at(p.root_col, p.root_row).gcol;
    }

    void on_columns(int64_t from, int64_t to, bool ground_only)
    {
        if (record == DRV_RECORD_NONE)
            return;
        cc_column_event_t ev;
        ev.from_gcol = from;
        ev.to_gcol = to;
        ev.ground_points_only = ground_only ? 1 : 0;
        ev.n_clusters_before = static_cast<int32_t>(clusters.size());
        events.push_back(ev);
        if (record < DRV_RECORD_FULL)
            return;
        for (int64_t g = from; g <= to; g++)
        {
            int local = static_cast<int>(g % ringcols);
            std::vector<drv_cell_t>& cells = ground_only ? ground_cells : cluster_cells;
            (ground_only ? ground_cols : cluster_cols).push_back(g);
            size_t base = cells.size();
            cells.resize(base + R);
            for (int r = 0; r < R; r++)
                snapshot(at(local, r), cells[base + r]);
        }
    }

    // ---- insertFiringIntoRangeImage cpp:105-292 ----
    void insert_firing(const cc_raw_point_t* pts, const Iso& pose)
    {
        sensor_pos_d[0] = pose.m[0][3];
        sensor_pos_d[1] = pose.m[1][3];
        sensor_pos_d[2] = pose.m[2][3];
        for (int i = 0; i < 3; i++)
            sensor_pos[i] = static_cast<float>(sensor_pos_d[i]);

        int64_t foremost = -1, rearmost = -1;
        const int64_t prev_rot = prev_rearmost / N;

        for (int row = 0; row < R; row++)
        {
            const cc_raw_point_t& raw = pts[row];
            const double px = raw.x, py = raw.y, pz = raw.z;
            if (std::isnan(px))
                continue;
            double po[3];
            pose.apply(px, py, pz, po);
            const double rel[3] = {po[0] - sensor_pos_d[0], po[1] - sensor_pos_d[1], po[2] - sensor_pos_d[2]};

            // azimuth from the SENSOR-frame coordinates (cpp:142)
            const float azimuth = std::atan2(static_cast<float>(py), static_cast<float>(px));
            const float inc_az =
                cfg.sensor_is_clockwise ? -azimuth + static_cast<float>(M_PI) : azimuth + static_cast<float>(M_PI);

            const int col_in_rot = static_cast<int>(inc_az / width);
            int64_t g = prev_rot * N + col_in_rot;
            const int prev_col_in_rot = static_cast<int>(prev_rearmost % N);
            const int diff = col_in_rot - prev_col_in_rot;
            const int half = N / 2;
            int rot_off = 0;
            if (diff < -half)
            {
                g += N;
                rot_off = 1;
            }
            else if (prev_rearmost > 0 && diff > half)
            {
                g -= N;
                rot_off = -1;
            }
            int local = static_cast<int>(g % ringcols);
            if (g < 0)
                continue; // the reference indexes range_image_ out of bounds here (UB); we drop the point
            Cell* cell = &at(local, row);

            const double cont_az = (2 * M_PI) * static_cast<double>(prev_rot + rot_off) + inc_az;

            // cell occupied -> try the next column (cpp:188-204)
            const float distance =
                static_cast<float>(std::sqrt(rel[0] * rel[0] + (rel[1] * rel[1] + rel[2] * rel[2])));
            if (!std::isnan(cell->distance) && !std::isnan(distance))
            {
                int next_local = local + 1;
                if (next_local >= ringcols)
                    next_local -= ringcols;
                Cell* next = &at(next_local, row);
                if (std::isnan(next->distance))
                {
                    cell = next;
                    local = next_local;
                    g++;
                }
            }
            // never overwrite a valid cell with NaN or a farther return (cpp:206-208)
            if (!std::isnan(cell->distance) && (std::isnan(distance) || distance >= cell->distance))
                continue;

            const bool too_far_behind = first_unfinished >= 0 && g < first_unfinished; // cpp:210-221
            if (!too_far_behind)
            {
                cell->x = static_cast<float>(po[0]);
                cell->y = static_cast<float>(po[1]);
                cell->z = static_cast<float>(po[2]);
                cell->firing_index = raw.firing_index;
                cell->intensity = raw.intensity;
                cell->stamp = raw.stamp;
                cell->distance = distance;
                cell->azimuth = azimuth;
                cell->inclination = std::asin(static_cast<float>(rel[2]) / cell->distance);
                cell->cont_az = cont_az;
                cell->gcol = g;
                cell->local_col = local;
                cell->row = row;
                cell->guid = raw.globally_unique_point_index;
            }
            if (rearmost < 0 || g < rearmost)
                rearmost = g;
            if (foremost < 0 || g > foremost)
                foremost = g;
        }

        if (rearmost >= 0 && foremost >= 0)
        {
            if ((foremost - rearmost) > N / 2) // firing straddles the negative x axis (cpp:252-261)
            {
                reset_required = true;
                return;
            }
            if (rearmost > prev_rearmost)
                prev_rearmost = rearmost;
            if (foremost > prev_foremost)
                prev_foremost = foremost;
        }
        if (prev_foremost < 0)
            return;
        if (ring_start == -1)
        {
            ring_start = prev_rearmost;
            first_unpublished = prev_rearmost;
        }
        if (prev_foremost > ring_end)
            ring_end = prev_foremost;
        if (first_unfinished == -1)
            first_unfinished = prev_rearmost;
        while (first_unfinished < prev_rearmost)
            segment_column(first_unfinished++, pose);
    }

    static void to2d(float x, float y, float z, float& ox, float& oy) // to2dInAzimuthPlane hpp:229-232
    {
        ox = std::sqrt(x * x + y * y);
        oy = z;
    }

    // ---- performGroundPointSegmentationForColumn cpp:294-624 ----
    void segment_column(int64_t gcol, const Iso& pose)
    {
        const int local = static_cast<int>(gcol % ringcols);
        if (!has_robot_tf)
            throw std::runtime_error("Transform robot frame from sensor frame was not set yet!");
        const Iso ego_from_odom = robot_from_sensor.mul(pose.inverse());
        const float height_sensor_to_ground =
            -static_cast<float>(robot_from_sensor.m[2][3]) + cfg.height_ref_to_ground_;

        bool first_obstacle_detected = false;
        bool first_point_found = false;
        float last_ground[3] = {0, 0, height_sensor_to_ground};
        float prev_pos[3] = {0, 0, 0};
        uint8_t prev_label = 0;
        float incl_prev_laser = 0;

        for (int row = R - 1; row >= 0; row--)
        {
            Cell& p = at(local, row);
            if (p.gcol != gcol && p.gcol != -1) // ring overrun (cpp:321-345)
                throw std::runtime_error(
                    "This column is not cleared. Probably this means the ring buffer is full or there "
                    "is some other issue with clearing (not cleared at all or written after clearing): " +
                    std::to_string(p.gcol) + ", " + std::to_string(gcol) + ", " + std::to_string(ringcols));
            p.gcol = gcol;
            p.local_col = local;

            const float incl_cur = p.inclination; // cpp:353-357
            const float d = incl_cur - incl_prev_laser;
            if (!std::isnan(d))
                incl_gap[row] = d;
            incl_prev_laser = incl_cur;

            if (std::isnan(p.distance)) // cpp:360-374
            {
                if (cfg.supplement_inclination_angle_for_nan_cells && row < R - 1)
                    p.inclination = at(local, row + 1).inclination + incl_gap[row];
                p.cont_az = (static_cast<double>(gcol) + 0.5) * width;
                continue;
            }
            if (cfg.fog_filtering_enabled && p.intensity < static_cast<uint8_t>(cfg.fog_filtering_intensity_below) &&
                p.distance < cfg.fog_filtering_distance_below && p.inclination > cfg.fog_filtering_inclination_above)
            {
                p.label = GP_FOG;
                p.dbg = C_LIGHTGRAY;
                continue;
            }
            double ego[3];
            ego_from_odom.apply(p.x, p.y, p.z, ego);
            if (ego[0] < cfg.length_ref_to_front_end_ && ego[0] > cfg.length_ref_to_rear_end_ &&
                ego[1] < cfg.width_ref_to_left_mirror_ && ego[1] > cfg.width_ref_to_right_mirror_ &&
                ego[2] < cfg.height_ref_to_maximum_ && ego[2] > cfg.height_ref_to_ground_)
            {
                p.label = GP_EGO_VEHICLE;
                p.dbg = C_VIOLET;
                continue;
            }
            const float cur[3] = {p.x - sensor_pos[0], p.y - sensor_pos[1], p.z - sensor_pos[2]};

            if (!first_point_found) // cpp:409-431
            {
                first_point_found = true;
                const float h = cur[2] - height_sensor_to_ground;
                if (h > cfg.first_ring_as_ground_min_allowed_z_diff && h < cfg.first_ring_as_ground_max_allowed_z_diff)
                {
                    p.label = GP_GROUND;
                    p.dbg = C_GRAY;
                    std::memcpy(last_ground, cur, sizeof(cur));
                    first_obstacle_detected = false;
                }
                else
                {
                    p.label = GP_OBSTACLE;
                    p.dbg = C_ORANGE;
                    first_obstacle_detected = true;
                }
                std::memcpy(prev_pos, cur, sizeof(cur));
                prev_label = p.dbg;
                continue;
            }

            float c2x, c2y, p2x, p2y, g2x, g2y;
            to2d(cur[0], cur[1], cur[2], c2x, c2y);
            to2d(prev_pos[0], prev_pos[1], prev_pos[2], p2x, p2y);
            const float ptc_x = c2x - p2x, ptc_y = c2y - p2y;
            const float slope_to_prev = ptc_y / ptc_x;
            bool flat_prev = std::abs(slope_to_prev) < cfg.max_slope && ptc_x > 0;
            flat_prev = flat_prev && (!cfg.use_terrain || ptc_x < 5);

            to2d(last_ground[0], last_ground[1], last_ground[2], g2x, g2y);
            const float gtc_x = c2x - g2x, gtc_y = c2y - g2y;
            const float slope_to_ground = gtc_y / gtc_x;
            const bool flat_ground = std::abs(slope_to_ground) < cfg.max_slope && gtc_x > 0;

            if (!first_obstacle_detected && flat_prev)
            {
                p.label = GP_GROUND;
                p.dbg = C_GREEN;
            }
            else if (!cfg.use_terrain) // the use_terrain branch is dead code in the reference (cpp:455-489)
            {
                if (first_obstacle_detected && flat_prev && flat_ground)
                {
                    p.label = GP_GROUND;
                    p.dbg = C_YELLOWGREEN;
                }
                else if (std::abs(gtc_x) < cfg.ground_because_close_to_last_certain_ground_max_dist_diff &&
                         std::abs(gtc_y) < cfg.ground_because_close_to_last_certain_ground_max_z_diff)
                {
                    p.label = GP_GROUND;
                    p.dbg = C_YELLOW;
                }
            }

            if (p.label != GP_GROUND) // cpp:508-536
            {
                p.label = GP_OBSTACLE;
                p.dbg = C_RED;
                int below = row + 1;
                while (below < R)
                {
                    Cell& q = at(local, below);
                    float q2x, q2y;
                    to2d(q.x - sensor_pos[0], q.y - sensor_pos[1], q.z - sensor_pos[2], q2x, q2y);
                    if (q.dbg == C_YELLOW ||
                        (q.label == GP_GROUND &&
                         std::abs(c2x - q2x) < cfg.obstacle_because_next_certain_obstacle_max_dist_diff))
                    {
                        if (q.label == GP_GROUND)
                        {
                            q.label = GP_OBSTACLE;
                            q.dbg = C_DARKRED;
                        }
                        below++;
                    }
                    else
                        break;
                }
            }
            first_obstacle_detected |= p.label == GP_OBSTACLE;

            if (p.dbg == C_GREEN || p.dbg == C_YELLOWGREEN) // cpp:541-561
            {
                if (slope_to_prev > cfg.last_ground_point_slope_higher_than &&
                    std::abs(ptc_x) < cfg.last_ground_point_distance_smaller_than && prev_label != C_YELLOW)
                    std::memcpy(last_ground, cur, sizeof(cur));
            }
            std::memcpy(prev_pos, cur, sizeof(cur));
            prev_label = p.dbg;
        }

        for (int row = R - 1; row >= 0; row--) // cpp:567-616
        {
            Cell& p = at(local, row);
            p.ignored = 0;
            if (std::isnan(p.distance))
            {
                p.ignored = 1;
                continue;
            }
            if (p.label != GP_OBSTACLE)
            {
                p.ignored = 1;
                continue;
            }
            if (p.distance < 1. * cfg.max_distance)
            {
                p.ignored = 1;
                continue;
            }
            if (cfg.ignore_points_with_too_big_inclination_angle_diff && row < (R - 1) &&
                std::atan2(cfg.max_distance, p.distance) < incl_gap[row])
            {
                p.ignored = 1;
                continue;
            }
            if (cfg.ignore_points_in_chessboard_pattern)
            {
                const bool column_even = p.gcol % 2 == 0;
                const bool row_even = row % 2 == 0;
                if ((column_even && !row_even) || (!column_even && row_even))
                {
                    p.ignored = 1;
                    continue;
                }
            }
        }

        on_columns(gcol, gcol, true);
        associate_column(gcol);
    }

    // ---- association cpp:638-835 ----
    uint32_t find_link(uint32_t root)
    {
        while (ring[root].link_parent != root)
        {
            ring[root].link_parent = ring[ring[root].link_parent].link_parent;
            root = ring[root].link_parent;
        }
        return root;
    }

    void traverse(Cell& p, uint32_t p_index, float mad, int first_local_col) // traverseFieldOfView cpp:698-771
    {
        int steps_back = static_cast<int>(std::ceil(mad / width));
        steps_back = std::min(steps_back, cfg.max_steps_in_row);
        int64_t other_col = p.local_col;
        for (int back = 0; back <= steps_back; back++)
        {
            for (int dir = -1; dir <= 1; dir += 2)
            {
                if (dir == 1 && back == 0)
                    continue;
                int steps_v = (dir == 1 || back == 0) ? 1 : 0;
                int other_row = (dir == 1 || back == 0) ? p.row + dir : p.row;
                while (other_row >= 0 && other_row < R && steps_v <= cfg.max_steps_in_column)
                {
                    const uint32_t o_index = static_cast<uint32_t>(other_col * R + other_row);
                    Cell& o = ring[o_index];
                    p.visited += 1;
                    if (std::abs(o.inclination - p.inclination) > mad)
                        break;
                    const bool same_root = o.root_row == p.root_row && o.root_col == p.root_col;
                    if (!o.ignored && (p.root_col == 0 || !same_root))
                    {
                        const float dx = p.x - o.x, dy = p.y - o.y, dz = p.z - o.z;
                        if (dx * dx + dy * dy + dz * dz < max_distance_squared) // cpp:638-641
                        {
                            if (p.root_col == -1)
                            {
                                // associatePointToPointTree cpp:643-673
                                const uint32_t root_index = static_cast<uint32_t>(o.root_col * R + o.root_row);
                                Cell& root = ring[root_index];
                                const uint32_t new_width = static_cast<uint32_t>(p.gcol - root.gcol + 1);
                                if (new_width <= static_cast<uint32_t>(N) && !root.finished)
                                {
                                    p.root_col = o.root_col;
                                    p.root_row = o.root_row;
                                    p.tree_id = root.gcol * R + root.row;
                                    ring[root.tree_tail].next_in_tree = p_index;
                                    root.tree_tail = p_index;
                                    o.num_children++; // point_other.child_points.emplace_back(...) cpp:663
                                    root.cluster_width = new_width;
                                    root.finished_at = std::max(root.finished_at, p.cont_az + mad);
                                    root.tree_num_points++;
                                }
                                else
                                    refused_joins++; // cpp:654-659
                            }
                            else
                            {
                                // associatePointTreeToPointTree cpp:675-696
                                const uint32_t ra = static_cast<uint32_t>(p.root_col * R + p.root_row);
                                const uint32_t rb = static_cast<uint32_t>(o.root_col * R + o.root_row);
                                if (!ring[ra].finished && !ring[rb].finished)
                                {
                                    const uint32_t a = find_link(ra), b = find_link(rb);
                                    if (a != b)
                                        ring[b].link_parent = a;
                                }
                                else
                                    refused_links++; // cpp:688-690
                            }
                        }
                    }
                    if (p.root_col != -1 && cfg.stop_after_association_enabled &&
                        steps_v >= cfg.stop_after_association_min_steps)
                        break;
                    other_row += dir;
                    steps_v++;
                }
            }
            if (p.root_col != -1 && cfg.stop_after_association_enabled && back >= cfg.stop_after_association_min_steps)
                break;
            if (other_col == first_local_col)
                break;
            other_col--;
            if (other_col < 0)
                other_col += ringcols;
        }
    }

    void associate_column(int64_t gcol) // associatePointsInColumn cpp:773-835
    {
        std::vector<uint32_t> new_trees;
        double min_az = std::numeric_limits<double>::max();
        const int first_local = static_cast<int>(first_unpublished % ringcols);
        const int local = static_cast<int>(gcol % ringcols);
        for (int row = 0; row < R; row++)
        {
            const uint32_t index = static_cast<uint32_t>(local * R + row);
            Cell& p = ring[index];
            if (p.cont_az < min_az)
                min_az = p.cont_az;
            if (p.ignored)
                continue;
            // cells of a column get row/local column only when a point was inserted; all non-ignored cells have one
            const float mad = std::asin(cfg.max_distance / p.distance);
            traverse(p, index, mad, first_local);
            if (p.root_col == -1)
            {
                p.root_col = local;
                p.root_row = row;
                p.tree_id = p.gcol * R + p.row;
                p.finished_at = p.cont_az + mad;
                p.cluster_width = 1;
                p.tree_num_points = 1;
                p.link_parent = index;
                p.tree_tail = index;
                p.next_in_tree = NONE;
                new_trees.push_back(index);
            }
        }
        finish_pass(gcol, new_trees, min_az);
    }

    // ---- findFinishedTreesAndAssignSameId cpp:837-974 ----
    void finish_pass(int64_t gcol, const std::vector<uint32_t>& new_trees, double min_az)
    {
        unfinished_trees.insert(unfinished_trees.end(), new_trees.begin(), new_trees.end());
        if (gcol % cfg.cluster_point_trees_every_nth_column != 0)
            return;

        // aggregate every link component over the unfinished list (the reference does this with one BFS per
        // component, cpp:851-907); the aggregate slot of a component is remembered on its representative
        struct Agg
        {
            int64_t min_col, max_col;
            uint32_t num_points;
            bool unfinished;
            bool done;
        };
        pass_counter++;
        std::vector<uint32_t> reps(unfinished_trees.size());
        std::vector<Agg> aggs;
        for (size_t i = 0; i < unfinished_trees.size(); i++)
        {
            const uint32_t rep = find_link(unfinished_trees[i]);
            reps[i] = rep;
            Cell& rc = ring[rep];
            if (rc.pass_id != pass_counter)
            {
                rc.pass_id = pass_counter;
                rc.pass_slot = static_cast<uint32_t>(aggs.size());
                aggs.push_back(Agg{std::numeric_limits<int64_t>::max(), 0, 0, false, false});
            }
            Agg& a = aggs[rc.pass_slot];
            const Cell& root = ring[unfinished_trees[i]];
            a.min_col = std::min(a.min_col, root.gcol);                                            // cpp:879
            a.max_col = std::max(a.max_col, root.gcol + static_cast<int64_t>(root.cluster_width)); // cpp:880-881
            if (root.finished_at > min_az)                                                         // cpp:884-885
                a.unfinished = true;
            a.num_points += root.tree_num_points;
        }

        std::vector<std::vector<uint32_t>> finished_cluster_trees;
        std::vector<uint64_t> finished_ids;
        std::vector<int> slot_to_cluster(aggs.size(), -1);
        for (size_t i = 0; i < unfinished_trees.size(); i++) // list order decides cluster id order (cpp:851, 939)
        {
            const uint32_t slot = ring[reps[i]].pass_slot;
            Agg& a = aggs[slot];
            if (!a.done)
            {
                a.done = true;
                if ((a.max_col - a.min_col) >= N) // forced finish of clusters spanning a rotation, cpp:909-919
                    a.unfinished = false;
                if (!a.unfinished)
                {
                    if (a.num_points > 5) // cpp:936-940
                    {
                        slot_to_cluster[slot] = static_cast<int>(finished_ids.size());
                        finished_ids.push_back(cluster_counter++);
                        finished_cluster_trees.emplace_back();
                    }
                }
            }
            if (!a.unfinished)
            {
                ring[unfinished_trees[i]].finished = 1; // cpp:928-934
                const int cl = slot_to_cluster[slot];
                if (cl >= 0)
                    finished_cluster_trees[cl].push_back(unfinished_trees[i]);
            }
        }

        // erase finished trees, minimum root column over the list BEFORE erasing (cpp:943-959)
        int64_t min_required = std::numeric_limits<int64_t>::max();
        size_t w = 0;
        for (size_t i = 0; i < unfinished_trees.size(); i++)
        {
            const Cell& root = ring[unfinished_trees[i]];
            if (root.gcol < min_required)
                min_required = root.gcol;
            if (!root.finished)
                unfinished_trees[w++] = unfinished_trees[i];
        }
        unfinished_trees.resize(w);
        if (min_required == std::numeric_limits<int64_t>::max())
            min_required = gcol + 1;

        publish(min_required, finished_ids, finished_cluster_trees);
    }

    // ---- collectPointsForCusterAndPublish cpp:976-1092 (single-threaded: the min-required list has one entry) ----
    void publish(int64_t min_required, const std::vector<uint64_t>& ids,
                 const std::vector<std::vector<uint32_t>>& trees_per_cluster)
    {
        for (size_t k = 0; k < ids.size(); k++)
        {
            uint64_t min_stamp = std::numeric_limits<uint64_t>::max(), max_stamp = 0;
            const int64_t offset = static_cast<int64_t>(cluster_points.size());
            int64_t n = 0;
            for (uint32_t root : trees_per_cluster[k])
            {
                for (uint32_t i = root; i != NONE; i = ring[i].next_in_tree)
                {
                    Cell& p = ring[i];
                    p.id = ids[k];
                    n++;
                    min_stamp = std::min(min_stamp, p.stamp);
                    max_stamp = std::max(max_stamp, p.stamp);
                    if (record != DRV_RECORD_NONE)
                    {
                        drv_cluster_point_t cp;
                        cp.gcol = p.gcol;
                        cp.globally_unique_point_index = p.guid;
                        cp.row = p.row;
                        cp.pad_ = 0;
                        cluster_points.push_back(cp);
                    }
                }
            }
            if (n > 20) // cpp:1023
            {
                const uint64_t stamp =
                    cfg.use_last_point_for_cluster_stamp ? max_stamp : min_stamp + (max_stamp - min_stamp) / 2;
                if (record != DRV_RECORD_NONE)
                {
                    drv_cluster_t c;
                    c.stamp = stamp;
                    c.id = ids[k];
                    c.point_offset = offset;
                    c.num_points = n;
                    c.event_index = static_cast<int64_t>(events.size());
                    clusters.push_back(c);
                }
            }
            else if (record != DRV_RECORD_NONE)
                cluster_points.resize(offset); // the callback is not invoked for <= 20 points
        }

        const int64_t old_start = ring_start;
        const int64_t old_unpublished = first_unpublished;
        first_unpublished = min_required;
        if (first_unpublished < old_unpublished)
            throw std::runtime_error("This shouldn't happen, ring buffer is not allowed to increase at the front: " +
                                     std::to_string(first_unpublished) + ", " + std::to_string(old_unpublished));
        ring_start = std::max<int64_t>(0, first_unpublished - N);
        on_columns(old_unpublished, first_unpublished - 1, false);
        clear_columns(old_start, ring_start - 1);
    }

    uint64_t pass_counter{0};
};

extern "C" {

const char* drv_impl_name(void)
{
    return "restatement";
}

drv_t* drv_create(void)
{
    drv* d = new drv();
    // defaults of hpp:24-87
    cc_config_t& c = d->cfg;
    std::memset(&c, 0, sizeof(c));
    c.sensor_is_clockwise = 1;
    c.num_columns = 1700;
    c.supplement_inclination_angle_for_nan_cells = 1;
    c.max_slope = 0.2f;
    c.first_ring_as_ground_max_allowed_z_diff = 0.4f;
    c.first_ring_as_ground_min_allowed_z_diff = -0.4f;
    c.last_ground_point_slope_higher_than = -0.1f;
    c.last_ground_point_distance_smaller_than = 5.f;
    c.ground_because_close_to_last_certain_ground_max_z_diff = 0.4f;
    c.ground_because_close_to_last_certain_ground_max_dist_diff = 2.0f;
    c.obstacle_because_next_certain_obstacle_max_dist_diff = 0.3f;
    c.terrain_max_allowed_z_diff = 0.4f;
    c.fog_filtering_intensity_below = 2;
    c.fog_filtering_distance_below = 18.f;
    c.fog_filtering_inclination_above = -0.06f;
    c.max_distance = 0.7f;
    c.max_steps_in_row = 20;
    c.max_steps_in_column = 20;
    c.stop_after_association_enabled = 1;
    c.stop_after_association_min_steps = 1;
    c.ignore_points_in_chessboard_pattern = 1;
    c.ignore_points_with_too_big_inclination_angle_diff = 1;
    c.cluster_point_trees_every_nth_column = 1;
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
        d->set_config(*cfg);
        d->reset(num_rows);
        if (robot_from_sensor)
        {
            d->robot_from_sensor = Iso::from12(robot_from_sensor);
            d->has_robot_tf = true;
        }
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
    d->set_config(*cfg);
    return 0;
}

// test statistics of the restatement only: refused joins / links so far (-1 in the builds of the reference itself)
void drv_refusals(drv_t* d, int64_t* joins, int64_t* links)
{
    *joins = d->refused_joins;
    *links = d->refused_links;
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
        {
            if (d->R != rows) // cpp:90-91
                throw std::runtime_error("The number of points in a firing has changed. This is probably a bug!");
            d->insert_firing(pts + static_cast<size_t>(k) * rows, Iso::from12(poses + 12 * k));
        }
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
    return d->reset_required ? 1 : 0;
}

int drv_num_rows(drv_t* d)
{
    return d->R;
}

int drv_ring_buffer_max_columns(drv_t* d)
{
    return d->ringcols;
}

int drv_prepare(drv_t* d, int n, int rows, const cc_raw_point_t* pts, const double* poses)
{
    d->prepared.assign(pts, pts + static_cast<size_t>(n) * rows);
    d->prepared_poses.assign(poses, poses + static_cast<size_t>(n) * 12);
    d->prepared_rows = rows;
    return 0;
}

double drv_run_prepared(drv_t* d, int from, int to, int64_t /*max_lag_columns*/)
{
    const int rows = d->prepared_rows;
    if (from < 0 || from > to || static_cast<size_t>(to) * rows > d->prepared.size())
    {
        d->error = "drv_run_prepared: bad range";
        return -1.;
    }
    auto t0 = std::chrono::steady_clock::now();
    if (drv_add_firings(d, to - from, rows, d->prepared.data() + static_cast<size_t>(from) * rows,
                        d->prepared_poses.data() + static_cast<size_t>(from) * 12))
        return -1.;
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int drv_run_prepared_latency(drv_t* d, int from, int to, double* out_us)
{
    const int rows = d->prepared_rows;
    if (from < 0 || from > to || static_cast<size_t>(to) * rows > d->prepared.size())
    {
        d->error = "drv_run_prepared_latency: bad range";
        return 1;
    }
    for (int k = from; k < to; k++)
    {
        const auto t0 = std::chrono::steady_clock::now();
        if (drv_add_firings(d, 1, rows, d->prepared.data() + static_cast<size_t>(k) * rows, d->prepared_poses.data() + static_cast<size_t>(k) * 12))
            return 1;
        out_us[k - from] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    return 0;
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
    d->events.clear();
    d->ground_cols.clear();
    d->cluster_cols.clear();
    d->ground_cells.clear();
    d->cluster_cells.clear();
    d->clusters.clear();
    d->cluster_points.clear();
}

void drv_set_callbacks(drv_t*, int) {}

// the caller excerpts need real `Point` objects; the restatement works on flat arrays and has none
int drv_has_caller_excerpts(void)
{
    return 0;
}
void drv_set_cloud_record(drv_t*, int) {}
int64_t drv_num_clouds(drv_t*)
{
    return 0;
}
int64_t drv_cloud_bytes(drv_t*)
{
    return 0;
}
void drv_get_clouds(drv_t*, drv_cloud_t*, uint8_t*) {}
void drv_kitti_begin(drv_t*, int, int, const int32_t*) {}
int64_t drv_kitti_get(drv_t*, int, uint8_t*, uint32_t*)
{
    return -1;
}

} // extern "C"
