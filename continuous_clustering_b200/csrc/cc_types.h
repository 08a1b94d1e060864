// cc_types.h -- device-side data layout of one sensor stream (one cc_handle_t).
//
// HBM layout: the continuous range image is a ring of `ringcols = 10 * num_columns` columns (cpp:17) kept as a
// column-major structure of arrays, cell = local_col * R + row (same order as cpp:181, so a window of
// consecutive columns of one field is one contiguous span and a warp reading one column is one coalesced
// request). One array per field group instead of the reference's 232-byte AoS `Point` (hpp:126-161).
#ifndef CC_TYPES_H
#define CC_TYPES_H

#include <stdint.h>

#include "cc_platform.h"

// label values: the PointCloudColors entries the reference uses as labels (general.hpp:208-357, hpp:15-22)
enum : uint8_t
{
    CC_DARKRED = 32,
    CC_GRAY = 53,
    CC_GREEN = 54,
    CC_LIGHTGRAY = 71,
    CC_MAGENTA = 85,
    CC_ORANGE = 105,
    CC_RED = 119,
    CC_VIOLET = 141,
    CC_WHITE = 143,
    CC_YELLOW = 145,
    CC_YELLOWGREEN = 146,
    CC_GP_UNKNOWN = CC_WHITE,
    CC_GP_GROUND = CC_GREEN,
    CC_GP_OBSTACLE = CC_RED,
    CC_GP_EGO_VEHICLE = CC_MAGENTA,
    CC_GP_FOG = CC_LIGHTGRAY
};

#define CC_NONE 0xffffffffu
#define CC_INVALID_CWR (-2147483647 - 1)
#define CC_COL_INF 0x3fffffffffffffffLL
#define CC_K1_POINTS_PER_CHUNK 8192 /* the insertion scan stages min(128, 8192 / rows) firings per cp.async group */
#define CC_K1_MAX_CHUNK 128
#define CC_LINK_SLOTS 4 /* tree<->tree link candidates kept per probed point before the overflow list is used */
#define CC_K1_SLOW_RUN 4 /* firings that go through the per-firing path after an irregular one */
#define CC_K1_WINDOW 64 /* columns of per-row occupancy history kept in shared memory by the insertion scan */

// device-detected conditions (CcDevState::error)
enum
{
    CC_DEV_OK = 0,
    CC_DEV_COLUMN_NOT_CLEARED = 1, // cpp:321-345
    CC_DEV_TOO_MANY_COLUMNS = 2,   // more new columns in one push than the handle was sized for
    CC_DEV_LIST_OVERFLOW = 3,      // unfinished-tree list / edge list / cluster buffers full
    CC_DEV_RING_START_DECREASED = 4 // cpp:1072-1075
};

struct CcDevCfg // plain copy of cc_config_t + derived values (cpp:13-17, 80, 302-303); passed by value to kernels
{
    int R, N, ringcols, half;
    int clockwise, supplement;
    float width;
    float max_slope, first_max, first_min, lg_slope, lg_dist, close_z, close_d, next_obst_d;
    int use_terrain;
    float h_max, h_ground, l_front, l_rear, w_left, w_right;
    int fog_enabled, fog_intensity;
    float fog_dist, fog_incl;
    float max_distance, max_distance_sq;
    int max_steps_row, max_steps_col, stop_enabled, stop_min_steps, chessboard, incl_rule, use_last_stamp, nth;
    float height_sensor_to_ground;
    int debug_flag_period; // test hook: treat every n-th column as flagged (0 = off)
    double robot_from_sensor[12];
};

struct CcDevState // persistent scalars of the stream, resident in HBM; copied to the host once per push
{
    // continuous range image generation (hpp:254-259)
    long long P;        // srig_previous_global_column_index_of_rearmost_laser
    long long foremost; // srig_previous_global_column_index_of_foremost_laser
    long long F;        // srig_first_unfinished_global_column_index
    long long ring_start, ring_end; // hpp:250-251
    long long first_unpub;          // sc_first_unpublished_global_column_index (hpp:270)
    unsigned long long cluster_counter; // sc_cluster_counter_ (hpp:274)
    double runmax_carry;                // max over all finish passes so far of the column's minimum azimuth
    int reset_required;
    int error;
    long long err_a, err_b;
    // this push
    long long colbase; // first column that went through segmentation in this push
    int ncols;         // number of such columns
    int n_ulist;       // unfinished point trees (sc_unfinished_point_trees_, hpp:273)
    int n_ulist_saved;
    int n_edges;       // tree<->tree link candidates found by the probe
    int n_probe;       // non-ignored points of the new columns (association probe work list)
    int n_heavy;       // of those, the points left to the warp-cooperative walk
    int n_flagged;     // columns with an association that the reference might have refused (cpp:654-659, 688-690)
    long long danger_col; // first column at which a cluster could be force-finished (cpp:909-919), CC_COL_INF if none
    long long forced_col; // last column whose exact finish pass force-finished a component in this push, -1 if none
    int abort;            // speculative commit must be rolled back
    int n_clusters, n_cluster_points;
    long long clear_from, clear_to;   // columns retired by this push [from, to)
    long long clear2_from, clear2_to; // columns retired by the push before (recycled at the start of the next push)
    long long seg_c0, seg_c1;       // column range of the running commit segment (inclusive)
    long long seg_first_unpub_old;
    long long gbase; // column of entry 0 of the per-root-column arrays
    long long scan_base;            // column that o_g is relative to (rearmost column when the push started)
    long long push_first_unpub_old; // first_unpub before the first finish pass of this push
    int sv_n_clusters, sv_n_cluster_points;
    int scan_fast_firings, scan_slow_firings, scan_fast_attempts; // insertion scan statistics of this push
    int scan_kbad;            // firings [0, scan_kbad) were resolved by the lite insertion path
    long long scan_lite_base; // column the lite arrays are relative to
    int scan_lite_firings;
    int ticket_gap;    // same for k_gap_scan (the last block chains the column chunks)
    int n_vfix; // points whose visit count was redone with the walk cut at the first unpublished column (d_visited_fix)
    int halted; // set when a push could not be committed speculatively: later pushes in flight skip themselves
};

struct CcFiringRecord // insertion scan -> K1b: how the points of one firing were resolved
{
    int mode; // 1: regular firing, column = unwrap(cwr) with the integers below; 0: per point in o_g / o_rot
    int goff, pc, rot, P;
    int pad_[3];
};

struct CcFiringSummary // k_prep -> lite insertion path
{
    int anchor;   // column-in-rotation of the firing's first valid row
    int rear_rel; // rearmost / foremost column of the firing relative to the anchor (wrapped into (-N/2, N/2])
    int fore_rel;
    int nvalid;   // valid points; -1 = a column-in-rotation outside [0, N]: per-firing path only
};

struct CcCluster // device -> host record of one finished cluster with more than 5 points (cpp:936-940)
{
    unsigned long long id;
    unsigned long long min_stamp, max_stamp;
    long long finish_col, min_col, max_col;
    unsigned int num_points, point_offset;
    unsigned int cursor, pad_;
};

struct CcClusterPoint
{
    long long gcol;
    int row;
    int pad_;
};

struct CcDevPtrs
{
    CcDevState* st;
    // ---- ring (ringcols * R cells) ----
    float4* pos;     // x, y, z (odom frame), distance                                   cpp:223-229
    float* azimuth;  // Point::azimuth_angle
    float* incl;     // Point::inclination_angle (NaN cells supplemented, cpp:364-369)
    double* cont_az; // Point::continuous_azimuth_angle
    uchar4* lab;     // ground_point_label, debug_ground_point_label, is_ignored, intensity
    unsigned long long* stamp;
    unsigned long long* guid;
    unsigned long long* firing_index;
    float4* assoc; // association view written by segmentation: x (NaN when is_ignored), y, z, inclination
    float* mad;    // asinf(max_distance / distance) of non-ignored cells (cpp:805), else 0
    unsigned int* tparent; // point tree: first-hit parent while probing, tree root once committed (tree_root_)
    unsigned int* tfirst;  // the first hit of the walk: the point whose child_points list holds this point (cpp:663)
    unsigned int* cparent; // union-find over tree roots (replaces associated_trees, hpp:149)
    unsigned long long* tfinish; // root: finished_at_continuous_azimuth_angle as ordered bits (hpp:146)
    long long* tmaxcol;          // root: last global column of the tree (root col + cluster_width - 1)
    unsigned int* tnpoints;      // root: tree_num_points
    unsigned int* tstate;        // root: 0 unfinished, else 1 + commit sequence number in which it finished
    unsigned int* tid;           // root: cluster id of its finished cluster (0 = none / <= 5 points)
    int* tslot;                  // root: cluster slot in the current push (-1 none)
    unsigned int* rootslot;      // root: index in the unfinished list
    unsigned int* cid;           // Point::id
    unsigned short* visited;     // Point::number_of_visited_neighbors
    unsigned char* vback;        // columns back the association walk of the point got (d_visited_fix)
    long long* slot_gcol;        // per ring column: global column segmented into it, -1 = cleared
    // ---- per-row carried state ----
    float* gap_state;     // sc_inclination_angles_between_lasers_ (hpp:275)
    long long* rowmax;    // per row: last column written by insertion
    // ---- per push staging (max_firings * R) ----
    const void* raw;      // cc_raw_point_t[n * R]
    const double* poses;  // [n][12]
    float4* s_pos;        // odom x, y, z, distance
    float* s_dist;        // distance again, contiguous (scan input)
    float* s_az;
    float* s_incl;
    float* s_incaz;
    double* s_ego;        // [max_firings][12] robot_from_sensor * odom_from_sensor^-1 of every firing (ego-box test)
    int* s_cwr;
    int* s_cwrT;          // the same, [row][max_firings]
    int* o_g;             // resolved global column relative to CcDevState::scan_base, INT_MIN = not stored
    int* o_rot;           // rotation index used for the continuous azimuth
    CcFiringRecord* firing_rec; // [max_firings]
    CcFiringSummary* lite_sum;  // [max_firings]
    int* lite_U;                // [max_firings] unwrapped anchor column, relative to scan_lite_base
    int* lite_P;                // [max_firings + 1] rearmost column so far before firing k
    int* lite_F;                // [max_firings + 1] foremost column so far before firing k
    int* lite_rowfront;         // [R] last stored column of the row within the lite prefix (INT_MIN none)
    // ---- per new column (maxcols) ----
    int* col_trigger;     // firing (index in this push) whose insertion completed the column (hpp:169-173)
    float* col_gap;       // [maxcols * R] value of sc_inclination_angles_between_lasers_ when the column is segmented
    float* gap_chunk_last;  // [maxcols / CC_GAP_CHUNK + 1][R] last valid value inside a chunk of columns
    float* gap_chunk_carry; // same shape: value carried into the chunk
    double* col_minaz;    // current_minimum_continuous_azimuth_angle (cpp:777, 791-793)
    double* col_runmax;   // running max of col_minaz including earlier pushes
    long long* col_first_unpub; // sc_first_unpublished_global_column_index after the column's pass
    int* probe_list;            // [maxcols * R] new-column cell (ci * R + row) of every non-ignored point
    int* heavy_list;            // [maxcols * R] subset of probe_list for k_probe_heavy
    unsigned char* col_flag;    // association of this column must be redone column-sequentially
    // ---- clustering scratch ----
    unsigned int* ulist;     // unfinished roots (current)
    unsigned int* ulist_new; // compaction target
    int* u_rep;              // per list entry: list index of its component representative
    unsigned long long* u_maxfinish;
    long long* u_mincol;
    long long* u_maxend;
    unsigned int* u_np;
    long long* u_finishcol;
    int* u_cluster;
    // snapshot of list-root state for rolling back a speculative commit
    unsigned int* sv_cparent;
    unsigned long long* sv_tfinish;
    long long* sv_tmaxcol;
    unsigned int* sv_tnpoints;
    unsigned int* edge_a; // link candidates (cell, cell)
    unsigned int* edge_b;
    long long* G;          // per root column (gbase + i): last column at which a tree rooted there is unfinished
    CcCluster* clusters;
    CcClusterPoint* cluster_points;
    int* n_new_ulist;
    unsigned long long* trace; // optional device-side timeline (cc_debug_trace), else null
    int cap_ulist, cap_edges, cap_clusters, cap_cluster_points, cap_G, maxcols, max_firings;
};

#endif
