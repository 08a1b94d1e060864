// cc_api.cu -- host side of the C ABI declared in include/cc_b200.h: owns the device memory of one sensor
// stream, stages firings, launches the kernels of cc_kernels.cuh on the handle's stream and brings the results
// of a push back to the host. There is no CPU implementation of the path in this library.
#include "../../include/cc_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cc_kernels.cuh"
#include "cc_eval.cuh"
#include "cc_kitti.cuh"
#include "cc_packets.cuh"

// optimistic prefix of a push's results brought to the host behind it: cluster records, and member lists for a quarter of
// the cells of the largest push (a synthetic street scene finishes ~8 % of a push's cells as cluster members)
static const int CC_PREFETCH_CLUSTERS = 1024;

namespace
{

struct Alloc
{
    void** slot;
};

} // namespace

struct cc_handle
{
    int device{0};
    cudaStream_t stream{nullptr};
    cudaEvent_t ev0{nullptr}, ev1{nullptr};
    std::string error;
    cc_config_t config{};
    bool config_set{false};
    bool is_reset{false};
    bool reset_required_cfg{false}; // set by cc_set_config (cpp:69-74)
    bool has_tf{false};
    double robot_from_sensor[12]{};
    int R{-1}, N{0}, ringcols{0};
    float width{0.f};
    int max_firings{4096};
    int maxcols{0};
    int gap_rows{-1};
    int debug_flag_period{0};
    unsigned long long* d_trace{nullptr}; // device-side timeline buffer (cc_debug_trace)
    bool trace_on{false};
    int tune{0}; // profiling aid: CC_B200_TUNE environment variable (bit 0: cooperative probe walk without masks; bit 2: tiled probe)
    bool label_prefetch{false};
    const uchar4* cur_labels{nullptr}; // labels of the last finished push (pinned slot buffer)
    int cur_label_cols{0};
    int used_exact_flag{0};
    size_t probe_smem{0}; // [block-scan scratch][running maxima of up to maxcols columns]
    size_t tile_smem_set{0};
    size_t ground_smem_set{0};
    size_t lite_smem_set{0};
    size_t fin_smem_set{0};
    CcDevPtrs d{};
    unsigned int* d_s_parent{nullptr};
    unsigned int* d_s_links{nullptr};
    // per-push buffers, double-buffered so that a push can be in flight while the previous one is collected
    struct Slot
    {
        // device
        CcDevState* d_state_snap{nullptr};
        long long* d_first_unpub{nullptr};
        CcCluster* d_clusters{nullptr};
        CcClusterPoint* d_points{nullptr};
        uchar4* d_labels{nullptr};
        // page-locked host
        CcDevState* h_state{nullptr};
        long long* h_first_unpub{nullptr};
        CcCluster* h_clusters{nullptr};
        CcClusterPoint* h_points{nullptr}; // one of the handle's three member-list buffers (borrowed)
        uchar4* h_labels{nullptr}; // one of the handle's three label buffers (borrowed)
        cudaEvent_t ev0{nullptr}, ev1{nullptr}, ready{nullptr}, done{nullptr};
        cudaEvent_t h2d{nullptr}, h2d0{nullptr}; // of the input buffer the push reads (borrowed, not owned)
        CcHostHeader* h_hdr{nullptr};            // page-locked: completion flag + device timestamps of a fused push
        bool fused{false}, want_fused{false};
        // the parameters the push was SUBMITTED under (cc_set_config / cc_set_robot_from_sensor may be called while it is
        // in flight or staged): used for its kernels, for a re-run after the exact path, and for its events and stamps
        CcDevCfg cfg;
        bool cfg_has_tf{false};
        unsigned int ticket{0};
        const void* src_points{nullptr}; // fused push: where the kernel fetches the firings from (page-locked host memory)
        const double* src_poses{nullptr};
        // the push occupying the slot
        int n{0};
        const void* in_points{nullptr}; // device pointers of the inputs
        const double* in_poses{nullptr};
        bool has_tf{false}, spec{false};
        uint64_t launches0{0}, launches1{0};
        int pre_cols{0}, pre_clusters{0}, pre_points{0};
    } slots[2];
    // Input staging for host pushes, one more buffer than pushes in flight: a push submitted while two are in flight
    // starts its host -> device copy at once and waits "staged"; its kernels are launched when the oldest push in flight
    // has been waited for. The copy of push k + 2 thereby overlaps the kernels of pushes k and k + 1.
    struct InBuf
    {
        unsigned char* d_raw{nullptr};
        double* d_poses{nullptr};
        void* h_raw{nullptr}; // page-locked staging for callers whose buffers are pageable
        double* h_poses{nullptr};
        cudaEvent_t h2d0{nullptr}, h2d{nullptr};
        int n{0};
        CcDevCfg cfg; // snapshot taken when the push was submitted (see Slot::cfg)
        bool cfg_has_tf{false};
        bool fused{false}; // no copy-engine transfer was queued: the fused kernel fetches src_* itself
        const void* src_points{nullptr};
        const double* src_poses{nullptr};
    } inbuf[3];
    int next_in{0};
    // page-locked label buffers: two pushes in flight + the labels of the last finished push, which stay valid until
    // the next cc_wait()
    uchar4* h_label_ring[3]{nullptr, nullptr, nullptr};
    CcClusterPoint* h_points_ring[3]{nullptr, nullptr, nullptr}; // same rotation: the member lists are handed out as views
    const cc_cluster_point_t* points_view{nullptr}; // member lists of the last finished push
    size_t n_events{0};                             // events of the last finished push (h->events keeps its capacity)
    int next_label{0};
    int staged{-1}; // input buffer of the staged push, -1 none
    int next_slot{0};
    int pending[2]{-1, -1}; // slots of the pushes in flight, oldest first
    int n_pending{0};
    cudaStream_t copy_stream{nullptr};
    // host -> device copies of the raw firings run on their own stream: the copy stream carries the result read-back of
    // the push before, which waits for that push's kernels; inputs queued behind it could not overlap them
    cudaStream_t in_stream{nullptr};
    // host-initiated reads of finished results (remainders beyond the prefetch, cc_read_columns): never queued behind
    // the copy stream's wait for a push that is still in flight
    cudaStream_t aux_stream{nullptr};
    std::vector<void*> allocs;       // freed on destroy / re-reset
    std::vector<void*> allocs_fixed; // independent of the ring size
    CcCell* h_export{nullptr}; // page-locked: cells of the last cc_export_columns / cc_read_columns call
    size_t export_cap{0};      // in cells
    unsigned char* h_cloud{nullptr}; // page-locked: payload of the last cc_pack_*_pointcloud2 call
    size_t cloud_cap{0};
    unsigned int* d_counts{nullptr}; // child counts of the columns being packed
    size_t counts_cap{0};
    unsigned long long* d_min_stamp{nullptr};
    unsigned long long* h_min_stamp{nullptr};
    CcPackRequest* h_requests{nullptr}; // page-locked request table of cc_pack_requests_pointcloud2 + min stamps behind it
    CcPackRequest* d_requests{nullptr};
    unsigned long long* d_req_stamps{nullptr};
    unsigned long long* h_req_stamps{nullptr};
    int requests_cap{0};
    const CcClusterPoint* last_points_dev{nullptr}; // member lists of the last finished push, on the device
    CcDevState* h_state{nullptr}; // pinned mirror (reset, column-sequential path)
    CcDevState state{};
    unsigned int seq{0};
    uint64_t launches{0};
    uint64_t launches_at_push_start{0};
    int sm_count{148};
    // fused single-launch path for short pushes (k_push_fused): one thread-block cluster, inputs read from and results
    // written to page-locked host memory by the kernel itself
    int fused_max{384};     // pushes of at most this many firings take it (CC_B200_FUSED_MAX; 0 = never): measured
                            // crossover with the kernel chain on a B200 at ~420 firings of 64 rows
    int fused_cluster{0};   // CTAs per cluster (0: not available on this device / configuration)
    int fused_threads{512};
    size_t fused_smem{0};
    int fin_cluster{0};       // CTAs of k_fin_cluster (0: one-CTA k_fin_all)
    size_t fin_cluster_smem{0};
    unsigned int ticket{0};
    int prefetch_points{16384};
    int occ_probe{8}, occ_probe_heavy{8}; // resident CTAs per SM of the two association kernels
    // results of the last push
    cc_batch_info_t info{};
    std::vector<cc_column_event_t> events;
    std::vector<cc_cluster_t> clusters;
    std::vector<cc_cluster_point_t> cluster_points;
    std::vector<long long> h_first_unpub;
    std::vector<unsigned char> h_flags;
    std::vector<CcCluster> h_clusters;
    // per-kernel timing
    bool timing{false};
    int n_timed{0};
    std::vector<cudaEvent_t> tev;
    std::vector<const char*> tnames;
};

static int timing_begin(cc_handle* h, const char* name);
static void timing_end(cc_handle* h, int i);
static void free_host_slots(cc_handle* h);

#define CC_CHECK(h, expr)                                                                                              \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t cc_e_ = (expr);                                                                                    \
        if (cc_e_ != cudaSuccess)                                                                                      \
        {                                                                                                              \
            (h)->error = std::string(#expr) + ": " + cudaGetErrorString(cc_e_);                                        \
            return CC_ERR_CUDA;                                                                                        \
        }                                                                                                              \
    } while (0)

#define CC_RUN(h, kernel, grid, block, smem, ...)                                                                      \
    do                                                                                                                 \
    {                                                                                                                  \
        const int cc_t_ = timing_begin((h), #kernel);                                                                  \
        CC_LAUNCH(kernel, grid, block, smem, (h)->stream, __VA_ARGS__);                                                \
        timing_end((h), cc_t_);                                                                                        \
        (h)->launches++;                                                                                               \
    } while (0)

// optional per-kernel CUDA-event timing of one push (cc_set_kernel_timing); off by default
static int timing_begin(cc_handle* h, const char* name)
{
    if (!h->timing || h->n_timed >= static_cast<int>(h->tev.size()) / 2)
        return -1;
    const int i = h->n_timed++;
    h->tnames[i] = name;
    cudaEventRecord(h->tev[2 * i], h->stream);
    return i;
}
static void timing_end(cc_handle* h, int i)
{
    if (i >= 0)
        cudaEventRecord(h->tev[2 * i + 1], h->stream);
}

static void free_host_slots(cc_handle* h)
{
    for (cc_handle::Slot& sl : h->slots)
    {
        for (void* q : {static_cast<void*>(sl.h_state), static_cast<void*>(sl.h_first_unpub), static_cast<void*>(sl.h_clusters),
                        static_cast<void*>(sl.h_hdr)})
            if (q)
                cudaFreeHost(q);
        sl.h_hdr = nullptr;
        sl.h_state = nullptr;
        sl.h_first_unpub = nullptr;
        sl.h_clusters = nullptr;
        sl.h_points = nullptr;
        sl.h_labels = nullptr;
    }
    for (uchar4*& q : h->h_label_ring)
    {
        if (q)
            cudaFreeHost(q);
        q = nullptr;
    }
    for (CcClusterPoint*& q : h->h_points_ring)
    {
        if (q)
            cudaFreeHost(q);
        q = nullptr;
    }
    for (cc_handle::InBuf& ib : h->inbuf)
    {
        if (ib.h_raw)
            cudaFreeHost(ib.h_raw);
        if (ib.h_poses)
            cudaFreeHost(ib.h_poses);
        ib.h_raw = nullptr;
        ib.h_poses = nullptr;
    }
}

static void free_list(std::vector<void*>& v)
{
    for (void* p : v)
        cudaFree(p);
    v.clear();
}

template<typename T>
static cudaError_t dev_alloc(cc_handle* h, std::vector<void*>& list, T** out, size_t count)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess)
        return e;
    list.push_back(p);
    *out = static_cast<T*>(p);
    return cudaSuccess;
}

static void fill_devcfg(const cc_handle* h, CcDevCfg& c)
{
    const cc_config_t& s = h->config;
    std::memset(&c, 0, sizeof(c));
    c.R = h->R;
    c.N = h->N;
    c.ringcols = h->ringcols;
    c.half = h->N / 2;
    c.clockwise = s.sensor_is_clockwise != 0;
    c.supplement = s.supplement_inclination_angle_for_nan_cells != 0;
    c.width = h->width;
    c.max_slope = s.max_slope;
    c.first_max = s.first_ring_as_ground_max_allowed_z_diff;
    c.first_min = s.first_ring_as_ground_min_allowed_z_diff;
    c.lg_slope = s.last_ground_point_slope_higher_than;
    c.lg_dist = s.last_ground_point_distance_smaller_than;
    c.close_z = s.ground_because_close_to_last_certain_ground_max_z_diff;
    c.close_d = s.ground_because_close_to_last_certain_ground_max_dist_diff;
    c.next_obst_d = s.obstacle_because_next_certain_obstacle_max_dist_diff;
    c.use_terrain = s.use_terrain != 0;
    c.h_max = s.height_ref_to_maximum_;
    c.h_ground = s.height_ref_to_ground_;
    c.l_front = s.length_ref_to_front_end_;
    c.l_rear = s.length_ref_to_rear_end_;
    c.w_left = s.width_ref_to_left_mirror_;
    c.w_right = s.width_ref_to_right_mirror_;
    c.fog_enabled = s.fog_filtering_enabled != 0;
    c.fog_intensity = static_cast<int>(static_cast<uint8_t>(s.fog_filtering_intensity_below));
    c.fog_dist = s.fog_filtering_distance_below;
    c.fog_incl = s.fog_filtering_inclination_above;
    c.max_distance = s.max_distance;
    c.max_distance_sq = s.max_distance * s.max_distance; // cpp:80
    c.max_steps_row = s.max_steps_in_row;
    c.max_steps_col = s.max_steps_in_column;
    c.stop_enabled = s.stop_after_association_enabled != 0;
    c.stop_min_steps = s.stop_after_association_min_steps;
    c.chessboard = s.ignore_points_in_chessboard_pattern != 0;
    c.incl_rule = s.ignore_points_with_too_big_inclination_angle_diff != 0;
    c.use_last_stamp = s.use_last_point_for_cluster_stamp != 0;
    c.nth = s.cluster_point_trees_every_nth_column > 0 ? s.cluster_point_trees_every_nth_column : 1;
    std::memcpy(c.robot_from_sensor, h->robot_from_sensor, sizeof(c.robot_from_sensor));
    c.debug_flag_period = h->debug_flag_period;
    c.height_sensor_to_ground = -static_cast<float>(h->robot_from_sensor[11]) + s.height_ref_to_ground_; // cpp:302-303
}

extern "C" {

const char* cc_version(void)
{
#ifdef CC_EMU
    return "continuous_clustering_b200 0.1 (CPU emulation build: tests only)";
#else
    return "continuous_clustering_b200 0.1 (sm_100a)";
#endif
}

void cc_config_default(cc_config_t* c)
{
    std::memset(c, 0, sizeof(*c));
    c->is_single_threaded = 0;
    c->sensor_is_clockwise = 1;
    c->num_columns = 1700;
    c->supplement_inclination_angle_for_nan_cells = 1;
    c->max_slope = 0.2f;
    c->first_ring_as_ground_max_allowed_z_diff = 0.4f;
    c->first_ring_as_ground_min_allowed_z_diff = -0.4f;
    c->last_ground_point_slope_higher_than = -0.1f;
    c->last_ground_point_distance_smaller_than = 5.f;
    c->ground_because_close_to_last_certain_ground_max_z_diff = 0.4f;
    c->ground_because_close_to_last_certain_ground_max_dist_diff = 2.0f;
    c->obstacle_because_next_certain_obstacle_max_dist_diff = 0.3f;
    c->use_terrain = 0;
    c->terrain_max_allowed_z_diff = 0.4f;
    c->fog_filtering_enabled = 0;
    c->fog_filtering_intensity_below = 2;
    c->fog_filtering_distance_below = 18.f;
    c->fog_filtering_inclination_above = -0.06f;
    c->max_distance = 0.7f;
    c->max_steps_in_row = 20;
    c->max_steps_in_column = 20;
    c->stop_after_association_enabled = 1;
    c->stop_after_association_min_steps = 1;
    c->ignore_points_in_chessboard_pattern = 1;
    c->ignore_points_with_too_big_inclination_angle_diff = 1;
    c->use_last_point_for_cluster_stamp = 0;
    c->cluster_point_trees_every_nth_column = 1;
}

cc_status_t cc_create(int device_ordinal, int max_firings_per_push, cc_handle_t** out)
{
    if (!out)
        return CC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device_ordinal < 0 || device_ordinal >= ndev)
        return CC_ERR_CUDA; // no CUDA device: there is no other way to run this path
    cc_handle* h = new cc_handle();
    h->device = device_ordinal;
    if (max_firings_per_push > 0)
        h->max_firings = std::min(max_firings_per_push, 8192); // per-firing scan arrays must fit one CTA's shared memory
    cc_config_default(&h->config);
    if (const char* t = std::getenv("CC_B200_TUNE"))
        h->tune = std::atoi(t);
    if (const char* t = std::getenv("CC_B200_FUSED_MAX"))
        h->fused_max = std::max(0, std::atoi(t));
    if (cudaSetDevice(h->device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&h->h_state), sizeof(CcDevState)) != cudaSuccess)
    {
        delete h;
        return CC_ERR_CUDA;
    }
    for (cc_handle::Slot& sl : h->slots)
        if (cudaEventCreate(&sl.ev0) != cudaSuccess || cudaEventCreate(&sl.ev1) != cudaSuccess ||
            cudaEventCreate(&sl.ready) != cudaSuccess || cudaEventCreate(&sl.done) != cudaSuccess ||
            false)
        {
            delete h;
            return CC_ERR_CUDA;
        }
    for (cc_handle::InBuf& ib : h->inbuf)
        if (cudaEventCreate(&ib.h2d) != cudaSuccess || cudaEventCreate(&ib.h2d0) != cudaSuccess)
        {
            delete h;
            return CC_ERR_CUDA;
        }
#ifndef CC_EMU
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess)
        h->sm_count = prop.multiProcessorCount;
    {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_probe_heavy, 64, cc_heavy_smem_bytes(64, 2)) == cudaSuccess && nb > 0)
            h->occ_probe_heavy = nb;
    }
#endif
    *out = h;
    return CC_OK;
}

void cc_destroy(cc_handle_t* h)
{
    if (!h)
        return;
    cudaSetDevice(h->device);
    if (h->stream)
        cudaStreamSynchronize(h->stream);
    free_list(h->allocs);
    free_list(h->allocs_fixed);
    if (h->copy_stream)
        cudaStreamSynchronize(h->copy_stream);
    free_host_slots(h);
    for (cc_handle::Slot& sl : h->slots)
        for (cudaEvent_t e : {sl.ev0, sl.ev1, sl.ready, sl.done})
            if (e)
                cudaEventDestroy(e);
    for (cc_handle::InBuf& ib : h->inbuf)
        for (cudaEvent_t e : {ib.h2d, ib.h2d0})
            if (e)
                cudaEventDestroy(e);
    if (h->copy_stream)
        cudaStreamDestroy(h->copy_stream);
    if (h->in_stream)
    {
        cudaStreamSynchronize(h->in_stream);
        cudaStreamDestroy(h->in_stream);
    }
    if (h->aux_stream)
        cudaStreamDestroy(h->aux_stream);
    if (h->h_state)
        cudaFreeHost(h->h_state);
    if (h->h_export)
        cudaFreeHost(h->h_export);
    if (h->h_cloud)
        cudaFreeHost(h->h_cloud);
    if (h->h_min_stamp)
        cudaFreeHost(h->h_min_stamp);
    if (h->d_counts)
        cudaFree(h->d_counts);
    if (h->d_min_stamp)
        cudaFree(h->d_min_stamp);
    if (h->h_requests)
        cudaFreeHost(h->h_requests);
    if (h->h_req_stamps)
        cudaFreeHost(h->h_req_stamps);
    if (h->d_requests)
        cudaFree(h->d_requests);
    if (h->d_req_stamps)
        cudaFree(h->d_req_stamps);
    if (h->d_trace)
        cudaFree(h->d_trace);
    if (h->ev0)
        cudaEventDestroy(h->ev0);
    if (h->ev1)
        cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : h->tev)
        cudaEventDestroy(e);
    if (h->stream)
        cudaStreamDestroy(h->stream);
    delete h;
}

const char* cc_last_error(const cc_handle_t* h)
{
    return h ? h->error.c_str() : "null handle";
}

cc_status_t cc_set_config(cc_handle_t* h, const cc_config_t* cfg)
{
    if (!h || !cfg)
        return CC_ERR_INVALID_ARGUMENT;
    // (the association probe keeps one running maximum per new column in shared memory: 8192 firings + num_columns
    // columns at most; sensors of the reference have 1024 .. 2200 columns per rotation)
    if (cfg->num_columns <= 1 || cfg->num_columns > 16384 || cfg->cluster_point_trees_every_nth_column <= 0)
    {
        h->error = "invalid configuration (num_columns must be in 2 .. 16384, cluster_point_trees_every_nth_column > 0)";
        return CC_ERR_INVALID_ARGUMENT;
    }
    // cpp:66-81
    if ((h->config.is_single_threaded != 0) != (cfg->is_single_threaded != 0))
        h->reset_required_cfg = true;
    if ((h->config.sensor_is_clockwise != 0) != (cfg->sensor_is_clockwise != 0))
        h->reset_required_cfg = true;
    if (h->config.num_columns != cfg->num_columns)
        h->reset_required_cfg = true;
    h->config = *cfg;
    h->config_set = true;
    return CC_OK;
}

int cc_reset_required(const cc_handle_t* h)
{
    return h && (h->reset_required_cfg || h->state.reset_required) ? 1 : 0;
}

cc_status_t cc_set_robot_from_sensor(cc_handle_t* h, const double m[12])
{
    if (!h || !m)
        return CC_ERR_INVALID_ARGUMENT;
    std::memcpy(h->robot_from_sensor, m, sizeof(h->robot_from_sensor));
    h->has_tf = true;
    return CC_OK;
}

int cc_has_robot_from_sensor(const cc_handle_t* h)
{
    return h && h->has_tf ? 1 : 0;
}

int cc_num_rows(const cc_handle_t* h)
{
    return h ? h->R : -1;
}
int cc_num_columns(const cc_handle_t* h)
{
    return h ? h->N : 0;
}
int cc_ring_buffer_max_columns(const cc_handle_t* h)
{
    return h ? h->ringcols : 0;
}
void* cc_stream(const cc_handle_t* h)
{
    return h ? static_cast<void*>(h->stream) : nullptr;
}
uint64_t cc_total_launches(const cc_handle_t* h)
{
    return h ? h->launches : 0;
}

static int scan_chunk(int R)
{
    int c = CC_K1_MAX_CHUNK;
    while (c > 1 && c * R > CC_K1_POINTS_PER_CHUNK)
        c >>= 1;
    return c;
}

static int scan_threads()
{
#ifdef CC_EMU
    return 1;
#else
    return 1024;
#endif
}

static int scan_smem_bytes(int R, int C = 0, int T = 0)
{
    C = C > 0 ? C : scan_chunk(R);
    T = T > 0 ? T : scan_threads();
    const int nparts = T / R > 0 ? (T / R < C ? T / R : C) : 1;
    size_t words = static_cast<size_t>(CC_K1_WINDOW) * R + 2 * R + 4 * static_cast<size_t>(C) * R + 2 * 2 * 32 +
                   static_cast<size_t>(C) * (R + 1) + 2 * C + 2 * (C + 1) + 4 * static_cast<size_t>(R) * nparts + 8;
    return static_cast<int>(words * 4);
}

static int grid_for(const cc_handle* h, long long work, int block)
{
    long long g = (work + block - 1) / block;
    const long long cap = static_cast<long long>(h->sm_count) * 8;
    if (g > cap)
        g = cap;
    if (g < 1)
        g = 1;
    return static_cast<int>(g);
}

static int fused_scan_chunk(int R)
{
    const int c = scan_chunk(R);
    return c < 32 ? c : 32;
}

// The fused single-launch push (k_push_fused): dynamic shared memory = the largest need of any stage at the fused
// block size; the largest cluster the device can co-schedule (16 CTAs needs the non-portable opt-in) is used.
static void configure_fused(cc_handle* h)
{
    h->fused_cluster = 0;
    if (h->fused_max <= 0)
        return;
#ifdef CC_EMU
    h->fused_threads = 1;
#else
    h->fused_threads = 512;
#endif
    const int T = h->fused_threads, R = h->R;
    const int warps = (T + CC_WARP - 1) / CC_WARP;
    if (h->fused_max > CC_WARP * CC_SMALL_PER)
        h->fused_max = CC_WARP * CC_SMALL_PER; // the warp-wide scans of the fused kernel
    size_t smem = cc_lite_small_smem_bytes(h->fused_max);
    smem = std::max(smem, static_cast<size_t>(scan_smem_bytes(R, fused_scan_chunk(R), T)));
    smem = std::max(smem, warps * cc_ground_warp_bytes(R));
    smem = std::max(smem, h->probe_smem);
    smem = std::max(smem, cc_heavy_smem_bytes(T, 2));
    smem = std::max(smem, static_cast<size_t>(warps) * CC_PROBE_PIPE * CC_WARP * sizeof(float4)); // d_visited_fix
    smem = std::max(smem, static_cast<size_t>(T) * 8 + static_cast<size_t>(h->d.cap_G) * sizeof(int));
    smem = std::max(smem, static_cast<size_t>(2 * 512 * sizeof(int)));
    if (smem > 200 * 1024)
        return; // such a configuration keeps the kernel chain
    h->fused_smem = smem;
#ifdef CC_EMU
    h->fused_cluster = 1;
#else
    if (cudaFuncSetAttribute(k_push_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_push_fused, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
    {
        cudaGetLastError();
        return;
    }
    int want = 16;
    if (const char* c = std::getenv("CC_B200_FUSED_CLUSTER"))
        want = std::max(1, std::min(16, std::atoi(c)));
    for (int cs = want; cs >= 1; cs >>= 1)
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(static_cast<unsigned>(cs), 1, 1);
        lc.blockDim = dim3(static_cast<unsigned>(T), 1, 1);
        lc.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, k_push_fused, &lc) == cudaSuccess && nc >= 1)
        {
            h->fused_cluster = cs;
            break;
        }
        cudaGetLastError();
    }
#endif
}

// k_fin_cluster: the finish pass of whole-push commits over one thread-block cluster
static void configure_fin_cluster(cc_handle* h)
{
    h->fin_cluster = 0;
    if (std::getenv("CC_B200_FIN_CLUSTER") && std::atoi(std::getenv("CC_B200_FIN_CLUSTER")) == 0)
        return;
    size_t smem = static_cast<size_t>(512) * 8 + static_cast<size_t>(h->d.cap_G) * sizeof(int);
    if (smem > 200 * 1024)
        smem = 200 * 1024;
    h->fin_cluster_smem = smem;
#ifdef CC_EMU
    h->fin_cluster = 1;
#else
    if (cudaFuncSetAttribute(k_fin_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_fin_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
    {
        cudaGetLastError();
        return;
    }
    for (int cs = 16; cs >= 2; cs >>= 1)
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(static_cast<unsigned>(cs), 1, 1);
        lc.blockDim = dim3(512, 1, 1);
        lc.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, k_fin_cluster, &lc) == cudaSuccess && nc >= 1)
        {
            h->fin_cluster = cs;
            break;
        }
        cudaGetLastError();
    }
#endif
}

static bool fused_eligible(const cc_handle* h, int n)
{
    return h->fused_cluster > 0 && n <= h->fused_max && !h->timing;
}

// ContinuousClustering::reset cpp:11-64
cc_status_t cc_reset(cc_handle_t* h, int num_rows)
{
    if (!h || num_rows <= 0 || num_rows > 256)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    CC_CHECK(h, cudaStreamSynchronize(h->stream));
    CC_CHECK(h, cudaStreamSynchronize(h->copy_stream));
    CC_CHECK(h, cudaStreamSynchronize(h->in_stream));
    h->n_pending = 0;
    h->next_slot = 0;
    h->staged = -1;
    const int N = h->config.num_columns;
    const bool realloc_ring = (num_rows != h->R) || (N != h->N) || h->allocs.empty();
    if (realloc_ring)
    {
        free_list(h->allocs);
        h->R = num_rows;
        h->N = N;
        h->ringcols = N * 10;
        const size_t cells = static_cast<size_t>(h->ringcols) * h->R;
        const size_t stage = static_cast<size_t>(h->max_firings) * h->R;
        h->maxcols = std::min(h->max_firings + N, 8 * N); // relative column arithmetic must stay inside one ring length
        CcDevPtrs& d = h->d;
        const int keep_gap_rows = h->gap_rows;
        float* keep_gap = d.gap_state;
        std::memset(&d, 0, sizeof(d));
        d.gap_state = keep_gap;
        (void)keep_gap_rows;
        std::vector<void*>& L = h->allocs;
        CC_CHECK(h, dev_alloc(h, L, &d.st, 1));
        CC_CHECK(h, dev_alloc(h, L, &d.pos, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.azimuth, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.incl, cells + 1));
        CC_CHECK(h, dev_alloc(h, L, &d.cont_az, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.lab, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.stamp, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.guid, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.firing_index, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.assoc, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.mad, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tparent, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tfirst, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.cparent, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tfinish, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tmaxcol, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tnpoints, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tstate, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tid, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.tslot, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.rootslot, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.cid, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.visited, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.vback, cells));
        CC_CHECK(h, dev_alloc(h, L, &d.slot_gcol, static_cast<size_t>(h->ringcols)));
        CC_CHECK(h, dev_alloc(h, L, &d.rowmax, static_cast<size_t>(h->R)));
        CC_CHECK(h, dev_alloc(h, L, &d.s_pos, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.s_dist, stage + static_cast<size_t>(CC_K1_MAX_CHUNK) * h->R));
        CC_CHECK(h, dev_alloc(h, L, &d.s_az, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.s_incl, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.s_incaz, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.s_ego, static_cast<size_t>(h->max_firings) * 12));
        CC_CHECK(h, dev_alloc(h, L, &d.s_cwr, stage + static_cast<size_t>(CC_K1_MAX_CHUNK) * h->R));
        CC_CHECK(h, dev_alloc(h, L, &d.s_cwrT, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.o_g, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.o_rot, stage));
        CC_CHECK(h, dev_alloc(h, L, &d.firing_rec, static_cast<size_t>(h->max_firings)));
        CC_CHECK(h, dev_alloc(h, L, &d.lite_sum, static_cast<size_t>(h->max_firings)));
        CC_CHECK(h, dev_alloc(h, L, &d.lite_U, static_cast<size_t>(h->max_firings)));
        CC_CHECK(h, dev_alloc(h, L, &d.lite_P, static_cast<size_t>(h->max_firings) + 1));
        CC_CHECK(h, dev_alloc(h, L, &d.lite_F, static_cast<size_t>(h->max_firings) + 1));
        CC_CHECK(h, dev_alloc(h, L, &d.lite_rowfront, static_cast<size_t>(h->R)));
        const size_t mc = static_cast<size_t>(h->maxcols);
        CC_CHECK(h, dev_alloc(h, L, &d.col_trigger, mc));
        CC_CHECK(h, dev_alloc(h, L, &d.col_gap, mc * h->R));
        CC_CHECK(h, dev_alloc(h, L, &d.gap_chunk_last, (mc / CC_GAP_CHUNK + 1) * h->R));
        CC_CHECK(h, dev_alloc(h, L, &d.gap_chunk_carry, (mc / CC_GAP_CHUNK + 1) * h->R));
        CC_CHECK(h, dev_alloc(h, L, &d.col_minaz, mc));
        CC_CHECK(h, dev_alloc(h, L, &d.col_runmax, mc));
        CC_CHECK(h, dev_alloc(h, L, &d.col_flag, mc));
        CC_CHECK(h, dev_alloc(h, L, &d.probe_list, mc * h->R));
        CC_CHECK(h, dev_alloc(h, L, &d.heavy_list, mc * h->R));
        CC_CHECK(h, dev_alloc(h, L, &h->d_s_parent, mc * h->R));
        CC_CHECK(h, dev_alloc(h, L, &h->d_s_links, mc * h->R * CC_LINK_SLOTS));
        d.cap_ulist = 1 << 18;
        d.cap_edges = 1 << 22;
        d.cap_clusters = 1 << 16;
        d.cap_cluster_points = static_cast<int>(std::min<size_t>(mc * h->R + (static_cast<size_t>(2) * N * h->R), 1u << 30));
        d.cap_G = h->maxcols + 4 * N;
        d.maxcols = h->maxcols;
        d.max_firings = h->max_firings;
        const size_t ul = static_cast<size_t>(d.cap_ulist);
        CC_CHECK(h, dev_alloc(h, L, &d.ulist, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.ulist_new, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_rep, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_maxfinish, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_mincol, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_maxend, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_np, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_finishcol, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.u_cluster, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.sv_cparent, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.sv_tfinish, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.sv_tmaxcol, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.sv_tnpoints, ul));
        CC_CHECK(h, dev_alloc(h, L, &d.edge_a, static_cast<size_t>(d.cap_edges)));
        CC_CHECK(h, dev_alloc(h, L, &d.edge_b, static_cast<size_t>(d.cap_edges)));
        CC_CHECK(h, dev_alloc(h, L, &d.G, static_cast<size_t>(d.cap_G)));
        CC_CHECK(h, dev_alloc(h, L, &d.n_new_ulist, 1));
        free_host_slots(h);
        h->prefetch_points = static_cast<int>(std::max<size_t>(16384, stage / 4));
        for (cc_handle::Slot& sl : h->slots)
        {
            CC_CHECK(h, dev_alloc(h, L, &sl.d_state_snap, 1));
            CC_CHECK(h, dev_alloc(h, L, &sl.d_first_unpub, mc));
            CC_CHECK(h, dev_alloc(h, L, &sl.d_clusters, static_cast<size_t>(d.cap_clusters)));
            CC_CHECK(h, dev_alloc(h, L, &sl.d_points, static_cast<size_t>(d.cap_cluster_points)));
            CC_CHECK(h, dev_alloc(h, L, &sl.d_labels, mc * h->R));
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&sl.h_state), sizeof(CcDevState)));
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&sl.h_first_unpub), mc * sizeof(long long)));
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&sl.h_clusters), CC_PREFETCH_CLUSTERS * sizeof(CcCluster)));
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&sl.h_hdr), sizeof(CcHostHeader)));
            std::memset(sl.h_hdr, 0, sizeof(CcHostHeader));
        }
        for (uchar4*& q : h->h_label_ring)
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&q), mc * h->R * sizeof(uchar4)));
        for (CcClusterPoint*& q : h->h_points_ring)
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&q), static_cast<size_t>(h->prefetch_points) * sizeof(CcClusterPoint)));
        for (cc_handle::InBuf& ib : h->inbuf)
        {
            CC_CHECK(h, dev_alloc(h, L, &ib.d_raw, stage * sizeof(cc_raw_point_t)));
            CC_CHECK(h, dev_alloc(h, L, &ib.d_poses, static_cast<size_t>(h->max_firings) * 12));
            CC_CHECK(h, cudaMallocHost(&ib.h_raw, stage * sizeof(cc_raw_point_t)));
            CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&ib.h_poses), static_cast<size_t>(h->max_firings) * 12 * sizeof(double)));
        }
        // The first DMA access to freshly page-locked pages runs at a third of the link rate on some hosts (measured:
        // ~20 GB/s against ~54 GB/s from the second access on, scripts/h2d_probe4.py): touch every page-locked buffer of
        // the handle once in both directions now, so that the first pushes do not pay for it on their critical path.
        {
            void* scratch = h->inbuf[0].d_raw;
            const size_t scratch_bytes = stage * sizeof(cc_raw_point_t);
            auto warm = [&](void* hp, size_t bytes) -> cudaError_t
            {
                bytes = std::min(bytes, scratch_bytes);
                cudaError_t e = cudaMemcpyAsync(hp, scratch, bytes, cudaMemcpyDeviceToHost, h->stream);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(scratch, hp, bytes, cudaMemcpyHostToDevice, h->stream);
                return e;
            };
            for (cc_handle::Slot& sl : h->slots)
            {
                CC_CHECK(h, warm(sl.h_first_unpub, mc * sizeof(long long)));
                CC_CHECK(h, warm(sl.h_clusters, CC_PREFETCH_CLUSTERS * sizeof(CcCluster)));
            }
            for (uchar4* q : h->h_label_ring)
                CC_CHECK(h, warm(q, mc * h->R * sizeof(uchar4)));
            for (CcClusterPoint* q : h->h_points_ring)
                CC_CHECK(h, warm(q, static_cast<size_t>(h->prefetch_points) * sizeof(CcClusterPoint)));
            for (cc_handle::InBuf& ib : h->inbuf)
            {
                CC_CHECK(h, warm(ib.h_raw, stage * sizeof(cc_raw_point_t)));
                CC_CHECK(h, warm(ib.h_poses, static_cast<size_t>(h->max_firings) * 12 * sizeof(double)));
            }
        }
        d.raw = h->inbuf[0].d_raw;
        d.poses = h->inbuf[0].d_poses;
        d.col_first_unpub = h->slots[0].d_first_unpub;
        d.clusters = h->slots[0].d_clusters;
        d.cluster_points = h->slots[0].d_points;
        CC_CHECK(h, cudaMemsetAsync(d.firing_index, 0, cells * sizeof(unsigned long long), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.cparent, 0, cells * sizeof(unsigned int), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.tfinish, 0, cells * sizeof(unsigned long long), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.tmaxcol, 0, cells * sizeof(long long), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.tnpoints, 0, cells * sizeof(unsigned int), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.tid, 0, cells * sizeof(unsigned int), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.tslot, 0xff, cells * sizeof(int), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.rootslot, 0, cells * sizeof(unsigned int), h->stream));
        CC_CHECK(h, cudaMemsetAsync(d.incl, 0, (cells + 1) * sizeof(float), h->stream));
    }
    // sc_inclination_angles_between_lasers_.resize(num_rows, NaN) only fills new elements (cpp:46): the values
    // survive a reset with an unchanged row count
    if (h->gap_rows != num_rows)
    {
        free_list(h->allocs_fixed);
        CC_CHECK(h, dev_alloc(h, h->allocs_fixed, &h->d.gap_state, static_cast<size_t>(num_rows)));
        std::vector<float> nan(num_rows, std::nanf(""));
        CC_CHECK(h, cudaMemcpy(h->d.gap_state, nan.data(), nan.size() * sizeof(float), cudaMemcpyHostToDevice));
        h->gap_rows = num_rows;
    }
    h->width = static_cast<float>(2 * M_PI) / static_cast<float>(N); // cpp:16
    CcDevCfg cfg;
    fill_devcfg(h, cfg);
    // clearColumns(0, ring_buffer_max_columns - 1) cpp:29
    CC_RUN(h, k_clear, grid_for(h, static_cast<long long>(h->ringcols) * h->R, 256), 256, 0, cfg, h->d, 0LL,
           static_cast<long long>(h->ringcols), 0);
    CcDevState s;
    std::memset(&s, 0, sizeof(s));
    s.P = 0;
    s.foremost = -1;
    s.F = -1;
    s.ring_start = -1;
    s.ring_end = -1;
    s.first_unpub = -1;
    s.cluster_counter = 1;
    s.runmax_carry = -1.0;
    s.colbase = -1;
    s.clear_from = -1;
    s.clear_to = -1;
    s.clear2_from = -1;
    s.clear2_to = -1;
    s.danger_col = CC_COL_INF;
    s.forced_col = -1;
    *h->h_state = s;
    CC_CHECK(h, cudaMemcpyAsync(h->d.st, h->h_state, sizeof(s), cudaMemcpyHostToDevice, h->stream));
    std::vector<long long> rm(h->R, -1);
    CC_CHECK(h, cudaMemcpyAsync(h->d.rowmax, rm.data(), rm.size() * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    CC_CHECK(h, cudaStreamSynchronize(h->stream));
    CC_CHECK(h, cudaGetLastError());
    h->state = s;
    h->has_tf = false; // cpp:38
    h->reset_required_cfg = false;
    h->is_reset = true;
    h->n_events = 0;
    h->clusters.clear();
    h->cluster_points.clear();
    h->points_view = nullptr;
    std::memset(&h->info, 0, sizeof(h->info));
    h->probe_smem = (32 + static_cast<size_t>(h->maxcols)) * sizeof(double);
#ifndef CC_EMU
    {
        if (h->probe_smem > 48 * 1024)
            CC_CHECK(h, cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->probe_smem)));
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_probe, 256, h->probe_smem) == cudaSuccess && nb > 0)
            h->occ_probe = nb;
    }
    {
        const int smem = scan_smem_bytes(h->R);
        if (smem > 48 * 1024)
            CC_CHECK(h, cudaFuncSetAttribute(k_insert_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
#endif
    configure_fused(h);
    configure_fin_cluster(h);
    return CC_OK;
}

// cudaEventSynchronize wakes up late (hundreds of microseconds) while other work keeps the device busy; a push is
// only a few hundred microseconds long, so poll instead
static bool g_block_on_events = std::getenv("CC_B200_BLOCKING_WAIT") != nullptr; // several ranks per socket: do not spin
static cudaError_t spin_wait(cudaEvent_t e)
{
#ifdef CC_EMU
    (void)e;
    return cudaSuccess;
#else
    if (g_block_on_events)
        return cudaEventSynchronize(e);
    cudaError_t r;
    while ((r = cudaEventQuery(e)) == cudaErrorNotReady)
    {
    }
    return r;
#endif
}

// a fused push signals completion through a flag in page-locked host memory (written after everything else it reports)
static cudaError_t spin_flag(cc_handle* h, const volatile unsigned int* flag, unsigned int ticket)
{
#ifdef CC_EMU
    (void)h;
    return *flag == ticket ? cudaSuccess : 1;
#else
    unsigned int spins = 0;
    while (*flag != ticket)
    {
        if ((++spins & 0x3fffu) == 0) // a failed launch / a fault never sets the flag
        {
            const cudaError_t q = cudaStreamQuery(h->stream);
            if (q != cudaErrorNotReady && *flag != ticket)
                return q == cudaSuccess ? cudaErrorUnknown : q;
        }
    }
    return cudaSuccess;
#endif
}

static cc_status_t fetch_state(cc_handle* h)
{
    CC_CHECK(h, cudaMemcpyAsync(h->h_state, h->d.st, sizeof(CcDevState), cudaMemcpyDeviceToHost, h->stream));
    CC_CHECK(h, cudaStreamSynchronize(h->stream));
    h->state = *h->h_state;
    return CC_OK;
}

static cc_status_t device_error_to_status(cc_handle* h)
{
    const CcDevState& s = h->state;
    switch (s.error)
    {
        case CC_DEV_OK:
            return CC_OK;
        case CC_DEV_COLUMN_NOT_CLEARED:
            h->error = "This column is not cleared. Probably this means the ring buffer is full or there "
                       "is some other issue with clearing (not cleared at all or written after clearing): " +
                       std::to_string(s.err_a) + ", " + std::to_string(s.err_b) + ", " + std::to_string(h->ringcols);
            return CC_ERR_COLUMN_NOT_CLEARED;
        case CC_DEV_TOO_MANY_COLUMNS:
            h->error = "more columns completed in one push than the handle was sized for";
            return CC_ERR_BATCH_TOO_LARGE;
        case CC_DEV_RING_START_DECREASED:
            h->error = "This shouldn't happen, ring buffer is not allowed to increase at the front: " +
                       std::to_string(s.err_a) + ", " + std::to_string(s.err_b);
            return CC_ERR_RING_START_DECREASED;
        default:
            h->error = "device work list overflow (unfinished trees / links / clusters)";
            return CC_ERR_INTERNAL;
    }
}

// finish passes for columns [ci0, ci1] (ci1 < 0: all new columns); `last` also closes the push (k_push_done)
static void launch_finish(cc_handle* h, const CcDevCfg& cfg, int ci0, int ci1, int guard, int exact, int last,
                          CcDevState* snap = nullptr, int careful_ci = -1)
{
    const unsigned int seq = ++h->seq;
#ifdef CC_EMU
    const int fin_threads = 1;
#else
    const int fin_threads = 1024;
#endif
    // scan scratch + room for the segment's running maxima / the prefix maxima of G (whichever phase is running)
    size_t fin_smem = fin_threads * sizeof(long long) +
                      std::max(static_cast<size_t>(h->maxcols) * sizeof(double), static_cast<size_t>(h->d.cap_G) * sizeof(int));
    if (fin_smem > 200 * 1024)
        fin_smem = 200 * 1024;
#ifndef CC_EMU
    if (fin_smem > 48 * 1024 && fin_smem != h->fin_smem_set)
    {
        cudaFuncSetAttribute(k_fin_all, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fin_smem));
        h->fin_smem_set = fin_smem;
    }
#endif
    // whole-push speculative pass: the list phases over a thread-block cluster (k_fin_cluster); everything else (commits of
    // sub-ranges and exact single-column passes of the split path): one CTA
    const bool clustered = guard == 1 && !exact && ci0 == 0 && ci1 < 0 && h->fin_cluster > 0;
    if (clustered)
    {
#ifdef CC_EMU
        const int t0 = timing_begin(h, "k_fin_all");
        CC_LAUNCH(k_fin_cluster, 1, 1, h->fin_cluster_smem, h->stream, cfg, h->d, seq, last, static_cast<int>(h->fin_cluster_smem), snap);
        timing_end(h, t0);
#else
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(static_cast<unsigned>(h->fin_cluster), 1, 1);
        lc.blockDim = dim3(512, 1, 1);
        lc.dynamicSmemBytes = h->fin_cluster_smem;
        lc.stream = h->stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(h->fin_cluster);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr;
        lc.numAttrs = 2;
        const int t0 = timing_begin(h, "k_fin_all");
        cudaLaunchKernelEx(&lc, k_fin_cluster, cfg, h->d, seq, last, static_cast<int>(h->fin_cluster_smem), snap);
        timing_end(h, t0);
#endif
        h->launches++;
    }
    else
        CC_RUN(h, k_fin_all, 1, fin_threads, fin_smem, cfg, h->d, ci0, ci1, seq, guard, exact, last, static_cast<int>(fin_smem), snap,
               careful_ci);
    CC_RUN(h, k_fin_label, h->sm_count * 8, 256, 0, cfg, h->d, seq, guard, clustered ? 1 : 0);
    // number_of_visited_neighbors of the few points whose walk went beyond the first unpublished column (exact counts)
    // (a small grid: the list is empty or short; a warp redoes one point in a few microseconds)
    // (not after the exact pass of a single column: k_careful's walk honours the stop itself, nothing is listed)
    if (exact)
        return;
    static const int vfix_grid = std::getenv("CC_B200_VFIX_GRID") ? std::max(1, std::atoi(std::getenv("CC_B200_VFIX_GRID"))) : 128;
    CC_RUN(h, k_visited_fix, vfix_grid, 64, cc_heavy_smem_bytes(64, 2), cfg, h->d, guard, snap, h->tune);
}

static void launch_commit(cc_handle* h, const CcDevCfg& cfg, int ci0, int ci1, int guard, bool snapshot)
{
    const int g = h->sm_count * 8;
    if (snapshot)
        CC_RUN(h, k_snapshot, h->sm_count * 2, 256, 0, h->d, guard);
    CC_RUN(h, k_commit_copy, g, 256, 0, cfg, h->d, h->d_s_parent, ci0, ci1, guard);
    CC_RUN(h, k_commit_roots, g, 256, 0, cfg, h->d, ci0, ci1, guard);
    CC_RUN(h, k_commit_links, g, 256, 0, cfg, h->d, h->d_s_parent, h->d_s_links, ci0, ci1, guard);
}

// Column-sequential exact path for pushes whose probe flagged a possibly refused association, that hit a cluster
// about to span a rotation, or that run finish passes only every n-th column.
static cc_status_t slow_path(cc_handle* h, const CcDevCfg& cfg)
{
    const int ncols = h->state.ncols;
    const long long colbase = h->state.colbase;
    const long long danger = h->state.danger_col;
    const bool all_careful = false; // (passes every n-th column are handled by the speculative commit too: cc_pass_at_or_after)
    h->h_flags.assign(ncols, 0);
    CC_CHECK(h, cudaMemcpyAsync(h->h_flags.data(), h->d.col_flag, ncols, cudaMemcpyDeviceToHost, h->stream));
    CC_CHECK(h, cudaStreamSynchronize(h->stream));
    // A forced finish (cpp:909-919) at column `danger` invalidates the probe's results only for the columns whose walk can
    // reach a cell of the force-finished component, i.e. the next max_steps_in_row columns (cpp:704-705): their hits on it
    // are refused by the reference and they found new trees. Those columns go through the exact path; the columns behind
    // them only meet cells associated after the forced finish and are committed speculatively again (and checked again:
    // the range's own finish pass aborts at the next dangerous column, which is handled the same way below).
    const int reach = std::max(0, cfg.max_steps_row);
    std::vector<unsigned char> careful(ncols, 0);
    for (int ci = 0; ci < ncols; ci++)
        careful[ci] = all_careful || h->h_flags[ci] || (colbase + ci >= danger && colbase + ci <= danger + reach);
    int ci = 0;
    while (ci < ncols)
    {
        if (careful[ci])
        {
            if ((colbase + ci) % cfg.nth == 0)
                launch_finish(h, cfg, ci, ci, 0, 1, 0, nullptr, ci); // k_careful's work at the head of the finish pass
            else
                CC_RUN(h, k_careful, 1, 1, 0, cfg, h->d, ci);
            ci++;
            if (!all_careful && ci < ncols && !careful[ci])
            {
                // end of a run of exact columns: a forced finish inside it (any component, not only the one that was
                // announced) reaches max_steps_in_row columns further
                cc_status_t s = fetch_state(h);
                if (s != CC_OK)
                    return s;
                const long long f = h->state.forced_col;
                if (f >= 0)
                    for (long long c = std::max<long long>(f + 1, colbase + ci); c <= f + reach && c < colbase + ncols; c++)
                        careful[static_cast<size_t>(c - colbase)] = 1;
            }
            continue;
        }
        int cj = ci;
        while (cj + 1 < ncols && !careful[cj + 1])
            cj++;
        launch_commit(h, cfg, ci, cj, 0, true);
        launch_finish(h, cfg, ci, cj, 2, 0, 0);
        cc_status_t s = fetch_state(h);
        if (s != CC_OK)
            return s;
        if (h->state.abort)
        {
            const long long d2 = h->state.danger_col;
            CC_RUN(h, k_restore, h->sm_count * 2, 256, 0, h->d);
            CC_RUN(h, k_restore_finish, 1, 1, 0, h->d);
            if (d2 >= colbase + ci && d2 <= colbase + cj)
            {
                // the range runs into a (further) dangerous column: exact path from there for the columns its forced finish
                // can affect, the part before it and the part behind are tried again as ranges
                for (long long c = d2; c <= d2 + reach && c <= colbase + cj; c++)
                    careful[static_cast<size_t>(c - colbase)] = 1;
            }
            else
                for (int c = ci; c <= cj; c++)
                    careful[c] = 1;
            continue;
        }
        ci = cj + 1;
    }
    h->used_exact_flag = 1;
    return CC_OK;
}

// state snapshot + optimistic prefix of the results of the push in `sl`, brought to the host on the copy stream so
// that the next push's kernels do not wait for the transfer (the device-side result arrays are per slot)
static cc_status_t enqueue_results(cc_handle* h, cc_handle::Slot& sl, bool state_snapshot_done = false)
{
    if (!state_snapshot_done) // the push's last finish pass already wrote it on the normal path
    {
        CC_LAUNCH(k_state_snapshot, 1, 128, 0, h->stream, h->d, sl.d_state_snap);
        h->launches++;
    }
    if (h->label_prefetch && sl.has_tf)
        CC_RUN(h, k_pack_labels, h->sm_count * 4, 256, 0, sl.cfg, h->d, sl.d_labels, h->maxcols);
    CC_CHECK(h, cudaEventRecord(sl.ready, h->stream));
    CC_CHECK(h, cudaStreamWaitEvent(h->copy_stream, sl.ready, 0));
    sl.pre_cols = std::min(h->maxcols, sl.n + 64);
    sl.pre_clusters = std::min(h->d.cap_clusters, CC_PREFETCH_CLUSTERS);
    // member lists: what a push of this size typically finishes, not the whole capacity
    sl.pre_points = std::min(std::min(h->d.cap_cluster_points, h->prefetch_points), std::max(16384, sl.n * h->R / 4));
    static const bool results_by_copy_engine = std::getenv("CC_B200_RESULT_COPIES") != nullptr; // A/B aid: the five transfers
    if (!results_by_copy_engine)
    {
        CcResultDst dst;
        dst.h_state = sl.h_state;
        dst.h_first_unpub = sl.h_first_unpub;
        dst.h_clusters = reinterpret_cast<CcCluster*>(sl.h_clusters);
        dst.h_points = reinterpret_cast<CcClusterPoint*>(sl.h_points);
        dst.h_labels = h->label_prefetch && sl.has_tf ? sl.h_labels : nullptr;
        dst.cap_cols = sl.pre_cols;
        dst.cap_clusters = sl.pre_clusters;
        dst.cap_points = sl.pre_points;
        dst.rows = h->R;
        static const int export_grid = std::getenv("CC_B200_EXPORT_GRID") ? std::max(1, std::atoi(std::getenv("CC_B200_EXPORT_GRID"))) : 16;
        CC_LAUNCH(k_results_to_host, export_grid, 512, 0, h->copy_stream, sl.d_state_snap, sl.d_first_unpub, sl.d_clusters, sl.d_points,
                  sl.d_labels, dst);
        h->launches++;
        CC_CHECK(h, cudaEventRecord(sl.done, h->copy_stream));
        return CC_OK;
    }
    CC_CHECK(h, cudaMemcpyAsync(sl.h_state, sl.d_state_snap, sizeof(CcDevState), cudaMemcpyDeviceToHost, h->copy_stream));
    CC_CHECK(h, cudaMemcpyAsync(sl.h_first_unpub, sl.d_first_unpub, sl.pre_cols * sizeof(long long),
                                cudaMemcpyDeviceToHost, h->copy_stream));
    CC_CHECK(h, cudaMemcpyAsync(sl.h_clusters, sl.d_clusters, sl.pre_clusters * sizeof(CcCluster),
                                cudaMemcpyDeviceToHost, h->copy_stream));
    CC_CHECK(h, cudaMemcpyAsync(sl.h_points, sl.d_points, sl.pre_points * sizeof(CcClusterPoint),
                                cudaMemcpyDeviceToHost, h->copy_stream));
    if (h->label_prefetch && sl.has_tf)
        CC_CHECK(h, cudaMemcpyAsync(sl.h_labels, sl.d_labels, static_cast<size_t>(sl.pre_cols) * h->R * sizeof(uchar4),
                                    cudaMemcpyDeviceToHost, h->copy_stream));
    CC_CHECK(h, cudaEventRecord(sl.done, h->copy_stream));
    return CC_OK;
}

static void bind_slot(cc_handle* h, const cc_handle::Slot& sl)
{
    h->d.trace = h->trace_on ? h->d_trace : nullptr;
    h->d.raw = sl.in_points;
    h->d.poses = sl.in_poses;
    h->d.col_first_unpub = sl.d_first_unpub;
    h->d.clusters = sl.d_clusters;
    h->d.cluster_points = sl.d_points;
}

// Enqueues every kernel of one push on the handle's stream (nothing here waits for the device).
static cc_status_t launch_push(cc_handle* h, cc_handle::Slot& sl)
{
    const CcDevCfg& cfg = sl.cfg;
    bind_slot(h, sl);
    const int n = sl.n;
    sl.launches0 = h->launches;
    h->n_timed = 0;
    sl.has_tf = sl.cfg_has_tf;
    sl.spec = true; // whole-push speculative commit (finish passes at every n-th column included)
    sl.fused = sl.want_fused; // decided when the push was submitted (its inputs were routed accordingly)
    if (sl.fused)
    {
        CcFusedArgs a;
        std::memset(&a, 0, sizeof(a));
        a.n = n;
        a.has_tf = sl.has_tf ? 1 : 0;
        a.spec = sl.spec ? 1 : 0;
        a.scan_chunk = fused_scan_chunk(h->R);
        a.tune = h->tune;
        a.team_warps = 2;
        a.pack_labels = h->label_prefetch && sl.has_tf ? 1 : 0;
        a.seq = ++h->seq;
        if (++h->ticket == 0)
            ++h->ticket;
        a.ticket = sl.ticket = h->ticket;
        a.src_raw = sl.src_points;
        a.src_poses = sl.src_poses;
        const bool staged = sl.src_points != sl.in_points; // host inputs: fetched into the device input buffer
        a.dst_raw = staged ? const_cast<void*>(sl.in_points) : nullptr;
        a.dst_poses = staged ? const_cast<double*>(sl.in_poses) : nullptr;
        a.s_parent = h->d_s_parent;
        a.s_links = h->d_s_links;
        a.snap = sl.d_state_snap;
        a.d_labels = sl.d_labels;
        a.h_hdr = sl.h_hdr;
        a.h_state = sl.h_state;
        a.h_first_unpub = sl.h_first_unpub;
        a.h_clusters = sl.h_clusters;
        a.h_points = sl.h_points;
        a.h_labels = sl.h_labels;
        sl.pre_cols = a.cap_cols = std::min(h->maxcols, sl.n + 64);
        sl.pre_clusters = a.cap_clusters = std::min(h->d.cap_clusters, CC_PREFETCH_CLUSTERS);
        sl.pre_points = a.cap_points = std::min(std::min(h->d.cap_cluster_points, h->prefetch_points), std::max(16384, sl.n * h->R / 4));
        a.smem_bytes = static_cast<int>(h->fused_smem);
#ifdef CC_EMU
        CC_LAUNCH(k_push_fused, 1, 1, h->fused_smem, h->stream, cfg, h->d, a);
#else
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(static_cast<unsigned>(h->fused_cluster), 1, 1);
        lc.blockDim = dim3(static_cast<unsigned>(h->fused_threads), 1, 1);
        lc.dynamicSmemBytes = h->fused_smem;
        lc.stream = h->stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(h->fused_cluster);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr;
        lc.numAttrs = 2;
        CC_CHECK(h, cudaLaunchKernelEx(&lc, k_push_fused, cfg, h->d, a));
#endif
        h->launches++;
        sl.launches1 = h->launches;
        return CC_OK;
    }

    CC_CHECK(h, cudaEventRecord(sl.ev0, h->stream));
    const int R = h->R;
    const long long pts = static_cast<long long>(n) * R;
    // recycles the columns retired two pushes ago, then prepares the firings
    CC_RUN(h, k_prep, std::max(grid_for(h, static_cast<long long>(n) * CC_WARP, 256), h->sm_count * 2), 256, 0, cfg, h->d, n);
    // lite insertion path (regular prefix of the push, grid-wide) ...
    {
#ifndef CC_EMU
        const int lite_threads = 1024;
#else
        const int lite_threads = 1;
#endif
        const size_t lite_smem = cc_lite_smem_bytes(n);
#ifndef CC_EMU
        if (lite_smem > 48 * 1024 && lite_smem > h->lite_smem_set)
        {
            CC_CHECK(h, cudaFuncSetAttribute(k_scan_lite, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lite_smem)));
            h->lite_smem_set = lite_smem;
        }
#endif
        CC_RUN(h, k_scan_lite, 1, lite_threads, lite_smem, cfg, h->d, n);
    }
    CC_RUN(h, k_scan_check, R, n > 2048 ? 512 : 256, 512 * sizeof(int), cfg, h->d, n);
    // ... then the single-CTA scan commits that prefix and resolves whatever is left
    const int scan_smem = scan_smem_bytes(R);
    CC_RUN(h, k_insert_scan, 1, scan_threads(), scan_smem, cfg, h->d, n, scan_chunk(R), 1);
    CC_RUN(h, k_scatter, grid_for(h, pts, 256), 256, 0, cfg, h->d, n);

    if (!sl.has_tf)
    {
        // the reference throws from the segmentation stage of the first completed column (cpp:298-299): nothing
        // after insertion runs; pushes queued behind this one must not run either
        CC_RUN(h, k_halt, 1, 1, 0, h->d, 1);
    }
    else
    {
        CC_RUN(h, k_gap_scan, h->sm_count, 256, 0, cfg, h->d); // its last block chains the column chunks
        const int gw = 4; // warps (columns) per block
        const size_t ground_smem = gw * cc_ground_warp_bytes(R);
#ifndef CC_EMU
        if (ground_smem > 48 * 1024 && ground_smem != h->ground_smem_set)
        {
            CC_CHECK(h, cudaFuncSetAttribute(k_ground, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ground_smem)));
            h->ground_smem_set = ground_smem;
        }
#endif
        // one resident wave of warp-per-column blocks covers 8 * 4 * SMs columns
        CC_RUN(h, k_ground, h->sm_count * 8, gw * CC_WARP, ground_smem, cfg, h->d, h->d_s_parent);
        // one resident wave each (the blocks loop over the work lists): a second wave would only repeat the prologue
        // (every CTA first computes the running maximum of the column minima in shared memory)
        // The list-driven thread-per-point probe (global loads, 8 lanes per warp) is the default: measured against the tiled
        // probe with its field of view staged in shared memory by bulk asynchronous copies (CC_B200_TUNE bit 2), it is the
        // faster one at 4096 firings per push (13.9 + 13.5 us with k_probe_heavy against 18.4 + 12.5 us, profiles/r02_probe_ab.md):
        // a tile's time is its latency chain -- column-minimum reduction, copy landing, walk -- over two resident waves.
        if (!(h->tune & 4))
            CC_RUN(h, k_probe, h->sm_count * h->occ_probe, 256, h->probe_smem, cfg, h->d, h->d_s_parent, h->d_s_links, sl.spec ? 1 : 0);
        else
        {
            // tiles of 256 / R columns, their field of view staged in shared memory by a bulk asynchronous copy; + 1 CTA
            // that publishes the running maxima
            const size_t tile_smem = cc_tile_smem_bytes(R, cfg.max_steps_row);
#ifndef CC_EMU
            if (tile_smem > 48 * 1024 && tile_smem > h->tile_smem_set)
            {
                CC_CHECK(h, cudaFuncSetAttribute(k_probe_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tile_smem)));
                h->tile_smem_set = tile_smem;
            }
#endif
            const int tiles = (std::min(h->maxcols, n + 64) + cc_tile_cols(R) - 1) / cc_tile_cols(R);
            const int workers = std::max(1, std::min(tiles, h->sm_count * 8));
            CC_RUN(h, k_probe_tile, workers + 1, CC_TILE_CELLS, tile_smem, cfg, h->d, h->d_s_parent, h->d_s_links, sl.spec ? 1 : 0);
        }
        CC_RUN(h, k_probe_heavy, h->sm_count * h->occ_probe_heavy, 64, cc_heavy_smem_bytes(64, 2), cfg, h->d,
               h->d_s_parent, h->d_s_links, h->tune);
        if (sl.spec)
        {
            launch_commit(h, cfg, 0, -1, 1, false);
            launch_finish(h, cfg, 0, -1, 1, 0, 1, sl.d_state_snap);
        }
        else
            CC_RUN(h, k_halt, 1, 1, 0, h->d, 0); // finish passes every n-th column: column-sequential path, on the host's cue
    }
    CC_CHECK(h, cudaEventRecord(sl.ev1, h->stream));
    sl.launches1 = h->launches;
    return enqueue_results(h, sl, sl.has_tf && sl.spec);
}

// Waits for the oldest push in flight, finishes it (column-sequential path if the speculative commit could not be
// used) and builds its results.
static cc_status_t finish_push(cc_handle* h)
{
    if (h->n_pending == 0)
    {
        h->error = "no push in flight";
        return CC_ERR_INVALID_ARGUMENT;
    }
    cc_handle::Slot& sl = h->slots[h->pending[0]];
    h->n_events = 0;
    h->clusters.clear();
    h->cluster_points.clear();
    h->points_view = nullptr;
    std::memset(&h->info, 0, sizeof(h->info));
    auto pop = [&]()
    {
        h->pending[0] = h->pending[1];
        h->n_pending--;
    };
    const bool timed_by_device = sl.fused;
    if (sl.fused)
        CC_CHECK(h, spin_flag(h, &sl.h_hdr->flag, sl.ticket));
    else
        CC_CHECK(h, spin_wait(sl.done));
    h->state = *sl.h_state;
    const CcDevCfg cfg = sl.cfg;
    cc_status_t es = device_error_to_status(h);
    if (es == CC_OK && !sl.has_tf && h->state.ncols > 0)
    {
        h->error = "Transform robot frame from sensor frame was not set yet!";
        es = CC_ERR_NO_ROBOT_TRANSFORM;
    }
    if (es != CC_OK)
    {
        // later pushes saw the halt flag and did nothing; they are dropped. The stream needs a reset, like the
        // reference after an exception escaped addFiring.
        h->n_pending = 0;
        return es;
    }
    if (sl.has_tf && h->state.ncols > 0 && (!sl.spec || h->state.n_flagged > 0 || h->state.abort))
    {
        // the speculative commit was not usable: pushes queued behind this one skipped themselves (halt flag).
        // Finish this push column-sequentially, then run them again.
        for (int i = 1; i < h->n_pending; i++)
        {
            cc_handle::Slot& o = h->slots[h->pending[i]];
            if (o.fused)
                CC_CHECK(h, spin_flag(h, &o.h_hdr->flag, o.ticket));
            else
                CC_CHECK(h, cudaEventSynchronize(o.done));
        }
        CC_CHECK(h, cudaStreamSynchronize(h->stream));
        sl.fused = false; // its results are collected through the copy stream below
        bind_slot(h, sl);
        CC_RUN(h, k_halt, 1, 1, 0, h->d, -1); // clear
        if (h->state.abort)
        {
            CC_RUN(h, k_restore, h->sm_count * 2, 256, 0, h->d);
            CC_RUN(h, k_restore_finish, 1, 1, 0, h->d);
        }
        cc_status_t s = slow_path(h, cfg);
        if (s != CC_OK)
        {
            h->n_pending = 0;
            return s;
        }
        CC_RUN(h, k_push_done, 1, 1, 0, h->d, 0);
        CC_CHECK(h, cudaEventRecord(sl.ev1, h->stream));
        sl.launches1 = h->launches;
        s = enqueue_results(h, sl);
        if (s != CC_OK)
            return s;
        for (int i = 1; i < h->n_pending; i++)
        {
            s = launch_push(h, h->slots[h->pending[i]]);
            if (s != CC_OK)
                return s;
        }
        CC_CHECK(h, spin_wait(sl.done));
        h->state = *sl.h_state;
        es = device_error_to_status(h);
        if (es != CC_OK)
        {
            h->n_pending = 0;
            return es;
        }
    }
    CC_CHECK(h, cudaGetLastError());

    // ---- results of the push ----
    const CcDevState& st = h->state;
    const int ncols = sl.has_tf ? st.ncols : 0;
    const int ncl = st.n_clusters, ncp = st.n_cluster_points;
    h->h_first_unpub.resize(ncols);
    h->h_clusters.resize(ncl);
    const bool points_in_place = ncp <= sl.pre_points; // the prefetch brought all member lists: hand out the pinned buffer
    if (!points_in_place)
        h->cluster_points.resize(ncp);
    static_assert(sizeof(CcClusterPoint) == sizeof(cc_cluster_point_t), "cluster point layout");
    {
        // what the prefetch already brought, then (rarely) the remainder from the slot's device arrays
        const int c0 = std::min(ncols, sl.pre_cols), l0 = std::min(ncl, sl.pre_clusters), p0 = std::min(ncp, sl.pre_points);
        if (c0)
            std::memcpy(h->h_first_unpub.data(), sl.h_first_unpub, c0 * sizeof(long long));
        if (l0)
            std::memcpy(h->h_clusters.data(), sl.h_clusters, l0 * sizeof(CcCluster));
        if (p0 && !points_in_place)
            std::memcpy(h->cluster_points.data(), sl.h_points, p0 * sizeof(CcClusterPoint));
        bool more = false;
        if (ncols > c0)
        {
            CC_CHECK(h, cudaMemcpyAsync(h->h_first_unpub.data() + c0, sl.d_first_unpub + c0,
                                        (ncols - c0) * sizeof(long long), cudaMemcpyDeviceToHost, h->aux_stream));
            more = true;
        }
        if (ncl > l0)
        {
            CC_CHECK(h, cudaMemcpyAsync(h->h_clusters.data() + l0, sl.d_clusters + l0, (ncl - l0) * sizeof(CcCluster),
                                        cudaMemcpyDeviceToHost, h->aux_stream));
            more = true;
        }
        if (ncp > p0)
        {
            CC_CHECK(h, cudaMemcpyAsync(h->cluster_points.data() + p0, sl.d_points + p0,
                                        (ncp - p0) * sizeof(CcClusterPoint), cudaMemcpyDeviceToHost, h->aux_stream));
            more = true;
        }
        if (more)
            CC_CHECK(h, cudaStreamSynchronize(h->aux_stream));
    }
    h->cur_labels = nullptr;
    h->cur_label_cols = 0;
    if (h->label_prefetch && sl.has_tf)
    {
        if (ncols > sl.pre_cols) // more columns than the prefetch covered: fetch the packed labels again, all of them
        {
            CC_CHECK(h, cudaMemcpyAsync(sl.h_labels, sl.d_labels, static_cast<size_t>(std::min(ncols, h->maxcols)) * h->R * sizeof(uchar4),
                                        cudaMemcpyDeviceToHost, h->aux_stream));
            CC_CHECK(h, cudaStreamSynchronize(h->aux_stream));
        }
        h->cur_labels = sl.h_labels;
        h->cur_label_cols = std::min(ncols, h->maxcols);
    }
    h->points_view = points_in_place ? reinterpret_cast<const cc_cluster_point_t*>(sl.h_points) : h->cluster_points.data();
    h->last_points_dev = sl.d_points;
    float ms = 0.f;
    if (timed_by_device)
        ms = static_cast<float>(static_cast<double>(sl.h_hdr->t_end_ns - sl.h_hdr->t_start_ns) * 1e-6);
    else
        cudaEventElapsedTime(&ms, sl.ev0, sl.ev1);

    // clusters in the order the reference would deliver them: by the column whose pass finished them
    std::stable_sort(h->h_clusters.begin(), h->h_clusters.end(),
                     [](const CcCluster& a, const CcCluster& b) { return a.finish_col < b.finish_col; });
    h->clusters.reserve(ncl);
    for (const CcCluster& c : h->h_clusters)
    {
        cc_cluster_t o;
        o.id = c.id;
        o.min_stamp = c.min_stamp;
        o.max_stamp = c.max_stamp;
        o.stamp = cfg.use_last_stamp ? c.max_stamp : c.min_stamp + (c.max_stamp - c.min_stamp) / 2; // cpp:1025-1028
        o.finished_at_gcol = c.finish_col;
        o.min_gcol = c.min_col;
        o.max_gcol = c.max_col;
        o.num_points = c.num_points;
        o.point_offset = c.point_offset;
        h->clusters.push_back(o);
    }
    // finished-column callbacks in the reference's single-threaded order (cpp:618-620, 1087-1089)
    long long fu_old = st.push_first_unpub_old;
    size_t next_cluster = 0;
    if (h->events.size() < static_cast<size_t>(ncols) * 2)
        h->events.resize(static_cast<size_t>(ncols) * 2);
    cc_column_event_t* ev_out = h->events.data();
    size_t n_ev = 0;
    for (int ci = 0; ci < ncols; ci++)
    {
        const long long c = st.colbase + ci;
        cc_column_event_t e;
        e.from_gcol = c;
        e.to_gcol = c;
        e.ground_points_only = 1;
        e.n_clusters_before = static_cast<int32_t>(next_cluster);
        ev_out[n_ev++] = e;
        if (c % cfg.nth != 0)
            continue;
        while (next_cluster < h->clusters.size() && h->clusters[next_cluster].finished_at_gcol <= c)
            next_cluster++;
        const long long fu = h->h_first_unpub[ci];
        e.from_gcol = fu_old;
        e.to_gcol = fu - 1;
        e.ground_points_only = 0;
        e.n_clusters_before = static_cast<int32_t>(next_cluster);
        ev_out[n_ev++] = e;
        fu_old = fu;
    }
    cc_batch_info_t& info = h->info;
    info.ground_from_gcol = ncols ? st.colbase : 0;
    info.ground_to_gcol = ncols ? st.colbase + ncols : 0;
    info.first_unpublished_gcol = st.first_unpub;
    info.ring_start_gcol = st.ring_start;
    info.ring_end_gcol = st.ring_end;
    info.cleared_from_gcol = st.clear_from;
    info.cleared_to_gcol = st.clear_to;
    h->n_events = n_ev;
    info.n_events = static_cast<int32_t>(n_ev);
    info.n_clusters = ncl;
    info.n_cluster_points = ncp;
    info.reset_required = st.reset_required;
    info.used_exact_path = h->used_exact_flag;
    h->used_exact_flag = 0;
    info.gpu_launches = static_cast<int32_t>(sl.launches1 - sl.launches0);
    info.device_ms = ms;
    info.slow_insert_firings = st.scan_slow_firings + st.scan_fast_firings; // everything the lite path did not take
    info.n_unfinished_trees = st.n_ulist;
    info.fused_launch = timed_by_device ? 1 : 0;
    info.visited_recounts = st.n_vfix;
    pop();
    return CC_OK;
}

int cc_max_firings_per_push(const cc_handle_t* h)
{
    if (!h)
        return 0;
    return h->N > 0 ? std::min(h->max_firings, 3 * h->N) : h->max_firings;
}

static cc_status_t check_push(cc_handle* h, int n, int rows)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    if (!h->is_reset)
    {
        h->error = "cc_reset() has not been called";
        return CC_ERR_NOT_RESET;
    }
    if (rows != h->R)
    {
        h->error = "The number of points in a firing has changed. This is probably a bug!"; // cpp:90-91
        return CC_ERR_ROW_COUNT_CHANGED;
    }
    if (n <= 0 || n > cc_max_firings_per_push(h))
    {
        // the ring keeps 10 rotations and recycles columns two pushes late: a push may span at most 3 rotations
        h->error = n <= 0 ? "empty push" : "too many firings in one push (limit: min(max_firings_per_push, 3 * num_columns))";
        return n <= 0 ? CC_ERR_INVALID_ARGUMENT : CC_ERR_BATCH_TOO_LARGE;
    }
    return CC_OK;
}

// Launches the kernels of the push whose inputs are in input buffer `ib` (or at the caller's device pointers).
static cc_status_t launch_from(cc_handle* h, int n, const void* d_points, const double* d_poses, const cc_handle::InBuf* ib)
{
    cc_handle::Slot& sl = h->slots[h->next_slot];
    sl.n = n;
    if (ib)
    {
        sl.cfg = ib->cfg;
        sl.cfg_has_tf = ib->cfg_has_tf;
    }
    else
    {
        fill_devcfg(h, sl.cfg);
        sl.cfg_has_tf = h->has_tf;
    }
    sl.in_points = d_points;
    sl.in_poses = d_poses;
    sl.h_labels = h->h_label_ring[h->next_label];
    sl.h_points = h->h_points_ring[h->next_label];
    h->next_label = (h->next_label + 1) % 3;
    sl.h2d = ib ? ib->h2d : nullptr;
    sl.h2d0 = ib ? ib->h2d0 : nullptr;
    // where a fused push fetches its firings from: the caller's device arrays, or page-locked host memory (no copy was
    // queued for such an input buffer); null: the copy engine brings them, kernel chain
    sl.src_points = ib ? (ib->fused ? ib->src_points : nullptr) : d_points;
    sl.src_poses = ib ? (ib->fused ? ib->src_poses : nullptr) : d_poses;
    sl.want_fused = ib ? ib->fused : fused_eligible(h, n);
    if (ib && !ib->fused)
        CC_CHECK(h, cudaStreamWaitEvent(h->stream, ib->h2d, 0));
    cc_status_t s = launch_push(h, sl);
    if (s != CC_OK)
        return s;
    h->pending[h->n_pending++] = h->next_slot;
    h->next_slot ^= 1;
    return CC_OK;
}

// The staged push is launched by the first submit / wait call AFTER the cc_wait() that made room for it, not by that
// cc_wait() itself: its first kernel recycles columns the finished push may just have reported, and the caller reads
// those between the two calls (cc_read_columns).
static cc_status_t launch_staged_if_room(cc_handle* h)
{
    if (h->staged < 0 || h->n_pending >= 2)
        return CC_OK;
    const cc_handle::InBuf& ib = h->inbuf[h->staged];
    h->staged = -1;
    return launch_from(h, ib.n, ib.d_raw, ib.d_poses, &ib);
}

static cc_status_t submit(cc_handle* h, int n, int rows, const void* points, const double* poses, bool device_inputs)
{
    cc_status_t s = check_push(h, n, rows);
    if (s != CC_OK)
        return s;
    if (!points || !poses)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    s = launch_staged_if_room(h);
    if (s != CC_OK)
        return s;
    if (h->staged >= 0 || (h->n_pending >= 2 && device_inputs))
    {
        h->error = device_inputs ? "two pushes are already in flight: call cc_wait() first"
                                 : "two pushes are in flight and a third is staged: call cc_wait() first";
        return CC_ERR_INVALID_ARGUMENT;
    }
    if (device_inputs)
        return launch_from(h, n, points, poses, nullptr);
    cc_handle::InBuf& ib = h->inbuf[h->next_in];
    const size_t pb = static_cast<size_t>(n) * rows * sizeof(cc_raw_point_t), qb = static_cast<size_t>(n) * 12 * sizeof(double);
    // page-locked caller buffers are copied straight to the device; anything else goes through the input buffer's own
    // pinned staging area first. The copy runs on its own stream: it overlaps the kernels of the pushes in flight.
    const void* src_pts = points;
    const void* src_poses = poses;
#ifndef CC_EMU
    cudaPointerAttributes attr;
    const bool pts_pinned = cudaPointerGetAttributes(&attr, points) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    const bool poses_pinned = cudaPointerGetAttributes(&attr, poses) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
#else
    const bool pts_pinned = false, poses_pinned = false;
#endif
    if (!pts_pinned)
    {
        std::memcpy(ib.h_raw, points, pb);
        src_pts = ib.h_raw;
    }
    if (!poses_pinned)
    {
        std::memcpy(ib.h_poses, poses, qb);
        src_poses = ib.h_poses;
    }
    ib.fused = fused_eligible(h, n);
    ib.src_points = src_pts;
    ib.src_poses = static_cast<const double*>(src_poses);
    if (!ib.fused)
    {
        CC_CHECK(h, cudaEventRecord(ib.h2d0, h->in_stream));
        static const int split = std::getenv("CC_B200_H2D_SPLIT") ? std::max(1, std::atoi(std::getenv("CC_B200_H2D_SPLIT"))) : 1;
        if (split > 1 && pb >= (1u << 20))
        {
            // experiment: the upper part of the copy on a second stream (a second copy engine)
            const size_t half = (pb / 2) & ~static_cast<size_t>(255);
            CC_CHECK(h, cudaStreamWaitEvent(h->aux_stream, ib.h2d0, 0));
            CC_CHECK(h, cudaMemcpyAsync(ib.d_raw + half, static_cast<const unsigned char*>(src_pts) + half, pb - half,
                                        cudaMemcpyHostToDevice, h->aux_stream));
            CC_CHECK(h, cudaEventRecord(h->ev1, h->aux_stream));
            CC_CHECK(h, cudaMemcpyAsync(ib.d_raw, src_pts, half, cudaMemcpyHostToDevice, h->in_stream));
            CC_CHECK(h, cudaStreamWaitEvent(h->in_stream, h->ev1, 0));
        }
        else
            CC_CHECK(h, cudaMemcpyAsync(ib.d_raw, src_pts, pb, cudaMemcpyHostToDevice, h->in_stream));
        CC_CHECK(h, cudaMemcpyAsync(ib.d_poses, src_poses, qb, cudaMemcpyHostToDevice, h->in_stream));
        CC_CHECK(h, cudaEventRecord(ib.h2d, h->in_stream));
    }
    ib.n = n;
    fill_devcfg(h, ib.cfg);
    ib.cfg_has_tf = h->has_tf;
    const int mine = h->next_in;
    h->next_in = (h->next_in + 1) % 3;
    if (h->n_pending >= 2)
    {
        h->staged = mine; // launched by the cc_wait() that makes room
        return CC_OK;
    }
    return launch_from(h, n, ib.d_raw, ib.d_poses, &ib);
}

cc_status_t cc_submit_firings(cc_handle_t* h, int n, int rows, const cc_raw_point_t* points, const double* poses)
{
    return submit(h, n, rows, points, poses, false);
}

cc_status_t cc_submit_firings_device(cc_handle_t* h, int n, int rows, const cc_raw_point_t* d_points, const double* d_poses)
{
    return submit(h, n, rows, d_points, d_poses, true);
}

cc_status_t cc_wait(cc_handle_t* h)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    cc_status_t s = launch_staged_if_room(h);
    if (s == CC_OK)
        s = finish_push(h);
    if (s != CC_OK)
        h->staged = -1; // dropped with the pushes in flight (the stream needs a reset)
    return s;
}

int cc_pending(const cc_handle_t* h)
{
    return h ? h->n_pending + (h->staged >= 0 ? 1 : 0) : 0;
}

static cc_status_t push_sync(cc_handle* h, int n, int rows, const void* points, const double* poses, bool device_inputs)
{
    if (h && (h->n_pending > 0 || h->staged >= 0))
    {
        h->error = "asynchronous pushes are in flight: call cc_wait() first";
        return CC_ERR_INVALID_ARGUMENT;
    }
    if (h && h->is_reset && rows == h->R && n == 0)
    {
        h->n_events = 0;
        h->clusters.clear();
        h->cluster_points.clear();
        h->points_view = nullptr;
    h->points_view = nullptr;
        std::memset(&h->info, 0, sizeof(h->info));
        return CC_OK;
    }
    cc_status_t s = submit(h, n, rows, points, poses, device_inputs);
    if (s != CC_OK)
        return s;
    return finish_push(h);
}

cc_status_t cc_push_firings(cc_handle_t* h, int n, int rows, const cc_raw_point_t* points, const double* poses)
{
    return push_sync(h, n, rows, points, poses, false);
}

cc_status_t cc_push_firings_device(cc_handle_t* h, int n, int rows, const cc_raw_point_t* d_points, const double* d_poses)
{
    return push_sync(h, n, rows, d_points, d_poses, true);
}

cc_status_t cc_get_batch_info(const cc_handle_t* h, cc_batch_info_t* out)
{
    if (!h || !out)
        return CC_ERR_INVALID_ARGUMENT;
    *out = h->info;
    return CC_OK;
}

} // extern "C"

template<typename T>
static cc_status_t copy_out(const std::vector<T>& v, T* out, int cap, int* n_out)
{
    if (cap < 0 || (cap > 0 && !out))
        return CC_ERR_INVALID_ARGUMENT;
    const int n = static_cast<int>(std::min<size_t>(v.size(), static_cast<size_t>(cap)));
    if (n)
        std::memcpy(out, v.data(), static_cast<size_t>(n) * sizeof(T));
    if (n_out)
        *n_out = n;
    return CC_OK;
}

extern "C" {

cc_status_t cc_get_result_views(const cc_handle_t* h, const cc_column_event_t** events, const cc_cluster_t** clusters,
                                const cc_cluster_point_t** points)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    if (events)
        *events = h->events.data();
    if (clusters)
        *clusters = h->clusters.data();
    if (points)
        *points = h->points_view;
    return CC_OK;
}

cc_status_t cc_get_column_events(const cc_handle_t* h, cc_column_event_t* out, int cap, int* n_out)
{
    if (!h || cap < 0 || (cap > 0 && !out))
        return CC_ERR_INVALID_ARGUMENT;
    const int n = static_cast<int>(std::min<size_t>(h->n_events, static_cast<size_t>(cap)));
    if (n)
        std::memcpy(out, h->events.data(), static_cast<size_t>(n) * sizeof(cc_column_event_t));
    if (n_out)
        *n_out = n;
    return CC_OK;
}
cc_status_t cc_get_clusters(const cc_handle_t* h, cc_cluster_t* out, int cap, int* n_out)
{
    return h ? copy_out(h->clusters, out, cap, n_out) : CC_ERR_INVALID_ARGUMENT;
}
cc_status_t cc_get_cluster_points(const cc_handle_t* h, cc_cluster_point_t* out, int cap, int* n_out)
{
    if (!h || cap < 0 || (cap > 0 && !out))
        return CC_ERR_INVALID_ARGUMENT;
    const int n = std::min(h->info.n_cluster_points, cap);
    if (n > 0 && h->points_view)
        std::memcpy(out, h->points_view, static_cast<size_t>(n) * sizeof(cc_cluster_point_t));
    if (n_out)
        *n_out = n > 0 ? n : 0;
    return CC_OK;
}

// what a caller reads from range_image_ inside a column callback (ros_utils.cpp:56-63, kitti_demo.cpp:183-216): one
// gather kernel writes packed records of the cells straight into a page-locked buffer of the handle
cc_status_t cc_export_columns(cc_handle_t* h, int64_t from, int64_t to, const cc_cell_t** cells)
{
    static_assert(sizeof(CcCell) == sizeof(cc_cell_t), "cell record layout");
    if (!h || !cells || !h->is_reset)
        return CC_ERR_INVALID_ARGUMENT;
    *cells = nullptr;
    if (to < from)
        return CC_OK;
    if (from < 0 || to - from + 1 > h->ringcols)
    {
        h->error = "cc_export_columns: range outside the ring";
        return CC_ERR_INVALID_ARGUMENT;
    }
    CC_CHECK(h, cudaSetDevice(h->device));
    const int ncols = static_cast<int>(to - from + 1);
    const size_t n = static_cast<size_t>(ncols) * h->R;
    if (n > h->export_cap)
    {
        if (h->h_export)
            cudaFreeHost(h->h_export);
        h->h_export = nullptr;
        h->export_cap = 0;
        const size_t cap = std::max<size_t>(n + n / 2, static_cast<size_t>(256) * h->R);
        CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_export), cap * sizeof(CcCell)));
        h->export_cap = cap;
    }
    // reads run on their own stream: the host has already seen every finished push complete, and they must not queue
    // behind the kernels of a push that is still in flight (which never touches columns already reported)
    CcDevCfg cfg;
    fill_devcfg(h, cfg);
    const int grid = std::max(1, std::min(h->sm_count * 8, static_cast<int>((n + 127) / 128)));
    CC_LAUNCH(k_export_cells, grid, 128, 0, h->aux_stream, cfg, h->d, static_cast<long long>(from), ncols, h->h_export);
    h->launches++;
    CC_CHECK(h, cudaStreamSynchronize(h->aux_stream));
    CC_CHECK(h, cudaGetLastError());
    *cells = reinterpret_cast<const cc_cell_t*>(h->h_export);
    return CC_OK;
}

// ---- PointCloud2 payloads packed on the device (ros_utils.cpp:11-77, 108-298) ----
static cc_status_t pack_cloud(cc_handle* h, int mode, int64_t from, int ncols, const CcClusterPoint* list, int npoints,
                              int64_t cfrom, int ccols, bool ground_only, uint64_t stamp, cc_cloud_view_t* out)
{
    CC_CHECK(h, cudaSetDevice(h->device));
    const int nwords = (ground_only ? CC_CLOUD_STEP_GROUND : CC_CLOUD_STEP_CLUSTER) / 4;
    const size_t total = mode == 0 ? static_cast<size_t>(ncols) * h->R : static_cast<size_t>(npoints);
    const size_t bytes = total * nwords * 4;
    if (bytes + 64 > h->cloud_cap)
    {
        if (h->h_cloud)
            cudaFreeHost(h->h_cloud);
        h->h_cloud = nullptr;
        h->cloud_cap = 0;
        const size_t cap = std::max<size_t>(bytes + bytes / 2 + 64, 1 << 20);
        CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_cloud), cap));
        h->cloud_cap = cap;
    }
    if (!h->d_min_stamp)
    {
        CC_CHECK(h, cudaMalloc(reinterpret_cast<void**>(&h->d_min_stamp), sizeof(unsigned long long)));
        CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_min_stamp), sizeof(unsigned long long)));
    }
    CcDevCfg cfg;
    fill_devcfg(h, cfg);
    const unsigned int* counts = nullptr;
    if (!ground_only && ccols > 0)
    {
        // child_points.size() of the cells (ros_utils.cpp:289): children name their parent
        const size_t n = static_cast<size_t>(ccols) * h->R;
        if (n > h->counts_cap)
        {
            if (h->d_counts)
                cudaFree(h->d_counts);
            h->d_counts = nullptr;
            h->counts_cap = 0;
            CC_CHECK(h, cudaMalloc(reinterpret_cast<void**>(&h->d_counts), (n + n / 2) * sizeof(unsigned int)));
            h->counts_cap = n + n / 2;
        }
        CC_CHECK(h, cudaMemsetAsync(h->d_counts, 0, n * sizeof(unsigned int), h->aux_stream));
        const int ahead = std::max(0, static_cast<int>(std::min<int64_t>(cfg.max_steps_row, h->state.ring_end - (cfrom + ccols - 1))));
        const int grid = std::max(1, std::min(h->sm_count * 8, static_cast<int>((n + 255) / 256)));
        CC_LAUNCH(k_child_counts, grid, 256, 0, h->aux_stream, cfg, h->d, static_cast<long long>(cfrom), ccols, ahead, h->d_counts);
        h->launches++;
        counts = h->d_counts;
    }
    *h->h_min_stamp = ~0ull;
    CC_CHECK(h, cudaMemcpyAsync(h->d_min_stamp, h->h_min_stamp, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->aux_stream));
#ifdef CC_EMU
    const int threads = 1;
#else
    const int threads = 128;
#endif
    const int warps = (threads + CC_WARP - 1) / CC_WARP;
    const int grid = std::max(1, std::min(h->sm_count * 8, static_cast<int>((total + threads - 1) / threads)));
    CC_LAUNCH(k_pack_cloud, grid, threads, static_cast<size_t>(warps) * CC_WARP * 29 * sizeof(unsigned int), h->aux_stream, cfg, h->d,
              mode, static_cast<long long>(from), ncols, list, npoints, nwords, static_cast<long long>(cfrom), counts, h->h_cloud,
              h->d_min_stamp);
    h->launches++;
    CC_CHECK(h, cudaMemcpyAsync(h->h_min_stamp, h->d_min_stamp, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->aux_stream));
    CC_CHECK(h, cudaStreamSynchronize(h->aux_stream));
    CC_CHECK(h, cudaGetLastError());
    out->data = h->h_cloud;
    out->data_size = bytes;
    out->point_step = static_cast<uint32_t>(nwords * 4);
    out->n_fields = ground_only ? 19 : 26;
    out->width = mode == 0 ? static_cast<uint32_t>(ncols) : static_cast<uint32_t>(npoints);
    out->height = mode == 0 ? static_cast<uint32_t>(h->R) : 1u;
    out->stamp_ns = mode == 0 ? (*h->h_min_stamp == ~0ull ? 0ull : *h->h_min_stamp) : stamp;
    return CC_OK;
}

cc_status_t cc_pack_columns_pointcloud2(cc_handle_t* h, int64_t from, int64_t to, int ground_points_only, cc_cloud_view_t* out)
{
    if (!h || !out || !h->is_reset)
        return CC_ERR_INVALID_ARGUMENT;
    std::memset(out, 0, sizeof(*out));
    if (to < from)
        return CC_OK; // columnToPointCloud returns no message for an empty range (ros_utils.cpp:40-42)
    if (from < 0 || to - from + 1 > h->ringcols)
    {
        h->error = "cc_pack_columns_pointcloud2: range outside the ring";
        return CC_ERR_INVALID_ARGUMENT;
    }
    const int ncols = static_cast<int>(to - from + 1);
    return pack_cloud(h, 0, from, ncols, nullptr, 0, from, ncols, ground_points_only != 0, 0, out);
}

cc_status_t cc_pack_cluster_pointcloud2(cc_handle_t* h, int cluster_index, cc_cloud_view_t* out)
{
    if (!h || !out || !h->is_reset)
        return CC_ERR_INVALID_ARGUMENT;
    std::memset(out, 0, sizeof(*out));
    if (cluster_index < 0 || cluster_index >= static_cast<int>(h->clusters.size()) || !h->last_points_dev)
    {
        h->error = "cc_pack_cluster_pointcloud2: no such cluster in the last finished push";
        return CC_ERR_INVALID_ARGUMENT;
    }
    const cc_cluster_t& c = h->clusters[cluster_index];
    const int ccols = static_cast<int>(c.max_gcol - c.min_gcol + 1);
    return pack_cloud(h, 1, 0, 0, h->last_points_dev + c.point_offset, static_cast<int>(c.num_points), c.min_gcol, ccols, false,
                      c.stamp, out);
}

cc_status_t cc_pack_requests_pointcloud2(cc_handle_t* h, int n, const cc_pack_request_t* requests, cc_cloud_view_t* out)
{
    if (!h || n < 0 || (n > 0 && (!requests || !out)) || !h->is_reset)
        return CC_ERR_INVALID_ARGUMENT;
    if (n == 0)
        return CC_OK;
    CC_CHECK(h, cudaSetDevice(h->device));
    if (n > h->requests_cap)
    {
        if (h->h_requests)
            cudaFreeHost(h->h_requests);
        if (h->h_req_stamps)
            cudaFreeHost(h->h_req_stamps);
        if (h->d_requests)
            cudaFree(h->d_requests);
        if (h->d_req_stamps)
            cudaFree(h->d_req_stamps);
        h->h_requests = nullptr;
        h->h_req_stamps = nullptr;
        h->d_requests = nullptr;
        h->d_req_stamps = nullptr;
        h->requests_cap = 0;
        const int cap = std::max(256, 2 * n);
        CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_requests), cap * sizeof(CcPackRequest)));
        CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_req_stamps), cap * sizeof(unsigned long long)));
        CC_CHECK(h, cudaMalloc(reinterpret_cast<void**>(&h->d_requests), cap * sizeof(CcPackRequest)));
        CC_CHECK(h, cudaMalloc(reinterpret_cast<void**>(&h->d_req_stamps), cap * sizeof(unsigned long long)));
        h->requests_cap = cap;
    }
    // layout: payloads one after the other (16-byte aligned), 32-point tasks numbered request by request
    size_t bytes = 0;
    int ntasks = 0;
    int64_t cmin = INT64_MAX, cmax = -1;
    for (int i = 0; i < n; i++)
    {
        const cc_pack_request_t& r = requests[i];
        CcPackRequest q;
        std::memset(&q, 0, sizeof(q));
        q.kind = r.kind;
        if (r.kind == 2)
        {
            if (r.cluster_index < 0 || r.cluster_index >= static_cast<int>(h->clusters.size()) || !h->last_points_dev)
            {
                h->error = "cc_pack_requests_pointcloud2: no such cluster in the last finished push";
                return CC_ERR_INVALID_ARGUMENT;
            }
            const cc_cluster_t& c = h->clusters[r.cluster_index];
            q.npoints = static_cast<int>(c.num_points);
            q.list_offset = static_cast<int>(c.point_offset);
            cmin = std::min(cmin, c.min_gcol);
            cmax = std::max(cmax, c.max_gcol);
        }
        else if (r.kind == 0 || r.kind == 1)
        {
            if (r.to_gcol >= r.from_gcol)
            {
                if (r.from_gcol < 0 || r.to_gcol - r.from_gcol + 1 > h->ringcols)
                {
                    h->error = "cc_pack_requests_pointcloud2: range outside the ring";
                    return CC_ERR_INVALID_ARGUMENT;
                }
                q.from = r.from_gcol;
                q.ncols = static_cast<int>(r.to_gcol - r.from_gcol + 1);
                q.npoints = q.ncols * h->R;
                if (r.kind == 1)
                {
                    cmin = std::min(cmin, r.from_gcol);
                    cmax = std::max(cmax, r.to_gcol);
                }
            }
        }
        else
            return CC_ERR_INVALID_ARGUMENT;
        q.out_offset = static_cast<long long>(bytes);
        q.first_task = ntasks;
        const size_t step = q.kind == 0 ? CC_CLOUD_STEP_GROUND : CC_CLOUD_STEP_CLUSTER;
        bytes += (static_cast<size_t>(q.npoints) * step + 15) / 16 * 16;
        ntasks += (q.npoints + CC_WARP - 1) / CC_WARP;
        h->h_requests[i] = q;
        h->h_req_stamps[i] = ~0ull;
    }
    if (bytes + 64 > h->cloud_cap)
    {
        if (h->h_cloud)
            cudaFreeHost(h->h_cloud);
        h->h_cloud = nullptr;
        h->cloud_cap = 0;
        const size_t cap = std::max<size_t>(bytes + bytes / 2 + 64, 1 << 20);
        CC_CHECK(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_cloud), cap));
        h->cloud_cap = cap;
    }
    CcDevCfg cfg;
    fill_devcfg(h, cfg);
    const unsigned int* counts = nullptr;
    if (cmax >= cmin && cmax >= 0)
    {
        const int ccols = static_cast<int>(cmax - cmin + 1);
        const size_t cn = static_cast<size_t>(ccols) * h->R;
        if (cn > h->counts_cap)
        {
            if (h->d_counts)
                cudaFree(h->d_counts);
            h->d_counts = nullptr;
            h->counts_cap = 0;
            CC_CHECK(h, cudaMalloc(reinterpret_cast<void**>(&h->d_counts), (cn + cn / 2) * sizeof(unsigned int)));
            h->counts_cap = cn + cn / 2;
        }
        CC_CHECK(h, cudaMemsetAsync(h->d_counts, 0, cn * sizeof(unsigned int), h->aux_stream));
        const int ahead = std::max(0, static_cast<int>(std::min<int64_t>(cfg.max_steps_row, h->state.ring_end - cmax)));
        const int grid = std::max(1, std::min(h->sm_count * 8, static_cast<int>((cn + 255) / 256)));
        CC_LAUNCH(k_child_counts, grid, 256, 0, h->aux_stream, cfg, h->d, static_cast<long long>(cmin), ccols, ahead, h->d_counts);
        h->launches++;
        counts = h->d_counts;
    }
    CC_CHECK(h, cudaMemcpyAsync(h->d_requests, h->h_requests, n * sizeof(CcPackRequest), cudaMemcpyHostToDevice, h->aux_stream));
    CC_CHECK(h, cudaMemcpyAsync(h->d_req_stamps, h->h_req_stamps, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->aux_stream));
    if (ntasks > 0)
    {
#ifdef CC_EMU
        const int threads = 1;
#else
        const int threads = 128;
#endif
        const int warps = (threads + CC_WARP - 1) / CC_WARP;
        const int grid = std::max(1, std::min(h->sm_count * 8, (ntasks + warps - 1) / warps));
        CC_LAUNCH(k_pack_requests, grid, threads, static_cast<size_t>(warps) * CC_WARP * 29 * sizeof(unsigned int), h->aux_stream, cfg,
                  h->d, h->d_requests, n, ntasks, h->last_points_dev, static_cast<long long>(cmin == INT64_MAX ? 0 : cmin), counts,
                  h->h_cloud, h->d_req_stamps);
        h->launches++;
    }
    CC_CHECK(h, cudaMemcpyAsync(h->h_req_stamps, h->d_req_stamps, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->aux_stream));
    CC_CHECK(h, cudaStreamSynchronize(h->aux_stream));
    CC_CHECK(h, cudaGetLastError());
    for (int i = 0; i < n; i++)
    {
        const CcPackRequest& q = h->h_requests[i];
        cc_cloud_view_t& v = out[i];
        std::memset(&v, 0, sizeof(v));
        if (q.npoints == 0 && q.kind != 2)
            continue;
        v.point_step = q.kind == 0 ? CC_CLOUD_STEP_GROUND : CC_CLOUD_STEP_CLUSTER;
        v.n_fields = q.kind == 0 ? 19 : 26;
        v.data = h->h_cloud + q.out_offset;
        v.data_size = static_cast<uint64_t>(q.npoints) * v.point_step;
        v.width = q.kind == 2 ? static_cast<uint32_t>(q.npoints) : static_cast<uint32_t>(q.ncols);
        v.height = q.kind == 2 ? 1u : static_cast<uint32_t>(h->R);
        v.stamp_ns = q.kind == 2 ? h->clusters[requests[i].cluster_index].stamp : (h->h_req_stamps[i] == ~0ull ? 0ull : h->h_req_stamps[i]);
    }
    return CC_OK;
}

cc_status_t cc_read_columns(cc_handle_t* h, int64_t from, int64_t to, const cc_column_fields_t* f)
{
    if (!h || !f || !h->is_reset)
        return CC_ERR_INVALID_ARGUMENT;
    if (to < from)
        return CC_OK;
    const cc_cell_t* cells = nullptr;
    cc_status_t s = cc_export_columns(h, from, to, &cells);
    if (s != CC_OK)
    {
        if (h->error.rfind("cc_export_columns", 0) == 0)
            h->error = "cc_read_columns: range outside the ring";
        return s;
    }
    const size_t n = static_cast<size_t>(to - from + 1) * h->R;
    for (size_t i = 0; i < n; i++)
    {
        const cc_cell_t& c = cells[i];
        if (f->xyz)
        {
            f->xyz[3 * i + 0] = c.x;
            f->xyz[3 * i + 1] = c.y;
            f->xyz[3 * i + 2] = c.z;
        }
        if (f->distance)
            f->distance[i] = c.distance;
        if (f->azimuth_angle)
            f->azimuth_angle[i] = c.azimuth_angle;
        if (f->inclination_angle)
            f->inclination_angle[i] = c.inclination_angle;
        if (f->continuous_azimuth_angle)
            f->continuous_azimuth_angle[i] = c.continuous_azimuth_angle;
        if (f->global_column_index)
            f->global_column_index[i] = c.global_column_index;
        if (f->stamp)
            f->stamp[i] = c.stamp;
        if (f->globally_unique_point_index)
            f->globally_unique_point_index[i] = c.globally_unique_point_index;
        if (f->firing_index)
            f->firing_index[i] = c.firing_index;
        if (f->intensity)
            f->intensity[i] = c.intensity;
        if (f->ground_point_label)
            f->ground_point_label[i] = c.ground_point_label;
        if (f->debug_ground_point_label)
            f->debug_ground_point_label[i] = c.debug_ground_point_label;
        if (f->is_ignored)
            f->is_ignored[i] = c.is_ignored;
        if (f->id)
            f->id[i] = c.id;
        if (f->tree_root_gcol)
            f->tree_root_gcol[i] = c.tree_root_gcol;
        if (f->tree_root_row)
            f->tree_root_row[i] = c.tree_root_row;
        if (f->finished_at_continuous_azimuth_angle)
            f->finished_at_continuous_azimuth_angle[i] = c.finished_at_continuous_azimuth_angle;
        if (f->tree_num_points)
            f->tree_num_points[i] = c.tree_num_points;
        if (f->cluster_width)
            f->cluster_width[i] = c.cluster_width;
        if (f->number_of_visited_neighbors)
            f->number_of_visited_neighbors[i] = c.number_of_visited_neighbors;
        if (f->first_parent_gcol)
            f->first_parent_gcol[i] = c.first_parent_gcol;
        if (f->first_parent_row)
            f->first_parent_row[i] = c.first_parent_row;
        if (f->belongs_to_finished_cluster)
            f->belongs_to_finished_cluster[i] = c.belongs_to_finished_cluster;
    }
    return CC_OK;
}

cc_status_t cc_set_label_prefetch(cc_handle_t* h, int enable)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    h->label_prefetch = enable != 0;
    return CC_OK;
}

cc_status_t cc_get_column_labels(const cc_handle_t* h, const uint8_t** labels, int* n_cols)
{
    if (!h || !labels || !n_cols)
        return CC_ERR_INVALID_ARGUMENT;
    *labels = reinterpret_cast<const uint8_t*>(h->cur_labels);
    *n_cols = h->cur_labels ? h->cur_label_cols : 0;
    return CC_OK;
}

cc_status_t cc_debug_slot_times(cc_handle_t* h, int slot, float out_ms[5])
{
    if (!h || slot < 0 || slot > 1 || !out_ms || !h->ev0)
        return CC_ERR_INVALID_ARGUMENT;
    cc_handle::Slot& sl = h->slots[slot];
    // relative to the handle's base event (recorded by cc_debug_slot_base): input copy start / end, kernels start /
    // end, results on the host; -1 where an event has not been recorded or has not completed
    cudaEvent_t ev[5] = {sl.h2d0, sl.h2d, sl.ev0, sl.ev1, sl.done};
    for (int i = 0; i < 5; i++)
    {
        float ms = -1.f;
        if (!ev[i] || cudaEventElapsedTime(&ms, h->ev0, ev[i]) != cudaSuccess)
            ms = -1.f;
        out_ms[i] = ms;
    }
    cudaGetLastError();
    return CC_OK;
}

cc_status_t cc_debug_slot_base(cc_handle_t* h)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    CC_CHECK(h, cudaDeviceSynchronize());
    CC_CHECK(h, cudaEventRecord(h->ev0, h->stream));
    CC_CHECK(h, cudaEventSynchronize(h->ev0));
    return CC_OK;
}

int cc_debug_event_query(cc_handle_t* h, int slot, int which)
{
    if (!h || slot < 0 || slot > 1)
        return -1;
    cc_handle::Slot& sl = h->slots[slot];
    cudaEvent_t e = which == 0 ? sl.ev0 : which == 1 ? sl.ev1 : which == 2 ? sl.ready : sl.done;
    return cudaEventQuery(e) == cudaSuccess ? 1 : 0;
}

cc_status_t cc_debug_trace(cc_handle_t* h, int enable)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    const size_t bytes = static_cast<size_t>(CC_TRACE_KERNELS) * CC_TRACE_BLOCKS * 2 * sizeof(unsigned long long);
    if (enable && !h->d_trace)
        CC_CHECK(h, cudaMalloc(reinterpret_cast<void**>(&h->d_trace), bytes));
    if (h->d_trace)
    {
        CC_CHECK(h, cudaStreamSynchronize(h->stream));
        CC_CHECK(h, cudaMemset(h->d_trace, 0, bytes));
    }
    h->trace_on = enable != 0;
    return CC_OK;
}

cc_status_t cc_debug_get_trace(cc_handle_t* h, char* names, int names_cap, uint64_t* out, int cap_kernels, int* n_out)
{
    if (!h || !out || !n_out || !h->d_trace)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    const size_t count = static_cast<size_t>(CC_TRACE_KERNELS) * CC_TRACE_BLOCKS * 2;
    std::vector<unsigned long long> t(count);
    CC_CHECK(h, cudaStreamSynchronize(h->stream));
    CC_CHECK(h, cudaMemcpy(t.data(), h->d_trace, count * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CC_CHECK(h, cudaMemset(h->d_trace, 0, count * sizeof(unsigned long long)));
    const int n = std::min<int>(cap_kernels, CC_KID_COUNT);
    for (int k = 0; k < n; k++)
    {
        // per kernel: first block entry, last block exit, longest single block, blocks seen
        uint64_t t0 = ~0ull, t1 = 0, longest = 0, blocks = 0;
        for (int b = 0; b < CC_TRACE_BLOCKS; b++)
        {
            const unsigned long long a = t[(static_cast<size_t>(k) * CC_TRACE_BLOCKS + b) * 2], z = t[(static_cast<size_t>(k) * CC_TRACE_BLOCKS + b) * 2 + 1];
            if (!a || !z)
                continue;
            blocks++;
            t0 = std::min<uint64_t>(t0, a);
            t1 = std::max<uint64_t>(t1, z);
            longest = std::max<uint64_t>(longest, z - a);
        }
        out[4 * k + 0] = blocks ? t0 : 0;
        out[4 * k + 1] = t1;
        out[4 * k + 2] = longest;
        out[4 * k + 3] = blocks;
    }
    if (names && names_cap > 0)
    {
        std::strncpy(names, CC_KERNEL_NAMES, static_cast<size_t>(names_cap) - 1);
        names[names_cap - 1] = 0;
    }
    *n_out = n;
    return CC_OK;
}

cc_status_t cc_debug_flag_columns(cc_handle_t* h, int period)
{
    if (!h || period < 0)
        return CC_ERR_INVALID_ARGUMENT;
    h->debug_flag_period = period;
    return CC_OK;
}

cc_status_t cc_set_kernel_timing(cc_handle_t* h, int enable)
{
    if (!h)
        return CC_ERR_INVALID_ARGUMENT;
    CC_CHECK(h, cudaSetDevice(h->device));
    if (enable && h->tev.empty())
    {
        h->tev.resize(2 * 64);
        h->tnames.resize(64);
        for (cudaEvent_t& e : h->tev)
            CC_CHECK(h, cudaEventCreate(&e));
    }
    h->timing = enable != 0;
    h->n_timed = 0;
    return CC_OK;
}

cc_status_t cc_get_kernel_timings(cc_handle_t* h, char* names, int names_cap, float* ms, int cap, int* n_out)
{
    if (!h || !n_out)
        return CC_ERR_INVALID_ARGUMENT;
    std::string all;
    int n = 0;
    for (int i = 0; i < h->n_timed && i < cap; i++)
    {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->tev[2 * i], h->tev[2 * i + 1]) != cudaSuccess)
            t = -1.f;
        ms[i] = t;
        all += h->tnames[i];
        all += ';';
        n++;
    }
    if (names && names_cap > 0)
    {
        std::strncpy(names, all.c_str(), static_cast<size_t>(names_cap) - 1);
        names[names_cap - 1] = 0;
    }
    *n_out = n;
    return CC_OK;
}

} // extern "C"

// ---- KITTI replay front-end on the device (kitti_loader.cpp:47-210, 297-328; kitti_demo.cpp:123-159, 369-403) ----
namespace
{
// KittiLoader::interpolate, kitti_loader.cpp:297-328: the pose of the sequence at `stamp` -- rotation by quaternion
// slerp (Eigen 3.3's published formulas: matrix -> quaternion by the trace branches, the (1 - epsilon) shortcut,
// quaternion -> matrix from the doubled products), translation linearly. f64 on the host, 3x4 row major.
struct KQuat
{
    double x, y, z, w;
};
static KQuat kq_from_matrix(const double* m)
{
    double c[4];
    double t = m[0] + m[5] + m[10];
    if (t > 0.)
    {
        t = std::sqrt(t + 1.0);
        c[3] = 0.5 * t;
        t = 0.5 / t;
        c[0] = (m[2 * 4 + 1] - m[1 * 4 + 2]) * t;
        c[1] = (m[0 * 4 + 2] - m[2 * 4 + 0]) * t;
        c[2] = (m[1 * 4 + 0] - m[0 * 4 + 1]) * t;
    }
    else
    {
        int i = 0;
        if (m[5] > m[0])
            i = 1;
        if (m[10] > m[i * 4 + i])
            i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i * 4 + i] - m[j * 4 + j] - m[k * 4 + k] + 1.0);
        c[i] = 0.5 * t;
        t = 0.5 / t;
        c[3] = (m[k * 4 + j] - m[j * 4 + k]) * t;
        c[j] = (m[j * 4 + i] + m[i * 4 + j]) * t;
        c[k] = (m[k * 4 + i] + m[i * 4 + k]) * t;
    }
    return KQuat{c[0], c[1], c[2], c[3]};
}
static KQuat kq_slerp(const KQuat& a, double t, const KQuat& b)
{
    const double one = 1.0 - 2.220446049250313e-16;
    const double d = (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    const double ad = std::fabs(d);
    double s0, s1;
    if (ad >= one)
    {
        s0 = 1.0 - t;
        s1 = t;
    }
    else
    {
        const double theta = std::acos(ad), st = std::sin(theta);
        s0 = std::sin((1.0 - t) * theta) / st;
        s1 = std::sin(t * theta) / st;
    }
    if (d < 0.)
        s1 = -s1;
    return KQuat{s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}
static void kq_to_matrix(const KQuat& q, double* m)
{
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    m[0] = 1.0 - (tyy + tzz);
    m[1] = txy - twz;
    m[2] = txz + twy;
    m[4] = txy + twz;
    m[5] = 1.0 - (txx + tzz);
    m[6] = tyz - twx;
    m[8] = txz - twy;
    m[9] = tyz + twx;
    m[10] = 1.0 - (txx + tyy);
}
static void kitti_interpolate(const std::vector<uint64_t>& stamps, const std::vector<double>& poses, uint64_t stamp, double* out)
{
    const size_t after = static_cast<size_t>(std::lower_bound(stamps.begin(), stamps.end(), stamp) - stamps.begin());
    if (after == stamps.size())
        std::memcpy(out, poses.data() + 12 * (after - 1), 12 * sizeof(double));
    else if (after == 0)
        std::memcpy(out, poses.data(), 12 * sizeof(double));
    else
    {
        const double* pb = poses.data() + 12 * (after - 1);
        const double* pa = poses.data() + 12 * after;
        const double f = static_cast<double>(stamp - stamps[after - 1]) / static_cast<double>(stamps[after] - stamps[after - 1]);
        kq_to_matrix(kq_slerp(kq_from_matrix(pb), f, kq_from_matrix(pa)), out);
        for (int i = 0; i < 3; i++)
            out[4 * i + 3] = (1 - f) * pb[4 * i + 3] + f * pa[4 * i + 3];
    }
}
static void host_iso_inverse(const double* m, double* r)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r[i * 4 + j] = m[j * 4 + i];
    for (int i = 0; i < 3; i++)
        r[i * 4 + 3] = -(r[i * 4 + 0] * m[3] + (r[i * 4 + 1] * m[7] + r[i * 4 + 2] * m[11]));
}
static void host_iso_mul(const double* a, const double* b, double* r)
{
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++)
            r[i * 4 + j] = a[i * 4 + 0] * b[j] + (a[i * 4 + 1] * b[4 + j] + a[i * 4 + 2] * b[8 + j]);
        r[i * 4 + 3] = (a[i * 4 + 0] * b[3] + (a[i * 4 + 1] * b[7] + a[i * 4 + 2] * b[11])) + a[i * 4 + 3];
    }
}
} // namespace

struct cc_kitti
{
    int device{0};
    int max_points{0};
    int max_bins{0};
    cudaStream_t stream{nullptr};
    CcKittiPtrs d{};
    float4* d_xyzi{nullptr};
    double* d_bin_tf{nullptr};
    CcRawPoint* d_firings{nullptr};
    double* d_poses{nullptr};
    double* h_poses{nullptr};  // page-locked [W][12]
    double* h_bin_tf{nullptr}; // page-locked [max_bins][12]
    int* h_row_start{nullptr}; // page-locked [H + 2]
    int last_n{0};
    std::vector<uint64_t> stamps;
    std::vector<double> poses;
    std::vector<void*> allocs;
};

extern "C" {

cc_status_t cc_kitti_create(int device_ordinal, int max_points_per_frame, cc_kitti_t** out)
{
    if (!out || max_points_per_frame <= 0)
        return CC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device_ordinal < 0 || device_ordinal >= ndev)
        return CC_ERR_CUDA;
    cc_kitti* k = new cc_kitti();
    k->device = device_ordinal;
    k->max_points = max_points_per_frame;
    k->max_bins = 1024; // a rotation of up to one second
    const size_t np = static_cast<size_t>(max_points_per_frame);
    const size_t cells = static_cast<size_t>(CC_KITTI_W) * CC_KITTI_H;
    auto alloc = [&](void** p, size_t bytes) -> bool
    {
        if (cudaMalloc(p, bytes) != cudaSuccess)
            return false;
        k->allocs.push_back(*p);
        return true;
    };
    bool ok = cudaSetDevice(device_ordinal) == cudaSuccess &&
              cudaStreamCreateWithFlags(&k->stream, cudaStreamNonBlocking) == cudaSuccess &&
              alloc(reinterpret_cast<void**>(&k->d_xyzi), np * sizeof(float4)) &&
              alloc(reinterpret_cast<void**>(&k->d.wrap_flag), np) &&
              alloc(reinterpret_cast<void**>(&k->d.seg_sums), ((np + CC_KITTI_SEG - 1) / CC_KITTI_SEG + 1) * sizeof(int)) &&
              alloc(reinterpret_cast<void**>(&k->d.laser_index), np) &&
              alloc(reinterpret_cast<void**>(&k->d.uncorrected), np * 3 * sizeof(float)) &&
              alloc(reinterpret_cast<void**>(&k->d.column), np * sizeof(int)) &&
              alloc(reinterpret_cast<void**>(&k->d.row_start), (CC_KITTI_H + 2) * sizeof(int)) &&
              alloc(reinterpret_cast<void**>(&k->d.cell_point), cells * sizeof(int)) &&
              alloc(reinterpret_cast<void**>(&k->d_bin_tf), static_cast<size_t>(k->max_bins) * 12 * sizeof(double)) &&
              alloc(reinterpret_cast<void**>(&k->d_firings), cells * sizeof(CcRawPoint)) &&
              alloc(reinterpret_cast<void**>(&k->d_poses), static_cast<size_t>(CC_KITTI_W) * 12 * sizeof(double)) &&
              cudaMallocHost(reinterpret_cast<void**>(&k->h_poses), static_cast<size_t>(CC_KITTI_W) * 12 * sizeof(double)) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void**>(&k->h_bin_tf), static_cast<size_t>(k->max_bins) * 12 * sizeof(double)) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void**>(&k->h_row_start), (CC_KITTI_H + 2) * sizeof(int)) == cudaSuccess;
    if (!ok)
    {
        cc_kitti_destroy(k);
        return CC_ERR_CUDA;
    }
    *out = k;
    return CC_OK;
}

void cc_kitti_destroy(cc_kitti_t* k)
{
    if (!k)
        return;
    cudaSetDevice(k->device);
    if (k->stream)
    {
        cudaStreamSynchronize(k->stream);
        cudaStreamDestroy(k->stream);
    }
    for (void* p : k->allocs)
        cudaFree(p);
    if (k->h_poses)
        cudaFreeHost(k->h_poses);
    if (k->h_bin_tf)
        cudaFreeHost(k->h_bin_tf);
    if (k->h_row_start)
        cudaFreeHost(k->h_row_start);
    delete k;
}

cc_status_t cc_kitti_set_poses(cc_kitti_t* k, int n_poses, const uint64_t* stamps, const double* poses12)
{
    if (!k || n_poses <= 0 || !stamps || !poses12)
        return CC_ERR_INVALID_ARGUMENT;
    k->stamps.assign(stamps, stamps + n_poses);
    k->poses.assign(poses12, poses12 + static_cast<size_t>(n_poses) * 12);
    return CC_OK;
}

cc_status_t cc_kitti_frame(cc_kitti_t* k, int n_points, const float* xyzi, uint64_t stamp_start, uint64_t stamp_end,
                           const double* frame_pose12, int sequence_index, int frame_index, cc_kitti_frame_t* out)
{
    if (!k || !out || n_points < 0 || n_points > k->max_points || (n_points > 0 && !xyzi) || !frame_pose12 || k->stamps.empty() ||
        stamp_end < stamp_start)
        return CC_ERR_INVALID_ARGUMENT;
    // undoEgoMotionCorrection's lookup table (kitti_loader.cpp:183-197): one transform per millisecond of the rotation
    const uint64_t bin_resolution = 1000000;
    const uint64_t duration = stamp_end - stamp_start;
    const int n_bins = static_cast<int>(std::ceil(static_cast<double>(duration) / static_cast<double>(bin_resolution)));
    if (n_bins > k->max_bins)
        return CC_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(k->device) != cudaSuccess)
        return CC_ERR_CUDA;
    cudaStream_t st = k->stream;
    if (n_points > 0 &&
        cudaMemcpyAsync(k->d_xyzi, xyzi, static_cast<size_t>(n_points) * sizeof(float4), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return CC_ERR_CUDA;
    for (int b = 0; b < n_bins; b++)
    {
        const uint64_t stamp_at_bin = stamp_start + static_cast<uint64_t>(b) * bin_resolution + (bin_resolution / 2);
        double pose[12], inv[12];
        kitti_interpolate(k->stamps, k->poses, stamp_at_bin, pose);
        host_iso_inverse(pose, inv);
        host_iso_mul(inv, frame_pose12, k->h_bin_tf + 12 * b);
    }
    if (n_bins > 0 && cudaMemcpyAsync(k->d_bin_tf, k->h_bin_tf, static_cast<size_t>(n_bins) * 12 * sizeof(double),
                                      cudaMemcpyHostToDevice, st) != cudaSuccess)
        return CC_ERR_CUDA;
    CcKittiPtrs d = k->d;
    d.xyzi = k->d_xyzi;
    d.n = n_points;
    d.bin_tf = k->d_bin_tf;
    d.n_bins = n_bins > 0 ? n_bins : 1;
    d.stamp_start = stamp_start;
    d.stamp_end = stamp_end;
    const int nseg = (n_points + CC_KITTI_SEG - 1) / CC_KITTI_SEG;
    const int g_seg = std::max(1, std::min(148 * 4, nseg));
    const int g_pts = std::max(1, std::min(148 * 8, (n_points + 255) / 256));
    CC_LAUNCH(k_kitti_flags, g_seg, 256, 0, st, d);
    CC_LAUNCH(k_kitti_rows, g_seg, 256, 0, st, d);
    CC_LAUNCH(k_kitti_undo_ego, g_pts, 256, 0, st, d);
    CC_LAUNCH(k_kitti_range_image, CC_KITTI_H, 256, 0, st, d);
    CC_LAUNCH(k_kitti_firings, 148 * 4, 256, 0, st, d, k->d_firings, sequence_index, frame_index);
    // the pose of every pseudo firing (kitti_demo.cpp:386-396) while the device works
    for (int col = 0; col < CC_KITTI_W; col++)
    {
        const double elapsed_ratio = static_cast<double>(col) / (CC_KITTI_W - 1);
        const double elapsed_time = static_cast<double>(duration) * elapsed_ratio;
        kitti_interpolate(k->stamps, k->poses, stamp_start + static_cast<uint64_t>(elapsed_time), k->h_poses + 12 * col);
    }
    if (cudaMemcpyAsync(k->d_poses, k->h_poses, static_cast<size_t>(CC_KITTI_W) * 12 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(k->h_row_start, d.row_start, (CC_KITTI_H + 2) * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return CC_ERR_CUDA;
    k->last_n = n_points;
    // the reference's sanity checks (kitti_loader.cpp:91-98): rows found, longest COMPLETED row
    int rows = 0, longest = 0;
    for (int r = 0; r <= CC_KITTI_H; r++)
        if (n_points > 0 && (r == 0 || k->h_row_start[r] < n_points))
            rows = r + 1;
    for (int r = 0; r + 1 < std::min(rows, CC_KITTI_H); r++)
        longest = std::max(longest, k->h_row_start[r + 1] - k->h_row_start[r]);
    out->n_firings = CC_KITTI_W;
    out->rows_per_firing = CC_KITTI_H;
    out->d_firings = reinterpret_cast<const cc_raw_point_t*>(k->d_firings);
    out->d_poses = k->d_poses;
    out->poses = k->h_poses;
    out->rows_found = rows;
    out->max_points_in_row = longest;
    return longest > CC_KITTI_W ? CC_ERR_INVALID_ARGUMENT : CC_OK; // "More points in a single row than expected", cpp:96-97
}

cc_status_t cc_kitti_read_debug(cc_kitti_t* k, uint8_t* laser_index, int32_t* cell_point, float* uncorrected_xyz, cc_raw_point_t* firings)
{
    if (!k)
        return CC_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(k->device) != cudaSuccess)
        return CC_ERR_CUDA;
    const size_t n = static_cast<size_t>(k->last_n), cells = static_cast<size_t>(CC_KITTI_W) * CC_KITTI_H;
    bool ok = true;
    if (laser_index && n)
        ok = ok && cudaMemcpy(laser_index, k->d.laser_index, n, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (cell_point)
        ok = ok && cudaMemcpy(cell_point, k->d.cell_point, cells * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (uncorrected_xyz && n)
        ok = ok && cudaMemcpy(uncorrected_xyz, k->d.uncorrected, n * 3 * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (firings)
        ok = ok && cudaMemcpy(firings, k->d_firings, cells * sizeof(CcRawPoint), cudaMemcpyDeviceToHost) == cudaSuccess;
    return ok ? CC_OK : CC_ERR_CUDA;
}

} // extern "C"

// ---- sensor packets -> firings on the device (ouster_input.hpp:105-181) ----
static_assert(sizeof(cc_ouster_format_t) == sizeof(CcOusterFormat), "cc_ouster_format_t and CcOusterFormat must match");
struct cc_ouster
{
    int device{0};
    int max_packets{0};
    cudaStream_t stream{nullptr};
    CcOusterFormat fmt{};
    int packet_size{0};
    bool has_lut{false};
    bool interrupt{true}; // OusterInput::interrupt_message after reset() (:97-101): the packet in flight is discarded
    unsigned long long firing_index{0};
    unsigned char* d_packets{nullptr};
    unsigned long long* d_recv{nullptr};
    float* d_direction{nullptr};
    float* d_offset{nullptr};
    int* d_firing_of_column{nullptr};
    int* d_n_firings{nullptr};
    CcRawPoint* d_firings{nullptr};
    unsigned long long* d_firing_stamp{nullptr};
    unsigned long long* h_firing_stamp{nullptr}; // page-locked
    int* h_n_firings{nullptr};                   // page-locked
    std::vector<void*> allocs;
};

extern "C" {

void cc_ouster_format_legacy(int pixels_per_column, int columns_per_frame, cc_ouster_format_t* f)
{
    // the LEGACY lidar data profile: 16 measurement blocks per packet; block = 16-byte header (timestamp u64,
    // measurement id u16, frame id u16, encoder u32), 12 bytes per pixel (range u32 of which 20 bits, flags in the top 4,
    // reflectivity u16, signal u16, near-infrared u16, 2 unused), 4-byte status (0xffffffff = valid)
    std::memset(f, 0, sizeof(*f));
    f->columns_per_packet = 16;
    f->pixels_per_column = pixels_per_column;
    f->columns_per_frame = columns_per_frame;
    f->packet_header_size = 0;
    f->col_header_size = 16;
    f->col_footer_size = 4;
    f->pixel_bytes = 12;
    f->col_measurement_id_offset = 8;
    f->col_status_offset = 16 + 12 * pixels_per_column;
    f->col_status_bytes = 4;
    f->range_offset = 0;
    f->range_bytes = 4;
    f->range_mask = 0x000fffffu;
    f->range_shift = 0;
    f->signal_offset = 6;
    f->signal_bytes = 2;
    f->signal_mask = 0;
    f->signal_shift = 0;
    f->offset_from_direction_table = 1;
}

cc_status_t cc_ouster_create(int device_ordinal, const cc_ouster_format_t* format, int max_packets_per_call, cc_ouster_t** out)
{
    if (!out || !format || max_packets_per_call <= 0 || format->columns_per_packet <= 0 || format->pixels_per_column <= 0 ||
        format->columns_per_frame <= 0 || format->pixel_bytes <= 0 || format->range_bytes < 1 || format->range_bytes > 4 ||
        format->signal_bytes < 1 || format->signal_bytes > 4 || (format->col_status_bytes != 2 && format->col_status_bytes != 4))
        return CC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device_ordinal < 0 || device_ordinal >= ndev)
        return CC_ERR_CUDA;
    cc_ouster* o = new cc_ouster();
    o->device = device_ordinal;
    o->max_packets = max_packets_per_call;
    std::memcpy(&o->fmt, format, sizeof(CcOusterFormat));
    const int H = format->pixels_per_column;
    o->packet_size = format->packet_header_size +
                     format->columns_per_packet * (format->col_header_size + H * format->pixel_bytes + format->col_footer_size);
    const size_t ncol = static_cast<size_t>(max_packets_per_call) * format->columns_per_packet;
    const size_t lut = static_cast<size_t>(format->columns_per_frame) * H * 3 * sizeof(float);
    auto alloc = [&](void** p, size_t bytes) -> bool
    {
        if (cudaMalloc(p, bytes) != cudaSuccess)
            return false;
        o->allocs.push_back(*p);
        return true;
    };
    bool ok = cudaSetDevice(device_ordinal) == cudaSuccess &&
              cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking) == cudaSuccess &&
              alloc(reinterpret_cast<void**>(&o->d_packets), static_cast<size_t>(max_packets_per_call) * o->packet_size) &&
              alloc(reinterpret_cast<void**>(&o->d_recv), static_cast<size_t>(max_packets_per_call) * sizeof(unsigned long long)) &&
              alloc(reinterpret_cast<void**>(&o->d_direction), lut) && alloc(reinterpret_cast<void**>(&o->d_offset), lut) &&
              alloc(reinterpret_cast<void**>(&o->d_firing_of_column), ncol * sizeof(int)) &&
              alloc(reinterpret_cast<void**>(&o->d_n_firings), sizeof(int)) &&
              alloc(reinterpret_cast<void**>(&o->d_firings), ncol * H * sizeof(CcRawPoint)) &&
              alloc(reinterpret_cast<void**>(&o->d_firing_stamp), ncol * sizeof(unsigned long long)) &&
              cudaMallocHost(reinterpret_cast<void**>(&o->h_firing_stamp), ncol * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void**>(&o->h_n_firings), sizeof(int)) == cudaSuccess;
    if (!ok)
    {
        cc_ouster_destroy(o);
        return CC_ERR_CUDA;
    }
    *out = o;
    return CC_OK;
}

void cc_ouster_destroy(cc_ouster_t* o)
{
    if (!o)
        return;
    cudaSetDevice(o->device);
    if (o->stream)
    {
        cudaStreamSynchronize(o->stream);
        cudaStreamDestroy(o->stream);
    }
    for (void* p : o->allocs)
        cudaFree(p);
    if (o->h_firing_stamp)
        cudaFreeHost(o->h_firing_stamp);
    if (o->h_n_firings)
        cudaFreeHost(o->h_n_firings);
    delete o;
}

int cc_ouster_packet_size(const cc_ouster_t* o)
{
    return o ? o->packet_size : 0;
}

cc_status_t cc_ouster_set_lut(cc_ouster_t* o, const float* direction, const float* offset)
{
    if (!o || !direction || !offset)
        return CC_ERR_INVALID_ARGUMENT;
    const size_t lut = static_cast<size_t>(o->fmt.columns_per_frame) * o->fmt.pixels_per_column * 3 * sizeof(float);
    if (cudaSetDevice(o->device) != cudaSuccess || cudaMemcpy(o->d_direction, direction, lut, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(o->d_offset, offset, lut, cudaMemcpyHostToDevice) != cudaSuccess)
        return CC_ERR_CUDA;
    o->has_lut = true;
    return CC_OK;
}

cc_status_t cc_ouster_reset(cc_ouster_t* o)
{
    if (!o)
        return CC_ERR_INVALID_ARGUMENT;
    o->firing_index = 0; // SensorInput::reset, sensor_input.hpp:15-19
    o->interrupt = true; // OusterInput::reset, ouster_input.hpp:97-101
    return CC_OK;
}

cc_status_t cc_ouster_decode(cc_ouster_t* o, int n_packets, const uint8_t* packets, const uint64_t* receive_stamps, cc_decoded_firings_t* out)
{
    if (!o || !out || n_packets < 0 || n_packets > o->max_packets || (n_packets > 0 && (!packets || !receive_stamps)) || !o->has_lut)
        return CC_ERR_INVALID_ARGUMENT;
    std::memset(out, 0, sizeof(*out));
    out->rows_per_firing = o->fmt.pixels_per_column;
    out->d_firings = reinterpret_cast<const cc_raw_point_t*>(o->d_firings);
    out->firing_stamps = reinterpret_cast<const uint64_t*>(o->h_firing_stamp);
    out->first_firing_index = o->firing_index;
    // after a reset the packet in flight is cut short: its first valid measurement block is parsed into a firing that
    // is thrown away and the rest of the packet is skipped (ouster_input.hpp:171-180) -- nothing of it is published
    if (o->interrupt && n_packets > 0)
    {
        o->interrupt = false;
        packets += o->packet_size;
        receive_stamps++;
        n_packets--;
    }
    if (n_packets == 0)
        return CC_OK;
    if (cudaSetDevice(o->device) != cudaSuccess)
        return CC_ERR_CUDA;
    cudaStream_t st = o->stream;
    if (cudaMemcpyAsync(o->d_packets, packets, static_cast<size_t>(n_packets) * o->packet_size, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(o->d_recv, receive_stamps, static_cast<size_t>(n_packets) * sizeof(uint64_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return CC_ERR_CUDA;
    CcOusterPtrs p{};
    p.packets = o->d_packets;
    p.n_packets = n_packets;
    p.packet_size = o->packet_size;
    p.receive_stamp = o->d_recv;
    p.direction = o->d_direction;
    p.offset = o->d_offset;
    p.firing_of_column = o->d_firing_of_column;
    p.n_firings = o->d_n_firings;
    p.first_firing_index = o->firing_index;
    p.firings = o->d_firings;
    p.firing_stamp = o->d_firing_stamp;
    const int total = n_packets * o->fmt.columns_per_packet * o->fmt.pixels_per_column;
    CC_LAUNCH(k_ouster_index, 1, 256, 0, st, o->fmt, p);
    CC_LAUNCH(k_ouster_decode, std::max(1, std::min(148 * 8, (total + 255) / 256)), 256, 0, st, o->fmt, p);
    const size_t ncol = static_cast<size_t>(n_packets) * o->fmt.columns_per_packet;
    if (cudaMemcpyAsync(o->h_n_firings, o->d_n_firings, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(o->h_firing_stamp, o->d_firing_stamp, ncol * sizeof(uint64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return CC_ERR_CUDA;
    out->n_firings = *o->h_n_firings;
    o->firing_index += static_cast<unsigned long long>(out->n_firings);
    return CC_OK;
}

cc_status_t cc_ouster_read_firings(cc_ouster_t* o, int n_firings, cc_raw_point_t* firings)
{
    if (!o || n_firings < 0 || (n_firings > 0 && !firings) ||
        static_cast<size_t>(n_firings) > static_cast<size_t>(o->max_packets) * o->fmt.columns_per_packet)
        return CC_ERR_INVALID_ARGUMENT;
    if (n_firings == 0)
        return CC_OK;
    if (cudaSetDevice(o->device) != cudaSuccess ||
        cudaMemcpy(firings, o->d_firings, static_cast<size_t>(n_firings) * o->fmt.pixels_per_column * sizeof(CcRawPoint),
                   cudaMemcpyDeviceToHost) != cudaSuccess)
        return CC_ERR_CUDA;
    return CC_OK;
}

} // extern "C"

extern "C" {

// ---- evaluation metrics on the device (kitti_evaluation.cpp:44-146) ----
struct cc_eval
{
    int device{0};
    int max_points{0};
    cudaStream_t stream{nullptr};
    CcEvalPtrs d{};
    unsigned short* d_sem{nullptr};
    unsigned char* d_ground{nullptr};
    unsigned int* d_gt{nullptr};
    unsigned int* d_det{nullptr};
    unsigned long long* h_counts{nullptr}; // page-locked: 4 counts + 2 doubles
    std::vector<void*> allocs;
};

cc_status_t cc_eval_create(int device_ordinal, int max_points, cc_eval_t** out)
{
    if (!out || max_points <= 0)
        return CC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device_ordinal < 0 || device_ordinal >= ndev)
        return CC_ERR_CUDA;
    cc_eval* e = new cc_eval();
    e->device = device_ordinal;
    e->max_points = max_points;
    int cap = 1024;
    while (cap < 2 * max_points)
        cap <<= 1;
    e->d.pair_cap = cap;
    e->d.marg_cap = cap;
    auto alloc = [&](void** p, size_t bytes) -> bool
    {
        if (cudaMalloc(p, bytes) != cudaSuccess)
            return false;
        e->allocs.push_back(*p);
        return true;
    };
    bool ok = cudaSetDevice(device_ordinal) == cudaSuccess &&
              cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess &&
              alloc(reinterpret_cast<void**>(&e->d_sem), max_points * sizeof(unsigned short)) &&
              alloc(reinterpret_cast<void**>(&e->d_ground), max_points) &&
              alloc(reinterpret_cast<void**>(&e->d_gt), max_points * sizeof(unsigned int)) &&
              alloc(reinterpret_cast<void**>(&e->d_det), max_points * sizeof(unsigned int)) &&
              alloc(reinterpret_cast<void**>(&e->d.pair_keys), cap * sizeof(unsigned long long)) &&
              alloc(reinterpret_cast<void**>(&e->d.pair_cnt), cap * sizeof(unsigned int)) &&
              alloc(reinterpret_cast<void**>(&e->d.marg_keys), cap * sizeof(unsigned long long)) &&
              alloc(reinterpret_cast<void**>(&e->d.marg_cnt), cap * sizeof(unsigned int)) &&
              alloc(reinterpret_cast<void**>(&e->d.counts), 4 * sizeof(unsigned long long) + 2 * sizeof(double)) &&
              cudaMallocHost(reinterpret_cast<void**>(&e->h_counts), 4 * sizeof(unsigned long long) + 2 * sizeof(double)) == cudaSuccess;
    if (!ok)
    {
        cc_eval_destroy(e);
        return CC_ERR_CUDA;
    }
    e->d.entropy = reinterpret_cast<double*>(e->d.counts + 4);
    *out = e;
    return CC_OK;
}

void cc_eval_destroy(cc_eval_t* e)
{
    if (!e)
        return;
    cudaSetDevice(e->device);
    if (e->stream)
    {
        cudaStreamSynchronize(e->stream);
        cudaStreamDestroy(e->stream);
    }
    for (void* p : e->allocs)
        cudaFree(p);
    if (e->h_counts)
        cudaFreeHost(e->h_counts);
    delete e;
}

cc_status_t cc_eval_frame(cc_eval_t* e, int n, const uint16_t* sem, const uint8_t* ground, const uint32_t* gt, const uint32_t* det,
                          cc_eval_result_t* out)
{
    if (!e || !out || n < 0 || n > e->max_points || (n > 0 && (!sem || !ground || !gt || !det)))
        return CC_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(e->device) != cudaSuccess)
        return CC_ERR_CUDA;
    cudaStream_t st = e->stream;
    bool ok = cudaMemcpyAsync(e->d_sem, sem, static_cast<size_t>(n) * sizeof(uint16_t), cudaMemcpyHostToDevice, st) == cudaSuccess &&
              cudaMemcpyAsync(e->d_ground, ground, static_cast<size_t>(n), cudaMemcpyHostToDevice, st) == cudaSuccess &&
              cudaMemcpyAsync(e->d_gt, gt, static_cast<size_t>(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, st) == cudaSuccess &&
              cudaMemcpyAsync(e->d_det, det, static_cast<size_t>(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (!ok)
        return CC_ERR_CUDA;
    CcEvalPtrs d = e->d;
    d.semantic_label = e->d_sem;
    d.is_ground = e->d_ground;
    d.gt_label = e->d_gt;
    d.det_label = e->d_det;
    d.n = n;
    const int g_tab = std::max(1, std::min(148 * 4, (d.pair_cap + 255) / 256));
    const int g_pts = std::max(1, std::min(148 * 8, (n + 255) / 256));
    CC_LAUNCH(k_eval_clear, g_tab, 256, 0, st, d);
    CC_LAUNCH(k_eval_count, g_pts, 256, 0, st, d);
    CC_LAUNCH(k_eval_entropy, g_tab, 256, 0, st, d);
    if (cudaMemcpyAsync(e->h_counts, d.counts, 4 * sizeof(unsigned long long) + 2 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return CC_ERR_CUDA;
    out->tp = static_cast<double>(e->h_counts[0]);
    out->fn = static_cast<double>(e->h_counts[1]);
    out->fp = static_cast<double>(e->h_counts[2]);
    out->tn = static_cast<double>(e->h_counts[3]);
    const double* ent = reinterpret_cast<const double*>(e->h_counts + 4);
    out->over_segmentation_entropy = ent[0];
    out->under_segmentation_entropy = ent[1];
    return CC_OK;
}

cc_status_t cc_selftest_math(int device_ordinal, int op, int n, const float* a, const float* b, float* out)
{
    if (n < 0 || !a || !out || (op == 0 && !b))
        return CC_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device_ordinal) != cudaSuccess)
        return CC_ERR_CUDA;
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    const size_t bytes = static_cast<size_t>(std::max(n, 1)) * sizeof(float);
    cc_status_t rc = CC_OK;
    if (cudaMalloc(reinterpret_cast<void**>(&da), bytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&db), bytes) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&dout), bytes) != cudaSuccess)
        rc = CC_ERR_CUDA;
    if (rc == CC_OK)
    {
        cudaMemcpy(da, a, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice);
        if (b)
            cudaMemcpy(db, b, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice);
        CC_LAUNCH(k_selftest_math, 1024, 256, 0, static_cast<cudaStream_t>(nullptr), op, n, da, db, dout);
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess)
            rc = CC_ERR_CUDA;
        cudaMemcpy(out, dout, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return rc;
}

} // extern "C"
