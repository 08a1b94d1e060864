// cc_kernels.cuh -- the sm_100a kernels of the per-column hot path. No tensor cores: the path has no dense
// contraction; it is streaming scans + a neighbourhood walk + union-find, bound by HBM/L2 latency.
//
// Reference functions replaced ("cpp:" = src/clustering/continuous_clustering.cpp of the reference):
//   k_prep (+ recycling of retired columns), k_scan_lite, k_scan_check, k_insert_scan, k_scatter
//                                        insertFiringIntoRangeImage, clearColumns cpp:105-292, 1094-1145
//   k_gap_scan + k_ground                performGroundPointSegmentationForColumn cpp:294-624
//   k_probe + k_probe_heavy + k_commit_* / k_careful
//                                        associatePointsInColumn, traverseFieldOfView,
//                                        associatePointToPointTree, associatePointTreeToPointTree cpp:638-835
//   k_fin_all + k_fin_label              findFinishedTreesAndAssignSameId        cpp:837-974 and the id / ring
//                                        bookkeeping half of collectPointsForCusterAndPublish cpp:976-1092
// Every kernel starts with CC_PDL_ENTER() (programmatic dependent launch, cc_platform.h) and an optional timeline
// stamp (CcTraceScope, cc_debug_trace).
//
// Everything is compiled with -fmad=false: the reference is built without FMA contraction (CMakeLists.txt:5-12)
// and ground labels / column indices must be bit-identical.
#ifndef CC_KERNELS_CUH
#define CC_KERNELS_CUH

#include <math.h>
#include <string.h>

#ifndef CC_EMU
#include <cuda_pipeline.h>
#endif

#include "cc_math.cuh"
#include "cc_types.h"

#define CC_DEV __device__ __forceinline__

struct CcRawPoint // layout of continuous_clustering::RawPoint (point_types.hpp:10-19) == cc_raw_point_t
{
    float x, y, z;
    unsigned int pad0;
    unsigned long long firing_index;
    unsigned char intensity;
    unsigned char pad1[7];
    unsigned long long stamp;
    unsigned long long guid;
};

// ---- optional device-side timeline (cc_debug_trace): thread 0 of every block stamps its entry (after the grid
//      dependency wait) and exit with the global nanosecond timer; the host reduces them to per-kernel start / end ----
#define CC_TRACE_BLOCKS 2048
#define CC_TRACE_KERNELS 56
enum CcKernelId
{
    CC_KID_prep,
    CC_KID_scan_lite,
    CC_KID_scan_check,
    CC_KID_insert_scan,
    CC_KID_scatter,
    CC_KID_gap_scan,
    CC_KID_ground,
    CC_KID_probe,
    CC_KID_probe_heavy,
    CC_KID_snapshot,
    CC_KID_restore,
    CC_KID_restore_finish,
    CC_KID_commit_copy,
    CC_KID_commit_roots,
    CC_KID_commit_links,
    CC_KID_careful,
    CC_KID_fin_all,
    CC_KID_fin_label,
    CC_KID_clear,
    CC_KID_push_done,
    CC_KID_halt,
    CC_KID_pack_labels,
    CC_KID_state_snapshot,
    CC_KID_ground_main,
    CC_KID_ground_tail,
    CC_KID_gap_main,
    CC_KID_gap_tail,
    CC_KID_fin_init,
    CC_KID_fin_agg,
    CC_KID_fin_decide,
    CC_KID_fin_mark,
    CC_KID_fin_copyback,
    CC_KID_fin_columns,
    CC_KID_lite_p1,
    CC_KID_lite_scan1,
    CC_KID_lite_p2,
    CC_KID_lite_scan2,
    CC_KID_lite_p3,
    CC_KID_check_loop,
    CC_KID_check_scan,
    CC_KID_g_pose,
    CC_KID_g_A,
    CC_KID_g_B,
    CC_KID_g_C,
    CC_KID_g_D,
    CC_KID_push_fused,
    CC_KID_fetch,
    CC_KID_export,
    CC_KID_visited_fix,
    CC_KID_COUNT
};
#define CC_KERNEL_NAMES "k_prep;k_scan_lite;k_scan_check;k_insert_scan;k_scatter;k_gap_scan;k_ground;k_probe;k_probe_heavy;k_snapshot;k_restore;k_restore_finish;k_commit_copy;k_commit_roots;k_commit_links;k_careful;k_fin_all;k_fin_label;k_clear;k_push_done;k_halt;k_pack_labels;k_state_snapshot;.ground_main;.ground_tail;.gap_main;.gap_tail;.fin_init;.fin_agg;.fin_decide;.fin_mark;.fin_copyback;.fin_columns;.lite_p1;.lite_scan1;.lite_p2;.lite_scan2;.lite_p3;.check_loop;.check_scan;.g_pose;.g_A;.g_B;.g_C;.g_D;k_push_fused;.fetch;.export;k_visited_fix"
// Every stage is a device function over a VIRTUAL grid: (bid, nb) is (blockIdx.x, gridDim.x) when the stage runs as its
// own kernel, and (rank of the CTA in its cluster, CTAs per cluster) when the stages of a whole push run inside the
// single fused kernel k_push_fused with cluster barriers between them. blockDim.x / threadIdx.x are always the CTA's own.
struct CcGrid
{
    int bid, nb;
};
CC_DEV CcGrid cc_grid()
{
    CcGrid g;
    g.bid = static_cast<int>(blockIdx.x);
    g.nb = static_cast<int>(gridDim.x);
    return g;
}

struct CcTraceScope
{
    unsigned long long* slot;
    __device__ __forceinline__ CcTraceScope(unsigned long long* trace, int kid, int bid) : slot(nullptr)
    {
#ifndef CC_EMU
        if (trace && threadIdx.x == 0)
        {
            const unsigned int b = bid < CC_TRACE_BLOCKS ? bid : CC_TRACE_BLOCKS - 1;
            slot = trace + (static_cast<size_t>(kid) * CC_TRACE_BLOCKS + b) * 2;
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            slot[0] = t;
        }
#endif
    }
    __device__ __forceinline__ void stop()
    {
#ifndef CC_EMU
        if (slot)
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            slot[1] = t;
            slot = nullptr;
        }
#endif
    }
    __device__ __forceinline__ ~CcTraceScope()
    {
#ifndef CC_EMU
        if (slot)
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            slot[1] = t;
        }
#endif
    }
};

// The fields almost every kernel needs before it can start, loaded together (one memory round trip instead of one per
// dependent branch). Stable for the duration of the kernels that use it.
struct CcHead
{
    int halted, ncols, n_flagged, abort, error;
    long long colbase;
};
CC_DEV CcHead cc_head(const CcDevState* st)
{
    CcHead h;
    h.halted = st->halted;
    h.ncols = st->ncols;
    h.n_flagged = st->n_flagged;
    h.abort = st->abort;
    h.error = st->error;
    h.colbase = st->colbase;
    return h;
}
CC_DEV bool cc_head_ok(const CcHead& h, int guard) // cc_spec_ok on the loaded fields
{
    if (h.ncols <= 0)
        return false;
    if (guard == 1)
        return h.n_flagged == 0 && h.abort == 0 && h.error == 0;
    if (guard == 2)
        return h.abort == 0;
    return true;
}

CC_DEV float cc_nanf()
{
    return ccm::u2f(0x7fc00000u);
}
CC_DEV bool cc_isnan(float x)
{
    return x != x;
}
CC_DEV unsigned long long cc_d2ord(double d) // order-preserving for non-negative doubles
{
    return static_cast<unsigned long long>(__double_as_longlong(d));
}
CC_DEV double cc_ord2d(unsigned long long u)
{
    return __longlong_as_double(static_cast<long long>(u));
}
CC_DEV unsigned int cc_vload(const unsigned int* p)
{
    return *reinterpret_cast<const volatile unsigned int*>(p);
}

CC_DEV int cc_warp_min(int v)
{
    return __reduce_min_sync(CC_FULL_MASK, v);
}
CC_DEV int cc_warp_max(int v)
{
    return __reduce_max_sync(CC_FULL_MASK, v);
}
CC_DEV double cc_warp_min_f64(double v)
{
    for (int o = CC_WARP / 2; o > 0; o >>= 1)
    {
        double w = __shfl_xor_sync(CC_FULL_MASK, v, o);
        v = (w < v) ? w : v;
    }
    return v;
}

// rigid transform helpers, double, evaluation order of the Eigen stand-in the oracle is built with
CC_DEV void cc_iso_apply(const double* m, double x, double y, double z, double* out)
{
    for (int i = 0; i < 3; i++)
        out[i] = ((m[i * 4 + 0] * x + m[i * 4 + 1] * y) + m[i * 4 + 2] * z) + m[i * 4 + 3];
}
CC_DEV void cc_iso_inverse(const double* m, double* r)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r[i * 4 + j] = m[j * 4 + i];
    for (int i = 0; i < 3; i++)
        r[i * 4 + 3] = -(r[i * 4 + 0] * m[3] + (r[i * 4 + 1] * m[7] + r[i * 4 + 2] * m[11]));
}
CC_DEV void cc_iso_mul(const double* a, const double* b, double* r)
{
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++)
            r[i * 4 + j] = a[i * 4 + 0] * b[j] + (a[i * 4 + 1] * b[4 + j] + a[i * 4 + 2] * b[8 + j]);
        r[i * 4 + 3] = (a[i * 4 + 0] * b[3] + (a[i * 4 + 1] * b[7] + a[i * 4 + 2] * b[11])) + a[i * 4 + 3];
    }
}

// exclusive block scan (Hillis-Steele in shared memory) with an ordered combine op(earlier, later)
struct CcOpLastValid
{
    CC_DEV float operator()(float a, float b) const { return b != b ? a : b; }
};
struct CcOpMaxF64
{
    CC_DEV double operator()(double a, double b) const { return a > b ? a : b; }
};
struct CcOpMaxI32
{
    CC_DEV int operator()(int a, int b) const { return a > b ? a : b; }
};
struct CcOpMaxI64
{
    CC_DEV long long operator()(long long a, long long b) const { return a > b ? a : b; }
};
struct CcOpAddI32
{
    CC_DEV int operator()(int a, int b) const { return a + b; }
};
struct CcOpAddI64
{
    CC_DEV long long operator()(long long a, long long b) const { return a + b; }
};
template<typename T>
CC_DEV T cc_shfl_up_any(T v, int off) // shuffle of a plain struct, word by word
{
    static_assert(sizeof(T) % 4 == 0, "word-sized types only");
    unsigned int w[sizeof(T) / 4];
    memcpy(w, &v, sizeof(T));
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(T) / 4); i++)
        w[i] = __shfl_up_sync(CC_FULL_MASK, w[i], off);
    memcpy(&v, w, sizeof(T));
    return v;
}
template<typename T>
CC_DEV T cc_shfl_any(T v, int src) // same for a broadcast from one lane
{
    static_assert(sizeof(T) % 4 == 0, "word-sized types only");
    unsigned int w[sizeof(T) / 4];
    memcpy(w, &v, sizeof(T));
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(T) / 4); i++)
        w[i] = __shfl_sync(CC_FULL_MASK, w[i], src);
    memcpy(&v, w, sizeof(T));
    return v;
}
// Two-level: shuffle scan inside every warp, the warp totals scanned by warp 0 (two block barriers in all; the block
// has at most 32 warps). `sm` needs room for one T per warp. op(identity, x) == x and op(x, identity) == x.
template<typename T, typename Op>
CC_DEV T cc_block_exclusive_scan(T* sm, T v, T identity, Op op)
{
    const int t = threadIdx.x, lane = t % CC_WARP, warp = t / CC_WARP;
    const int nw = (blockDim.x + CC_WARP - 1) / CC_WARP;
    T x = v;
    for (int off = 1; off < CC_WARP; off <<= 1)
    {
        const T y = cc_shfl_up_any(x, off);
        if (lane >= off)
            x = op(y, x);
    }
    if (lane == CC_WARP - 1)
        sm[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
        T xi = lane < nw ? sm[lane] : identity;
        for (int off = 1; off < CC_WARP; off <<= 1)
        {
            const T y = cc_shfl_up_any(xi, off);
            if (lane >= off)
                xi = op(y, xi);
        }
        T ex = cc_shfl_up_any(xi, 1);
        if (lane == 0)
            ex = identity;
        if (lane < nw)
            sm[lane] = ex;
    }
    __syncthreads();
    const T warp_prefix = sm[warp];
    T e = cc_shfl_up_any(x, 1);
    if (lane == 0)
        e = identity;
    const T excl = op(warp_prefix, e);
    __syncthreads();
    return excl;
}

// finish passes run at the columns that are multiples of cluster_point_trees_every_nth_column (cpp:841)
CC_DEV long long cc_pass_at_or_after(long long c, int nth)
{
    return nth <= 1 ? c : ((c + nth - 1) / nth) * nth;
}
CC_DEV long long cc_pass_at_or_before(long long c, int nth)
{
    return nth <= 1 ? c : c - (c % nth);
}

CC_DEV int cc_local_col(long long g, int ringcols)
{
    return static_cast<int>(g % ringcols);
}

// =====================================================================================================
// K0  per raw point: rigid transform, azimuth -> column within rotation, distance, inclination (cpp:125-151,
//     189, 232). Fully parallel, one thread per (firing, row).
// =====================================================================================================
CC_DEV int cc_wrapdiff(int d, int N) // representative of d (mod N) in (-N/2, N/2]
{
    if (2 * d > N)
        d -= N;
    else if (2 * d <= -N)
        d += N;
    return d;
}

CC_DEV void d_clear(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, long long from, long long to);

CC_DEV void d_prep(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, int n_firings)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_prep, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    // K5 first: clearColumns (cpp:1094-1145) for the columns retired TWO pushes ago. Recycling is deferred so that every
    // column a push reports through a finished-column event can still be read by the caller after the push has been
    // waited for, even while the next push is already in flight (the reference's callbacks read range_image_ before
    // clearColumns runs, cpp:1087-1091). Pure stores: they drain while the firings below are being prepared.
    d_clear(g, cfg, p, p.st->clear2_from, p.st->clear2_to);
    if (g.bid == 0 && threadIdx.x == 0)
        p.st->scan_kbad = 0x7fffffff; // lowered by the scan stages to the first firing that is not regular
    // one warp per firing, lanes over rows: per-point staging + the firing's summary for the lite insertion path
    // (anchor = column-in-rotation of its first valid row; rearmost / foremost column relative to the anchor)
    const int R = cfg.R;
    const float pi_f = static_cast<float>(M_PI);
    const int lane = threadIdx.x % CC_WARP;
    const int gwarp = (g.bid * blockDim.x + threadIdx.x) / CC_WARP;
    const int nw = (g.nb * blockDim.x + CC_WARP - 1) / CC_WARP;
    for (int k = gwarp; k < n_firings; k += nw)
    {
        const double* pose = p.poses + 12 * k;
        unsigned int kmin = 0xffffffffu;
        int nvalid = 0, weird = 0;
        for (int row = lane; row < R; row += CC_WARP)
        {
            const int idx = k * R + row;
            const CcRawPoint* raw = reinterpret_cast<const CcRawPoint*>(p.raw) + idx;
            const float fx = raw->x, fy = raw->y, fz = raw->z;
            p.o_g[idx] = -0x7fffffff - 1;
            if (cc_isnan(fx))
            {
                p.s_cwr[idx] = CC_INVALID_CWR;
                p.s_cwrT[static_cast<size_t>(row) * p.max_firings + k] = CC_INVALID_CWR;
                continue;
            }
            double po[3];
            cc_iso_apply(pose, static_cast<double>(fx), static_cast<double>(fy), static_cast<double>(fz), po);
            const double rx = po[0] - pose[3], ry = po[1] - pose[7], rz = po[2] - pose[11];
            const float az = ccm::atan2f_glibc(fy, fx); // sensor-frame azimuth, cpp:142
            const float incaz = cfg.clockwise ? -az + pi_f : az + pi_f;
            const int cwr = static_cast<int>(ccm::div_rn(incaz, cfg.width));
            const float dist = static_cast<float>(sqrt(rx * rx + (ry * ry + rz * rz)));
            p.s_pos[idx] = make_float4(static_cast<float>(po[0]), static_cast<float>(po[1]), static_cast<float>(po[2]), dist);
            p.s_dist[idx] = dist;
            p.s_az[idx] = az;
            p.s_incaz[idx] = incaz;
            p.s_incl[idx] = ccm::asinf_glibc(ccm::div_rn(static_cast<float>(rz), dist));
            p.s_cwr[idx] = cwr;
            p.s_cwrT[static_cast<size_t>(row) * p.max_firings + k] = cwr; // row-major copy for the per-row check
            if (cwr != CC_INVALID_CWR)
            {
                nvalid++;
                if (cwr < 0 || cwr > cfg.N || cwr >= (1 << 20))
                    weird = 1;
                const unsigned int key = (static_cast<unsigned int>(row) << 20) | (static_cast<unsigned int>(cwr) & 0xfffffu);
                kmin = key < kmin ? key : kmin;
            }
        }
        kmin = __reduce_min_sync(CC_FULL_MASK, kmin);
        nvalid = __reduce_add_sync(CC_FULL_MASK, nvalid);
        weird = __reduce_or_sync(CC_FULL_MASK, weird);
        int rear = 0, fore = 0;
        const int anchor = static_cast<int>(kmin & 0xfffffu);
        if (nvalid > 0 && !weird)
        {
            int lmin = 0x7fffffff, lmax = -0x7fffffff - 1;
            for (int row = lane; row < R; row += CC_WARP)
            {
                const int cw = p.s_cwr[k * R + row]; // written by this very lane above
                if (cw != CC_INVALID_CWR)
                {
                    const int d = cc_wrapdiff(cw - anchor, cfg.N);
                    lmin = d < lmin ? d : lmin;
                    lmax = d > lmax ? d : lmax;
                }
            }
            rear = cc_warp_min(lmin);
            fore = cc_warp_max(lmax);
        }
        // robot_from_sensor * odom_from_sensor^-1 of the firing (the ego-box transform of the columns it completes,
        // cpp:300-301), one element per lane in the evaluation order of cc_iso_inverse + cc_iso_mul
        for (int e = lane; e < 12; e += CC_WARP)
        {
            const int i = e / 4, j = e % 4;
            const double* ra = cfg.robot_from_sensor + i * 4;
            double v;
            if (j < 3)
                v = ra[0] * pose[j * 4 + 0] + (ra[1] * pose[j * 4 + 1] + ra[2] * pose[j * 4 + 2]);
            else
            {
                double tr[3];
                for (int q = 0; q < 3; q++)
                    tr[q] = -(pose[q] * pose[3] + (pose[4 + q] * pose[7] + pose[8 + q] * pose[11]));
                v = (ra[0] * tr[0] + (ra[1] * tr[1] + ra[2] * tr[2])) + ra[3];
            }
            p.s_ego[static_cast<size_t>(k) * 12 + e] = v;
        }
        if (lane == 0)
        {
            CcFiringSummary fs;
            fs.anchor = anchor;
            fs.rear_rel = rear;
            fs.fore_rel = fore;
            fs.nvalid = weird ? -1 : nvalid; // -1: the firing must go through the per-firing path
            p.lite_sum[k] = fs;
        }
    }
}
__global__ void k_prep(CcDevCfg cfg, CcDevPtrs p, int n_firings)
{
    CC_PDL_ENTER();
    d_prep(cc_grid(), cfg, p, n_firings);
}

// =====================================================================================================
// K1-lite  the insertion scan when every firing of (a prefix of) the push is REGULAR (see k_insert_scan), done
//     grid-wide instead of on one SM:
//     k_scan_lite  (1 warp)      unwrapped column of every firing's anchor = prefix sum of wrapped anchor deltas;
//                                rearmost column before every firing = prefix maximum; straddle / margin checks
//     k_scan_check (block / row) per row the columns are strictly increasing and beyond the row's front
//     (k_insert_scan then applies the new row fronts and, if firings remain, the distance write-through)
//     The first irregular firing (scan_kbad) is exact; k_insert_scan then commits the prefix and processes the rest.
// =====================================================================================================
struct CcAnchorSeg // a run of firings: first / last valid anchor and the unwrapped distance between them
{
    int has, first_cw, last_cw, off;
};
struct CcOpAnchorSeg
{
    int N;
    CC_DEV CcAnchorSeg operator()(const CcAnchorSeg& a, const CcAnchorSeg& b) const
    {
        if (!a.has)
            return b;
        if (!b.has)
            return a;
        CcAnchorSeg r;
        r.has = 1;
        r.first_cw = a.first_cw;
        r.last_cw = b.last_cw;
        r.off = a.off + cc_wrapdiff(b.first_cw - a.last_cw, N) + b.off;
        return r;
    }
};
struct CcMaxPair
{
    int p, f;
};
struct CcOpMaxPair
{
    CC_DEV CcMaxPair operator()(const CcMaxPair& a, const CcMaxPair& b) const
    {
        CcMaxPair r;
        r.p = a.p > b.p ? a.p : b.p;
        r.f = a.f > b.f ? a.f : b.f;
        return r;
    }
};

CC_DEV int cc_lite_slot(int k) // one slot of padding per 8 firings
{
    return k + (k >> 3);
}
static inline __host__ __device__ size_t cc_lite_smem_bytes(int n)
{
    return 1024 + (static_cast<size_t>(n) + (n >> 3) + 8) * sizeof(CcFiringSummary);
}
#ifdef CC_EMU
#define CC_LITE_PER 8192 /* the emulation runs the block as one thread */
#else
#define CC_LITE_PER 8 /* firings per thread of k_scan_lite: 8192 firings per push with 1024 threads */
#endif

CC_DEV void d_scan_lite(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, int n)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_scan_lite, g.bid);
    if (g.bid != 0 || p.st->halted)
        return;
    CC_SMEM(smem);
    // every thread owns a few consecutive firings; the cross-thread parts are block scans
    CcAnchorSeg* sm_seg = reinterpret_cast<CcAnchorSeg*>(smem);
    CcMaxPair* sm_max = reinterpret_cast<CcMaxPair*>(smem);
    __shared__ int sh_kbad, sh_first_cw, sh_has;
    CcDevState* st = p.st;
    const int T = blockDim.x, t = threadIdx.x;
    const int N = cfg.N, half = cfg.half;
    const int NOT_SET = -0x7fffffff - 1;
    const long long base = st->P;
    const bool ok_state = st->F >= 0 && st->foremost >= 0 && base > 0 && st->ring_start != -1;
    const int pc0 = static_cast<int>(base % N);
    const int Fm0 = static_cast<int>(st->foremost - base);
    const int colbase_rel = static_cast<int>(st->F - base);
    const bool ok = ok_state && (n + T - 1) / T <= CC_LITE_PER;
    if (t == 0)
        sh_kbad = ok ? n : 0;
    __syncthreads();
    if (!ok)
    {
        if (t == 0)
        {
            st->scan_kbad = 0;
            st->scan_lite_base = base;
        }
        return;
    }
    const int per = (n + T - 1) / T;
    const int a = t * per < n ? t * per : n, b = (a + per < n) ? a + per : n;
    int kbad = n;
    // The per-firing summaries come in and the results go out through shared memory: a thread owns CONSECUTIVE firings
    // (128 bytes apart per thread: one memory sector per thread and access), so the global accesses are done by all
    // threads side by side instead (slot padding keeps the per-thread accesses off the same banks).
    CcFiringSummary* stage = reinterpret_cast<CcFiringSummary*>(smem + 1024);
    for (int k0 = 0; k0 < n; k0 += 8 * T) // the loads of a thread's eight firings are issued together: one round trip
    {
        CcFiringSummary v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (k0 + u * T + t < n)
                v[u] = p.lite_sum[k0 + u * T + t];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (k0 + u * T + t < n)
                stage[cc_lite_slot(k0 + u * T + t)] = v[u];
    }
    __syncthreads();
    // every thread keeps its (at most CC_LITE_PER) firings in registers across the three passes
    CcFiringSummary fs[CC_LITE_PER];
    int U[CC_LITE_PER], lP[CC_LITE_PER], lF[CC_LITE_PER];
#pragma unroll
    for (int u = 0; u < CC_LITE_PER; u++)
    {
        fs[u].anchor = fs[u].rear_rel = fs[u].fore_rel = fs[u].nvalid = 0;
        if (a + u < b)
            fs[u] = stage[cc_lite_slot(a + u)]; // staged with coalesced loads above
    }
    CcTraceScope cc_tr_l1(p.trace, CC_KID_lite_p1, g.bid);
    // pass 1: anchor columns relative to the thread's first valid firing
    CcAnchorSeg mine;
    mine.has = 0;
    mine.first_cw = 0;
    mine.last_cw = 0;
    mine.off = 0;
#pragma unroll
    for (int u = 0; u < CC_LITE_PER; u++)
    {
        if (a + u < b)
        {
            if (fs[u].nvalid < 0)
                kbad = a + u < kbad ? a + u : kbad;
            if (fs[u].nvalid > 0)
            {
                if (!mine.has)
                {
                    mine.has = 1;
                    mine.first_cw = fs[u].anchor;
                }
                else
                    mine.off += cc_wrapdiff(fs[u].anchor - mine.last_cw, N);
                mine.last_cw = fs[u].anchor;
            }
        }
        U[u] = mine.off;
    }
    CcAnchorSeg ident;
    ident.has = 0;
    ident.first_cw = ident.last_cw = ident.off = 0;
    CcOpAnchorSeg op;
    op.N = N;
    cc_tr_l1.stop();
    CcTraceScope cc_tr_s1(p.trace, CC_KID_lite_scan1, g.bid);
    const CcAnchorSeg before = cc_block_exclusive_scan(sm_seg, mine, ident, op);
    cc_tr_s1.stop();
    CcTraceScope cc_tr_l2(p.trace, CC_KID_lite_p2, g.bid);
    if (t == T - 1)
    {
        const CcAnchorSeg all = op(before, mine);
        sh_has = all.has;
        sh_first_cw = all.first_cw;
    }
    __syncthreads();
    // the reference's unwrap of the push's first valid firing against the rearmost column (cpp:152-175)
    int U0 = 0;
    if (sh_has)
    {
        const int cw = sh_first_cw, diff = cw - pc0;
        U0 = -pc0 + cw;
        if (diff < -half)
            U0 += N;
        else if (diff > half)
            U0 -= N;
    }
    const int mybase = U0 + (before.has && mine.has ? before.off + cc_wrapdiff(mine.first_cw - before.last_cw, N) : 0);
    // pass 2: absolute anchors, per-firing rearmost / foremost, thread-local exclusive prefix maxima
    CcMaxPair run;
    run.p = NOT_SET;
    run.f = NOT_SET;
#pragma unroll
    for (int u = 0; u < CC_LITE_PER; u++)
    {
        lP[u] = run.p;
        lF[u] = run.f;
        if (a + u < b && fs[u].nvalid > 0)
        {
            U[u] = mybase + U[u];
            const int rear = U[u] + fs[u].rear_rel, fore = U[u] + fs[u].fore_rel;
            if (fore - rear > N / 2) // cpp:252-261
                kbad = a + u < kbad ? a + u : kbad;
            run.p = rear > run.p ? rear : run.p;
            run.f = fore > run.f ? fore : run.f;
        }
    }
    CcMaxPair seed;
    seed.p = 0;   // rearmost column at the start of the push (relative: 0)
    seed.f = Fm0; // foremost column at the start of the push
    cc_tr_l2.stop();
    CcTraceScope cc_tr_s2(p.trace, CC_KID_lite_scan2, g.bid);
    CcMaxPair pre = cc_block_exclusive_scan(sm_max, run, seed, CcOpMaxPair());
    cc_tr_s2.stop();
    CcTraceScope cc_tr_l3(p.trace, CC_KID_lite_p3, g.bid);
    pre = CcOpMaxPair()(pre, seed);
    // pass 3: rearmost / foremost so far before every firing; unwrap margins
#pragma unroll
    for (int u = 0; u < CC_LITE_PER; u++)
    {
        if (a + u < b)
        {
            const int k = a + u;
            const int Pk = lP[u] > pre.p ? lP[u] : pre.p, Fk = lF[u] > pre.f ? lF[u] : pre.f;
            CcFiringSummary o; // (rearmost so far, foremost so far, unwrapped anchor) of firing k, copied out below
            o.anchor = Pk;
            o.rear_rel = Fk;
            o.fore_rel = U[u];
            o.nvalid = 0;
            stage[cc_lite_slot(k)] = o;
            if (fs[u].nvalid > 0)
            {
                const int rear = U[u] + fs[u].rear_rel, fore = U[u] + fs[u].fore_rel;
                if (!(rear - Pk > -half && fore - Pk < half))
                    kbad = k < kbad ? k : kbad; // the unwrap of some point could differ from the reference's
                const int Pnext = rear > Pk ? rear : Pk;
                if (Pnext - colbase_rel > p.maxcols)
                    kbad = k < kbad ? k : kbad; // the per-firing path raises the error
            }
        }
    }
    if (a < n && b == n)
    {
        p.lite_P[n] = run.p > pre.p ? run.p : pre.p;
        p.lite_F[n] = run.f > pre.f ? run.f : pre.f;
    }
    if (kbad < n)
        atomicMin(&sh_kbad, kbad);
    __syncthreads();
    for (int k = t; k < n; k += T)
    {
        const CcFiringSummary o = stage[cc_lite_slot(k)];
        p.lite_P[k] = o.anchor;
        p.lite_F[k] = o.rear_rel;
        p.lite_U[k] = o.fore_rel;
    }
    if (t == 0)
    {
        st->scan_kbad = sh_kbad;
        st->scan_lite_base = base;
    }
}
__global__ void __launch_bounds__(1024) k_scan_lite(CcDevCfg cfg, CcDevPtrs p, int n)
{
    CC_PDL_ENTER();
    d_scan_lite(cc_grid(), cfg, p, n);
}

struct CcOpLastSetI32
{
    CC_DEV int operator()(int a, int b) const { return b == (-0x7fffffff - 1) ? a : b; }
};

CC_DEV void d_scan_check(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, int n)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_scan_check, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    CC_SMEM(smem);
    int* sm = reinterpret_cast<int*>(smem);
    const int R = cfg.R, N = cfg.N;
    const CcDevState* st = p.st;
    const int kmax0 = st->scan_kbad < n ? st->scan_kbad : n; // read before any row of this grid lowers it
    const int NOT_SET = -0x7fffffff - 1;
    const int T = blockDim.x, t = threadIdx.x;
    __shared__ int sh_front;
  for (int row = g.bid; row < R; row += g.nb) // one CTA per row (the fused kernel has fewer CTAs than rows)
  {
    const int kmax = kmax0;
    const int seg = (kmax + T - 1) / T;
    const int a = t * seg < kmax ? t * seg : kmax, b = (a + seg < kmax) ? a + seg : kmax;
    int first = NOT_SET, firstk = kmax, last = NOT_SET, kbad = n, gmax = NOT_SET;
    __syncthreads(); // the previous row's sh_front has been read
    if (t == 0)
        sh_front = NOT_SET;
    __syncthreads();
    CcTraceScope cc_tr_cl(p.trace, CC_KID_check_loop, g.bid);
    for (int kb = a; kb < b; kb += 8) // loads of 8 firings issued together
    {
        int cw[8], U[8], an[8], P[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            const int k = kb + u;
            cw[u] = CC_INVALID_CWR;
            U[u] = an[u] = P[u] = 0;
            if (k < b)
            {
                cw[u] = p.s_cwrT[static_cast<size_t>(row) * p.max_firings + k];
                U[u] = p.lite_U[k];
                an[u] = p.lite_sum[k].anchor;
                P[u] = p.lite_P[k];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
            const int k = kb + u;
            if (cw[u] == CC_INVALID_CWR)
                continue;
            const int g = U[u] + cc_wrapdiff(cw[u] - an[u], N);
            if (first == NOT_SET)
            {
                first = g;
                firstk = k;
            }
            else if (g <= last)
                kbad = k < kbad ? k : kbad;
            last = g;
            if (g >= P[u]) // stored (not too far behind, cpp:210-221): candidate for the row's new front
                gmax = g > gmax ? g : gmax;
        }
    }
    cc_tr_cl.stop();
    CcTraceScope cc_tr_cs(p.trace, CC_KID_check_scan, g.bid);
    int prev = cc_block_exclusive_scan(sm, last, NOT_SET, CcOpLastSetI32());
    cc_tr_cs.stop();
    if (prev == NOT_SET)
    {
        long long rel = p.rowmax[row] - st->scan_lite_base; // the row's front
        prev = rel < -0x3fffffff ? -0x3fffffff : static_cast<int>(rel);
    }
    if (first != NOT_SET && first <= prev)
        kbad = firstk < kbad ? firstk : kbad;
    if (kbad < n)
        atomicMin(&p.st->scan_kbad, kbad);
    // the row's front after the lite prefix, valid when the whole prefix [0, kmax) survives every row's check
    gmax = cc_warp_max(gmax);
    if ((t % CC_WARP) == 0 && gmax != NOT_SET)
        atomicMax(&sh_front, gmax);
    __syncthreads();
    if (t == 0)
        p.lite_rowfront[row] = sh_front;
  }
}
__global__ void k_scan_check(CcDevCfg cfg, CcDevPtrs p, int n)
{
    CC_PDL_ENTER();
    d_scan_check(cc_grid(), cfg, p, n);
}

// ---- the two stages above for SHORT pushes, in one phase of the fused kernel: every CTA repeats the (tiny) per-firing
//      scan in its warp 0 with the per-firing values in shared memory -- nobody waits for one CTA to publish them -- and
//      then checks its rows, one warp per row. Pushes of up to CC_WARP * CC_SMALL_PER firings. ----
#ifdef CC_EMU
#define CC_SMALL_PER 8192 /* the emulation runs a warp as one lane */
#else
#define CC_SMALL_PER 16
#endif
template<typename T, typename Op>
CC_DEV T cc_warp_exclusive_scan(T v, T identity, Op op, int lane)
{
    T x = v;
    for (int off = 1; off < CC_WARP; off <<= 1)
    {
        const T y = cc_shfl_up_any(x, off);
        if (lane >= off)
            x = op(y, x);
    }
    T e = cc_shfl_up_any(x, 1);
    if (lane == 0)
        e = identity;
    return e;
}
static inline __host__ __device__ size_t cc_lite_small_smem_bytes(int n)
{
    return static_cast<size_t>(n) * (sizeof(CcFiringSummary) + 3 * sizeof(int)) + 16;
}

CC_DEV void d_lite_check_small(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, int n)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_scan_lite, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    CC_SMEM(smem);
    CcFiringSummary* fs = reinterpret_cast<CcFiringSummary*>(smem);
    int* smU = reinterpret_cast<int*>(fs + n);
    int* smP = smU + n;     // [n + 1]
    int* smF = smP + n + 1; // [n + 1]
    __shared__ int sh_kbad;
    CcDevState* st = p.st;
    const int T = blockDim.x, t = threadIdx.x, lane = t % CC_WARP, warp = t / CC_WARP;
    const int nwarps = (T + CC_WARP - 1) / CC_WARP;
    const int R = cfg.R, N = cfg.N, half = cfg.half;
    const int NOT_SET = -0x7fffffff - 1;
    const long long base = st->P;
    const bool ok_state = st->F >= 0 && st->foremost >= 0 && base > 0 && st->ring_start != -1;
    const int pc0 = static_cast<int>(base % N);
    const int Fm0 = static_cast<int>(st->foremost - base);
    const int colbase_rel = static_cast<int>(st->F - base);
    const bool ok = ok_state && n <= CC_WARP * CC_SMALL_PER;
    for (int k = t; k < n; k += T)
        fs[k] = p.lite_sum[k];
    __syncthreads();
    if (warp == 0)
    {
        int kbad = ok ? n : 0;
        if (ok)
        {
            // same three passes as d_scan_lite, one lane per run of consecutive firings
            const int per = (n + CC_WARP - 1) / CC_WARP;
            const int a = lane * per < n ? lane * per : n, b = (a + per < n) ? a + per : n;
            CcAnchorSeg mine;
            mine.has = 0;
            mine.first_cw = mine.last_cw = mine.off = 0;
            for (int k = a; k < b; k++)
            {
                if (fs[k].nvalid < 0)
                    kbad = k < kbad ? k : kbad;
                if (fs[k].nvalid > 0)
                {
                    if (!mine.has)
                    {
                        mine.has = 1;
                        mine.first_cw = fs[k].anchor;
                    }
                    else
                        mine.off += cc_wrapdiff(fs[k].anchor - mine.last_cw, N);
                    mine.last_cw = fs[k].anchor;
                }
                smU[k] = mine.off;
            }
            CcAnchorSeg ident;
            ident.has = 0;
            ident.first_cw = ident.last_cw = ident.off = 0;
            CcOpAnchorSeg op;
            op.N = N;
            const CcAnchorSeg before = cc_warp_exclusive_scan(mine, ident, op, lane);
            CcAnchorSeg all = op(before, mine);
            all = cc_shfl_any(all, CC_WARP - 1);
            int U0 = 0; // the reference's unwrap of the push's first valid firing against the rearmost column (cpp:152-175)
            if (all.has)
            {
                const int cw = all.first_cw, diff = cw - pc0;
                U0 = -pc0 + cw;
                if (diff < -half)
                    U0 += N;
                else if (diff > half)
                    U0 -= N;
            }
            const int mybase = U0 + (before.has && mine.has ? before.off + cc_wrapdiff(mine.first_cw - before.last_cw, N) : 0);
            CcMaxPair run;
            run.p = NOT_SET;
            run.f = NOT_SET;
            for (int k = a; k < b; k++)
            {
                smP[k] = run.p; // exclusive running maxima inside the lane's run
                smF[k] = run.f;
                if (fs[k].nvalid > 0)
                {
                    const int U = mybase + smU[k];
                    smU[k] = U;
                    const int rear = U + fs[k].rear_rel, fore = U + fs[k].fore_rel;
                    if (fore - rear > N / 2) // cpp:252-261
                        kbad = k < kbad ? k : kbad;
                    run.p = rear > run.p ? rear : run.p;
                    run.f = fore > run.f ? fore : run.f;
                }
            }
            CcMaxPair seed;
            seed.p = 0;   // rearmost column at the start of the push (relative: 0)
            seed.f = Fm0; // foremost column at the start of the push
            CcMaxPair pre = cc_warp_exclusive_scan(run, seed, CcOpMaxPair(), lane);
            pre = CcOpMaxPair()(pre, seed);
            for (int k = a; k < b; k++)
            {
                const int Pk = smP[k] > pre.p ? smP[k] : pre.p, Fk = smF[k] > pre.f ? smF[k] : pre.f;
                smP[k] = Pk;
                smF[k] = Fk;
                if (fs[k].nvalid > 0)
                {
                    const int rear = smU[k] + fs[k].rear_rel, fore = smU[k] + fs[k].fore_rel;
                    if (!(rear - Pk > -half && fore - Pk < half))
                        kbad = k < kbad ? k : kbad; // the unwrap of some point could differ from the reference's
                    const int Pnext = rear > Pk ? rear : Pk;
                    if (Pnext - colbase_rel > p.maxcols)
                        kbad = k < kbad ? k : kbad; // the per-firing path raises the error
                }
                if (g.bid == 0) // one copy for the stages that follow (scatter, insertion commit)
                {
                    p.lite_P[k] = Pk;
                    p.lite_F[k] = Fk;
                    p.lite_U[k] = smU[k];
                }
            }
            if (a < n && b == n)
            {
                smP[n] = run.p > pre.p ? run.p : pre.p;
                smF[n] = run.f > pre.f ? run.f : pre.f;
                if (g.bid == 0)
                {
                    p.lite_P[n] = smP[n];
                    p.lite_F[n] = smF[n];
                }
            }
            kbad = cc_warp_min(kbad);
        }
        if (lane == 0)
        {
            sh_kbad = kbad;
            if (g.bid == 0)
                st->scan_lite_base = base;
        }
    }
    __syncthreads();
    const int kmax = sh_kbad;
    int kbad_all = warp == 0 ? kmax : n;
    // ---- per row: columns strictly increasing and beyond the row's front (d_scan_check), one warp per row ----
    for (int row = g.bid * nwarps + warp; row < R; row += g.nb * nwarps)
    {
        const int seg = (kmax + CC_WARP - 1) / CC_WARP;
        const int a = lane * seg < kmax ? lane * seg : kmax, b = (a + seg < kmax) ? a + seg : kmax;
        int first = NOT_SET, firstk = kmax, last = NOT_SET, kbad = n, gmax = NOT_SET;
        const int* cwT = p.s_cwrT + static_cast<size_t>(row) * p.max_firings;
        for (int kb = a; kb < b; kb += 8) // loads of 8 firings issued together
        {
            int cw[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                cw[u] = kb + u < b ? cwT[kb + u] : CC_INVALID_CWR;
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                const int k = kb + u;
                if (cw[u] == CC_INVALID_CWR)
                    continue;
                const int gc = smU[k] + cc_wrapdiff(cw[u] - fs[k].anchor, N);
                if (first == NOT_SET)
                {
                    first = gc;
                    firstk = k;
                }
                else if (gc <= last)
                    kbad = k < kbad ? k : kbad;
                last = gc;
                if (gc >= smP[k]) // stored (not too far behind, cpp:210-221): candidate for the row's new front
                    gmax = gc > gmax ? gc : gmax;
            }
        }
        int prev = cc_warp_exclusive_scan(last, NOT_SET, CcOpLastSetI32(), lane);
        if (prev == NOT_SET)
        {
            long long rel = p.rowmax[row] - base; // the row's front
            prev = rel < -0x3fffffff ? -0x3fffffff : static_cast<int>(rel);
        }
        if (first != NOT_SET && first <= prev)
            kbad = firstk < kbad ? firstk : kbad;
        kbad = cc_warp_min(kbad);
        gmax = cc_warp_max(gmax);
        if (lane == 0)
            p.lite_rowfront[row] = gmax;
        kbad_all = kbad < kbad_all ? kbad : kbad_all;
    }
    if (lane == 0 && kbad_all < 0x7fffffff)
        atomicMin(&st->scan_kbad, kbad_all); // starts out at INT_MAX (d_prep)
}

// =====================================================================================================
// K1  insertion scan: the part of insertFiringIntoRangeImage that is sequential over firings -- rotation
//     unwrapping against the previous rearmost column (cpp:119-175), the cell-collision rule (cpp:188-208),
//     the "too far behind" cut (cpp:210-221), rearmost/foremost tracking, the straddle check (cpp:252-261)
//     and which firing completes which column (cpp:289-291). One CTA of 1024 threads.
//
//     Firings are staged CC chunk by chunk into shared memory with cp.async, one chunk ahead. A run of firings is
//     REGULAR when (a) every point unwraps to the same column for any rearmost column the run can reach, (b) per
//     row the columns are strictly increasing and beyond the row's front -- so every cell is empty when it is
//     hit and the collision rule never fires -- and (c) no firing straddles the -x axis. Then the firings are
//     independent given the running rearmost column, which is a prefix maximum: the whole run is resolved
//     data-parallel (phases A..D). The first firing that breaks regularity is found exactly; a few firings from
//     there go through the per-firing path (exact for everything), then the fast path resumes.
//     The per-row occupancy of the last CC_K1_WINDOW columns lives in shared memory; distances are written
//     through to the ring so that K1b can tell which writer of a cell won (the last writer is the closest, cpp:207).
// =====================================================================================================
struct CcScanState // uniform across the CTA, kept in registers by every thread
{
    int Prel, Frel, Fmrel, colbase_rel, prev_rot, pc;
    bool f_init, fm_init;
    long long ring_start, ring_end, first_unpub;
    int reset_required, error;
};

CC_DEV void d_insert_scan(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, int n_firings, int C, int after_lite)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_insert_scan, g.bid);
    if (g.bid != 0 || p.st->halted)
        return;
    CC_SMEM(smem);
    const int R = cfg.R, N = cfg.N, W = CC_K1_WINDOW, ringcols = cfg.ringcols;
    const int T = blockDim.x, tid = threadIdx.x, lane = tid % CC_WARP, warp = tid / CC_WARP;
    const int nwarps = (T + CC_WARP - 1) / CC_WARP;
    const int row_warps = (R + CC_WARP - 1) / CC_WARP < nwarps ? (R + CC_WARP - 1) / CC_WARP : nwarps;
    const int GS = R + 1;
    // threads per row in the row phases; every (row, part) owns `seg` consecutive firings of a run
    const int nparts = T / R > 0 ? (T / R < C ? T / R : C) : 1;
    const int seg = (C + nparts - 1) / nparts;

    float* wdist = reinterpret_cast<float*>(smem);           // [W][R] occupancy window
    int* rmx = reinterpret_cast<int*>(wdist + static_cast<size_t>(W) * R); // [R] row front
    int* rm_new = rmx + R;                                    // [R]
    int* st_cwr = rm_new + R;                                 // [2][C][R] staged column-in-rotation
    float* st_dist = reinterpret_cast<float*>(st_cwr + 2 * C * R);
    int* red = reinterpret_cast<int*>(st_dist + 2 * C * R);  // [2][32][2]
    int* G = red + 2 * 2 * 32;                                // [C][GS]
    int* f_rear = G + C * GS;                                 // [C]
    int* f_fore = f_rear + C;                                 // [C]
    int* f_P = f_fore + C;                                    // [C + 1] rearmost-so-far before firing k
    int* f_F = f_P + C + 1;                                   // [C + 1] foremost-so-far before firing k
    int* seg_first = f_F + C + 1;                             // [R][nparts] first valid column of a segment
    int* seg_firstk = seg_first + R * nparts;                 // its firing
    int* seg_last = seg_firstk + R * nparts;                  // last valid column of a segment
    int* seg_stored = seg_last + R * nparts;                  // last stored column of a segment
    int* f_misc = seg_stored + R * nparts;                    // [0] k_bad
    CcDevState* st = p.st;
    float* ring_w = &p.pos[0].w; // distance of ring cell i is ring_w[4 * i]

    const long long base = st->P;
    const int base_slot = static_cast<int>(base & (W - 1));
    const int base_local = static_cast<int>(base % ringcols);
    const int neg_limit = base > 0x3fffffff ? -0x7fffffff : -static_cast<int>(base); // g_rel < neg_limit <=> g < 0
    const int NOT_SET = -0x7fffffff - 1;
    const int half = cfg.half;
    const float nanv = cc_nanf();
    const bool p_positive_at_start = base > 0;

    const int n_chunks = (n_firings + C - 1) / C;
    // firings [0, k_start) were resolved by the lite path (k_scan_lite / _check / _apply)
    int k_start = 0;
    if (after_lite)
    {
        k_start = st->scan_kbad;
        k_start = k_start < 0 ? 0 : (k_start > n_firings ? n_firings : k_start);
    }
    const int chunk_start = k_start / C;
    auto prefetch = [&](int chunk)
    {
        if (chunk < n_chunks)
        {
            const int buf = chunk & 1;
            const size_t src = static_cast<size_t>(chunk) * C * R;
            for (int i = tid * 4; i < C * R; i += T * 4)
            {
                __pipeline_memcpy_async(st_cwr + buf * C * R + i, p.s_cwr + src + i, 16);
                __pipeline_memcpy_async(st_dist + buf * C * R + i, p.s_dist + src + i, 16);
            }
        }
        __pipeline_commit();
    };
    prefetch(chunk_start);

    if (after_lite && k_start > 0)
    {
        // apply the lite prefix: new row fronts and -- only when firings remain for the paths below, which test cell
        // occupancy -- the distance write-through of its stored points. A push that is regular as a whole hits every
        // cell once and finds it empty, so k_scatter needs no winner test and nothing is written here.
        if (k_start >= n_firings)
        {
            for (int row = tid; row < R; row += T)
            {
                const int f = p.lite_rowfront[row];
                if (f != NOT_SET && base + f > p.rowmax[row])
                    p.rowmax[row] = base + f;
            }
        }
        else
        {
            for (int row = tid; row < R; row += T)
                rm_new[row] = NOT_SET;
            __syncthreads();
            const int total = k_start * R;
            for (int idx = tid; idx < total; idx += T)
            {
                const int k = idx / R, row = idx - k * R;
                const int cw = p.s_cwr[idx];
                if (cw == CC_INVALID_CWR)
                    continue;
                const int g = p.lite_U[k] + cc_wrapdiff(cw - p.lite_sum[k].anchor, N);
                if (g < p.lite_P[k])
                    continue; // too far behind (cpp:210-221)
                int local = base_local + g;
                local = local < 0 ? local + ringcols : (local >= ringcols ? local - ringcols : local);
                p.pos[static_cast<size_t>(local) * R + row].w = p.s_dist[idx];
                atomicMax(&rm_new[row], g);
            }
            __syncthreads();
            for (int row = tid; row < R; row += T)
                if (rm_new[row] != NOT_SET && base + rm_new[row] > p.rowmax[row])
                    p.rowmax[row] = base + rm_new[row];
        }
        __syncthreads();
    }

    for (int row = tid; row < R; row += T)
    {
        const long long rm = p.rowmax[row];
        long long rel = rm - base;
        if (rel < -0x3fffffff)
            rel = -0x3fffffff;
        rmx[row] = static_cast<int>(rel);
    }
    if (k_start < n_firings) // occupancy window of the last W columns of every row, from the ring
        for (int i = tid; i < R * W; i += T)
        {
            const int row = i % R;
            const long long c = p.rowmax[row] - i / R;
            if (c >= 0)
                wdist[(c & (W - 1)) * R + row] = p.pos[static_cast<size_t>(cc_local_col(c, ringcols)) * R + row].w;
        }

    CcScanState s;
    {
        const long long F64 = st->F, Fm64 = st->foremost;
        s.f_init = F64 >= 0;
        s.fm_init = Fm64 >= 0;
        s.Prel = 0;
        s.Frel = s.f_init ? static_cast<int>(F64 - base) : 0;
        s.Fmrel = s.fm_init ? static_cast<int>(Fm64 - base) : 0;
        s.colbase_rel = s.Frel;
        s.prev_rot = static_cast<int>(base / N);
        s.pc = static_cast<int>(base % N);
        s.ring_start = st->ring_start;
        s.ring_end = st->ring_end;
        s.first_unpub = st->first_unpub;
        s.reset_required = st->reset_required;
        s.error = 0;
    }
    if (k_start > 0)
    {
        // commit the lite prefix: columns completed by each firing (cpp:289-291), rearmost / foremost so far
        for (int k = tid; k < k_start; k += T)
            for (int c = p.lite_P[k]; c < p.lite_P[k + 1]; c++)
                p.col_trigger[c - s.colbase_rel] = k;
        const int Pnew = p.lite_P[k_start], Fnew = p.lite_F[k_start];
        s.pc += Pnew - s.Prel;
        while (s.pc >= N)
        {
            s.pc -= N;
            s.prev_rot++;
        }
        s.Prel = Pnew;
        s.Frel = Pnew;
        s.Fmrel = Fnew;
        if (base + s.Fmrel > s.ring_end)
            s.ring_end = base + s.Fmrel;
    }
    __syncthreads();

    bool stop = false;
    int n_fast = 0, n_slow = 0, n_attempts = 0;
    int slow_run = CC_K1_SLOW_RUN; // grows while fast attempts keep failing early (dense collisions)
    for (int chunk = chunk_start; chunk < n_chunks && !stop && k_start < n_firings; chunk++)
    {
        prefetch(chunk + 1);
        __pipeline_wait_prior(1);
        __syncthreads();
        const int k0 = chunk * C;
        const int kc = (n_firings - k0) < C ? (n_firings - k0) : C;
        const int cbuf = (chunk & 1) * C * R;
        int ka = chunk == chunk_start ? k_start - k0 : 0; // next firing of the chunk to process
        while (ka < kc && !stop)
        {
            int kgood = ka; // firings [ka, kgood) are resolved by the fast path
            const bool try_fast = s.f_init && s.fm_init && (p_positive_at_start || s.Prel > 0) && s.ring_start != -1;
            if (try_fast)
            {
                const int kb = kc;
                if (tid == 0)
                {
                    f_misc[0] = kb;
                }
                for (int row = tid; row < R; row += T)
                    rm_new[row] = rmx[row];
                __syncthreads();
                // ---- phase A+B: one warp per firing, lanes over rows: unwrap every point against the state at the
                //      start of the run; rearmost / foremost column of the firing by warp reduction ----
                {
                    const int goff = s.Prel - s.pc;
                    int kbad = kb;
                    for (int k = ka + warp; k < kb; k += nwarps)
                    {
                        int lmin = 0x7fffffff, lmax = NOT_SET;
                        const int* cwp = st_cwr + cbuf + k * R;
                        int* gp = G + k * GS;
                        for (int row = lane; row < R; row += CC_WARP)
                        {
                            const int cw = cwp[row];
                            int g = NOT_SET;
                            if (cw != CC_INVALID_CWR)
                            {
                                const int diff = cw - s.pc;
                                g = goff + cw;
                                if (diff < -half)
                                    g += N;
                                else if (diff > half)
                                    g -= N;
                                if (g < neg_limit)
                                    kbad = k < kbad ? k : kbad;
                                lmin = g < lmin ? g : lmin;
                                lmax = g > lmax ? g : lmax;
                            }
                            gp[row] = g;
                        }
                        lmin = cc_warp_min(lmin);
                        lmax = cc_warp_max(lmax);
                        if (lane == 0)
                        {
                            f_rear[k] = lmin;
                            f_fore[k] = lmax;
                        }
                    }
                    kbad = cc_warp_min(kbad);
                    if (lane == 0 && kbad < kb)
                        atomicMin(&f_misc[0], kbad);
                }
                __syncthreads();
                // ---- phase S: rearmost / foremost so far before every firing = prefix maxima (cpp:263-266) ----
                if (warp == 0)
                {
                    const int n = kb - ka;
                    const int per = (n + CC_WARP - 1) / CC_WARP;
                    const int a = ka + lane * per, b = (a + per < kb) ? a + per : kb;
                    int mP = NOT_SET, mF = NOT_SET, kbad = kb, gmin = 0x7fffffff;
                    for (int k = a; k < b; k++)
                    {
                        const int rear = f_rear[k], fore = f_fore[k];
                        if (rear != 0x7fffffff)
                        {
                            if (fore - rear > N / 2) // cpp:252-261
                                kbad = k < kbad ? k : kbad;
                            mP = rear > mP ? rear : mP;
                            mF = fore > mF ? fore : mF;
                            gmin = rear < gmin ? rear : gmin;
                        }
                    }
                    // exclusive scan of the per-lane maxima
                    int eP = mP, eF = mF;
                    for (int o = 1; o < CC_WARP; o <<= 1)
                    {
                        const int vP = __shfl_up_sync(CC_FULL_MASK, eP, o), vF = __shfl_up_sync(CC_FULL_MASK, eF, o);
                        if (lane >= o)
                        {
                            eP = vP > eP ? vP : eP;
                            eF = vF > eF ? vF : eF;
                        }
                    }
                    int pP = __shfl_up_sync(CC_FULL_MASK, eP, 1), pF = __shfl_up_sync(CC_FULL_MASK, eF, 1);
                    if (lane == 0)
                    {
                        pP = NOT_SET;
                        pF = NOT_SET;
                    }
                    int Pcur = pP > s.Prel ? pP : s.Prel, Fcur = pF > s.Fmrel ? pF : s.Fmrel;
                    for (int k = a; k < b; k++)
                    {
                        f_P[k] = Pcur;
                        f_F[k] = Fcur;
                        const int rear = f_rear[k], fore = f_fore[k];
                        if (rear != 0x7fffffff)
                        {
                            Pcur = rear > Pcur ? rear : Pcur;
                            Fcur = fore > Fcur ? fore : Fcur;
                        }
                        if (Pcur - s.colbase_rel > p.maxcols)
                            kbad = k < kbad ? k : kbad; // the per-firing path raises the error
                    }
                    if (a < kb && b == kb) // the lane whose range ends the run publishes the totals
                    {
                        f_P[kb] = Pcur;
                        f_F[kb] = Fcur;
                    }
                    kbad = cc_warp_min(kbad);
                    gmin = cc_warp_min(gmin);
                    const int gmax = cc_warp_max(mF);
                    if (lane == 0)
                    {
                        if (gmax != NOT_SET && (gmax - s.Prel >= half || gmin - s.Prel <= -half || gmax - gmin >= half))
                            kbad = ka; // the unwrap decision could depend on how far the rearmost column moves
                        if (kbad < kb)
                            atomicMin(&f_misc[0], kbad);
                    }
                }
                __syncthreads();
                // ---- phase C: per (row, segment): columns strictly increasing inside the segment; its last stored
                //      column (stored = not "too far behind", cpp:210-221) ----
                for (int item = tid; item < R * nparts; item += T)
                {
                    const int row = item % R, part = item / R;
                    const int a = ka + part * seg, b = (a + seg < kb) ? a + seg : kb;
                    int first = NOT_SET, firstk = kb, last = NOT_SET, kbad = kb, last_stored = NOT_SET;
                    for (int k = a; k < b; k++)
                    {
                        const int g = G[k * GS + row];
                        if (g != NOT_SET)
                        {
                            if (first == NOT_SET)
                            {
                                first = g;
                                firstk = k;
                            }
                            else if (g <= last)
                                kbad = k < kbad ? k : kbad;
                            last = g;
                            if (g >= f_P[k])
                                last_stored = g;
                        }
                    }
                    seg_first[item] = first;
                    seg_firstk[item] = firstk;
                    seg_last[item] = last;
                    seg_stored[item] = last_stored;
                    if (kbad < kb)
                        atomicMin(&f_misc[0], kbad);
                }
                __syncthreads();
                // ---- phase C1b: first column of a segment beyond everything before it (and the row's front) ----
                for (int item = tid; item < R * nparts; item += T)
                {
                    const int row = item % R, part = item / R;
                    const int first = seg_first[item];
                    if (first == NOT_SET)
                        continue;
                    int prev = rmx[row];
                    for (int q = part - 1; q >= 0; q--)
                    {
                        const int l = seg_last[q * R + row];
                        if (l != NOT_SET)
                        {
                            prev = l;
                            break;
                        }
                    }
                    if (first <= prev)
                        atomicMin(&f_misc[0], seg_firstk[item]);
                    const int ls = seg_stored[item];
                    if (ls != NOT_SET)
                        atomicMax(&rm_new[row], ls); // new front of the row (valid when the whole run is regular)
                }
                __syncthreads();
                kgood = f_misc[0];
                n_attempts++;
                n_fast += kgood - ka;
                slow_run = (kgood - ka) < 8 ? (slow_run * 2 < 64 ? slow_run * 2 : 64) : CC_K1_SLOW_RUN;
                if (kgood > ka)
                {
                    if (kgood < kb)
                    {
                        // the run ends early: recompute the new fronts for [ka, kgood) only
                        __syncthreads();
                        for (int row = tid; row < R; row += T)
                            rm_new[row] = rmx[row];
                        __syncthreads();
                        for (int item = tid; item < R * nparts; item += T)
                        {
                            const int row = item % R, part = item / R;
                            const int a = ka + part * seg;
                            int b = (a + seg < kb) ? a + seg : kb;
                            b = b < kgood ? b : kgood;
                            int last = NOT_SET;
                            for (int k = a; k < b; k++)
                            {
                                const int g = G[k * GS + row];
                                if (g != NOT_SET && g >= f_P[k])
                                    last = g;
                            }
                            if (last != NOT_SET)
                                atomicMax(&rm_new[row], last);
                        }
                        __syncthreads();
                    }
                    // ---- phase C2b: the window cells between the old and the new front start out empty ----
                    for (int item = tid; item < R * nparts; item += T)
                    {
                        const int row = item % R, part = item / R;
                        const int rn = rm_new[row];
                        int lo = rmx[row] + 1;
                        if (lo < rn - W + 1)
                            lo = rn - W + 1;
                        for (int c = lo + part; c <= rn; c += nparts)
                            wdist[((c + base_slot) & (W - 1)) * R + row] = nanv;
                    }
                    __syncthreads();
                    // ---- phase D: every stored point: occupancy window + distance write-through to the ring. The
                    //      resolved columns themselves are recomputed by K1b from the per-firing record. ----
                    for (int k = ka + warp; k < kgood; k += nwarps)
                    {
                        const int Pk = f_P[k];
                        const int* gp = G + k * GS;
                        const float* dp = st_dist + cbuf + k * R;
                        for (int row = lane; row < R; row += CC_WARP)
                        {
                            const int g = gp[row];
                            if (g != NOT_SET && g >= Pk)
                            {
                                const float d = dp[row];
                                if (g > rm_new[row] - W)
                                    wdist[((g + base_slot) & (W - 1)) * R + row] = d;
                                int local = base_local + g;
                                local = local < 0 ? local + ringcols : (local >= ringcols ? local - ringcols : local);
                                ring_w[static_cast<size_t>(static_cast<unsigned int>(local * R + row)) * 4] = d;
                            }
                        }
                        if (lane == 0)
                        {
                            CcFiringRecord rec;
                            rec.mode = 1;
                            rec.goff = s.Prel - s.pc;
                            rec.pc = s.pc;
                            rec.rot = s.prev_rot;
                            rec.P = Pk;
                            rec.pad_[0] = rec.pad_[1] = rec.pad_[2] = 0;
                            p.firing_rec[k0 + k] = rec;
                        }
                    }
                    // columns completed by each firing (cpp:289-291)
                    for (int k = ka + tid; k < kgood; k += T)
                        for (int c = f_P[k]; c < f_P[k + 1]; c++)
                            p.col_trigger[c - s.colbase_rel] = k0 + k;
                    const int Pnew = f_P[kgood], Fnew = f_F[kgood];
                    __syncthreads();
                    for (int row = tid; row < R; row += T)
                        rmx[row] = rm_new[row];
                    s.pc += Pnew - s.Prel;
                    while (s.pc >= N)
                    {
                        s.pc -= N;
                        s.prev_rot++;
                    }
                    s.Prel = Pnew;
                    s.Frel = Pnew;
                    s.Fmrel = Fnew;
                    if (base + s.Fmrel > s.ring_end)
                        s.ring_end = base + s.Fmrel;
                }
                __syncthreads();
            }
            ka = kgood;
            if (ka >= kc)
                break;

            // ---- per-firing path: exact for everything; a few firings, then the fast path is tried again ----
            const int kslow_end = (ka + slow_run < kc) ? ka + slow_run : kc;
            for (int k = ka; k < kslow_end; k++)
            {
                if (tid == 0)
                    p.firing_rec[k0 + k].mode = 0; // resolved columns of this firing are in o_g / o_rot
                const int sbase = cbuf + k * R;
                const bool p_positive = p_positive_at_start || s.Prel > 0;
                const int goff = s.Prel - s.pc;
                int lmin = 0x7fffffff, lmax = NOT_SET;
                for (int row = tid; row < R; row += T)
                {
                    const int cw = st_cwr[sbase + row];
                    if (cw == CC_INVALID_CWR)
                        continue;
                    const float d = st_dist[sbase + row];
                    int g = goff + cw;
                    const int diff = cw - s.pc;
                    int rot = s.prev_rot;
                    if (diff < -half)
                    {
                        g += N;
                        rot++;
                    }
                    else if (p_positive && diff > half)
                    {
                        g -= N;
                        rot--;
                    }
                    if (g < neg_limit)
                        continue; // reference: out-of-bounds access (undefined); the point is dropped here
                    int rm = rmx[row];
                    const int slot = ((g + base_slot) & (W - 1)) * R + row;
                    float cd = nanv;
                    if (g <= rm)
                    {
                        if (g > rm - W)
                            cd = wdist[slot];
                        else
                        {
                            int local = base_local + g;
                            local = local < 0 ? local + ringcols : (local >= ringcols ? local - ringcols : local);
                            cd = p.pos[static_cast<size_t>(local) * R + row].w;
                        }
                    }
                    int slot_w = slot;
                    if (!cc_isnan(cd) && !cc_isnan(d))
                    {
                        const int g1 = g + 1;
                        const int slot1 = ((g1 + base_slot) & (W - 1)) * R + row;
                        float nd = nanv;
                        if (g1 <= rm)
                        {
                            if (g1 > rm - W)
                                nd = wdist[slot1];
                            else
                            {
                                int local = base_local + g1;
                                local = local < 0 ? local + ringcols : (local >= ringcols ? local - ringcols : local);
                                nd = p.pos[static_cast<size_t>(local) * R + row].w;
                            }
                        }
                        if (cc_isnan(nd))
                        {
                            g = g1;
                            cd = nd;
                            slot_w = slot1;
                        }
                    }
                    if (!cc_isnan(cd) && (cc_isnan(d) || d >= cd))
                        continue;
                    const bool too_far_behind = s.f_init && g < s.Frel;
                    if (!too_far_behind)
                    {
                        if (g > rm)
                        {
                            int lo = rm + 1;
                            if (lo < g - W + 1)
                                lo = g - W + 1;
                            for (int c = lo; c < g; c++)
                                wdist[((c + base_slot) & (W - 1)) * R + row] = nanv;
                            rmx[row] = g;
                            rm = g;
                        }
                        if (g > rm - W)
                            wdist[slot_w] = d;
                        int local = base_local + g;
                        local = local < 0 ? local + ringcols : (local >= ringcols ? local - ringcols : local);
                        p.pos[static_cast<size_t>(local) * R + row].w = d; // write-through
                        const int idx = (k0 + k) * R + row;
                        p.o_g[idx] = g;
                        p.o_rot[idx] = rot;
                    }
                    lmin = g < lmin ? g : lmin;
                    lmax = g > lmax ? g : lmax;
                }
                int wmin = 0x7fffffff, wmax = NOT_SET;
                if (warp < row_warps)
                {
                    wmin = cc_warp_min(lmin);
                    wmax = cc_warp_max(lmax);
                }
                if (nwarps > 1)
                {
                    int* r = red + (k & 1) * (2 * 32);
                    if (lane == 0 && warp < row_warps)
                    {
                        r[2 * warp] = wmin;
                        r[2 * warp + 1] = wmax;
                    }
                    __syncthreads();
                    wmin = 0x7fffffff;
                    wmax = NOT_SET;
                    for (int w = 0; w < row_warps; w++)
                    {
                        const int a = r[2 * w], b = r[2 * w + 1];
                        wmin = a < wmin ? a : wmin;
                        wmax = b > wmax ? b : wmax;
                    }
                }
                if (wmin != 0x7fffffff)
                {
                    const int rear = wmin, fore = wmax;
                    if (fore - rear > N / 2) // cpp:252-261
                    {
                        s.reset_required = 1;
                        continue;
                    }
                    if (rear > s.Prel)
                    {
                        s.pc += rear - s.Prel;
                        while (s.pc >= N)
                        {
                            s.pc -= N;
                            s.prev_rot++;
                        }
                        s.Prel = rear;
                    }
                    if (!s.fm_init || fore > s.Fmrel)
                    {
                        s.Fmrel = fore;
                        s.fm_init = true;
                    }
                }
                if (!s.fm_init)
                    continue;
                if (s.ring_start == -1)
                {
                    s.ring_start = base + s.Prel;
                    s.first_unpub = base + s.Prel;
                }
                if (base + s.Fmrel > s.ring_end)
                    s.ring_end = base + s.Fmrel;
                if (!s.f_init)
                {
                    s.f_init = true;
                    s.Frel = s.Prel;
                    s.colbase_rel = s.Frel;
                }
                // columns [F, P) are complete: this firing's pose drives their segmentation (cpp:289-291)
                if (s.Prel > s.Frel)
                {
                    if (s.Prel - s.colbase_rel > p.maxcols)
                    {
                        s.error = CC_DEV_TOO_MANY_COLUMNS;
                        stop = true;
                        break;
                    }
                    for (int c = s.Frel + tid; c < s.Prel; c += T)
                        p.col_trigger[c - s.colbase_rel] = k0 + k;
                    s.Frel = s.Prel;
                }
            }
            __syncthreads();
            n_slow += kslow_end - ka;
            ka = kslow_end;
        }
    }
    __syncthreads();
    for (int row = tid; row < R; row += T)
    {
        const int rm = rmx[row];
        if (rm > -0x3fffffff)
            p.rowmax[row] = base + rm;
    }
    if (tid == 0)
    {
        st->scan_base = base;
        st->scan_lite_firings = k_start;
        st->scan_fast_firings = n_fast;
        st->scan_slow_firings = n_slow;
        st->scan_fast_attempts = n_attempts;
        st->P = base + s.Prel;
        st->foremost = s.fm_init ? base + s.Fmrel : -1;
        st->F = s.f_init ? base + s.Frel : -1;
        st->ring_start = s.ring_start;
        st->ring_end = s.ring_end;
        st->first_unpub = s.first_unpub;
        st->reset_required = s.reset_required;
        st->colbase = s.f_init ? base + s.colbase_rel : -1;
        st->ncols = s.f_init ? s.Frel - s.colbase_rel : 0;
        if (s.error)
            st->error = s.error;
        st->clear2_from = st->clear_from; // retired by the previous push: recycled at the start of the next one
        st->clear2_to = st->clear_to;
        st->clear_from = s.ring_start;
        st->clear_to = s.ring_start;
        st->push_first_unpub_old = s.first_unpub;
        st->n_edges = 0;
        st->n_flagged = 0;
        st->n_probe = 0;
        st->n_heavy = 0;
        st->danger_col = CC_COL_INF;
        st->forced_col = -1;
        st->abort = 0;
        st->n_clusters = 0;
        st->n_cluster_points = 0;
        st->n_vfix = 0;
    }
}
__global__ void __launch_bounds__(1024) k_insert_scan(CcDevCfg cfg, CcDevPtrs p, int n_firings, int C, int after_lite)
{
    CC_PDL_ENTER();
    d_insert_scan(cc_grid(), cfg, p, n_firings, C, after_lite);
}

// =====================================================================================================
// K1b scatter the staged fields of every stored point into the ring (cpp:223-237). Parallel; a point whose
//     cell was re-written by a closer return of a later firing (cpp:207) sees a different distance and skips.
// =====================================================================================================
CC_DEV void d_scatter(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, int n_firings)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_scatter, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    const int total = n_firings * cfg.R;
    const int scan_kbad = p.st->scan_kbad;
    // A push that is regular as a whole needs nothing from the insertion scan's commit (which may be running beside this
    // stage in the fused kernel and rewrites scan_base): the column base is the one the lite stages left.
    const long long scan_base = scan_kbad >= n_firings ? p.st->scan_lite_base : p.st->scan_base;
    for (int idx = g.bid * blockDim.x + threadIdx.x; idx < total; idx += g.nb * blockDim.x)
    {
        const int k = idx / cfg.R, row = idx - k * cfg.R;
        int grel, rot;
        bool winner_test = true;
        if (k < scan_kbad)
        {
            winner_test = scan_kbad < n_firings; // a push that is regular as a whole has one writer per cell
            // firing resolved by the lite path: same integers as k_scan_check / k_scan_apply
            const int cw = p.s_cwr[idx];
            if (cw == CC_INVALID_CWR)
                continue;
            grel = p.lite_U[k] + cc_wrapdiff(cw - p.lite_sum[k].anchor, cfg.N);
            if (grel < p.lite_P[k])
                continue; // too far behind (cpp:210-221)
            rot = static_cast<int>((scan_base + grel - cw) / cfg.N); // exact: column == rot * N + cw
        }
        else if (p.firing_rec[k].mode)
        {
            // regular firing: recompute the unwrap of the scan (same integers) instead of reading per-point outputs
            const CcFiringRecord rec = p.firing_rec[k];
            const int cw = p.s_cwr[idx];
            if (cw == CC_INVALID_CWR)
                continue;
            const int diff = cw - rec.pc;
            grel = rec.goff + cw;
            rot = rec.rot;
            if (diff < -cfg.half)
            {
                grel += cfg.N;
                rot++;
            }
            else if (diff > cfg.half)
            {
                grel -= cfg.N;
                rot--;
            }
            if (grel < rec.P)
                continue; // too far behind (cpp:210-221)
        }
        else
        {
            grel = p.o_g[idx];
            if (grel == -0x7fffffff - 1)
                continue;
            rot = p.o_rot[idx];
        }
        const long long gc = scan_base + grel;
        const size_t cell = static_cast<size_t>(cc_local_col(gc, cfg.ringcols)) * cfg.R + row;
        const float4 sp = p.s_pos[idx];
        if (winner_test && ccm::f2u(p.pos[cell].w) != ccm::f2u(sp.w))
            continue;
        const CcRawPoint* raw = reinterpret_cast<const CcRawPoint*>(p.raw) + idx;
        p.pos[cell] = sp;
        p.azimuth[cell] = p.s_az[idx];
        p.incl[cell] = p.s_incl[idx];
        p.cont_az[cell] = (2 * M_PI) * static_cast<double>(rot) + static_cast<double>(p.s_incaz[idx]);
        uchar4 l = p.lab[cell];
        l.w = raw->intensity;
        p.lab[cell] = l;
        p.stamp[cell] = raw->stamp;
        p.guid[cell] = raw->guid;
        p.firing_index[cell] = raw->firing_index;
    }
}
__global__ void k_scatter(CcDevCfg cfg, CcDevPtrs p, int n_firings)
{
    CC_PDL_ENTER();
    d_scatter(cc_grid(), cfg, p, n_firings);
}

// =====================================================================================================
// K2a  sc_inclination_angles_between_lasers_ (cpp:353-357): per row, the last non-NaN inclination difference
//      to the laser below seen in any column so far. Columns are cut into chunks of CC_GAP_CHUNK: every (chunk, row)
//      resolves its columns locally with coalesced reads (NaN where the chunk has not seen a valid value yet) and
//      leaves its last valid value; the block that finishes last chains the chunks per row (warp scan), seeded with
//      the value carried from earlier pushes. k_ground takes the chunk's carry-in where the local value is NaN.
// =====================================================================================================
#define CC_GAP_CHUNK 32
#define CC_GAP_BATCH 16 /* columns whose loads are issued together */

CC_DEV float cc_ldcg_f32(const float* q)
{
#ifdef CC_EMU
    return *q;
#else
    return __ldcg(q);
#endif
}

CC_DEV void d_gap_main(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, const CcHead& hd)
{
    const int R = cfg.R;
    const int ncols = hd.ncols;
    const long long colbase = hd.colbase;
    const int T = blockDim.x, t = threadIdx.x;
    const int nchunks = (ncols + CC_GAP_CHUNK - 1) / CC_GAP_CHUNK;
    const float nanv = cc_nanf();
    const int base_local = ncols > 0 ? cc_local_col(colbase, cfg.ringcols) : 0;
    const int items = nchunks * R; // (chunk, row), rows fastest: a warp reads consecutive rows of one column
    CcTraceScope cc_tr_gmain(p.trace, CC_KID_gap_main, g.bid);
    for (int it = g.bid * T + t; it < items; it += g.nb * T)
    {
        const int chunk = it / R, row = it - chunk * R;
        const int c0 = chunk * CC_GAP_CHUNK;
        float last = nanv;
        for (int jb = 0; jb < CC_GAP_CHUNK && c0 + jb < ncols; jb += CC_GAP_BATCH)
        {
            float a[CC_GAP_BATCH], b[CC_GAP_BATCH];
            int local = base_local + c0 + jb;
            if (local >= cfg.ringcols)
                local -= cfg.ringcols;
#pragma unroll
            for (int j = 0; j < CC_GAP_BATCH; j++)
            {
                const bool in = c0 + jb + j < ncols;
                const size_t cell = static_cast<size_t>(local) * R + row;
                a[j] = in ? p.incl[cell] : nanv;
                b[j] = (row == R - 1) ? 0.f : (in ? p.incl[cell + 1] : nanv);
                local = local + 1 == cfg.ringcols ? 0 : local + 1;
            }
#pragma unroll
            for (int j = 0; j < CC_GAP_BATCH; j++)
            {
                const float d = a[j] - b[j];
                if (!cc_isnan(d))
                    last = d;
                if (c0 + jb + j < ncols)
                    p.col_gap[static_cast<size_t>(c0 + jb + j) * R + row] = last;
            }
        }
        p.gap_chunk_last[it] = last;
    }
}

// single CTA: chains the chunks per row, seeded with the value carried from earlier pushes
CC_DEV void d_gap_tail(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, const CcHead& hd)
{
    const int R = cfg.R;
    const int ncols = hd.ncols;
    const int T = blockDim.x, t = threadIdx.x;
    const int nchunks = (ncols + CC_GAP_CHUNK - 1) / CC_GAP_CHUNK;
    const float nanv = cc_nanf();
    CcTraceScope cc_tr_gtail(p.trace, CC_KID_gap_tail, g.bid);
    // chain the chunks: tiles of chunks are staged in shared memory with independent coalesced loads (one memory
    // round trip per tile); every row is then walked by `parts` threads, each over a contiguous range of chunks
    __shared__ float sh_tile[8192];
    __shared__ float sh_part[1024];
    const int tile_chunks = 8192 / R;
    const int parts = T / R > 0 ? (T / R < 1024 / R ? T / R : 1024 / R) : 1;
    for (int cb = 0; cb < nchunks; cb += tile_chunks)
    {
        const int nc = (nchunks - cb) < tile_chunks ? (nchunks - cb) : tile_chunks;
        const int n = nc * R;
        const float* src = p.gap_chunk_last + static_cast<size_t>(cb) * R;
        for (int i0 = 0; i0 < n; i0 += 16 * T)
        {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; u++)
                v[u] = i0 + u * T + t < n ? cc_ldcg_f32(src + i0 + u * T + t) : nanv;
#pragma unroll
            for (int u = 0; u < 16; u++)
                if (i0 + u * T + t < n)
                    sh_tile[i0 + u * T + t] = v[u];
        }
        __syncthreads();
        const int per = (nc + parts - 1) / parts;
        for (int it = t; it < parts * R; it += T)
        {
            const int part = it / R, row = it - part * R;
            const int ca = part * per < nc ? part * per : nc, cz = (ca + per < nc) ? ca + per : nc;
            float last = nanv;
            for (int c = ca; c < cz; c++)
            {
                const float v = sh_tile[c * R + row];
                if (!cc_isnan(v))
                    last = v;
            }
            sh_part[it] = last;
        }
        __syncthreads();
        for (int it = t; it < parts * R; it += T)
        {
            const int part = it / R, row = it - part * R;
            const int ca = part * per < nc ? part * per : nc, cz = (ca + per < nc) ? ca + per : nc;
            float carry = p.gap_state[row];
            for (int q = 0; q < part; q++)
            {
                const float v = sh_part[q * R + row];
                if (!cc_isnan(v))
                    carry = v;
            }
            for (int c = ca; c < cz; c++)
            {
                const float v = sh_tile[c * R + row];
                sh_tile[c * R + row] = carry;
                if (!cc_isnan(v))
                    carry = v;
            }
            if (part == parts - 1)
                sh_part[it] = carry; // the row's value after the tile (its own entry is no longer needed)
        }
        __syncthreads();
        for (int row = t; row < R; row += T)
            p.gap_state[row] = sh_part[(parts - 1) * R + row];
        for (int i = t; i < n; i += T)
            p.gap_chunk_carry[static_cast<size_t>(cb) * R + i] = sh_tile[i];
        __syncthreads();
    }
}

// The same in ONE phase for short pushes (fused kernel): a warp per row, every lane a run of consecutive columns (two
// independent loads per column issued together), the lanes chained by a warp scan seeded with the value carried from
// earlier pushes. Leaves fully resolved values in col_gap (and NaN in the chunk carries d_ground falls back to).
CC_DEV void d_gap_rows(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, const CcHead& hd)
{
    CcTraceScope cc_tr_gmain(p.trace, CC_KID_gap_main, g.bid);
    const int R = cfg.R;
    const int ncols = hd.ncols;
    if (ncols <= 0)
        return;
    const int T = blockDim.x, lane = threadIdx.x % CC_WARP, warp = threadIdx.x / CC_WARP;
    const int nwarps = (T + CC_WARP - 1) / CC_WARP;
    const float nanv = cc_nanf();
    const int base_local = cc_local_col(hd.colbase, cfg.ringcols);
    const int nchunks = (ncols + CC_GAP_CHUNK - 1) / CC_GAP_CHUNK;
    const int per = (ncols + CC_WARP - 1) / CC_WARP;
    const bool one_pass = per <= 8; // the lane's values stay in registers between the two halves
    for (int row = g.bid * nwarps + warp; row < R; row += g.nb * nwarps)
    {
        const int a = lane * per < ncols ? lane * per : ncols, b = (a + per < ncols) ? a + per : ncols;
        const float state = p.gap_state[row]; // issued together with the cells below
        // difference to the laser below for up to 8 columns from c0 (NaN outside the lane's run), loads issued together
        auto diffs = [&](int c0, float* d)
        {
            float x[8], y[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                x[u] = y[u] = nanv;
                if (c0 + u < b)
                {
                    int local = base_local + c0 + u;
                    if (local >= cfg.ringcols)
                        local -= cfg.ringcols;
                    const size_t cell = static_cast<size_t>(local) * R + row;
                    x[u] = p.incl[cell];
                    y[u] = row == R - 1 ? 0.f : p.incl[cell + 1];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                d[u] = x[u] - y[u];
        };
        float keep[8];
        float last = nanv; // last valid value of the lane's run
        if (one_pass)
        {
            diffs(a, keep);
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                if (!cc_isnan(keep[u]))
                    last = keep[u];
                keep[u] = last; // value after this column, still without what came before the run
            }
        }
        else
            for (int c0 = a; c0 < b; c0 += 8)
            {
                float d[8];
                diffs(c0, d);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (!cc_isnan(d[u]))
                        last = d[u];
            }
        float carry = cc_warp_exclusive_scan(last, nanv, CcOpLastValid(), lane);
        if (cc_isnan(carry))
            carry = state;
        if (one_pass)
        {
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (a + u < b)
                    p.col_gap[static_cast<size_t>(a + u) * R + row] = cc_isnan(keep[u]) ? carry : keep[u];
        }
        else
        {
            float run = carry;
            for (int c0 = a; c0 < b; c0 += 8)
            {
                float d[8];
                diffs(c0, d);
#pragma unroll
                for (int u = 0; u < 8; u++)
                {
                    if (!cc_isnan(d[u]))
                        run = d[u];
                    if (c0 + u < b)
                        p.col_gap[static_cast<size_t>(c0 + u) * R + row] = run;
                }
            }
        }
        // the row's value after the push
        const float mine_or_carry = cc_isnan(last) ? carry : last;
        const float fin = cc_shfl_any(mine_or_carry, CC_WARP - 1);
        if (lane == 0)
            p.gap_state[row] = fin;
        for (int ch = lane; ch < nchunks; ch += CC_WARP)
            p.gap_chunk_carry[static_cast<size_t>(ch) * R + row] = nanv;
    }
}

__global__ void __launch_bounds__(256) k_gap_scan(CcDevCfg cfg, CcDevPtrs p)
{
    CC_PDL_ENTER();
    const CcGrid g = cc_grid();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_gap_scan, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    d_gap_main(g, cfg, p, hd);
    // the block that finishes last chains the chunks
    __shared__ int sh_last;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        const int ticket = atomicAdd(&p.st->ticket_gap, 1);
        sh_last = ticket == g.nb - 1;
    }
    __syncthreads();
    if (!sh_last)
        return;
    __threadfence();
    if (threadIdx.x == 0)
        p.st->ticket_gap = 0;
    d_gap_tail(g, cfg, p, hd);
}

// =====================================================================================================
// K2b  ground-point segmentation of one column per warp (cpp:294-624). Lanes stage the column into shared
//      memory (classification of every cell, projection into the azimuth plane, ego-box test in double), lane 0
//      runs the bottom-to-top label state machine on the staged values, then all lanes derive is_ignored
//      (cpp:567-616) and write the association view of the column.
// =====================================================================================================
// per-warp shared memory of k_ground (R rows): by-row staging, the compacted sequence of regular points the label
// state machine walks, inclinations, labels, cell classes and the bitmap of regular rows
struct CcGroundRow
{
    float c2x, c2y; // position w.r.t. the sensor in the azimuth plane (hpp:229-232)
    unsigned int flags;
    float gap;
};
#define CC_GF_FIRST 1u /* first regular point of the column (cpp:409-431) */
#define CC_GF_FLAT 2u  /* first point: inside the first-ring bounds; else: flat w.r.t. the previous point (cpp:434-441) */
#define CC_GF_LG 4u    /* slope / distance part of the last-certain-ground update rule (cpp:542-551) */
static inline __host__ __device__ size_t cc_ground_warp_bytes(int R)
{
    return (static_cast<size_t>(R) * (2 * sizeof(CcGroundRow) + sizeof(float4) + sizeof(double) + sizeof(float) + sizeof(unsigned short) + 2) +
            12 * sizeof(double) + 28 * sizeof(unsigned int) + 15) / 16 * 16;
}

CC_DEV double cc_ldcg_f64(const double* q)
{
#ifdef CC_EMU
    return *q;
#else
    return __ldcg(q);
#endif
}

// `d_labels` / `h_labels` (or null): the packed labels of the column (cc_set_label_prefetch) are also written to the
// push's label buffers, device and page-locked host, for columns below `label_cap` (fused kernel).
template<bool EXPORT_LABELS>
CC_DEV void d_ground(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, uchar4* d_labels = nullptr,
                     uchar4* h_labels = nullptr, int label_cap = 0)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_ground, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    CC_SMEM(smem);
    const int R = cfg.R;
    const int warps_per_block = blockDim.x / CC_WARP;
    const int wib = threadIdx.x / CC_WARP, lane = threadIdx.x % CC_WARP;
    unsigned char* wbase = smem + static_cast<size_t>(wib) * cc_ground_warp_bytes(R);
    CcGroundRow* s = reinterpret_cast<CcGroundRow*>(wbase);  // by row
    CcGroundRow* cseq = s + R;                               // regular points, bottom row first; flags carry the row
    float4* s_q = reinterpret_cast<float4*>(cseq + R);       // the cells as inserted: x, y, z, distance
    double* s_caz = reinterpret_cast<double*>(s_q + R);      // continuous azimuth
    double* s_ego = s_caz + R;                               // robot_from_sensor * odom_from_sensor^-1 of the column (3 x 4)
    float* s_incl = reinterpret_cast<float*>(s_ego + 12);
    unsigned int* vm = reinterpret_cast<unsigned int*>(s_incl + R); // bitmap of regular rows (8 words)
    unsigned int* fm = vm + 8;  // by position in cseq: flat w.r.t. the previous point
    unsigned int* lm = fm + 8;  // by position in cseq: passes the slope / distance part of the last-ground update rule
    unsigned int* cst = lm + 8; // state handed from C1 to C3
    unsigned short* s_lab = reinterpret_cast<unsigned short*>(cst + 4); // label | debug label << 8
    unsigned char* s_cls = reinterpret_cast<unsigned char*>(s_lab + R);
    unsigned char* s_int = s_cls + R; // intensity
    const int ncols = hd.ncols;
    const long long colbase = hd.colbase;
    const float nanv = cc_nanf();
    const int nwords = (R + 31) / 32;
    const float hsg = cfg.height_sensor_to_ground;

    CcTraceScope cc_tr_main_obj(p.trace, CC_KID_ground_main, g.bid);
    for (int ci = g.bid * warps_per_block + wib; ci < ncols; ci += g.nb * warps_per_block)
    {
        const long long gcol = colbase + ci;
        const int local = cc_local_col(gcol, cfg.ringcols);
        const size_t base = static_cast<size_t>(local) * R;
        CcTraceScope cc_tr_pose(p.trace, CC_KID_g_pose, g.bid);
        // ---- A0: everything the column needs from global memory, issued before anything waits: the firing that
        //      completed the column (its pose is one more dependent round trip) and the cells ----
        const int trig = p.col_trigger[ci];
        const long long slot_before = lane == 0 ? p.slot_gcol[local] : -1;
        for (int row = lane; row < R; row += CC_WARP)
        {
            const float4 q = p.pos[base + row];
            const float incl = p.incl[base + row];
            const unsigned char intensity = p.lab[base + row].w;
            const double caz = p.cont_az[base + row];
            float gap = p.col_gap[static_cast<size_t>(ci) * R + row];
            if (cc_isnan(gap)) // nothing valid in the column's chunk so far: what the chunks before it left
                gap = p.gap_chunk_carry[static_cast<size_t>(ci / CC_GAP_CHUNK) * R + row];
            s_q[row] = q;
            s_caz[row] = caz;
            s_incl[row] = incl;
            s_int[row] = intensity;
            s[row].gap = gap;
        }
        const double* pose = p.poses + 12 * trig;
        const float spx = static_cast<float>(pose[3]), spy = static_cast<float>(pose[7]),
                    spz = static_cast<float>(pose[11]);
        // ego-box transform of the column (cpp:300-301): prepared per firing by d_prep
        for (int e = lane; e < 12; e += CC_WARP)
            s_ego[e] = p.s_ego[static_cast<size_t>(trig) * 12 + e];
        if (lane == 0)
        {
            if (slot_before != -1)
            {
                p.st->error = CC_DEV_COLUMN_NOT_CLEARED;
                p.st->err_a = slot_before;
                p.st->err_b = gcol;
            }
            for (int w = 0; w < 24; w++)
                vm[w] = 0u; // vm, fm, lm
        }

        __syncwarp();
        cc_tr_pose.stop();
        CcTraceScope cc_tr_A(p.trace, CC_KID_g_A, g.bid);
        // ---- A: classify every cell, project it into the azimuth plane, bitmap of the regular rows ----
        for (int row0 = 0; row0 < R; row0 += CC_WARP)
        {
            const int row = row0 + lane;
            const bool in = row < R;
            unsigned char cls = 0;
            if (in)
            {
                const float4 q = s_q[row]; // staged by this very lane
                const float incl = s_incl[row];
                const unsigned char intensity = s_int[row];
                cls = 3;
                unsigned short lab = CC_GP_UNKNOWN | (CC_WHITE << 8);
                if (cc_isnan(q.w))
                    cls = 0;
                else if (cfg.fog_enabled && intensity < static_cast<unsigned char>(cfg.fog_intensity) &&
                         q.w < cfg.fog_dist && incl > cfg.fog_incl)
                {
                    cls = 1;
                    lab = CC_GP_FOG | (CC_LIGHTGRAY << 8); // cpp:376-382
                }
                else
                {
                    double e[3];
                    cc_iso_apply(s_ego, static_cast<double>(q.x), static_cast<double>(q.y), static_cast<double>(q.z), e);
                    if (e[0] < cfg.l_front && e[0] > cfg.l_rear && e[1] < cfg.w_left && e[1] > cfg.w_right &&
                        e[2] < cfg.h_max && e[2] > cfg.h_ground)
                    {
                        cls = 2;
                        lab = CC_GP_EGO_VEHICLE | (CC_VIOLET << 8); // cpp:396-402
                    }
                }
                const float x = q.x - spx, y = q.y - spy, z = q.z - spz;
                s[row].c2x = ccm::sqrt_rn(x * x + y * y);
                s[row].c2y = z;
                s[row].flags = 0u;
                s_lab[row] = lab;
                s_cls[row] = cls;
            }
            const unsigned int m = __ballot_sync(CC_FULL_MASK, in && cls == 3);
            if (lane == 0)
                vm[row0 >> 5] |= m << (row0 & 31);
        }
        __syncwarp();

        cc_tr_A.stop();
        CcTraceScope cc_tr_B(p.trace, CC_KID_g_B, g.bid);
        // ---- B: everything of the label rules that does not depend on carried state, for all rows at once: the
        //      previous regular point (fog / ego / empty cells never become "previous", cpp:360-404), the slope to it,
        //      the compacted bottom-to-top sequence; and the inclination supplement of runs of empty cells ----
        int nregular = 0;
        for (int w = 0; w < nwords; w++)
            nregular += __popc(vm[w]);
        for (int row = lane; row < R; row += CC_WARP)
        {
            const unsigned char cls = s_cls[row];
            if (cls == 3)
            {
                int w = row >> 5;
                unsigned int m = vm[w] & ~((2u << (row & 31)) - 1u); // regular rows below this one (higher index)
                int rank = __popc(m);
                while (!m && ++w < nwords)
                    m = vm[w];
                for (int w2 = (row >> 5) + 1; w2 < nwords; w2++)
                    rank += __popc(vm[w2]);
                const CcGroundRow me = s[row];
                unsigned int f = 0u;
                if (!m)
                {
                    const float h = me.c2y - hsg; // cpp:412-414
                    f = CC_GF_FIRST | ((h > cfg.first_min && h < cfg.first_max) ? CC_GF_FLAT : 0u);
                }
                else
                {
                    const int pv = w * 32 + __ffs(m) - 1;
                    const float ptc_x = me.c2x - s[pv].c2x, ptc_y = me.c2y - s[pv].c2y;
                    const float slope_to_prev = ccm::div_rn(ptc_y, ptc_x);
                    bool flat_prev = fabsf(slope_to_prev) < cfg.max_slope && ptc_x > 0;
                    flat_prev = flat_prev && (!cfg.use_terrain || ptc_x < 5);
                    if (flat_prev)
                        f |= CC_GF_FLAT;
                    if (slope_to_prev > cfg.lg_slope && fabsf(ptc_x) < cfg.lg_dist)
                        f |= CC_GF_LG;
                }
                CcGroundRow c = me;
                c.flags = f | (static_cast<unsigned int>(row) << 16);
                cseq[rank] = c;
                if (f & CC_GF_FLAT)
                    atomicOr(&fm[rank >> 5], 1u << (rank & 31));
                if (f & CC_GF_LG)
                    atomicOr(&lm[rank >> 5], 1u << (rank & 31));
            }
            else if (cls == 0 && cfg.supplement && (row == R - 1 || s_cls[row + 1] != 0))
            {
                // bottom cell of a run of empty cells: the supplement chains upwards through the run (cpp:364-369)
                float v = s_incl[row];
                if (row < R - 1)
                {
                    v = s_incl[row + 1] + s[row].gap;
                    s_incl[row] = v;
                }
                for (int r2 = row - 1; r2 >= 0 && s_cls[r2] == 0; r2--)
                {
                    v = v + s[r2].gap;
                    s_incl[r2] = v;
                }
            }
        }
        __syncwarp();

        cc_tr_B.stop();
        CcTraceScope cc_tr_C(p.trace, CC_KID_g_C, g.bid);
        // ---- C: the sequential label state machine (cpp:305-565) over the regular points only ----
        // C1 (lane 0): up to the first obstacle a flat point is GREEN whatever came before, so the walk jumps from one
        //     non-flat point to the next with bit operations on the flat / last-ground-candidate masks
        // C2 (all lanes): GREEN labels of that prefix
        // C3 (lane 0): the rest, point by point
        if (lane == 0 && nregular > 0)
        {
            bool fod = false, prev_yellow = false;
            float lg_x = 0.f, lg_y = hsg; // to2d of last_ground_position_wrt_sensor
            {
                const CcGroundRow cur = cseq[0]; // always the FIRST point (cpp:409-431)
                const int row = static_cast<int>(cur.flags >> 16);
                if (cur.flags & CC_GF_FLAT)
                {
                    s_lab[row] = CC_GP_GROUND | (CC_GRAY << 8);
                    lg_x = cur.c2x;
                    lg_y = cur.c2y;
                }
                else
                {
                    s_lab[row] = CC_GP_OBSTACLE | (CC_ORANGE << 8);
                    fod = true;
                }
            }
            int i = 1;
            while (!fod && i < nregular)
            {
                // next point at index >= i that is not flat w.r.t. its predecessor
                int j = nregular;
                for (int w = i >> 5; w < nwords; w++)
                {
                    unsigned int m = ~fm[w];
                    if (w == (i >> 5))
                        m &= ~((1u << (i & 31)) - 1u);
                    if (m)
                    {
                        const int cand = w * 32 + __ffs(m) - 1;
                        j = cand < nregular ? cand : nregular;
                        break;
                    }
                }
                if (j > i)
                {
                    // points [i, j) are GREEN; the last of them that passes the update rule (cpp:542-551) becomes the
                    // last certain ground point (only point i can have a YELLOW predecessor)
                    int k = -1;
                    for (int w = (j - 1) >> 5; w >= (i >> 5) && k < 0; w--)
                    {
                        unsigned int m = lm[w];
                        if (w == ((j - 1) >> 5) && ((j - 1) & 31) != 31)
                            m &= (2u << ((j - 1) & 31)) - 1u;
                        if (w == (i >> 5))
                        {
                            m &= ~((1u << (i & 31)) - 1u);
                            if (prev_yellow)
                                m &= ~(1u << (i & 31));
                        }
                        if (m)
                            k = w * 32 + 31 - __clz(static_cast<int>(m));
                    }
                    if (k >= 0)
                    {
                        lg_x = cseq[k].c2x;
                        lg_y = cseq[k].c2y;
                    }
                    prev_yellow = false;
                    i = j;
                    if (i >= nregular)
                        break;
                }
                // point i is not flat: YELLOW if close to the last certain ground, else the first obstacle
                const CcGroundRow cur = cseq[i];
                const float gtc_x = cur.c2x - lg_x, gtc_y = cur.c2y - lg_y;
                if (!cfg.use_terrain && fabsf(gtc_x) < cfg.close_d && fabsf(gtc_y) < cfg.close_z)
                {
                    s_lab[cur.flags >> 16] = CC_GP_GROUND | (CC_YELLOW << 8);
                    prev_yellow = true;
                    i++;
                }
                else
                    break; // handled by C3
            }
            cst[0] = i;
            cst[1] = (fod ? 1u : 0u) | (prev_yellow ? 2u : 0u);
            cst[2] = ccm::f2u(lg_x);
            cst[3] = ccm::f2u(lg_y);
        }
        __syncwarp();
        if (nregular > 0)
        {
            const int npre = static_cast<int>(cst[0]);
            for (int k = 1 + lane; k < npre; k += CC_WARP)
            {
                const unsigned int f = cseq[k].flags;
                if (f & CC_GF_FLAT)
                    s_lab[f >> 16] = CC_GP_GROUND | (CC_GREEN << 8);
            }
        }
        __syncwarp();
        // C3 (all lanes): the points from the first obstacle on, 32 at a time. Once an obstacle has been seen the only
        //     state the rules carry from point to point is the last certain ground position (and whether the previous
        //     point was YELLOW, which the neighbouring lane knows), and few points ever move it: every lane evaluates
        //     its point against the current position, the points up to and including the first one that moves it are
        //     final, and the next round starts behind it. The relabel walks of the round's obstacles (cpp:513-535)
        //     cover disjoint cells -- a walk never passes another obstacle -- so they run side by side.
        if (nregular > 0 && static_cast<int>(cst[0]) < nregular)
        {
            int i = static_cast<int>(cst[0]);
            bool fod = (cst[1] & 1u) != 0, prev_yellow = (cst[1] & 2u) != 0;
            float lg_x = ccm::u2f(cst[2]), lg_y = ccm::u2f(cst[3]);
            const bool terrain = cfg.use_terrain != 0;
            const float ms = cfg.max_slope;
            while (i < nregular)
            {
                const int k = i + lane;
                const bool has = k < nregular;
                CcGroundRow cur;
                cur.c2x = cur.c2y = cur.gap = 0.f;
                cur.flags = 0u;
                if (has)
                    cur = cseq[k];
                const int row = static_cast<int>(cur.flags >> 16);
                const float c2x = cur.c2x, c2y = cur.c2y;
                const bool flat_prev = (cur.flags & CC_GF_FLAT) != 0;
                const float gtc_x = c2x - lg_x, gtc_y = c2y - lg_y;
                const float agx = fabsf(gtc_x), agy = fabsf(gtc_y);
                // |RN(gtc_y / gtc_x)| < max_slope, decided without the division unless the quotient is within 1e-6
                // (relative) of the threshold
                const float t = ms * agx;
                const bool t_ok = t > 1e-30f && t < 1e30f;
                bool slope_ok = t_ok && agy < t * 0.999999f;
                // the point C1 stopped at is an obstacle by construction (not flat, not close to the last ground): it is
                // the only one that can still see fod == false, and only in lane 0 of the first round
                const bool my_fod = fod || lane > 0;
                const bool cand = my_fod && flat_prev && gtc_x > 0 && !terrain;
                if (has && cand && !(t_ok && (slope_ok || agy > t * 1.000001f)))
                    slope_ok = fabsf(ccm::div_rn(gtc_y, gtc_x)) < ms;
                const bool green = !my_fod && flat_prev;
                const bool yellowgreen = cand && slope_ok;
                const bool yellow = !terrain && agx < cfg.close_d && agy < cfg.close_z;
                const bool is_yellow = !green && !yellowgreen && yellow;
                const unsigned int lab = green         ? (CC_GP_GROUND | (CC_GREEN << 8))
                                         : yellowgreen ? (CC_GP_GROUND | (CC_YELLOWGREEN << 8))
                                         : yellow      ? (CC_GP_GROUND | (CC_YELLOW << 8))
                                                       : (CC_GP_OBSTACLE | (CC_RED << 8));
                // was the previous point YELLOW? (cpp:545: such a point does not move the last ground position)
                const unsigned int ymask = __ballot_sync(CC_FULL_MASK, has && is_yellow);
                const bool py = lane == 0 ? prev_yellow : ((ymask >> (lane - 1)) & 1u) != 0;
                const bool upd = has && (green || yellowgreen) && (cur.flags & CC_GF_LG) && !py; // cpp:541-561
                const unsigned int umask = __ballot_sync(CC_FULL_MASK, upd);
                int ncommit = nregular - i < CC_WARP ? nregular - i : CC_WARP;
                if (umask)
                    ncommit = __ffs(umask); // up to and including the first point that moves the last ground position
                const bool mine = lane < ncommit;
                if (mine)
                    s_lab[row] = static_cast<unsigned short>(lab);
                __syncwarp();
                if (mine && !(green || yellowgreen || yellow)) // cpp:508-536
                {
                    int below = row + 1;
                    while (below < R)
                    {
                        const unsigned int ql = s_lab[below];
                        const bool is_ground = (ql & 0xffu) == CC_GP_GROUND;
                        if ((ql >> 8) == CC_YELLOW || (is_ground && fabsf(c2x - s[below].c2x) < cfg.next_obst_d))
                        {
                            if (is_ground)
                                s_lab[below] = CC_GP_OBSTACLE | (CC_DARKRED << 8);
                            below++;
                        }
                        else
                            break;
                    }
                }
                __syncwarp();
                const int last = ncommit - 1;
                if (umask)
                {
                    lg_x = __shfl_sync(CC_FULL_MASK, c2x, last);
                    lg_y = __shfl_sync(CC_FULL_MASK, c2y, last);
                    prev_yellow = false;
                }
                else
                    prev_yellow = ((ymask >> last) & 1u) != 0;
                // any committed obstacle sets first_obstacle_detected; the first point of C3 always is one
                fod = true;
                i += ncommit;
            }
        }
        __syncwarp();

        cc_tr_C.stop();
        CcTraceScope cc_tr_D(p.trace, CC_KID_g_D, g.bid);
        // ---- D: is_ignored (cpp:567-616) + association view ----
        double min_az = 1.7976931348623157e308;
        for (int row = lane; row < R; row += CC_WARP)
        {
            const size_t cell = base + row;
            const float4 q = s_q[row];
            const unsigned short lab = s_lab[row];
            const unsigned char label = static_cast<unsigned char>(lab & 0xff);
            const float gap = s[row].gap;
            bool ignored = false;
            if (cc_isnan(q.w))
                ignored = true;
            else if (label != CC_GP_OBSTACLE)
                ignored = true;
            else if (q.w < cfg.max_distance)
                ignored = true;
            else if (cfg.incl_rule && row < R - 1 && ccm::atan2f_glibc(cfg.max_distance, q.w) < gap)
                ignored = true;
            else if (cfg.chessboard)
            {
                const bool column_even = (gcol % 2) == 0, row_even = (row % 2) == 0;
                if (column_even != row_even)
                    ignored = true;
            }
            const float incl = s_incl[row];
            double caz;
            if (s_cls[row] == 0)
            {
                caz = (static_cast<double>(gcol) + 0.5) * static_cast<double>(cfg.width); // cpp:371-372
                p.cont_az[cell] = caz;
                p.incl[cell] = incl;
            }
            else
                caz = s_caz[row];
            if (caz < min_az)
                min_az = caz;
            const uchar4 packed = make_uchar4(label, static_cast<unsigned char>(lab >> 8), ignored ? 1 : 0, s_int[row]);
            p.lab[cell] = packed;
            if (EXPORT_LABELS && d_labels && ci < label_cap)
            {
                d_labels[static_cast<size_t>(ci) * R + row] = packed;
                h_labels[static_cast<size_t>(ci) * R + row] = packed;
            }
            p.assoc[cell] = make_float4(ignored ? nanv : q.x, q.y, q.z, incl);
            p.mad[cell] = ignored ? 0.f : ccm::asinf_glibc(ccm::div_rn(cfg.max_distance, q.w));
            if (!ignored)
                s_cls[row] |= 0x80; // read back by this very lane below
            else // never takes part in the association: no parent, nothing visited
            {
                s_parent[static_cast<size_t>(ci) * R + row] = CC_NONE;
                p.visited[cell] = 0;
                p.vback[cell] = 0;
            }
        }
        // the column's non-ignored points join the list the association probe works through (one atomic per column)
        int ntake = 0;
        for (int row0 = 0; row0 < R; row0 += CC_WARP)
        {
            const int row = row0 + lane;
            ntake += __popc(__ballot_sync(CC_FULL_MASK, row < R && (s_cls[row < R ? row : 0] & 0x80) != 0));
        }
        if (ntake)
        {
            int pos0 = 0;
            if (lane == 0)
                pos0 = atomicAdd(&p.st->n_probe, ntake);
            pos0 = __shfl_sync(CC_FULL_MASK, pos0, 0);
            for (int row0 = 0; row0 < R; row0 += CC_WARP)
            {
                const int row = row0 + lane;
                const bool take = row < R && (s_cls[row < R ? row : 0] & 0x80) != 0;
                const unsigned int m = __ballot_sync(CC_FULL_MASK, take);
                if (take)
                    p.probe_list[pos0 + __popc(m & ((1u << lane) - 1u))] = ci * R + row;
                pos0 += __popc(m);
            }
        }
        min_az = cc_warp_min_f64(min_az);
        if (lane == 0)
        {
            p.col_minaz[ci] = min_az;
            p.slot_gcol[local] = gcol;
            p.col_flag[ci] = 0;
        }
        __syncwarp();
    }

}
__global__ void __launch_bounds__(128, 8) k_ground(CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent)
{
    CC_PDL_ENTER();
    d_ground<false>(cc_grid(), cfg, p, s_parent);
}

// Running maximum of the columns' minimum azimuth (the value every finish pass compares against, cpp:884-885), continued
// across pushes: every CTA of the association probe computes it for itself in shared memory (one coalesced read of the
// per-column minima and a block scan), instead of all of them waiting for one CTA to do it. `sh` holds ncols doubles
// behind `part` (block-scan scratch, one double per warp). CTA 0 also leaves it in col_runmax for the kernels after.
CC_DEV void d_runmax_block(const CcDevPtrs& p, int ncols, double* part, double* sh, bool write_out)
{
    const int T = blockDim.x, t = threadIdx.x;
    const double carry = p.st->runmax_carry;
    for (int i0 = 0; i0 < ncols; i0 += 16 * T)
    {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; u++)
            v[u] = i0 + u * T + t < ncols ? p.col_minaz[i0 + u * T + t] : -1.0;
#pragma unroll
        for (int u = 0; u < 16; u++)
            if (i0 + u * T + t < ncols)
                sh[i0 + u * T + t] = v[u];
    }
    __syncthreads();
    const int seg = ((ncols + T - 1) / T) | 1; // odd: the threads' segments start in different banks
    const int lo = t * seg < ncols ? t * seg : ncols, hi = (lo + seg < ncols) ? lo + seg : ncols;
    double m = -1.0;
    for (int i = lo; i < hi; i++)
        m = sh[i] > m ? sh[i] : m;
    double pre = cc_block_exclusive_scan(part, m, -1.0, CcOpMaxF64());
    pre = carry > pre ? carry : pre;
    for (int i = lo; i < hi; i++)
    {
        const double v = sh[i];
        pre = v > pre ? v : pre;
        sh[i] = pre;
    }
    __syncthreads();
    if (write_out)
        for (int i = t; i < ncols; i += T)
            p.col_runmax[i] = sh[i];
}

CC_DEV void d_snapshot(const CcGrid g, CcDevPtrs p, int spec);

// =====================================================================================================
// K3a  association probe (cpp:698-835): every non-ignored cell of the new columns walks its field of view and
//      records its first hit (the tree it joins) and every later hit (tree<->tree links). The walk is purely
//      geometric as long as the reference never REFUSES an association (cpp:654-659, 688-690); every hit that
//      could have been refused -- its target's finish azimuth is not beyond the largest column-minimum azimuth
//      seen before this column, or its tree is already finished -- flags the column for the column-sequential
//      exact path. No persistent state is modified here.
// =====================================================================================================
#define CC_PROBE_PIPE 6   /* vertical runs fetched ahead per warp by the cooperative walk */
#define CC_PROBE_BUDGET 12 /* cells a point may visit on the thread-per-point path before the warp takes it over */
#ifdef CC_EMU
#define CC_PROBE_PPW 1
#else
#define CC_PROBE_PPW 8 /* points per warp on the thread-per-point path: one round of the resident warps at 4096 columns */
#endif

// Warp-cooperative walk of ONE point with all lanes: every vertical run of the walk (cpp:716-750) is evaluated 32 cells
// at once (inclination break, 3-D distance predicate) and the sequential early-exit rules are applied to the ballot
// masks. The walk order, the visit count and which hit is first are those of the reference. The cells of the next
// CC_PROBE_PIPE runs are fetched ahead with cp.async into a per-warp ring in shared memory (lane j holds cell j of a
// run): an isolated point walks its whole window, 2 * max_steps_in_row + 1 runs.
// `max_back` >= 0 (count-only mode, d_visited_fix): the walk stops after that many columns back -- the reference's stop
// at the first unpublished column (cpp:762-763) -- and only number_of_visited_neighbors is written.
CC_DEV void d_probe_coop(const CcDevCfg& cfg, const CcDevPtrs& p, unsigned int* s_parent, unsigned int* s_links, float4* ring,
                         int base_local, long long colbase, int lane, int pidx, int max_back = -1)
{
    const bool count_only = max_back >= 0;
    const int R = cfg.R, msr = cfg.max_steps_row;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int pci = pidx / R, prow = pidx - pci * R;
    int plocal = base_local + pci;
    if (plocal >= cfg.ringcols)
        plocal -= cfg.ringcols;
    const unsigned int pq = static_cast<unsigned int>(plocal) * R + prow;
    const float4 pa = p.assoc[pq];
    const float ax = pa.x, ay = pa.y, az = pa.z, aw = pa.w;
    const float mad = p.mad[pq];
    const double prev_runmax = pci > 0 ? p.col_runmax[pci - 1] : p.st->runmax_carry;
    int steps_back = static_cast<int>(ceilf(ccm::div_rn(mad, cfg.width)));
    steps_back = steps_back < msr ? steps_back : msr;
    steps_back = steps_back < 0 ? 0 : steps_back;
    if (count_only && steps_back > max_back)
        steps_back = max_back;
    // run r of the walk order (cpp:707-719): r = 0 is the own column upwards; then for every column further back the
    // upward run (from the own row) and the downward run
    const int nruns = 1 + 2 * steps_back;
    int reached = 0; // columns back whose runs were walked
    auto issue = [&](int r)
    {
        if (r < nruns)
        {
            const int back = (r + 1) >> 1, dir = (r == 0 || (r & 1)) ? -1 : 1;
            const int start_step = (dir == 1 || back == 0) ? 1 : 0;
            const int start_row = (dir == 1 || back == 0) ? prow + dir : prow;
            int n = cfg.max_steps_col - start_step + 1;
            const int room = dir < 0 ? start_row + 1 : R - start_row;
            n = n < room ? n : room;
            int olocal = plocal - back;
            if (olocal < 0)
                olocal += cfg.ringcols;
            if (lane < n)
                __pipeline_memcpy_async(ring + (r % CC_PROBE_PIPE) * CC_WARP + lane,
                                        p.assoc + static_cast<size_t>(olocal) * R + (start_row + dir * lane), 16);
        }
        __pipeline_commit();
    };
    for (int r = 0; r < CC_PROBE_PIPE; r++)
        issue(r);
    unsigned int first = CC_NONE;
    int visited = 0, nlinks = 0;
    bool flagged = false;
    for (int r = 0; r < nruns; r++)
    {
        const int back = (r + 1) >> 1, dir = (r == 0 || (r & 1)) ? -1 : 1;
        reached = back;
        int olocal = plocal - back;
        if (olocal < 0)
            olocal += cfg.ringcols; // cpp:768-769
        {
            const int start_step = (dir == 1 || back == 0) ? 1 : 0;
            const int start_row = (dir == 1 || back == 0) ? prow + dir : prow;
            int n = cfg.max_steps_col - start_step + 1; // cells of this vertical run
            const int room = dir < 0 ? start_row + 1 : R - start_row;
            n = n < room ? n : room;
            __pipeline_wait_prior(CC_PROBE_PIPE - 1); // run r has landed (every lane reads only what it fetched itself)
            bool run_done = false;
            for (int jb = 0; jb < n && !run_done; jb += CC_WARP)
            {
                const int cnt = (n - jb) < CC_WARP ? (n - jb) : CC_WARP;
                const int j = jb + lane;
                const bool in = lane < cnt;
                const int orow = start_row + dir * j;
                const unsigned int o = static_cast<unsigned int>(olocal) * R + (in ? orow : 0);
                float4 b = make_float4(cc_nanf(), 0.f, 0.f, cc_nanf());
                if (in)
                    b = jb == 0 ? ring[(r % CC_PROBE_PIPE) * CC_WARP + lane] : p.assoc[o];
                const bool brk = in && fabsf(b.w - aw) > mad; // cpp:728-729 (NaN never breaks)
                bool hit = false;
                if (in && !brk && !cc_isnan(b.x))
                {
                    const float dx = ax - b.x, dy = ay - b.y, dz = az - b.z;
                    hit = dx * dx + dy * dy + dz * dz < cfg.max_distance_sq; // cpp:638-641
                }
                const unsigned int brk_mask = __ballot_sync(CC_FULL_MASK, brk);
                const unsigned int hit_mask = __ballot_sync(CC_FULL_MASK, hit);
                const int limit = brk_mask ? __ffs(brk_mask) - 1 : cnt; // cells before the break are processed
                const unsigned int below_limit = limit >= 32 ? 0xffffffffu : ((1u << limit) - 1u);
                // early stop once associated (cpp:747-749): after the first cell with steps >= min steps
                int jstop = 0x7fffffff;
                if (cfg.stop_enabled)
                {
                    const int need = cfg.stop_min_steps - start_step - jb; // lane index where steps >= min
                    if (first != CC_NONE)
                        jstop = need > 0 ? need : 0;
                    else if (hit_mask & below_limit)
                    {
                        const int jh = __ffs(hit_mask & below_limit) - 1;
                        jstop = jh > need ? jh : need;
                    }
                }
                int processed = limit; // number of cells whose association step runs
                bool stopped = false;
                if (jstop < limit)
                {
                    processed = jstop + 1;
                    stopped = true;
                }
                visited += stopped ? processed : (limit < cnt ? limit + 1 : cnt);
                const unsigned int proc_mask = processed >= 32 ? 0xffffffffu : ((1u << processed) - 1u);
                unsigned int hp = hit_mask & proc_mask;
                if (hp)
                {
                    if (!count_only && ((hp >> lane) & 1u)) // every hit lane checks its own target
                    {
                        // could the reference have refused this hit?
                        const double finish_o = p.cont_az[o] + static_cast<double>(p.mad[o]);
                        if (finish_o <= prev_runmax)
                            flagged = true;
                        if (back > pci)
                        {
                            const unsigned int ro = p.tparent[o];
                            if (ro == CC_NONE || p.tstate[ro] != 0)
                                flagged = true;
                        }
                    }
                    if (first == CC_NONE)
                    {
                        const int jf = __ffs(hp) - 1;
                        first = __shfl_sync(CC_FULL_MASK, o, jf);
                        hp &= hp - 1;
                    }
                    // the remaining hits are tree<->tree link candidates (cpp:740-741)
                    if (!count_only && ((hp >> lane) & 1u))
                    {
                        const int slot = nlinks + __popc(hp & lt_mask);
                        if (slot < CC_LINK_SLOTS)
                            s_links[static_cast<size_t>(pidx) * CC_LINK_SLOTS + slot] = o;
                        else
                        {
                            const int e = atomicAdd(&p.st->n_edges, 1);
                            if (e < p.cap_edges)
                            {
                                p.edge_a[e] = pq;
                                p.edge_b[e] = o;
                            }
                            else
                                p.st->error = CC_DEV_LIST_OVERFLOW;
                        }
                    }
                    nlinks += __popc(hp);
                }
                run_done = stopped || limit < cnt;
            }
        }
        issue(r + CC_PROBE_PIPE); // into the slot just consumed
        if ((r == 0 || !(r & 1)) && first != CC_NONE && cfg.stop_enabled && back >= cfg.stop_min_steps)
            break; // cpp:757-759, after both runs of a column
    }
    __pipeline_wait_prior(0); // nothing in flight when the next point starts filling the ring
    if (count_only)
    {
        if (lane == 0)
        {
            p.visited[pq] = static_cast<unsigned short>(visited);
            p.vback[pq] = 0;
        }
        return;
    }
    const bool any_flag = __ballot_sync(CC_FULL_MASK, flagged) != 0 ||
                          (cfg.debug_flag_period > 0 && ((colbase + pci) % cfg.debug_flag_period) == 0);
    for (int jl = nlinks + lane; jl < CC_LINK_SLOTS; jl += CC_WARP)
        s_links[static_cast<size_t>(pidx) * CC_LINK_SLOTS + jl] = CC_NONE;
    if (lane == 0)
    {
        s_parent[pidx] = first == CC_NONE ? pq : first;
        p.visited[pq] = static_cast<unsigned short>(visited);
        p.vback[pq] = static_cast<unsigned char>(reached);
        if (any_flag)
        {
            p.col_flag[pci] = 1;
            atomicAdd(&p.st->n_flagged, 1);
        }
    }
}

// K3a: every non-ignored point of the new columns (the list k_ground compacted). Most points visit a handful of cells
// (the first neighbour above or in the previous column associates them and the walk stops, cpp:747-759): those are
// walked by ONE thread each, literally as the reference does, with a budget of CC_PROBE_BUDGET cells. A point that
// exceeds the budget (sparse surroundings: the walk covers up to (2 * max_steps_in_row + 1) * max_steps_in_column
// cells) or finds more link candidates than it has slots is redone from scratch by a whole warp (k_probe_heavy).
CC_DEV void d_probe(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, unsigned int* s_links, int do_snapshot)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_probe, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    if (do_snapshot) // list-root state before the speculative commit (rolled back if it aborts); the probe itself
        d_snapshot(g, p, 0); // modifies no persistent state
    const int R = cfg.R;
    const int ncols = hd.ncols;
    const long long colbase = hd.colbase;
    const int base_local = ncols > 0 ? cc_local_col(colbase, cfg.ringcols) : 0;
    const int lane = threadIdx.x % CC_WARP;
    const int npoints = p.st->n_probe < p.maxcols * R ? p.st->n_probe : p.maxcols * R;
    const int gwarp = (g.bid * blockDim.x + threadIdx.x) / CC_WARP;
    const int nwarps = (g.nb * blockDim.x + CC_WARP - 1) / CC_WARP;
    const int ngroups = (npoints + CC_PROBE_PPW - 1) / CC_PROBE_PPW;
    CC_SMEM(smem);
    double* rm_part = reinterpret_cast<double*>(smem);
    double* rm = rm_part + 32; // running maximum of the column minima, columns of this push
    const int ncols_rm = ncols < p.maxcols ? ncols : p.maxcols;
    d_runmax_block(p, ncols_rm, rm_part, rm, g.bid == 0);
    const double rm_carry = p.st->runmax_carry;
    for (int grp = gwarp; grp < ngroups; grp += nwarps)
    {
        // the points of a warp are far apart in the list: points that need the cooperative walk come in runs (the
        // neighbouring cells of a sparse region), which one warp would have to work through one after the other
        const int pi = lane * ngroups + grp;
        const bool has = lane < CC_PROBE_PPW && pi < npoints;
        bool heavy = false;
        int pidx = 0;
        if (has)
        {
            pidx = p.probe_list[pi];
            const int pci = pidx / R, prow = pidx - pci * R;
            int plocal = base_local + pci;
            if (plocal >= cfg.ringcols)
                plocal -= cfg.ringcols;
            const unsigned int pq = static_cast<unsigned int>(plocal) * R + prow;
            const float4 a = p.assoc[pq];
            const float mad = p.mad[pq];
            const double prev_runmax = pci > 0 ? rm[pci - 1] : rm_carry;
            int steps_back = static_cast<int>(ceilf(ccm::div_rn(mad, cfg.width)));
            steps_back = steps_back < cfg.max_steps_row ? steps_back : cfg.max_steps_row;
            unsigned int first = CC_NONE, l0 = CC_NONE, l1 = CC_NONE, l2 = CC_NONE, l3 = CC_NONE;
            static_assert(CC_LINK_SLOTS == 4, "four link registers below");
            int nl = 0, visited = 0;
            bool flagged = false;
            int ocol = plocal, reached = 0;
            for (int back = 0; back <= steps_back; back++)
            {
                reached = back;
                for (int dir = -1; dir <= 1 && !heavy; dir += 2)
                {
                    if (dir == 1 && back == 0)
                        continue;
                    int steps_v = (dir == 1 || back == 0) ? 1 : 0;
                    int orow = (dir == 1 || back == 0) ? prow + dir : prow;
                    while (orow >= 0 && orow < R && steps_v <= cfg.max_steps_col)
                    {
                        if (visited >= CC_PROBE_BUDGET)
                        {
                            heavy = true;
                            break;
                        }
                        const unsigned int o = static_cast<unsigned int>(ocol) * R + orow;
                        const float4 b = p.assoc[o];
                        visited++;
                        if (fabsf(b.w - a.w) > mad) // cpp:728-729 (NaN never breaks)
                            break;
                        if (!cc_isnan(b.x))
                        {
                            const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
                            if (dx * dx + dy * dy + dz * dz < cfg.max_distance_sq) // cpp:638-641
                            {
                                // could the reference have refused this hit?
                                const double finish_o = p.cont_az[o] + static_cast<double>(p.mad[o]);
                                if (finish_o <= prev_runmax)
                                    flagged = true;
                                if (back > pci)
                                {
                                    const unsigned int ro = p.tparent[o];
                                    if (ro == CC_NONE || p.tstate[ro] != 0)
                                        flagged = true;
                                }
                                if (first == CC_NONE)
                                    first = o;
                                else if (nl == 0)
                                    l0 = o, nl = 1;
                                else if (nl == 1)
                                    l1 = o, nl = 2;
                                else if (nl == 2)
                                    l2 = o, nl = 3;
                                else if (nl == 3)
                                    l3 = o, nl = 4;
                                else
                                {
                                    heavy = true; // the overflow list is the cooperative walk's business
                                    break;
                                }
                            }
                        }
                        if (first != CC_NONE && cfg.stop_enabled && steps_v >= cfg.stop_min_steps) // cpp:747-749
                            break;
                        orow += dir;
                        steps_v++;
                    }
                }
                if (heavy)
                    break;
                if (first != CC_NONE && cfg.stop_enabled && back >= cfg.stop_min_steps) // cpp:757-759
                    break;
                ocol--;
                if (ocol < 0)
                    ocol += cfg.ringcols; // cpp:768-769
            }
            if (!heavy)
            {
                unsigned int* lk = s_links + static_cast<size_t>(pidx) * CC_LINK_SLOTS;
                lk[0] = l0;
                lk[1] = l1;
                lk[2] = l2;
                lk[3] = l3;
                s_parent[pidx] = first == CC_NONE ? pq : first;
                p.visited[pq] = static_cast<unsigned short>(visited);
                p.vback[pq] = static_cast<unsigned char>(reached);
                if (flagged || (cfg.debug_flag_period > 0 && ((colbase + pci) % cfg.debug_flag_period) == 0))
                {
                    p.col_flag[pci] = 1;
                    atomicAdd(&p.st->n_flagged, 1);
                }
            }
        }
        // points for the cooperative walk go to the list k_probe_heavy works through, one warp per point
        const unsigned int hm = __ballot_sync(CC_FULL_MASK, has && heavy);
        if (hm)
        {
            int pos0 = 0;
            if (lane == __ffs(hm) - 1)
                pos0 = atomicAdd(&p.st->n_heavy, __popc(hm));
            pos0 = __shfl_sync(CC_FULL_MASK, pos0, __ffs(hm) - 1);
            if (has && heavy)
                p.heavy_list[pos0 + __popc(hm & ((1u << lane) - 1u))] = pidx;
        }
    }
}
__global__ void __launch_bounds__(256) k_probe(CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, unsigned int* s_links, int do_snapshot)
{
    CC_PDL_ENTER();
    d_probe(cc_grid(), cfg, p, s_parent, s_links, do_snapshot);
}

// K3a (tiled): the same walk, one thread per CELL of a tile of consecutive new columns, with the tile's whole field of
// view -- the tile's columns and the max_steps_in_row columns before them, one contiguous span of `assoc` in the ring
// (two where the ring wraps) -- staged in shared memory by ONE bulk asynchronous copy (cp.async.bulk + mbarrier): the
// dependent loads of a walk then cost a shared-memory access instead of an L2 round trip, all 32 lanes of a warp walk
// (they are neighbouring rows of one column: similar walks), and the budget before a point is handed to the
// cooperative walk is CC_TILE_BUDGET cells instead of CC_PROBE_BUDGET. While the copy is in flight the CTA reduces the
// column minima before its tile (the validator's running maximum). One extra CTA leaves the running maxima of all
// columns in col_runmax for the kernels after this one. Results are those of d_probe, cell for cell.
#define CC_TILE_CELLS 256 /* cells (threads) per tile: 256 / R columns */
#define CC_TILE_BUDGET 16
static inline __host__ __device__ int cc_tile_cols(int R)
{
    return CC_TILE_CELLS / R > 0 ? CC_TILE_CELLS / R : 1;
}
static inline __host__ __device__ size_t cc_tile_smem_bytes(int R, int max_steps_row)
{
    return static_cast<size_t>(cc_tile_cols(R) + max_steps_row) * R * (sizeof(float4) + sizeof(double) + sizeof(float)) +
           (32 + 64) * sizeof(double) + 16;
}
#ifndef CC_EMU
CC_DEV unsigned int cc_smem_addr(const void* q)
{
    return static_cast<unsigned int>(__cvta_generic_to_shared(q));
}
CC_DEV void cc_mbar_init(unsigned long long* bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cc_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
CC_DEV void cc_mbar_expect_tx(unsigned long long* bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cc_smem_addr(bar)), "r"(bytes) : "memory");
}
CC_DEV void cc_bulk_g2s(void* dst_smem, const void* src_global, unsigned int bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cc_smem_addr(dst_smem)),
                 "l"(src_global), "r"(bytes), "r"(cc_smem_addr(bar))
                 : "memory");
}
CC_DEV void cc_mbar_wait(unsigned long long* bar, unsigned int parity)
{
    unsigned int done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(cc_smem_addr(bar)), "r"(parity)
                     : "memory");
}
#endif

CC_DEV void d_probe_tile(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, unsigned int* s_links, int do_snapshot)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_probe, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    if (do_snapshot) // list-root state before the speculative commit (rolled back if it aborts)
        d_snapshot(g, p, 0);
    const int R = cfg.R, msr = cfg.max_steps_row;
    const int ncols = hd.ncols < p.maxcols ? hd.ncols : p.maxcols;
    const long long colbase = hd.colbase;
    const int base_local = ncols > 0 ? cc_local_col(colbase, cfg.ringcols) : 0;
    const int T = blockDim.x, t = threadIdx.x;
    const double rm_carry = p.st->runmax_carry;
    CC_SMEM(smem);
    const int tile_cols = cc_tile_cols(R), wcols = tile_cols + msr;
    float4* win = reinterpret_cast<float4*>(smem);                                  // [wcols][R] x, y, z, inclination
    double* win_caz = reinterpret_cast<double*>(win + static_cast<size_t>(wcols) * R); // [wcols][R] continuous azimuth
    double* part = win_caz + static_cast<size_t>(wcols) * R; // block-scan scratch (32) + per-tile maxima (64)
    double* rm_tile = part + 32; // running maximum BEFORE every column of the tile
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(rm_tile + 64);
    float* win_mad = reinterpret_cast<float*>(bar + 2);                             // [wcols][R] maximum azimuth difference
    const int n_workers = g.nb > 1 ? g.nb - 1 : 1;
    if (g.nb > 1 && g.bid == g.nb - 1)
    {
        // the extra CTA: running maximum of the column minima of every column of the push, for k_probe_heavy and the finish pass
        const int per = (ncols + T - 1) / T;
        const int lo = t * per < ncols ? t * per : ncols, hi = lo + per < ncols ? lo + per : ncols;
        double m = -1.0;
        for (int i = lo; i < hi; i++)
        {
            const double v = p.col_minaz[i];
            m = v > m ? v : m;
        }
        double pre = cc_block_exclusive_scan(part, m, -1.0, CcOpMaxF64());
        pre = rm_carry > pre ? rm_carry : pre;
        for (int i = lo; i < hi; i++)
        {
            const double v = p.col_minaz[i];
            pre = v > pre ? v : pre;
            p.col_runmax[i] = pre;
        }
        return;
    }
    const int ntiles = (ncols + tile_cols - 1) / tile_cols;
#ifndef CC_EMU
    if (t == 0)
        cc_mbar_init(bar, 1);
    __syncthreads();
#endif
    unsigned int phase = 0;
    for (int tile = g.bid; tile < ntiles; tile += n_workers)
    {
        const int ci0 = tile * tile_cols;
        // ---- the tile's field of view: ring columns [w0, w0 + wcols), wrapped ----
        int w0 = base_local + ci0 - msr;
        w0 %= cfg.ringcols;
        if (w0 < 0)
            w0 += cfg.ringcols;
        const int n1 = wcols < cfg.ringcols - w0 ? wcols : cfg.ringcols - w0; // columns up to the end of the ring
#ifdef CC_EMU
        for (int i = t; i < wcols * R; i += T)
        {
            const int wc = i / R;
            const int lc = wc < n1 ? w0 + wc : wc - n1;
            win[i] = p.assoc[static_cast<size_t>(lc) * R + (i - wc * R)];
            win_caz[i] = p.cont_az[static_cast<size_t>(lc) * R + (i - wc * R)];
            win_mad[i] = p.mad[static_cast<size_t>(lc) * R + (i - wc * R)];
        }
#else
        if (t == 0)
        {
            const unsigned int cells1 = static_cast<unsigned int>(n1 * R), cells2 = static_cast<unsigned int>((wcols - n1) * R);
            cc_mbar_expect_tx(bar, static_cast<unsigned int>(wcols * R * (sizeof(float4) + sizeof(double) + sizeof(float))));
            cc_bulk_g2s(win, p.assoc + static_cast<size_t>(w0) * R, cells1 * sizeof(float4), bar);
            cc_bulk_g2s(win_caz, p.cont_az + static_cast<size_t>(w0) * R, cells1 * sizeof(double), bar);
            cc_bulk_g2s(win_mad, p.mad + static_cast<size_t>(w0) * R, cells1 * sizeof(float), bar);
            if (cells2)
            {
                cc_bulk_g2s(win + cells1, p.assoc, cells2 * sizeof(float4), bar);
                cc_bulk_g2s(win_caz + cells1, p.cont_az, cells2 * sizeof(double), bar);
                cc_bulk_g2s(win_mad + cells1, p.mad, cells2 * sizeof(float), bar);
            }
        }
#endif
        // ---- while it lands: the largest column minimum before the tile ----
        {
            double m = -1.0;
            for (int i = t; i < ci0; i += T)
            {
                const double v = p.col_minaz[i];
                m = v > m ? v : m;
            }
            const double pre = cc_block_exclusive_scan(part, m, -1.0, CcOpMaxF64());
            if (t == T - 1)
            {
                double run = pre > m ? pre : m;
                run = rm_carry > run ? rm_carry : run;
                for (int k = 0; k < tile_cols && k < 64; k++)
                {
                    rm_tile[k] = run;
                    if (ci0 + k < ncols)
                    {
                        const double v = p.col_minaz[ci0 + k];
                        run = v > run ? v : run;
                    }
                }
            }
        }
        __syncthreads();
#ifndef CC_EMU
        cc_mbar_wait(bar, phase);
        phase ^= 1u;
#endif
        for (int cell0 = 0; cell0 < tile_cols * R; cell0 += T) // (whole warps stay in the loop: ballot below)
        {
            const int cell = cell0 + t;
            const int tcol = cell / R, prow = cell - tcol * R;
            const int pci = ci0 + tcol;
            bool heavy = false;
            const int pidx = pci * R + prow;
            bool has = cell < tile_cols * R && pci < ncols;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has)
            {
                a = win[static_cast<size_t>(msr + tcol) * R + prow];
                has = !cc_isnan(a.x);
            }
            if (has)
            {
                int plocal = base_local + pci;
                if (plocal >= cfg.ringcols)
                    plocal -= cfg.ringcols;
                const unsigned int pq = static_cast<unsigned int>(plocal) * R + prow;
                const float mad = win_mad[static_cast<size_t>(msr + tcol) * R + prow];
                const double prev_runmax = rm_tile[tcol < 64 ? tcol : 63];
                int steps_back = static_cast<int>(ceilf(ccm::div_rn(mad, cfg.width)));
                steps_back = steps_back < msr ? steps_back : msr;
                unsigned int first = CC_NONE, l0 = CC_NONE, l1 = CC_NONE, l2 = CC_NONE, l3 = CC_NONE;
                int nl = 0, visited = 0;
                bool flagged = false;
                int ocol = plocal, reached = 0;
                for (int back = 0; back <= steps_back; back++)
                {
                    reached = back;
                    const size_t wofs = static_cast<size_t>(msr + tcol - back) * R;
                    const float4* wcol = win + wofs;
                    for (int dir = -1; dir <= 1 && !heavy; dir += 2)
                    {
                        if (dir == 1 && back == 0)
                            continue;
                        int steps_v = (dir == 1 || back == 0) ? 1 : 0;
                        int orow = (dir == 1 || back == 0) ? prow + dir : prow;
                        while (orow >= 0 && orow < R && steps_v <= cfg.max_steps_col)
                        {
                            if (visited >= CC_TILE_BUDGET)
                            {
                                heavy = true;
                                break;
                            }
                            const float4 b = wcol[orow];
                            visited++;
                            if (fabsf(b.w - a.w) > mad) // cpp:728-729 (NaN never breaks)
                                break;
                            if (!cc_isnan(b.x))
                            {
                                const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
                                if (dx * dx + dy * dy + dz * dz < cfg.max_distance_sq) // cpp:638-641
                                {
                                    const unsigned int o = static_cast<unsigned int>(ocol) * R + orow;
                                    // could the reference have refused this hit?
                                    const double finish_o = win_caz[wofs + orow] + static_cast<double>(win_mad[wofs + orow]);
                                    if (finish_o <= prev_runmax)
                                        flagged = true;
                                    if (back > pci)
                                    {
                                        const unsigned int ro = p.tparent[o];
                                        if (ro == CC_NONE || p.tstate[ro] != 0)
                                            flagged = true;
                                    }
                                    if (first == CC_NONE)
                                        first = o;
                                    else if (nl == 0)
                                        l0 = o, nl = 1;
                                    else if (nl == 1)
                                        l1 = o, nl = 2;
                                    else if (nl == 2)
                                        l2 = o, nl = 3;
                                    else if (nl == 3)
                                        l3 = o, nl = 4;
                                    else
                                    {
                                        heavy = true; // the overflow list is the cooperative walk's business
                                        break;
                                    }
                                }
                            }
                            if (first != CC_NONE && cfg.stop_enabled && steps_v >= cfg.stop_min_steps) // cpp:747-749
                                break;
                            orow += dir;
                            steps_v++;
                        }
                    }
                    if (heavy)
                        break;
                    if (first != CC_NONE && cfg.stop_enabled && back >= cfg.stop_min_steps) // cpp:757-759
                        break;
                    ocol--;
                    if (ocol < 0)
                        ocol += cfg.ringcols; // cpp:768-769
                }
                if (!heavy)
                {
                    unsigned int* lk = s_links + static_cast<size_t>(pidx) * CC_LINK_SLOTS;
                    lk[0] = l0;
                    lk[1] = l1;
                    lk[2] = l2;
                    lk[3] = l3;
                    s_parent[pidx] = first == CC_NONE ? pq : first;
                    p.visited[pq] = static_cast<unsigned short>(visited);
                    p.vback[pq] = static_cast<unsigned char>(reached);
                    if (flagged || (cfg.debug_flag_period > 0 && ((colbase + pci) % cfg.debug_flag_period) == 0))
                    {
                        p.col_flag[pci] = 1;
                        atomicAdd(&p.st->n_flagged, 1);
                    }
                }
            }
            // points for the cooperative walk go to the list k_probe_heavy works through
#ifdef CC_EMU
            if (has && heavy)
                p.heavy_list[atomicAdd(&p.st->n_heavy, 1)] = pidx;
#else
            const unsigned int hm = __ballot_sync(CC_FULL_MASK, has && heavy);
            if (hm)
            {
                const int lane = t % CC_WARP;
                int pos0 = 0;
                if (lane == __ffs(hm) - 1)
                    pos0 = atomicAdd(&p.st->n_heavy, __popc(hm));
                pos0 = __shfl_sync(CC_FULL_MASK, pos0, __ffs(hm) - 1);
                if (has && heavy)
                    p.heavy_list[pos0 + __popc(hm & ((1u << lane) - 1u))] = pidx;
            }
#endif
        }
        __syncthreads(); // the window is overwritten by the next tile's copy
    }
}
__global__ void __launch_bounds__(CC_TILE_CELLS) k_probe_tile(CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, unsigned int* s_links,
                                                            int do_snapshot)
{
    CC_PDL_ENTER();
    d_probe_tile(cc_grid(), cfg, p, s_parent, s_links, do_snapshot);
}

// K3a': the points the thread-per-point path gave up on, one CTA (two warps) per point. When a vertical run fits one warp step
// (max_steps_in_column < 32) and the window has at most 64 runs:
//   phase 1  the warps of the CTA share the runs of the window: independent loads, inclination-break and distance
//            predicates of every run at once, their ballot masks left in shared memory;
//   phase 2  warp 0 applies the sequential rules to the masks. As long as the point is not associated a run without a
//            hit before its break only adds to the visit count, so the lanes handle those "quiet" runs in parallel and
//            the literal rules (cpp:725-759) start at the first run with a hit.
// A point without neighbours -- the worst case, it walks its whole window -- costs one memory round trip and a few
// hundred instructions instead of ~100 dependent instructions per run. Other configurations: warp 0 walks alone
// (d_probe_coop).
// The warps of a CTA work in TEAMS of `team_warps` warps, one point per team at a time: the whole CTA (2 warps) when
// the stage runs as its own kernel, pairs of warps of a wide CTA inside the fused kernel.
static inline __host__ __device__ size_t cc_heavy_smem_bytes(int block_threads, int team_warps)
{
    const int nwarps = (block_threads + CC_WARP - 1) / CC_WARP;
    const int teams = nwarps / (team_warps > 0 ? team_warps : 1) > 0 ? nwarps / (team_warps > 0 ? team_warps : 1) : 1;
    return static_cast<size_t>(teams) * (CC_PROBE_PIPE * CC_WARP * sizeof(float4) + 128 * sizeof(unsigned int));
}
template<bool TEAMS>
CC_DEV void cc_team_sync(int team_warps, int nwarps, int team)
{
#ifndef CC_EMU
    if (!TEAMS || team_warps >= nwarps)
        __syncthreads();
    else if (team_warps == 1)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(team_warps * CC_WARP) : "memory");
#else
    (void)team_warps, (void)nwarps, (void)team;
#endif
}

// TEAMS = false: the team is the whole CTA (the stage as its own kernel: plain block barriers, no named barriers reserved)
// COUNT: count-only mode of k_visited_fix -- the list is (heavy_list[e], probe_list[e]) = (point, columns back at which the
// reference's walk stops, cpp:762-763), n_vfix entries; only number_of_visited_neighbors is written.
template<bool TEAMS, bool COUNT = false>
CC_DEV void d_probe_heavy(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, unsigned int* s_links, int tune, int team_warps)
{
    CcTraceScope cc_trace_scope(p.trace, COUNT ? CC_KID_visited_fix : CC_KID_probe_heavy, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    const int R = cfg.R, msr = cfg.max_steps_row;
    const int ncols = hd.ncols;
    const long long colbase = hd.colbase;
    const int base_local = ncols > 0 ? cc_local_col(colbase, cfg.ringcols) : 0;
    const int lane = threadIdx.x % CC_WARP;
    const int cta_warps = (blockDim.x + CC_WARP - 1) / CC_WARP;
    const int nwarps = TEAMS && team_warps < cta_warps ? team_warps : cta_warps; // warps of my team
    const int teams_per_cta = cta_warps / nwarps;
    const int team = (threadIdx.x / CC_WARP) / nwarps, warp = (threadIdx.x / CC_WARP) % nwarps;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int n_listed = COUNT ? p.st->n_vfix : p.st->n_heavy;
    const int nheavy = n_listed < p.maxcols * R ? n_listed : p.maxcols * R;
    CC_SMEM(smem);
    float4* ring = reinterpret_cast<float4*>(smem) + (TEAMS ? static_cast<size_t>(team) * CC_PROBE_PIPE * CC_WARP : 0);
    __shared__ unsigned int sh_masks[128]; // the whole CTA is one team: fixed addresses
    unsigned int* brk_m = TEAMS ? reinterpret_cast<unsigned int*>(reinterpret_cast<float4*>(smem) +
                                                                  static_cast<size_t>(teams_per_cta) * CC_PROBE_PIPE * CC_WARP) + team * 128
                                : sh_masks;
    unsigned int* hit_m = brk_m + 64;
    const bool masks_ok = CC_WARP == 32 && cfg.max_steps_col < 32 && 2 * msr + 1 <= 64 && !(tune & 1);
    if (team >= teams_per_cta)
        return; // warps beyond the last full team
    for (int hi = g.bid * teams_per_cta + team; hi < nheavy; hi += g.nb * teams_per_cta)
    {
        const int pidx = p.heavy_list[hi];
        const int max_back = COUNT ? p.probe_list[hi] : -1;
        if (!masks_ok)
        {
            if (warp == 0)
                d_probe_coop(cfg, p, s_parent, s_links, ring, base_local, colbase, lane, pidx, max_back);
            continue;
        }
        const int pci = pidx / R, prow = pidx - pci * R;
        int plocal = base_local + pci;
        if (plocal >= cfg.ringcols)
            plocal -= cfg.ringcols;
        const unsigned int pq = static_cast<unsigned int>(plocal) * R + prow;
        const float4 pa = p.assoc[pq];
        const float ax = pa.x, ay = pa.y, az = pa.z, aw = pa.w;
        const float mad = p.mad[pq];
        int steps_back = static_cast<int>(ceilf(ccm::div_rn(mad, cfg.width)));
        steps_back = steps_back < msr ? steps_back : msr;
        steps_back = steps_back < 0 ? 0 : steps_back;
        if (COUNT && steps_back > max_back)
            steps_back = max_back;
        const int nruns = 1 + 2 * steps_back; // run r: r = 0 own column upwards; odd r: column (r + 1) / 2 back, upwards
                                              // from the own row; even r: same column downwards (cpp:707-719)
        // ---- phase 1 ----
        for (int r0 = 0; r0 < nruns; r0 += 8 * nwarps)
        {
            float4 b[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                const int r = r0 + warp + u * nwarps;
                b[u] = make_float4(cc_nanf(), 0.f, 0.f, cc_nanf());
                if (r < nruns)
                {
                    const int back = (r + 1) >> 1, dir = (r == 0 || (r & 1)) ? -1 : 1;
                    const int start_step = (dir == 1 || back == 0) ? 1 : 0;
                    const int start_row = (dir == 1 || back == 0) ? prow + dir : prow;
                    int n = cfg.max_steps_col - start_step + 1;
                    const int room = dir < 0 ? start_row + 1 : R - start_row;
                    n = n < room ? n : room;
                    int olocal = plocal - back;
                    if (olocal < 0)
                        olocal += cfg.ringcols; // cpp:768-769
                    if (lane < n)
                        b[u] = p.assoc[static_cast<size_t>(olocal) * R + (start_row + dir * lane)];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                const int r = r0 + warp + u * nwarps;
                // lanes beyond the run hold NaN: neither a break nor a hit
                const bool brk = fabsf(b[u].w - aw) > mad; // cpp:728-729 (NaN never breaks)
                bool hit = false;
                if (!brk && !cc_isnan(b[u].x))
                {
                    const float dx = ax - b[u].x, dy = ay - b[u].y, dz = az - b[u].z;
                    hit = dx * dx + dy * dy + dz * dz < cfg.max_distance_sq; // cpp:638-641
                }
                const unsigned int bm = __ballot_sync(CC_FULL_MASK, brk);
                const unsigned int hm = __ballot_sync(CC_FULL_MASK, hit);
                if (lane == 0 && r < nruns)
                {
                    brk_m[r] = bm;
                    hit_m[r] = hm;
                }
            }
        }
        cc_team_sync<TEAMS>(nwarps, cta_warps, team);
        // ---- phase 2 ----
        if (warp == 0)
        {
            const double prev_runmax = pci > 0 ? p.col_runmax[pci - 1] : p.st->runmax_carry;
            unsigned int first = CC_NONE;
            int visited = 0, nlinks = 0;
            bool flagged = false;
            // quiet prefix: runs without a hit before their break, summed by the lanes
            int q = nruns;
            for (int half = 0; half < 2 && q == nruns; half++)
            {
                const int r = half * 32 + lane;
                bool loud = false;
                int contrib = 0;
                if (r < nruns)
                {
                    const int back = (r + 1) >> 1, dir = (r == 0 || (r & 1)) ? -1 : 1;
                    const int start_step = (dir == 1 || back == 0) ? 1 : 0;
                    const int start_row = (dir == 1 || back == 0) ? prow + dir : prow;
                    int n = cfg.max_steps_col - start_step + 1;
                    const int room = dir < 0 ? start_row + 1 : R - start_row;
                    n = n < room ? n : room;
                    if (n > 0)
                    {
                        const unsigned int brk_mask = brk_m[r];
                        const int limit = brk_mask ? __ffs(brk_mask) - 1 : n;
                        const unsigned int below_limit = limit >= 32 ? 0xffffffffu : ((1u << limit) - 1u);
                        loud = (hit_m[r] & below_limit) != 0;
                        contrib = limit < n ? limit + 1 : n;
                    }
                }
                const unsigned int loud_mask = __ballot_sync(CC_FULL_MASK, loud);
                if (loud_mask)
                {
                    const int l = __ffs(loud_mask) - 1;
                    q = half * 32 + l;
                    visited += __reduce_add_sync(CC_FULL_MASK, lane < l ? contrib : 0);
                }
                else
                    visited += __reduce_add_sync(CC_FULL_MASK, contrib);
            }
            int reached = q >> 1; // columns back whose runs were walked (the quiet prefix is runs [0, q))
            for (int r = q; r < nruns; r++)
            {
                const int back = (r + 1) >> 1, dir = (r == 0 || (r & 1)) ? -1 : 1;
                reached = back;
                const int start_step = (dir == 1 || back == 0) ? 1 : 0;
                const int start_row = (dir == 1 || back == 0) ? prow + dir : prow;
                int n = cfg.max_steps_col - start_step + 1; // cells of this vertical run
                const int room = dir < 0 ? start_row + 1 : R - start_row;
                n = n < room ? n : room;
                if (n > 0)
                {
                    const int cnt = n;
                    const unsigned int brk_mask = brk_m[r], hit_mask = hit_m[r];
                    const int limit = brk_mask ? __ffs(brk_mask) - 1 : cnt; // cells before the break are processed
                    const unsigned int below_limit = limit >= 32 ? 0xffffffffu : ((1u << limit) - 1u);
                    // early stop once associated (cpp:747-749): after the first cell with steps >= min steps
                    int jstop = 0x7fffffff;
                    if (cfg.stop_enabled)
                    {
                        const int need = cfg.stop_min_steps - start_step; // lane index where steps >= min
                        if (first != CC_NONE)
                            jstop = need > 0 ? need : 0;
                        else if (hit_mask & below_limit)
                        {
                            const int jh = __ffs(hit_mask & below_limit) - 1;
                            jstop = jh > need ? jh : need;
                        }
                    }
                    int processed = limit; // number of cells whose association step runs
                    bool stopped = false;
                    if (jstop < limit)
                    {
                        processed = jstop + 1;
                        stopped = true;
                    }
                    visited += stopped ? processed : (limit < cnt ? limit + 1 : cnt);
                    const unsigned int proc_mask = processed >= 32 ? 0xffffffffu : ((1u << processed) - 1u);
                    unsigned int hp = hit_mask & proc_mask;
                    if (hp)
                    {
                        int olocal = plocal - back;
                        if (olocal < 0)
                            olocal += cfg.ringcols;
                        const unsigned int o = static_cast<unsigned int>(olocal) * R + (lane < cnt ? start_row + dir * lane : 0);
                        if (!COUNT && ((hp >> lane) & 1u)) // every hit lane checks its own target
                        {
                            // could the reference have refused this hit?
                            const double finish_o = p.cont_az[o] + static_cast<double>(p.mad[o]);
                            if (finish_o <= prev_runmax)
                                flagged = true;
                            if (back > pci)
                            {
                                const unsigned int ro = p.tparent[o];
                                if (ro == CC_NONE || p.tstate[ro] != 0)
                                    flagged = true;
                            }
                        }
                        if (first == CC_NONE)
                        {
                            const int jf = __ffs(hp) - 1;
                            first = __shfl_sync(CC_FULL_MASK, o, jf);
                            hp &= hp - 1;
                        }
                        // the remaining hits are tree<->tree link candidates (cpp:740-741)
                        if (!COUNT && ((hp >> lane) & 1u))
                        {
                            const int slot = nlinks + __popc(hp & lt_mask);
                            if (slot < CC_LINK_SLOTS)
                                s_links[static_cast<size_t>(pidx) * CC_LINK_SLOTS + slot] = o;
                            else
                            {
                                const int e = atomicAdd(&p.st->n_edges, 1);
                                if (e < p.cap_edges)
                                {
                                    p.edge_a[e] = pq;
                                    p.edge_b[e] = o;
                                }
                                else
                                    p.st->error = CC_DEV_LIST_OVERFLOW;
                            }
                        }
                        nlinks += __popc(hp);
                    }
                }
                if ((r == 0 || !(r & 1)) && first != CC_NONE && cfg.stop_enabled && back >= cfg.stop_min_steps)
                    break; // cpp:757-759, after both runs of a column
            }
            if (COUNT)
            {
                if (lane == 0)
                    p.visited[pq] = static_cast<unsigned short>(visited);
            }
            else
            {
                const bool any_flag = __ballot_sync(CC_FULL_MASK, flagged) != 0 ||
                                      (cfg.debug_flag_period > 0 && ((colbase + pci) % cfg.debug_flag_period) == 0);
                for (int jl = nlinks + lane; jl < CC_LINK_SLOTS; jl += CC_WARP)
                    s_links[static_cast<size_t>(pidx) * CC_LINK_SLOTS + jl] = CC_NONE;
                if (lane == 0)
                {
                    s_parent[pidx] = first == CC_NONE ? pq : first;
                    p.visited[pq] = static_cast<unsigned short>(visited);
                    p.vback[pq] = static_cast<unsigned char>(reached);
                    if (any_flag)
                    {
                        p.col_flag[pci] = 1;
                        atomicAdd(&p.st->n_flagged, 1);
                    }
                }
            }
        }
        cc_team_sync<TEAMS>(nwarps, cta_warps, team); // the masks are rewritten for the next point
    }
}
__global__ void __launch_bounds__(64, 16) k_probe_heavy(CcDevCfg cfg, CcDevPtrs p, unsigned int* s_parent, unsigned int* s_links, int tune)
{
    CC_PDL_ENTER();
    d_probe_heavy<false>(cc_grid(), cfg, p, s_parent, s_links, tune, 2);
}

// ---- union-find over tree roots (lock-free, ECL-CC style: hook the larger index under the smaller) ----
CC_DEV unsigned int cc_uf_find(unsigned int* parent, unsigned int x)
{
    unsigned int cur = cc_vload(parent + x);
    if (cur != x)
    {
        unsigned int prev = x, next;
        while (cur != (next = cc_vload(parent + cur)))
        {
            parent[prev] = next; // path halving; any ancestor is a valid parent
            prev = cur;
            cur = next;
        }
    }
    return cur;
}
CC_DEV void cc_uf_union(unsigned int* parent, unsigned int a, unsigned int b)
{
    while (true)
    {
        a = cc_uf_find(parent, a);
        b = cc_uf_find(parent, b);
        if (a == b)
            return;
        if (a > b)
        {
            const unsigned int t = a;
            a = b;
            b = t;
        }
        if (atomicCAS(parent + b, b, a) == b)
            return;
    }
}

// Speculative kernels run only while no column is flagged and the commit was not aborted.
// guard 0: always run; 1: whole-push speculative commit (skip when any column is flagged, the commit was aborted
// or an error was raised); 2: commit of a sub-range in the split path (skip only after an abort)
CC_DEV bool cc_spec_ok(const CcDevState* st, int guard)
{
    if (st->ncols <= 0)
        return false;
    if (guard == 1)
        return st->n_flagged == 0 && st->abort == 0 && st->error == 0;
    if (guard == 2)
        return st->abort == 0;
    return true;
}

// =====================================================================================================
// K3b  commit of probed columns [ci0, ci1]: (1) copy probe parents, initialise new roots and append them to the
//      unfinished list; (2) resolve every point's tree root by pointer chasing with compression and add its
//      contribution to the root (finished_at, width, tree_num_points: cpp:661-672); (3) apply tree<->tree
//      links (cpp:675-696) to the union-find.
// =====================================================================================================
CC_DEV void d_snapshot(const CcGrid g, CcDevPtrs p, int spec)
{
    if (!cc_spec_ok(p.st, spec))
        return;
    const int n = p.st->n_ulist;
    for (int i = g.bid * blockDim.x + threadIdx.x; i < n; i += g.nb * blockDim.x)
    {
        const unsigned int r = p.ulist[i];
        p.sv_cparent[i] = p.cparent[r];
        p.sv_tfinish[i] = p.tfinish[r];
        p.sv_tmaxcol[i] = p.tmaxcol[r];
        p.sv_tnpoints[i] = p.tnpoints[r];
    }
    if (g.bid == 0 && threadIdx.x == 0)
    {
        p.st->n_ulist_saved = n;
        p.st->sv_n_clusters = p.st->n_clusters;
        p.st->sv_n_cluster_points = p.st->n_cluster_points;
    }
}

__global__ void k_snapshot(CcDevPtrs p, int guard)
{
    CC_PDL_ENTER();
    const CcGrid g = cc_grid();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_snapshot, g.bid);
    if (p.st->halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    d_snapshot(g, p, guard);
}

CC_DEV void d_restore(const CcGrid g, CcDevPtrs p)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_restore, g.bid);
    const int n = p.st->n_ulist_saved;
    for (int i = g.bid * blockDim.x + threadIdx.x; i < n; i += g.nb * blockDim.x)
    {
        const unsigned int r = p.ulist[i];
        p.cparent[r] = p.sv_cparent[i];
        p.tfinish[r] = p.sv_tfinish[i];
        p.tmaxcol[r] = p.sv_tmaxcol[i];
        p.tnpoints[r] = p.sv_tnpoints[i];
        p.tstate[r] = 0u; // list entries are unfinished trees (the fused kernel marks while it still decides)
    }
}
__global__ void k_restore(CcDevPtrs p)
{
    CC_PDL_ENTER();
    d_restore(cc_grid(), p);
}

CC_DEV void d_restore_finish(const CcGrid g, CcDevPtrs p)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_restore_finish, g.bid);
    p.st->n_ulist = p.st->n_ulist_saved;
    p.st->abort = 0;
    p.st->danger_col = CC_COL_INF; // (the host has read it; the next attempt on a sub-range reports its own)
    p.st->n_clusters = p.st->sv_n_clusters;
    p.st->n_cluster_points = p.st->sv_n_cluster_points;
}
__global__ void k_restore_finish(CcDevPtrs p)
{
    CC_PDL_ENTER();
    d_restore_finish(cc_grid(), p);
}

CC_DEV void d_commit_copy(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, const unsigned int* s_parent, int ci0, int ci1, int spec)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_commit_copy, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    if (!cc_head_ok(hd, spec))
        return;
    const int R = cfg.R;
    if (ci1 < 0)
        ci1 = hd.ncols - 1;
    const long long colbase = hd.colbase;
    const int total = (ci1 - ci0 + 1) * R;
    const int lane = threadIdx.x % CC_WARP;
    const unsigned int lt_mask = (1u << lane) - 1u;
    // warps stay converged: the new roots of a warp's cells take their places in the unfinished list with one atomic
    for (int i0 = g.bid * blockDim.x + threadIdx.x - lane; i0 < total; i0 += g.nb * blockDim.x)
    {
        const int i = i0 + lane;
        bool is_root = false;
        unsigned int q = 0u;
        if (i < total)
        {
            const int ci = ci0 + i / R, row = i % R;
            const long long gcol = colbase + ci;
            q = static_cast<unsigned int>(cc_local_col(gcol, cfg.ringcols)) * R + row;
            const unsigned int par = s_parent[static_cast<size_t>(ci) * R + row];
            p.tparent[q] = par;
            p.tfirst[q] = par; // the point whose child list the reference appends this one to (cpp:663)
            if (par == q) // new point tree (cpp:808-826)
            {
                is_root = true;
                p.cparent[q] = q;
                p.tfinish[q] = cc_d2ord(p.cont_az[q] + static_cast<double>(p.mad[q]));
                p.tmaxcol[q] = gcol;
                p.tnpoints[q] = 1;
                p.tstate[q] = 0;
                p.tid[q] = 0;
                p.tslot[q] = -1;
            }
        }
        const unsigned int roots = __ballot_sync(CC_FULL_MASK, is_root);
        if (roots)
        {
            int pos0 = 0;
            if (lane == __ffs(roots) - 1)
                pos0 = atomicAdd(&p.st->n_ulist, __popc(roots));
            pos0 = __shfl_sync(CC_FULL_MASK, pos0, __ffs(roots) - 1);
            if (is_root)
            {
                const int pos = pos0 + __popc(roots & lt_mask);
                if (pos < p.cap_ulist)
                {
                    p.ulist[pos] = q;
                    p.rootslot[q] = static_cast<unsigned int>(pos);
                }
                else
                    p.st->error = CC_DEV_LIST_OVERFLOW;
            }
        }
    }
}
__global__ void k_commit_copy(CcDevCfg cfg, CcDevPtrs p, const unsigned int* s_parent, int ci0, int ci1, int spec)
{
    CC_PDL_ENTER();
    d_commit_copy(cc_grid(), cfg, p, s_parent, ci0, ci1, spec);
}

// 64-bit maximum over the lanes of `grp` (all lanes of the warp call this; lanes outside the group pass anything)
CC_DEV unsigned long long cc_group_max_u64(bool mine, unsigned long long v)
{
    const unsigned int hi = __reduce_max_sync(CC_FULL_MASK, mine ? static_cast<unsigned int>(v >> 32) : 0u);
    const bool top = mine && static_cast<unsigned int>(v >> 32) == hi;
    const unsigned int lo = __reduce_max_sync(CC_FULL_MASK, top ? static_cast<unsigned int>(v) : 0u);
    return (static_cast<unsigned long long>(hi) << 32) | lo;
}

CC_DEV void d_commit_roots(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, int ci0, int ci1, int spec)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_commit_roots, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    if (!cc_head_ok(hd, spec))
        return;
    const int R = cfg.R;
    if (ci1 < 0)
        ci1 = hd.ncols - 1;
    const long long colbase = hd.colbase;
    const int total = (ci1 - ci0 + 1) * R;
    const int lane = threadIdx.x % CC_WARP;
    // warps stay converged: the contributions of a warp's cells to one root (cells of one object in neighbouring rows
    // mostly share it) are combined before they touch the root's words in memory
    for (int i0 = g.bid * blockDim.x + threadIdx.x - lane; i0 < total; i0 += g.nb * blockDim.x)
    {
        const int i = i0 + lane;
        unsigned int r = CC_NONE;
        unsigned long long fin = 0ull;
        int ci = 0;
        if (i < total)
        {
            ci = ci0 + i / R;
            const int row = i % R;
            const unsigned int q = static_cast<unsigned int>(cc_local_col(colbase + ci, cfg.ringcols)) * R + row;
            r = cc_vload(p.tparent + q);
            if (r == q)
                r = CC_NONE; // a root contributes nothing to itself
            if (r != CC_NONE)
            {
                // concurrent pointer jumping: every hop is published in the cell's own entry (always an ancestor, finally
                // the root), so the walks of the other cells of the tree shortcut through it
                unsigned int n;
                while ((n = cc_vload(p.tparent + r)) != r)
                {
                    *reinterpret_cast<volatile unsigned int*>(p.tparent + q) = n;
                    r = n;
                }
                fin = cc_d2ord(p.cont_az[q] + static_cast<double>(p.mad[q]));
            }
        }
        unsigned int remaining = __ballot_sync(CC_FULL_MASK, r != CC_NONE);
        while (remaining)
        {
            const int leader = __ffs(remaining) - 1;
            const unsigned int key = __shfl_sync(CC_FULL_MASK, r, leader);
            const bool mine = r == key;
            const unsigned int grp = __ballot_sync(CC_FULL_MASK, mine);
            const unsigned long long gfin = cc_group_max_u64(mine, fin);
            const int gci = cc_warp_max(mine ? ci : -0x7fffffff - 1);
            if (lane == leader)
            {
                atomicMax(p.tfinish + key, gfin);
                atomicMax(p.tmaxcol + key, colbase + gci);
                atomicAdd(p.tnpoints + key, static_cast<unsigned int>(__popc(grp)));
            }
            remaining &= ~grp;
        }
    }
}
__global__ void k_commit_roots(CcDevCfg cfg, CcDevPtrs p, int ci0, int ci1, int spec)
{
    CC_PDL_ENTER();
    d_commit_roots(cc_grid(), cfg, p, ci0, ci1, spec);
}

CC_DEV unsigned int cc_tree_root(const unsigned int* tparent, unsigned int x)
{
    unsigned int r = cc_vload(tparent + x), n;
    while ((n = cc_vload(tparent + r)) != r)
        r = n;
    return r;
}

// `chase`: the tree roots of both ends are found here (pointer chase to the fixed point) instead of being read from
// tparent: the stage can then share a phase with d_commit_roots (whose pointer jumping only ever stores ancestors).
CC_DEV void d_commit_links(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, const unsigned int* s_parent, const unsigned int* s_links,
                               int ci0, int ci1, int spec, bool chase = false)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_commit_links, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    if (!cc_head_ok(hd, spec))
        return;
    const int R = cfg.R;
    if (ci1 < 0)
        ci1 = hd.ncols - 1;
    {
        const long long colbase = hd.colbase;
        const int total = (ci1 - ci0 + 1) * R * CC_LINK_SLOTS;
        for (int i = g.bid * blockDim.x + threadIdx.x; i < total; i += g.nb * blockDim.x)
        {
            const int cell = i / CC_LINK_SLOTS;
            const size_t idx = static_cast<size_t>(ci0) * R + cell;
            if (s_parent[idx] == CC_NONE)
                continue; // is_ignored: its slots were not written
            const unsigned int o = s_links[idx * CC_LINK_SLOTS + (i % CC_LINK_SLOTS)];
            if (o == CC_NONE)
                continue;
            const int ci = ci0 + cell / R, row = cell % R;
            const unsigned int q = static_cast<unsigned int>(cc_local_col(colbase + ci, cfg.ringcols)) * R + row;
            const unsigned int ra = chase ? cc_tree_root(p.tparent, q) : cc_vload(p.tparent + q);
            const unsigned int rb = chase ? cc_tree_root(p.tparent, o) : cc_vload(p.tparent + o);
            if (ra != rb)
                cc_uf_union(p.cparent, ra, rb);
        }
    }
    int n = p.st->n_edges;
    n = n < p.cap_edges ? n : p.cap_edges;
    const int base_local = cc_local_col(hd.colbase, cfg.ringcols);
    for (int e = g.bid * blockDim.x + threadIdx.x; e < n; e += g.nb * blockDim.x)
    {
        const unsigned int q = p.edge_a[e], o = p.edge_b[e];
        int ci = static_cast<int>(q / R) - base_local;
        if (ci < 0)
            ci += cfg.ringcols;
        if (ci < ci0 || ci > ci1)
            continue;
        const unsigned int ra = chase ? cc_tree_root(p.tparent, q) : cc_vload(p.tparent + q);
        const unsigned int rb = chase ? cc_tree_root(p.tparent, o) : cc_vload(p.tparent + o);
        if (ra != rb)
            cc_uf_union(p.cparent, ra, rb);
    }
}
__global__ void k_commit_links(CcDevCfg cfg, CcDevPtrs p, const unsigned int* s_parent, const unsigned int* s_links,
                               int ci0, int ci1, int spec)
{
    CC_PDL_ENTER();
    d_commit_links(cc_grid(), cfg, p, s_parent, s_links, ci0, ci1, spec);
}

// =====================================================================================================
// K3c  exact column-sequential association of ONE column (cpp:773-835 verbatim semantics, including refused
//      associations and the stop at the first unpublished column). One thread; used only for columns the probe
//      flagged and while a cluster is about to span a full rotation.
// =====================================================================================================
CC_DEV void d_careful(const CcDevCfg& cfg, const CcDevPtrs& p, int ci) // one thread
{
    const int R = cfg.R;
    CcDevState* st = p.st;
    const long long gcol = st->colbase + ci;
    const int local = cc_local_col(gcol, cfg.ringcols);
    const int first_local = cc_local_col(st->first_unpub, cfg.ringcols);
    for (int row = 0; row < R; row++)
    {
        const unsigned int q = static_cast<unsigned int>(local) * R + row;
        p.tparent[q] = CC_NONE;
        p.tfirst[q] = CC_NONE;
        p.visited[q] = 0;
        p.vback[q] = 0; // the walk below honours the stop at the first unpublished column itself
    }
    for (int row = 0; row < R; row++)
    {
        const unsigned int q = static_cast<unsigned int>(local) * R + row;
        const float4 a = p.assoc[q];
        if (cc_isnan(a.x))
            continue;
        const float mad = p.mad[q];
        const double my_finish = p.cont_az[q] + static_cast<double>(mad);
        int steps_back = static_cast<int>(ceilf(ccm::div_rn(mad, cfg.width)));
        steps_back = steps_back < cfg.max_steps_row ? steps_back : cfg.max_steps_row;
        unsigned int root = CC_NONE;
        int visited = 0;
        int other_col = local;
        for (int back = 0; back <= steps_back; back++)
        {
            for (int dir = -1; dir <= 1; dir += 2)
            {
                if (dir == 1 && back == 0)
                    continue;
                int steps_v = (dir == 1 || back == 0) ? 1 : 0;
                int orow = (dir == 1 || back == 0) ? row + dir : row;
                while (orow >= 0 && orow < R && steps_v <= cfg.max_steps_col)
                {
                    const unsigned int o = static_cast<unsigned int>(other_col) * R + orow;
                    const float4 b = p.assoc[o];
                    visited++;
                    if (fabsf(b.w - a.w) > mad)
                        break;
                    if (!cc_isnan(b.x))
                    {
                        const unsigned int ro = p.tparent[o];
                        if (ro != CC_NONE && ro != root)
                        {
                            const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
                            if (dx * dx + dy * dy + dz * dz < cfg.max_distance_sq)
                            {
                                if (root == CC_NONE)
                                {
                                    // associatePointToPointTree cpp:643-673
                                    const long long rootcol = p.slot_gcol[ro / R];
                                    const unsigned int new_width = static_cast<unsigned int>(gcol - rootcol + 1);
                                    if (new_width <= static_cast<unsigned int>(cfg.N) && p.tstate[ro] == 0)
                                    {
                                        root = ro;
                                        p.tparent[q] = ro;
                                        p.tfirst[q] = o; // cpp:663
                                        p.tmaxcol[ro] = gcol;
                                        const unsigned long long f = cc_d2ord(my_finish);
                                        if (f > p.tfinish[ro])
                                            p.tfinish[ro] = f;
                                        p.tnpoints[ro] += 1;
                                    }
                                }
                                else if (p.tstate[root] == 0 && p.tstate[ro] == 0) // cpp:675-696
                                    cc_uf_union(p.cparent, root, ro);
                            }
                        }
                    }
                    if (root != CC_NONE && cfg.stop_enabled && steps_v >= cfg.stop_min_steps)
                        break;
                    orow += dir;
                    steps_v++;
                }
            }
            if (root != CC_NONE && cfg.stop_enabled && back >= cfg.stop_min_steps)
                break;
            if (other_col == first_local)
                break;
            other_col--;
            if (other_col < 0)
                other_col += cfg.ringcols;
        }
        p.visited[q] = static_cast<unsigned short>(visited);
        if (root == CC_NONE)
        {
            p.tparent[q] = q;
            p.tfirst[q] = q;
            p.cparent[q] = q;
            p.tfinish[q] = cc_d2ord(my_finish);
            p.tmaxcol[q] = gcol;
            p.tnpoints[q] = 1;
            p.tstate[q] = 0;
            p.tid[q] = 0;
            p.tslot[q] = -1;
            const int pos = st->n_ulist;
            if (pos < p.cap_ulist)
            {
                p.ulist[pos] = q;
                p.rootslot[q] = static_cast<unsigned int>(pos);
                st->n_ulist = pos + 1;
            }
            else
                st->error = CC_DEV_LIST_OVERFLOW;
        }
    }
}
__global__ void k_careful(CcDevCfg cfg, CcDevPtrs p, int ci)
{
    CC_PDL_ENTER();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_careful, static_cast<int>(blockIdx.x));
    if (blockIdx.x != 0 || threadIdx.x != 0)
        return;
    d_careful(cfg, p, ci);
}

// =====================================================================================================
// K4  finish detection over the unfinished point trees for the passes of columns [c0, c1] (cpp:837-974) and
//     the id / ring bookkeeping of the publish stage (cpp:1035-1083).
//     spec = 1: c0..c1 is a whole range of columns processed at once; a component finishes at the first pass
//               whose column minimum azimuth reaches the component's largest finished_at (binary search on the
//               running maximum); anything that would need the reference's forced finish (cpp:909-919) aborts.
//     spec = 0: c0 == c1, exact single pass including the forced finish.
// =====================================================================================================
CC_DEV void d_fin_init(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, int ci0, int ci1, int spec)
{
    if (!cc_spec_ok(p.st, spec))
        return;
    CcDevState* st = p.st;
    if (ci1 < 0)
        ci1 = st->ncols - 1;
    const int n = st->n_ulist < p.cap_ulist ? st->n_ulist : p.cap_ulist;
    const long long gbase = st->first_unpub;
    const long long c1 = st->colbase + ci1;
    long long glen = c1 - gbase + 1;
    if (glen > p.cap_G)
        glen = p.cap_G;
    const int tid = g.bid * blockDim.x + threadIdx.x, nt = g.nb * blockDim.x;
    for (int i = tid; i < n; i += nt)
    {
        p.u_maxfinish[i] = 0ull;
        p.u_mincol[i] = CC_COL_INF;
        p.u_maxend[i] = -1;
        p.u_np[i] = 0u;
        p.u_finishcol[i] = CC_COL_INF;
        p.u_cluster[i] = -1;
    }
    for (long long j = tid; j < glen; j += nt)
        p.G[j] = -1;
    if (tid == 0)
    {
        *p.n_new_ulist = 0;
        st->seg_c0 = st->colbase + ci0;
        st->seg_c1 = c1;
        st->seg_first_unpub_old = st->first_unpub;
        st->gbase = gbase;
        if (c1 - gbase + 1 > p.cap_G)
            st->error = CC_DEV_LIST_OVERFLOW;
    }
}

CC_DEV void d_fin_agg(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, int spec)
{
    if (!cc_spec_ok(p.st, spec))
        return;
    const int n = p.st->n_ulist < p.cap_ulist ? p.st->n_ulist : p.cap_ulist;
    const int tid = g.bid * blockDim.x + threadIdx.x, nt = g.nb * blockDim.x;
    // four list entries per thread at a time: their loads are issued together (the list is a few thousand entries long
    // and this runs in one CTA, so every dependent round trip counts once per batch instead of once per entry)
    for (int i0 = tid; i0 < n; i0 += 4 * nt)
    {
        unsigned int root[4], cp[4], np[4];
        unsigned long long tf[4];
        long long gc[4], mc[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            root[k] = i0 + k * nt < n ? p.ulist[i0 + k * nt] : CC_NONE;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            cp[k] = np[k] = 0u;
            tf[k] = 0ull;
            gc[k] = mc[k] = 0;
            if (root[k] != CC_NONE)
            {
                // the root's own aggregates do not depend on the find: issued with its first hop
                tf[k] = p.tfinish[root[k]];
                gc[k] = p.slot_gcol[root[k] / cfg.R];
                mc[k] = p.tmaxcol[root[k]];
                np[k] = p.tnpoints[root[k]];
                cp[k] = cc_vload(p.cparent + root[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            if (root[k] == CC_NONE)
                continue;
            const int i = i0 + k * nt;
            const unsigned int rep = cp[k] == root[k] ? root[k] : cc_uf_find(p.cparent, root[k]);
            const int j = rep == root[k] ? i : static_cast<int>(p.rootslot[rep]); // a list root's slot is its list position
            p.u_rep[i] = j;
            atomicMax(p.u_maxfinish + j, tf[k]);
            atomicMin(p.u_mincol + j, gc[k]);
            atomicMax(p.u_maxend + j, mc[k] + 1);
            atomicAdd(p.u_np + j, np[k]);
        }
    }
}

CC_DEV void d_fin_decide(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, int guard, int exact, const double* runmax_s)
{
    if (!cc_spec_ok(p.st, guard))
        return;
    const int spec = !exact;
    // running maximum of the segment's columns: shared-memory copy when the caller staged one (index 0 = column c0)
    const double* rmx = runmax_s ? runmax_s - (p.st->seg_c0 - p.st->colbase) : p.col_runmax;
    CcDevState* st = p.st;
    const int n = st->n_ulist < p.cap_ulist ? st->n_ulist : p.cap_ulist;
    const long long c0 = st->seg_c0, c1 = st->seg_c1, colbase = st->colbase;
    const int tid_ = g.bid * blockDim.x + threadIdx.x, nt_ = g.nb * blockDim.x;
    // four entries per thread at a time, their loads issued together (see d_fin_agg)
    for (int i0 = tid_; i0 < n; i0 += 4 * nt_)
    {
      int rep_[4];
      unsigned long long mf_[4];
      long long me_[4], mn_[4];
      unsigned int np_[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
          const int i = i0 + k * nt_;
          rep_[k] = -1;
          mf_[k] = 0ull;
          me_[k] = mn_[k] = 0;
          np_[k] = 0u;
          if (i < n)
          {
              rep_[k] = p.u_rep[i];
              mf_[k] = p.u_maxfinish[i];
              me_[k] = p.u_maxend[i];
              mn_[k] = p.u_mincol[i];
              np_[k] = p.u_np[i];
          }
      }
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
        const int i = i0 + k * nt_;
        if (rep_[k] != i)
            continue;
        const double F = cc_ord2d(mf_[k]);
        const long long maxend = me_[k], mincol = mn_[k];
        long long finish_col = CC_COL_INF;
        if (spec)
        {
            if (maxend - mincol >= cfg.N)
            {
                // a partial component could be force-finished from column mincol + N - 1 on (cpp:909-919)
                atomicMin(&st->danger_col, mincol + cfg.N - 1);
                st->abort = 1;
                continue;
            }
            const long long lastcol = maxend - 1;
            const long long s = cc_pass_at_or_after(lastcol > c0 ? lastcol : c0, cfg.nth);
            const long long c1p = cc_pass_at_or_before(c1, cfg.nth); // last pass column of the range
            // first pass column c >= s with runmax(c) >= F
            if (s <= c1p && rmx[c1p - colbase] >= F)
            {
                long long lo = s, hi = c1p;
                while (lo < hi)
                {
                    const long long mid = (lo + hi) >> 1;
                    if (rmx[mid - colbase] >= F)
                        hi = mid;
                    else
                        lo = mid + 1;
                }
                lo = cc_pass_at_or_after(lo, cfg.nth); // (<= c1p: the running maximum is monotone)
                if (!(p.col_minaz[lo - colbase] >= F))
                {
                    // the running maximum was reached before the component was complete while this column's own
                    // minimum is still behind it: needs the exact pass-by-pass rule
                    st->abort = 1;
                    atomicMin(&st->danger_col, lo);
                    continue;
                }
                finish_col = lo;
            }
        }
        else
        {
            const double min_az = p.col_minaz[c1 - colbase];
            const bool unfinished = F > min_az;                 // cpp:884-885
            const bool exceeds = (maxend - mincol) >= cfg.N;    // cpp:909-919
            if (!unfinished || exceeds)
                finish_col = c1;
            if (unfinished && exceeds) // a forced finish: associations of the next max_steps_in_row columns may be refused
                atomicMax(&st->forced_col, c1);
        }
        if (finish_col != CC_COL_INF)
        {
            p.u_finishcol[i] = finish_col;
            const unsigned int np = np_[k];
            if (np > 5) // cpp:936-940
            {
                const int slot = atomicAdd(&st->n_clusters, 1);
                const int off = atomicAdd(&st->n_cluster_points, static_cast<int>(np));
                if (slot < p.cap_clusters && off + static_cast<int>(np) <= p.cap_cluster_points)
                {
                    CcCluster c;
                    c.id = st->cluster_counter + static_cast<unsigned long long>(slot);
                    c.min_stamp = ~0ull;
                    c.max_stamp = 0ull;
                    c.finish_col = finish_col;
                    c.min_col = mincol;
                    c.max_col = maxend - 1;
                    c.num_points = np;
                    c.point_offset = static_cast<unsigned int>(off);
                    c.cursor = 0;
                    c.pad_ = 0;
                    p.clusters[slot] = c;
                    p.u_cluster[i] = slot;
                }
                else
                    st->error = CC_DEV_LIST_OVERFLOW;
            }
        }
      }
    }
}

CC_DEV void d_fin_mark(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, unsigned int seq, int spec)
{
    if (!cc_spec_ok(p.st, spec))
        return;
    CcDevState* st = p.st;
    const int n = st->n_ulist < p.cap_ulist ? st->n_ulist : p.cap_ulist;
    const long long gbase = st->gbase;
    const unsigned long long counter = st->cluster_counter;
    const int tid = g.bid * blockDim.x + threadIdx.x, nt = g.nb * blockDim.x;
    const int lane = threadIdx.x % CC_WARP;
    // four entries per thread at a time (loads issued together); the entries that stay unfinished take their places in
    // the compacted list with one atomic per warp. The loop bound is uniform per warp.
    for (int i0 = tid - lane; i0 < n; i0 += 4 * nt)
    {
        int j[4];
        unsigned int root[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const int i = i0 + lane + k * nt;
            j[k] = i < n ? p.u_rep[i] : -1;
            root[k] = i < n ? p.ulist[i] : CC_NONE;
        }
        long long fc[4], gi[4];
        int slot[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            fc[k] = CC_COL_INF;
            gi[k] = -1;
            slot[k] = -1;
            if (j[k] >= 0)
            {
                fc[k] = p.u_finishcol[j[k]];
                slot[k] = p.u_cluster[j[k]];
                gi[k] = p.slot_gcol[root[k] / cfg.R] - gbase;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const bool valid = j[k] >= 0;
            const bool finished = valid && fc[k] != CC_COL_INF;
            if (finished)
            {
                p.tstate[root[k]] = 1u + seq;
                p.tslot[root[k]] = slot[k];
                p.tid[root[k]] = slot[k] >= 0 ? static_cast<unsigned int>(counter + static_cast<unsigned long long>(slot[k])) : 0u;
            }
            if (valid && gi[k] >= 0 && gi[k] < p.cap_G)
                atomicMax(p.G + gi[k], finished ? fc[k] : CC_COL_INF);
            const bool keep = valid && !finished;
            const unsigned int km = __ballot_sync(CC_FULL_MASK, keep);
            if (km)
            {
                int pos0 = 0;
                if (lane == __ffs(km) - 1)
                    pos0 = atomicAdd(p.n_new_ulist, __popc(km));
                pos0 = __shfl_sync(CC_FULL_MASK, pos0, __ffs(km) - 1);
                if (keep)
                    p.ulist_new[pos0 + __popc(km & ((1u << lane) - 1u))] = root[k];
            }
        }
    }
}
// Decision and marking in ONE phase (fused kernel, whole-push speculative commit only): every list entry derives its
// component's decision from the representative's aggregates itself -- the same value for every member -- instead of
// waiting a phase for the representative to publish it; only the representative allocates the cluster record. The roots
// do not learn their cluster slot this way: d_fin_label(via_rep) looks it up through the representative.
CC_DEV void d_fin_decide_mark(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, unsigned int seq)
{
    if (!cc_spec_ok(p.st, 1))
        return;
    CcDevState* st = p.st;
    const int n = st->n_ulist < p.cap_ulist ? st->n_ulist : p.cap_ulist;
    const long long c0 = st->seg_c0, c1 = st->seg_c1, colbase = st->colbase, gbase = st->gbase;
    const double* rmx = p.col_runmax;
    const int tid = g.bid * blockDim.x + threadIdx.x, nt = g.nb * blockDim.x;
    const int lane = threadIdx.x % CC_WARP;
    const double rm_last = rmx[c1 - colbase];
    for (int i0 = tid - lane; i0 < n; i0 += nt) // uniform per warp (compaction below)
    {
        const int i = i0 + lane;
        const bool valid = i < n;
        long long finish_col = CC_COL_INF;
        unsigned int root = CC_NONE;
        long long gi = -1;
        if (valid)
        {
            const int j = p.u_rep[i];
            root = p.ulist[i];
            const unsigned long long mf = p.u_maxfinish[j];
            const long long maxend = p.u_maxend[j], mincol = p.u_mincol[j];
            const unsigned int np = p.u_np[j];
            gi = p.slot_gcol[root / cfg.R] - gbase;
            const double F = cc_ord2d(mf);
            bool bad = false;
            if (maxend - mincol >= cfg.N)
            {
                // a partial component could be force-finished from column mincol + N - 1 on (cpp:909-919)
                if (i == j)
                {
                    atomicMin(&st->danger_col, mincol + cfg.N - 1);
                    st->abort = 1;
                }
                bad = true;
            }
            else if (cc_pass_at_or_after((maxend - 1) > c0 ? (maxend - 1) : c0, cfg.nth) <= cc_pass_at_or_before(c1, cfg.nth) &&
                     rmx[cc_pass_at_or_before(c1, cfg.nth) - colbase] >= F)
            {
                const long long lastcol = maxend - 1;
                long long lo = cc_pass_at_or_after(lastcol > c0 ? lastcol : c0, cfg.nth), hi = cc_pass_at_or_before(c1, cfg.nth);
                while (lo < hi) // first pass column c >= lo with runmax(c) >= F
                {
                    const long long mid = (lo + hi) >> 1;
                    if (rmx[mid - colbase] >= F)
                        hi = mid;
                    else
                        lo = mid + 1;
                }
                lo = cc_pass_at_or_after(lo, cfg.nth); // (still inside the range: the running maximum is monotone)
                if (!(p.col_minaz[lo - colbase] >= F))
                {
                    // the running maximum was reached before the component was complete while this column's own minimum
                    // is still behind it: needs the exact pass-by-pass rule
                    if (i == j)
                    {
                        st->abort = 1;
                        atomicMin(&st->danger_col, lo);
                    }
                    bad = true;
                }
                else
                    finish_col = lo;
            }
            if (!bad && finish_col != CC_COL_INF && i == j)
            {
                p.u_finishcol[i] = finish_col;
                if (np > 5) // cpp:936-940
                {
                    const int slot = atomicAdd(&st->n_clusters, 1);
                    const int off = atomicAdd(&st->n_cluster_points, static_cast<int>(np));
                    if (slot < p.cap_clusters && off + static_cast<int>(np) <= p.cap_cluster_points)
                    {
                        CcCluster c;
                        c.id = st->cluster_counter + static_cast<unsigned long long>(slot);
                        c.min_stamp = ~0ull;
                        c.max_stamp = 0ull;
                        c.finish_col = finish_col;
                        c.min_col = mincol;
                        c.max_col = maxend - 1;
                        c.num_points = np;
                        c.point_offset = static_cast<unsigned int>(off);
                        c.cursor = 0;
                        c.pad_ = 0;
                        p.clusters[slot] = c;
                        p.u_cluster[i] = slot;
                    }
                    else
                        st->error = CC_DEV_LIST_OVERFLOW;
                }
            }
        }
        const bool finished = valid && finish_col != CC_COL_INF;
        if (finished)
            p.tstate[root] = 1u + seq;
        if (valid && gi >= 0 && gi < p.cap_G)
            atomicMax(p.G + gi, finished ? finish_col : CC_COL_INF);
        const bool keep = valid && !finished;
        const unsigned int km = __ballot_sync(CC_FULL_MASK, keep);
        if (km)
        {
            int pos0 = 0;
            if (lane == __ffs(km) - 1)
                pos0 = atomicAdd(p.n_new_ulist, __popc(km));
            pos0 = __shfl_sync(CC_FULL_MASK, pos0, __ffs(km) - 1);
            if (keep)
                p.ulist_new[pos0 + __popc(km & ((1u << lane) - 1u))] = root;
        }
    }
}

CC_DEV void d_fin_copyback(const CcGrid g, const CcDevPtrs& p, int spec)
{
    if (!cc_spec_ok(p.st, spec))
        return;
    const int n = *p.n_new_ulist;
    const int tid = g.bid * blockDim.x + threadIdx.x, nt = g.nb * blockDim.x;
    for (int i0 = tid; i0 < n; i0 += 4 * nt)
    {
        unsigned int root[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            root[k] = i0 + k * nt < n ? p.ulist_new[i0 + k * nt] : CC_NONE;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (root[k] != CC_NONE)
            {
                p.ulist[i0 + k * nt] = root[k];
                p.rootslot[root[k]] = static_cast<unsigned int>(i0 + k * nt);
            }
    }
}

// per column: sc_first_unpublished_global_column_index after its pass = the smallest root column among the trees
// that were still unfinished when the pass started (cpp:943-959), or column + 1. G[r] = last column at which
// some tree rooted in column gbase + r is unfinished; with PG = prefix max of G, the answer for column c is the
// first r with PG[r] >= c. Single block.
// `nslices` > 1 (k_fin_cluster): every CTA of the cluster builds the prefix maxima in its own shared memory and answers
// every nslices-th group of columns; CTA `slice` 0 also does the end-of-pass bookkeeping. Falls back to one CTA when the
// prefix maxima do not fit shared memory (they are then built in place in global memory).
CC_DEV void d_fin_columns(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, int spec, int smem_ints, int slice = 0,
                          int nslices = 1)
{
    if (!cc_spec_ok(p.st, spec))
        return;
    if (nslices <= 1 && g.bid != 0)
        return;
    CC_SMEM(smem);
    long long* part = reinterpret_cast<long long*>(smem);
    CcDevState* st = p.st;
    const long long gbase = st->gbase, c0 = st->seg_c0, c1 = st->seg_c1, colbase = st->colbase;
    long long glen = c1 - gbase + 1;
    if (glen > p.cap_G)
        glen = p.cap_G;
    const int T = blockDim.x, t = threadIdx.x;
    // the prefix maxima live in shared memory when they fit (binary searches below: 11 dependent reads per column),
    // as columns relative to gbase (CC_COL_INF -> INT_MAX)
    int* pg = reinterpret_cast<int*>(part + T);
    const bool in_smem = glen <= smem_ints;
    if (!in_smem && slice != 0)
        return;
    const bool sliced = nslices > 1 && in_smem;
    if (in_smem)
    {
        // G staged once with coalesced loads issued together, as columns relative to gbase (monotone encoding, so the
        // prefix maximum can be taken on the encoded values); contiguous segment per thread + block scan, in place
        const int n = static_cast<int>(glen);
        for (int i0 = 0; i0 < n; i0 += 8 * T)
        {
            long long v[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                v[u] = i0 + u * T + t < n ? p.G[i0 + u * T + t] : -1;
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (i0 + u * T + t < n)
                {
                    const long long rel = v[u] - gbase;
                    pg[i0 + u * T + t] = v[u] < 0 ? -1 : (rel > 0x7ffffffe ? 0x7fffffff : static_cast<int>(rel));
                }
        }
        __syncthreads();
        const int seg = ((n + T - 1) / T) | 1; // odd: the threads' segments start in different banks
        const int lo = t * seg < n ? t * seg : n, hi = (lo + seg < n) ? lo + seg : n;
        int m = -1;
        for (int j = lo; j < hi; j++)
            m = pg[j] > m ? pg[j] : m;
        int pre = cc_block_exclusive_scan(reinterpret_cast<int*>(part), m, -1, CcOpMaxI32());
        for (int j = lo; j < hi; j++)
        {
            pre = pg[j] > pre ? pg[j] : pre;
            pg[j] = pre;
        }
    }
    else
    {
        const long long chunk = (glen + T - 1) / T;
        const long long lo = t * chunk, hi = (lo + chunk < glen) ? lo + chunk : glen;
        long long m = -1;
        for (long long j = lo; j < hi; j++)
            m = p.G[j] > m ? p.G[j] : m;
        long long pre = cc_block_exclusive_scan(part, m, -1LL, CcOpMaxI64());
        for (long long j = lo; j < hi; j++)
        {
            pre = p.G[j] > pre ? p.G[j] : pre;
            p.G[j] = pre;
        }
    }
    __syncthreads();
    // first r in [0, c - gbase] with PG[r] >= c, else c + 1 - gbase
    auto first_unpublished_after = [&](long long c) -> long long
    {
        long long a = 0, b = c - gbase + 1;
        if (b > glen)
            b = glen;
        if (in_smem)
        {
            const int crel = static_cast<int>(c - gbase);
            while (a < b)
            {
                const long long mid = (a + b) >> 1;
                if (pg[mid] >= crel)
                    b = mid;
                else
                    a = mid + 1;
            }
        }
        else
            while (a < b)
            {
                const long long mid = (a + b) >> 1;
                if (p.G[mid] >= c)
                    b = mid;
                else
                    a = mid + 1;
            }
        const long long fu = gbase + a;
        return fu > c + 1 ? c + 1 : fu;
    };
    const long long c_first = c0 + t + (sliced ? static_cast<long long>(slice) * T : 0);
    const long long c_step = static_cast<long long>(sliced ? nslices : 1) * T;
    // with passes only every n-th column (cpp:841) a column between two passes keeps what the pass before it left
    const long long fu_before = st->seg_first_unpub_old; // first unpublished column when this pass started (d_fin_init)
    for (long long c = c_first; c <= c1; c += c_step)
    {
        const long long cp = cc_pass_at_or_before(c, cfg.nth);
        p.col_first_unpub[c - colbase] = cp >= c0 ? first_unpublished_after(cp) : fu_before;
    }
    // the bookkeeping below needs the answer for the last pass column: thread 0 of slice 0 computes it itself (another CTA
    // may own it)
    long long fu_last = 0;
    if (t == 0 && slice == 0)
    {
        const long long cp = cc_pass_at_or_before(c1, cfg.nth);
        fu_last = c1 >= c0 ? (cp >= c0 ? first_unpublished_after(cp) : fu_before) : p.col_first_unpub[c1 - colbase];
    }
    __syncthreads();
    if (t == 0 && slice == 0)
    {
        const long long fu = fu_last;
        if (fu < st->first_unpub)
        {
            st->error = CC_DEV_RING_START_DECREASED; // cpp:1072-1075
            st->err_a = fu;
            st->err_b = st->first_unpub;
        }
        st->first_unpub = fu;
        st->ring_start = fu - cfg.N > 0 ? fu - cfg.N : 0; // cpp:1079
        st->runmax_carry = p.col_runmax[c1 - colbase];
        st->n_ulist = *p.n_new_ulist;
    }
}

CC_DEV void d_push_done(const CcDevPtrs& p, int guard);
CC_DEV void d_state_snapshot(const CcDevPtrs& p, CcDevState* dst)
{
    const int n = static_cast<int>(sizeof(CcDevState) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(p.st);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        d[i] = src[i];
}

// All list-sized phases of a finish pass in ONE CTA (the unfinished-tree list holds 10^2..10^4 entries: a single
// CTA with block-wide barriers between the phases is faster than six dependent launches). The fused kernel runs the
// same phases over all CTAs of its cluster with cluster barriers in between.
__global__ void __launch_bounds__(1024) k_fin_all(CcDevCfg cfg, CcDevPtrs p, int ci0, int ci1, unsigned int seq, int guard,
                                                  int exact, int last, int smem_bytes, CcDevState* snap, int careful_ci)
{
    CC_PDL_ENTER();
    const CcGrid g = cc_grid();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_fin_all, g.bid);
    if (g.bid != 0)
        return;
    if (p.st->halted) // an earlier push in flight could not be committed speculatively (see k_halt)
    {
        if (snap)
            d_state_snapshot(p, snap);
        return;
    }
    if (careful_ci >= 0) // split path: the exact association of the column (K3c) shares the launch with its finish pass
    {
        if (threadIdx.x == 0)
        {
            CcTraceScope cc_tr_c(p.trace, CC_KID_careful, g.bid);
            d_careful(cfg, p, careful_ci);
        }
        __syncthreads();
    }
    CC_SMEM(smem);
    // shared memory: [block-scan scratch: T x 8 B][running maximum of the segment's columns | prefix maxima of G]
    const int T = blockDim.x;
    const int spare = (smem_bytes - T * 8) / 8; // doubles (or pairs of ints) that fit behind the scan scratch
    double* runmax_s = nullptr;
    {
        CcTraceScope cc_tr_ph(p.trace, CC_KID_fin_init, g.bid);
        // the running maxima of the segment's columns are staged beside the initialisation (no barrier in between:
        // the segment bounds are computed here the way d_fin_init stores them)
        if (cc_spec_ok(p.st, guard))
        {
            const int ci1_eff = ci1 < 0 ? p.st->ncols - 1 : ci1;
            const int nseg = ci1_eff - ci0 + 1;
            if (nseg > 0 && nseg <= spare)
            {
                runmax_s = reinterpret_cast<double*>(smem) + T;
                for (int i = threadIdx.x; i < nseg; i += T)
                    runmax_s[i] = p.col_runmax[ci0 + i];
            }
        }
        d_fin_init(g, cfg, p, ci0, ci1, guard);
    }
    __syncthreads();
    {
        CcTraceScope cc_tr_ph(p.trace, CC_KID_fin_agg, g.bid);
        d_fin_agg(g, cfg, p, guard);
    }
    __syncthreads();
    {
        CcTraceScope cc_tr_ph(p.trace, CC_KID_fin_decide, g.bid);
        d_fin_decide(g, cfg, p, guard, exact, runmax_s);
    }
    __syncthreads();
    {
        CcTraceScope cc_tr_ph(p.trace, CC_KID_fin_mark, g.bid);
        d_fin_mark(g, cfg, p, seq, guard);
    }
    __syncthreads();
    {
        CcTraceScope cc_tr_ph(p.trace, CC_KID_fin_copyback, g.bid);
        d_fin_copyback(g, p, guard);
    }
    __syncthreads();
    {
        CcTraceScope cc_tr_ph(p.trace, CC_KID_fin_columns, g.bid);
        d_fin_columns(g, cfg, p, guard, spare * 2);
    }
    __syncthreads();
    if (last && threadIdx.x == 0)
        d_push_done(p, guard);
    if (snap) // copy of the stream state at the end of the push (what the host reads while the next push already runs)
    {
        __syncthreads();
        d_state_snapshot(p, snap);
    }
}

// Point::id of every member of a cluster finished in this commit (cpp:1005) + the member list and stamp range the
// host needs for the finished-cluster callback (cpp:1007-1028).
// `spans` (shared memory, 2 * span_cap ints, or null): SPAN MODE for short pushes -- instead of testing every unpublished
// cell (up to a rotation of columns, whatever the size of the push) only the column spans [min_col, max_col] of the
// clusters this push finished are visited; a cell inside the spans of several clusters is taken by the one it belongs
// to. Falls back to the scan when the spans together are no smaller than the unpublished range.
// `via_rep` (fused kernel: decision and marking share one phase, so the roots do not carry their cluster slot): the slot
// is read from the component representative's list entry, the id from the cluster record. `h_points` (or null): member
// list entries are also written straight to page-locked host memory (first `h_cap` entries).
// Both are compiled in only with FUSED. The other instantiation (the stage as its own kernel, which visits every cell of
// the commit's columns anyway) also lists the points whose visit count has to be redone (d_visited_fix).
CC_DEV int cc_visited_max_back(const CcDevCfg& cfg, const CcDevPtrs& p, long long gcol, size_t cell, long long c0, long long fu0,
                               long long colbase)
{
    // >= 0: the walk of this point went beyond the first unpublished column (cpp:762-763) and has to be cut there
    const int reached = p.vback[cell];
    if (reached <= 0)
        return -1;
    const long long fu = gcol == c0 ? fu0 : p.col_first_unpub[gcol - 1 - colbase];
    if (fu >= 0 && gcol - reached < fu)
        return gcol > fu ? static_cast<int>(gcol - fu) : 0;
    return -1;
}
template<bool FUSED>
CC_DEV void d_fin_label(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, unsigned int seq, int spec, int* spans_ = nullptr,
                        int span_cap = 0, bool via_rep_ = false, CcClusterPoint* h_points_ = nullptr, int h_cap = 0)
{
    int* spans = FUSED ? spans_ : nullptr;
    const bool via_rep = via_rep_;
    CcClusterPoint* h_points = FUSED ? h_points_ : nullptr;
    CcTraceScope cc_trace_scope(p.trace, CC_KID_fin_label, g.bid);
    const CcHead hd = cc_head(p.st);
    if (hd.halted)
        return; // an earlier push in flight could not be committed speculatively (see k_halt)
    if (!cc_head_ok(hd, spec))
        return;
    const CcDevState* st = p.st;
    const int R = cfg.R;
    const long long gbase = st->gbase, c1 = st->seg_c1;
    long long total = (c1 - gbase + 1) * R;
    const int lane = threadIdx.x % CC_WARP;
    const unsigned int lt_mask = (1u << lane) - 1u;
    int nspan = 0; // > 0: span mode
    if (spans)
    {
        // every CTA builds the same table: exclusive prefix sum of the clusters' span sizes in cells + their first columns
        const int T = blockDim.x, t = threadIdx.x;
        const int ncl = st->n_clusters < p.cap_clusters ? st->n_clusters : p.cap_clusters;
        __shared__ int sh_span_total;
        __shared__ long long sh_scan[32];
        if (ncl == 0)
            return;
        if (ncl <= span_cap && ncl <= T)
        {
            int cells = 0, mincol_rel = 0;
            if (t < ncl)
            {
                const long long mn = p.clusters[t].min_col, mx = p.clusters[t].max_col;
                cells = static_cast<int>((mx - mn + 1) * R);
                mincol_rel = static_cast<int>(mn - gbase);
            }
            const long long off = cc_block_exclusive_scan(sh_scan, static_cast<long long>(cells), 0LL, CcOpAddI64());
            if (t < ncl)
            {
                spans[t] = static_cast<int>(off < 0x7fffffff ? off : 0x7fffffff);
                spans[span_cap + t] = mincol_rel;
            }
            if (t == ncl - 1)
                sh_span_total = off + cells < 0x7fffffff ? static_cast<int>(off + cells) : 0x7fffffff;
            __syncthreads();
            if (static_cast<long long>(sh_span_total) < total)
            {
                nspan = ncl;
                total = sh_span_total;
            }
        }
    }
    // warps stay converged: the members of one cluster among a warp's cells reserve their slots in the cluster's point
    // list and update its stamp range with one atomic each
    for (long long i0 = g.bid * blockDim.x + threadIdx.x - lane; i0 < total; i0 += static_cast<long long>(g.nb) * blockDim.x)
    {
        const long long i = i0 + lane;
        int slot = -1;
        long long gcol = 0;
        int row = 0;
        unsigned long long stamp = 0ull;
        if (i < total)
        {
            int want = -1; // span mode: the cluster whose span this work item belongs to
            if (nspan)
            {
                int a = 0, b = nspan - 1; // last k with spans[k] <= i
                while (a < b)
                {
                    const int mid = (a + b + 1) >> 1;
                    if (spans[mid] <= static_cast<int>(i))
                        a = mid;
                    else
                        b = mid - 1;
                }
                want = a;
                const int local_i = static_cast<int>(i) - spans[a];
                gcol = gbase + spans[span_cap + a] + local_i / R;
                row = local_i % R;
            }
            else
            {
                gcol = gbase + i / R;
                row = static_cast<int>(i % R);
            }
            const size_t cell = static_cast<size_t>(cc_local_col(gcol, cfg.ringcols)) * R + row;
            if (!FUSED && gcol >= st->seg_c0)
            {
                const int mb = cc_visited_max_back(cfg, p, gcol, cell, st->seg_c0, st->seg_first_unpub_old, hd.colbase);
                if (mb >= 0)
                {
                    // (the probe's work lists are free again: the probes of this push are done)
                    const int pos = atomicAdd(&p.st->n_vfix, 1);
                    p.heavy_list[pos] = static_cast<int>(gcol - hd.colbase) * R + row;
                    p.probe_list[pos] = mb;
                }
            }
            if (p.slot_gcol[cell / R] == gcol)
            {
                const unsigned int root = p.tparent[cell];
                if (root != CC_NONE && p.tstate[root] == 1u + seq)
                {
                    slot = via_rep ? p.u_cluster[p.u_rep[p.rootslot[root]]] : p.tslot[root];
                    if (want >= 0 && slot != want)
                        slot = -1; // labelled by the work item of its own cluster's span
                    if (slot >= 0)
                    {
                        p.cid[cell] = via_rep ? static_cast<unsigned int>(p.clusters[slot].id) : p.tid[root];
                        stamp = p.stamp[cell];
                    }
                }
            }
        }
        unsigned int remaining = __ballot_sync(CC_FULL_MASK, slot >= 0);
        while (remaining)
        {
            const int leader = __ffs(remaining) - 1;
            const int key = __shfl_sync(CC_FULL_MASK, slot, leader);
            const bool mine = slot == key;
            const unsigned int grp = __ballot_sync(CC_FULL_MASK, mine);
            const unsigned long long smax = cc_group_max_u64(mine, stamp);
            const unsigned long long smin = ~cc_group_max_u64(mine, ~stamp);
            CcCluster* c = p.clusters + key;
            unsigned int pos0 = 0u;
            if (lane == leader)
            {
                pos0 = atomicAdd(&c->cursor, static_cast<unsigned int>(__popc(grp)));
                atomicMin(&c->min_stamp, smin);
                atomicMax(&c->max_stamp, smax);
            }
            pos0 = __shfl_sync(CC_FULL_MASK, pos0, leader);
            if (mine)
            {
                const unsigned int pos = pos0 + static_cast<unsigned int>(__popc(grp & lt_mask));
                if (pos < c->num_points)
                {
                    CcClusterPoint cp;
                    cp.gcol = gcol;
                    cp.row = row;
                    cp.pad_ = 0;
                    p.cluster_points[c->point_offset + pos] = cp;
                    if (h_points && c->point_offset + pos < static_cast<unsigned int>(h_cap))
                        h_points[c->point_offset + pos] = cp;
                }
            }
            remaining &= ~grp;
        }
    }
}
// number_of_visited_neighbors, exactly: the reference's walk stops at the first unpublished column (cpp:762-763), which
// for column c is where the finish pass of column c - 1 left it -- known only now, after the finish passes of the commit.
// The (geometric) probes walked without that stop and left how many columns back they got; the few points that went
// beyond the stop (mostly the first points of a new object while nothing else is unfinished) are counted again, a warp
// per point, with the walk cut there. Such points come in runs of neighbouring cells: the cells a warp inspects are
// spread over the range (stride = number of warps) so that a run is shared by many warps.
// `ring`: CC_PROBE_PIPE * CC_WARP float4 of shared memory per warp.
// from_list: the points were listed by d_fin_label (heavy_list / probe_list, n_vfix entries); else the cells of the
// commit's columns are scanned here (fused kernel: short pushes).
CC_DEV void d_visited_fix(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, int spec, float4* ring_all, bool from_list)
{
    const CcHead hd = cc_head(p.st);
    if (hd.halted || !cc_head_ok(hd, spec))
        return;
    const CcDevState* st = p.st;
    const int R = cfg.R;
    if (from_list)
    {
        const int n = st->n_vfix;
        const int bl = cc_local_col(hd.colbase, cfg.ringcols);
        const int ln = threadIdx.x % CC_WARP, wp = threadIdx.x / CC_WARP;
        const int nw = (blockDim.x + CC_WARP - 1) / CC_WARP;
        float4* rg = ring_all + static_cast<size_t>(wp) * CC_PROBE_PIPE * CC_WARP;
        for (int e = g.bid * nw + wp; e < n; e += g.nb * nw)
            d_probe_coop(cfg, p, nullptr, nullptr, rg, bl, hd.colbase, ln, p.heavy_list[e], p.probe_list[e]);
        return;
    }
    const long long c0 = st->seg_c0, c1 = st->seg_c1, colbase = hd.colbase;
    const long long fu0 = st->seg_first_unpub_old;
    const int base_local = cc_local_col(colbase, cfg.ringcols);
    const int lane = threadIdx.x % CC_WARP, warp = threadIdx.x / CC_WARP;
    const int nwarps = (blockDim.x + CC_WARP - 1) / CC_WARP;
    float4* ring = ring_all + static_cast<size_t>(warp) * CC_PROBE_PIPE * CC_WARP;
    const long long total = (c1 - c0 + 1) * R;
    const long long W = static_cast<long long>(g.nb) * nwarps, w = static_cast<long long>(g.bid) * nwarps + warp;
    for (long long j0 = 0; j0 * W < total; j0 += CC_WARP) // cell (j0 + lane) * W + w
    {
        const long long i = (j0 + lane) * W + w;
        int max_back = -1, pidx = 0;
        if (i < total)
        {
            const long long gcol = c0 + i / R;
            const int row = static_cast<int>(i % R);
            const size_t cell = static_cast<size_t>(cc_local_col(gcol, cfg.ringcols)) * R + row;
            max_back = cc_visited_max_back(cfg, p, gcol, cell, c0, fu0, colbase);
            pidx = static_cast<int>(gcol - colbase) * R + row;
        }
        unsigned int todo = __ballot_sync(CC_FULL_MASK, max_back >= 0);
        if (todo && lane == 0)
            atomicAdd(&p.st->n_vfix, __popc(todo));
        while (todo)
        {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int mb = __shfl_sync(CC_FULL_MASK, max_back, src), px = __shfl_sync(CC_FULL_MASK, pidx, src);
            d_probe_coop(cfg, p, nullptr, nullptr, ring, base_local, colbase, lane, px, mb);
        }
    }
}

// via_rep: the finish pass before it was k_fin_cluster (the roots do not carry their cluster slot, see d_fin_decide_mark)
__global__ void k_fin_label(CcDevCfg cfg, CcDevPtrs p, unsigned int seq, int spec, int via_rep)
{
    CC_PDL_ENTER();
    d_fin_label<false>(cc_grid(), cfg, p, seq, spec, nullptr, 0, via_rep != 0);
}

// The listed points are recounted by the mask algorithm of k_probe_heavy (a two-warp CTA per point: every vertical run of the cut
// window evaluated at once, the sequential rules applied to the ballot masks) -- a point that walked its whole window costs one
// memory round trip instead of one per run.
__global__ void __launch_bounds__(64) k_visited_fix(CcDevCfg cfg, CcDevPtrs p, int spec, CcDevState* snap, int tune)
{
    CC_PDL_ENTER();
    const CcGrid g = cc_grid();
    if (snap && g.bid == 0 && threadIdx.x == 0)
        snap->n_vfix = p.st->n_vfix; // the list was filled after the finish pass copied the state
    const CcHead hd = cc_head(p.st);
    if (hd.halted || !cc_head_ok(hd, spec))
        return;
    d_probe_heavy<false, true>(g, cfg, p, nullptr, nullptr, tune, 2);
}

// =====================================================================================================
// K5  clearColumns (cpp:1094-1145) for the columns that left the ring in this push.
// =====================================================================================================
// mode 0: explicit range [from, to) (reset); mode 1: the range retired TWO pushes ago. Recycling is deferred so that
// every column a push reports through a finished-column event can still be read by the caller after the push has
// been waited for, even while the next push is already in flight (the reference's callbacks read range_image_
// before clearColumns runs, cpp:1087-1091).
CC_DEV void d_clear(const CcGrid g, const CcDevCfg& cfg, const CcDevPtrs& p, long long from, long long to)
{
    if (from < 0)
        from = 0;
    if (to <= from)
        return;
    const int R = cfg.R;
    const long long total = (to - from) * R;
    const float nanv = cc_nanf();
    for (long long i = g.bid * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(g.nb) * blockDim.x)
    {
        const long long gcol = from + i / R;
        const int row = static_cast<int>(i % R);
        const int local = cc_local_col(gcol, cfg.ringcols);
        const size_t cell = static_cast<size_t>(local) * R + row;
        p.pos[cell] = make_float4(nanv, nanv, nanv, nanv);
        p.azimuth[cell] = nanv;
        p.incl[cell] = nanv;
        p.cont_az[cell] = __longlong_as_double(0x7ff8000000000000LL);
        p.lab[cell] = make_uchar4(CC_GP_UNKNOWN, CC_WHITE, 0, 0);
        p.stamp[cell] = 0ull;
        p.guid[cell] = ~0ull;
        p.assoc[cell] = make_float4(nanv, nanv, nanv, nanv);
        p.mad[cell] = 0.f;
        p.tparent[cell] = CC_NONE;
        p.tfirst[cell] = CC_NONE;
        p.cid[cell] = 0u;
        p.visited[cell] = 0;
        p.vback[cell] = 0;
        p.tstate[cell] = 0u;
        if (row == 0)
            p.slot_gcol[local] = -1;
    }
}

// explicit range [from, to) (reset). The columns retired during normal operation are recycled by k_prep.
CC_DEV void d_clear_range(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, long long from, long long to, int mode)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_clear, g.bid);
    CcDevState* st = p.st;
    if (mode)
    {
        if (st->halted)
            return;
        from = st->clear2_from;
        to = st->clear2_to;
    }
    d_clear(g, cfg, p, from, to);
}
__global__ void k_clear(CcDevCfg cfg, CcDevPtrs p, long long from, long long to, int mode)
{
    CC_PDL_ENTER();
    d_clear_range(cc_grid(), cfg, p, from, to, mode);
}

// end of a push: remember the range of columns that left the ring (recycled at the start of the next push) and
// advance sc_cluster_counter_ (cpp:939) by the ids handed out. guard as in cc_spec_ok.
CC_DEV void d_push_done(const CcDevPtrs& p, int guard);
__global__ void k_push_done(CcDevPtrs p, int guard)
{
    CC_PDL_ENTER();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_push_done, static_cast<int>(blockIdx.x));
    d_push_done(p, guard);
}
CC_DEV void d_push_done(const CcDevPtrs& p, int guard)
{
    CcDevState* st = p.st;
    if (guard == 1 && (st->error != 0 || st->n_flagged != 0 || st->abort != 0))
    {
        if (st->ncols > 0 || st->error != 0)
            st->halted = 1; // pushes already queued behind this one must not run before the host has finished it
        return;
    }
    if (st->ring_start > st->clear_from)
        st->clear_to = st->ring_start;
    st->cluster_counter += static_cast<unsigned long long>(st->n_clusters);
}

// mode 1: halt if this push completed columns (no robot transform: the reference throws); 0: halt if it has new
// columns (finish passes only every n-th column: always column-sequential); -1: clear the flag
CC_DEV void d_halt(const CcGrid g, CcDevPtrs p, int mode)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_halt, g.bid);
    CcDevState* st = p.st;
    if (mode < 0)
        st->halted = 0;
    else if (!st->halted && (st->ncols > 0 || st->error != 0))
        st->halted = 1;
}
__global__ void k_halt(CcDevPtrs p, int mode)
{
    CC_PDL_ENTER();
    d_halt(cc_grid(), p, mode);
}

// labels (ground label, debug label, is_ignored, intensity) of the push's new columns, packed contiguously for one
// device->host copy next to the other results of the push (cc_set_label_prefetch)
CC_DEV void d_pack_labels(const CcGrid g, CcDevCfg cfg, CcDevPtrs p, uchar4* out, int cap_cols)
{
    CcTraceScope cc_trace_scope(p.trace, CC_KID_pack_labels, g.bid);
    const CcDevState* st = p.st;
    if (st->halted)
        return;
    int ncols = st->ncols;
    ncols = ncols < cap_cols ? ncols : cap_cols;
    const long long colbase = st->colbase;
    const int R = cfg.R;
    const long long total = static_cast<long long>(ncols) * R;
    for (long long i = g.bid * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(g.nb) * blockDim.x)
    {
        const long long ci = i / R;
        const int row = static_cast<int>(i - ci * R);
        out[i] = p.lab[static_cast<size_t>(cc_local_col(colbase + ci, cfg.ringcols)) * R + row];
    }
}
__global__ void k_pack_labels(CcDevCfg cfg, CcDevPtrs p, uchar4* out, int cap_cols)
{
    CC_PDL_ENTER();
    d_pack_labels(cc_grid(), cfg, p, out, cap_cols);
}

// The results of a push -- stream state, first unpublished column per column, cluster records, member lists, packed
// labels -- written straight into the push's page-locked HOST buffers by one kernel on the copy stream: exactly as many
// bytes as the push produced (the copy engine moved capacity-sized prefixes in five transfers, ~150 us behind the last
// kernel; this is one launch and ~1.4 MB of stores over the link at 4096 firings per push).
struct CcResultDst
{
    CcDevState* h_state;
    long long* h_first_unpub;
    CcCluster* h_clusters;
    CcClusterPoint* h_points;
    uchar4* h_labels; // or null
    int cap_cols, cap_clusters, cap_points, rows;
};
CC_DEV void cc_copy_out(void* dst, const void* src, size_t bytes, size_t tid, size_t nt)
{
    // 16-byte pieces (every source array is 16-byte aligned and the records are 8 or 16 bytes), then the 4-byte tail
    const size_t n16 = bytes / 16;
    const float4* s16 = reinterpret_cast<const float4*>(src);
    float4* d16 = reinterpret_cast<float4*>(dst);
    for (size_t i = tid; i < n16; i += nt)
        d16[i] = s16[i];
    const unsigned int* s4 = reinterpret_cast<const unsigned int*>(src);
    unsigned int* d4 = reinterpret_cast<unsigned int*>(dst);
    for (size_t i = n16 * 4 + tid; i < bytes / 4; i += nt)
        d4[i] = s4[i];
}
__global__ void k_results_to_host(const CcDevState* snap, const long long* d_first_unpub, const CcCluster* d_clusters,
                                  const CcClusterPoint* d_points, const uchar4* d_labels, CcResultDst dst)
{
    CC_PDL_ENTER();
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, nt = static_cast<size_t>(gridDim.x) * blockDim.x;
    const int ncols = snap->ncols < dst.cap_cols ? snap->ncols : dst.cap_cols;
    const int ncl = snap->n_clusters < dst.cap_clusters ? snap->n_clusters : dst.cap_clusters;
    const int ncp = snap->n_cluster_points < dst.cap_points ? snap->n_cluster_points : dst.cap_points;
    cc_copy_out(dst.h_state, snap, sizeof(CcDevState), tid, nt);
    if (ncols > 0)
        cc_copy_out(dst.h_first_unpub, d_first_unpub, static_cast<size_t>(ncols) * sizeof(long long), tid, nt);
    if (ncl > 0)
        cc_copy_out(dst.h_clusters, d_clusters, static_cast<size_t>(ncl) * sizeof(CcCluster), tid, nt);
    if (ncp > 0)
        cc_copy_out(dst.h_points, d_points, static_cast<size_t>(ncp) * sizeof(CcClusterPoint), tid, nt);
    if (dst.h_labels && ncols > 0)
        cc_copy_out(dst.h_labels, d_labels, static_cast<size_t>(ncols) * dst.rows * sizeof(uchar4), tid, nt);
}

// copy of the stream state at the end of a push (what the host reads while the next push already runs)
__global__ void k_state_snapshot(CcDevPtrs p, CcDevState* dst)
{
    CC_PDL_ENTER();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_state_snapshot, static_cast<int>(blockIdx.x));
    d_state_snapshot(p, dst);
}

// =====================================================================================================
// K*  the whole push in ONE launch (short pushes: latency mode). One thread-block cluster; the stages above run one
//     after the other over the CTAs of the cluster with a hardware cluster barrier (release / acquire, ~0.2 us) where the
//     multi-kernel path has a kernel boundary (~1 us + ramp-up each, and ~4 us of host launch time each). The firings are
//     read straight from page-locked HOST memory (no copy-engine hop) and the results of the push -- state, per-column
//     first-unpublished, cluster records, member lists, packed labels -- are written straight back to page-locked host
//     memory, followed by a flag the host polls. Same device functions, same results as the kernel chain.
// =====================================================================================================
struct CcHostHeader // page-locked, written by the device at the very end of a fused push
{
    unsigned int flag; // ticket of the push once everything else is visible
    unsigned int pad_;
    unsigned long long t_start_ns, t_end_ns; // %globaltimer at kernel entry (after the grid dependency) / exit
};

struct CcFusedArgs
{
    int n, has_tf, spec, scan_chunk, tune, team_warps, pack_labels;
    unsigned int seq, ticket;
    const void* src_raw;     // cc_raw_point_t[n * R]: page-locked host memory (or device memory)
    const double* src_poses; // [n][12]
    void* dst_raw;           // device staging the stages read (null: read src directly, device-resident inputs)
    double* dst_poses;
    unsigned int* s_parent;
    unsigned int* s_links;
    CcDevState* snap; // device copy of the state at the end of the push
    uchar4* d_labels;
    // results, page-locked host memory
    CcHostHeader* h_hdr;
    CcDevState* h_state;
    long long* h_first_unpub;
    CcCluster* h_clusters;
    CcClusterPoint* h_points;
    uchar4* h_labels;
    int cap_cols, cap_clusters, cap_points, smem_bytes;
};

CC_DEV CcGrid cc_grid_cluster()
{
    CcGrid g;
#ifdef CC_EMU
    g.bid = 0;
    g.nb = 1;
#else
    unsigned int r, n;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
    g.bid = static_cast<int>(r);
    g.nb = static_cast<int>(n);
#endif
    return g;
}
// barrier over all threads of the cluster; global-memory writes before it are visible to every CTA after it
CC_DEV void cc_cluster_sync()
{
#ifndef CC_EMU
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
    asm volatile("" ::: "memory"); // the stages stay in program order for the host compiler as well
#endif
}
CC_DEV void cc_fence_system()
{
#ifndef CC_EMU
    __threadfence_system();
#endif
}
CC_DEV unsigned long long cc_globaltimer()
{
#ifdef CC_EMU
    return 0ull;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}

#ifdef CC_EMU
struct uint4
{
    unsigned int x, y, z, w;
};
#endif

// 16-byte copy over the virtual grid (bytes is a multiple of 16, both pointers 16-byte aligned)
CC_DEV void d_copy16(const CcGrid g, void* dst, const void* src, size_t bytes)
{
    const size_t n = bytes / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const size_t stride = static_cast<size_t>(g.nb) * blockDim.x;
    for (size_t i0 = static_cast<size_t>(g.bid) * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride)
    {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * stride < n)
                v[u] = s4[i0 + u * stride];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * stride < n)
                d4[i0 + u * stride] = v[u];
    }
}

__global__ void __launch_bounds__(512, 1) k_push_fused(CcDevCfg cfg, CcDevPtrs p, CcFusedArgs a)
{
    CC_PDL_ENTER();
    const CcGrid g = cc_grid_cluster();
    const unsigned long long t_start = cc_globaltimer();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_push_fused, g.bid);
    const int n = a.n;
    // ---- inputs: page-locked host memory -> device staging ----
    if (a.dst_raw)
    {
        CcTraceScope tr(p.trace, CC_KID_fetch, g.bid);
        d_copy16(g, a.dst_raw, a.src_raw, static_cast<size_t>(n) * cfg.R * sizeof(CcRawPoint));
        d_copy16(g, a.dst_poses, a.src_poses, static_cast<size_t>(n) * 12 * sizeof(double));
        tr.stop();
        cc_cluster_sync();
    }
    // ---- insertion (cpp:105-292) ----
    d_prep(g, cfg, p, n);
    cc_cluster_sync();
    d_lite_check_small(g, cfg, p, n);
    cc_cluster_sync();
    if (p.st->scan_kbad >= n && g.nb > 1)
    {
        // the push is regular as a whole: the insertion scan only commits it (one CTA) and the scatter needs nothing
        // from that commit -- side by side
        if (g.bid == 0)
            d_insert_scan(g, cfg, p, n, a.scan_chunk, 1);
        else
        {
            CcGrid gs;
            gs.bid = g.bid - 1;
            gs.nb = g.nb - 1;
            d_scatter(gs, cfg, p, n);
        }
    }
    else
    {
        d_insert_scan(g, cfg, p, n, a.scan_chunk, 1);
        cc_cluster_sync();
        d_scatter(g, cfg, p, n);
    }
    cc_cluster_sync();
    if (!a.has_tf)
    {
        // the reference throws from the segmentation stage of the first completed column (cpp:298-299)
        if (g.bid == 0 && threadIdx.x == 0)
            d_halt(g, p, 1);
    }
    else
    {
        // ---- ground segmentation (cpp:294-624) ----
        {
            const CcHead hd = cc_head(p.st);
            if (!hd.halted)
                d_gap_rows(g, cfg, p, hd);
        }
        cc_cluster_sync();
        d_ground<true>(g, cfg, p, a.s_parent, a.pack_labels ? a.d_labels : nullptr, a.h_labels, a.cap_cols);
        cc_cluster_sync();
        // ---- association (cpp:638-835) ----
        d_probe(g, cfg, p, a.s_parent, a.s_links, a.spec);
        cc_cluster_sync();
        if (p.st->n_heavy > 0) // (the same value in every CTA: read after the barrier)
        {
            d_probe_heavy<true>(g, cfg, p, a.s_parent, a.s_links, a.tune, a.team_warps);
            cc_cluster_sync();
        }
        if (a.spec)
        {
            d_commit_copy(g, cfg, p, a.s_parent, 0, -1, 1);
            cc_cluster_sync();
            // ---- tree roots + tree<->tree links in one phase (the links chase the roots themselves), together with the
            //      initialisation of the finish pass ----
            const bool fin = !p.st->halted;
            if (fin)
            {
                CcTraceScope tr(p.trace, CC_KID_fin_init, g.bid);
                d_fin_init(g, cfg, p, 0, -1, 1);
            }
            d_commit_roots(g, cfg, p, 0, -1, 1);
            d_commit_links(g, cfg, p, a.s_parent, a.s_links, 0, -1, 1, true);
            cc_cluster_sync();
            // ---- finish detection (cpp:837-974), list phases over all CTAs ----
            if (fin)
            {
                {
                    CcTraceScope tr(p.trace, CC_KID_fin_agg, g.bid);
                    d_fin_agg(g, cfg, p, 1);
                }
                cc_cluster_sync();
                {
                    CcTraceScope tr(p.trace, CC_KID_fin_decide, g.bid);
                    d_fin_decide_mark(g, cfg, p, a.seq);
                }
                cc_cluster_sync();
                // CTA 0: per-column first-unpublished + end-of-push bookkeeping; CTA 1: the compacted list goes back;
                // the others label the members of the finished clusters (the labelling reads gbase / seg_c1 / n_clusters
                // and the cluster records, which neither of the two modifies)
                const int nlab = g.nb >= 4 ? g.nb - 2 : (g.nb >= 2 ? g.nb - 1 : 1);
                const int first_lab = g.nb - nlab;
                if (g.bid == 0)
                {
                    CcTraceScope tr(p.trace, CC_KID_fin_columns, g.bid);
                    d_fin_columns(g, cfg, p, 1, (a.smem_bytes - static_cast<int>(blockDim.x) * 8) / 4);
                    __syncthreads();
                    if (threadIdx.x == 0)
                        d_push_done(p, 1);
                    __syncthreads();
                }
                if (g.bid == (g.nb >= 4 ? 1 : 0))
                {
                    CcTraceScope tr(p.trace, CC_KID_fin_copyback, g.bid);
                    CcGrid g1;
                    g1.bid = 0;
                    g1.nb = 1;
                    d_fin_copyback(g1, p, 1);
                }
                if (g.bid >= first_lab)
                {
                    CcGrid gl;
                    gl.bid = g.bid - first_lab;
                    gl.nb = nlab;
                    CC_SMEM(smem_l);
                    __syncthreads(); // (a CTA that also ran the tail: its shared memory is free again)
                    d_fin_label<true>(gl, cfg, p, a.seq, 1, reinterpret_cast<int*>(smem_l), 512, true, a.h_points, a.cap_points);
                }
            }
            else if (g.bid == 0 && threadIdx.x == 0)
                d_push_done(p, 1);
        }
        else if (g.bid == 0 && threadIdx.x == 0)
            d_halt(g, p, 0); // finish passes every n-th column: column-sequential path, on the host's cue
    }
    // ---- results straight to the host: what is left after the stages wrote theirs (labels: d_ground, member lists:
    //      d_fin_label) -- the per-column first-unpublished columns, and behind one more barrier (stamp ranges are final
    //      once every labelling CTA is done) the cluster records and the state ----
    // One system-scope fence for the whole push, by the thread that raises the flag: what the other CTAs wrote to host
    // memory is ordered before it by the cluster barrier (release / acquire) and the fence's cumulativity.
    cc_cluster_sync();
    if (a.has_tf && a.spec && (g.nb == 1 || g.bid != 0))
    {
        CcGrid gv = g;
        if (g.nb > 1)
        {
            gv.bid = g.bid - 1;
            gv.nb = g.nb - 1;
        }
        CC_SMEM(smem_v);
        d_visited_fix(gv, cfg, p, 1, reinterpret_cast<float4*>(smem_v), false);
        __syncthreads();
    }
    if (g.bid == 0)
    {
        CcTraceScope tr(p.trace, CC_KID_export, g.bid);
        const CcDevState* st = p.st;
        const bool ok = !st->halted && a.has_tf;
        if (ok)
        {
            int ncols = st->ncols < a.cap_cols ? st->ncols : a.cap_cols;
            ncols = ncols > 0 ? ncols : 0;
            const int ncl = st->n_clusters < a.cap_clusters ? st->n_clusters : a.cap_clusters;
            const int T = blockDim.x, t = threadIdx.x;
            for (int i = t; i < ncols; i += T)
                a.h_first_unpub[i] = p.col_first_unpub[i];
            const int words = ncl * static_cast<int>(sizeof(CcCluster) / 8);
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.clusters);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.h_clusters);
            for (int i = t; i < words; i += T)
                dst[i] = src[i];
        }
        d_state_snapshot(p, a.snap);
        d_state_snapshot(p, a.h_state);
        __syncthreads();
    }
    if (g.bid == 0 && threadIdx.x == 0)
    {
        a.h_hdr->t_start_ns = t_start;
        a.h_hdr->t_end_ns = cc_globaltimer();
        cc_fence_system();
        *reinterpret_cast<volatile unsigned int*>(&a.h_hdr->flag) = a.ticket;
    }
}

// =====================================================================================================
// K6  publish side: what callers read from `range_image_` inside the callbacks, gathered ON THE DEVICE.
//     k_export_cells   every field of `Point` (hpp:126-161) of the cells of a column range as one packed 128-byte record
//                      per cell, written straight to page-locked host memory: replaces 16 field-wise copies + a host
//                      scatter per cc_read_columns call.
// =====================================================================================================
struct alignas(16) CcCell // == cc_cell_t (include/cc_b200.h)
{
    float x, y, z, distance;
    float azimuth_angle, inclination_angle;
    double continuous_azimuth_angle;
    long long global_column_index; // -1 in columns that have not been segmented (cleared value, cpp:1117)
    unsigned long long stamp, globally_unique_point_index, firing_index;
    unsigned long long id;
    double finished_at_continuous_azimuth_angle; // tree roots only, else 0 (cleared value)
    long long tree_root_gcol;                    // -1 = not associated
    long long first_parent_gcol;                 // the point whose child_points list holds this point (cpp:663), -1 = none
    unsigned int tree_num_points, cluster_width; // tree roots only
    int tree_root_row, first_parent_row;
    unsigned short number_of_visited_neighbors, pad0_;
    unsigned char intensity, ground_point_label, debug_ground_point_label, is_ignored;
    unsigned char belongs_to_finished_cluster, pad1_[7];
};
static_assert(sizeof(CcCell) == 128, "cc_cell_t layout");

CC_DEV CcCell cc_gather_cell(const CcDevCfg& cfg, const CcDevPtrs& p, long long gcol, int row)
{
    const int R = cfg.R;
    const int local = cc_local_col(gcol, cfg.ringcols);
    const size_t cell = static_cast<size_t>(local) * R + row;
    CcCell c;
    const float4 q = p.pos[cell];
    const uchar4 l = p.lab[cell];
    // NaN bit patterns: the reference's NaNs all descend from std::nanf("") / std::nan("") (0x7fc00000 / 0x7ff8...0), which
    // x86 arithmetic propagates unchanged; the GPU's arithmetic produces 0x7fffffff instead. Published bytes (PointCloud2
    // payloads) carry the reference's pattern.
    auto canon = [](float v) { return v != v ? ccm::u2f(0x7fc00000u) : v; };
    auto canon64 = [](double v) { return v != v ? __longlong_as_double(0x7ff8000000000000LL) : v; };
    c.x = canon(q.x);
    c.y = canon(q.y);
    c.z = canon(q.z);
    c.distance = canon(q.w);
    c.azimuth_angle = canon(p.azimuth[cell]);
    c.inclination_angle = canon(p.incl[cell]);
    c.continuous_azimuth_angle = canon64(p.cont_az[cell]);
    c.global_column_index = p.slot_gcol[local] == gcol ? gcol : -1; // refilled by segmentation (cpp:347-350)
    c.stamp = p.stamp[cell];
    c.globally_unique_point_index = p.guid[cell];
    c.firing_index = p.firing_index[cell];
    c.id = p.cid[cell];
    const unsigned int root = p.tparent[cell];
    const bool is_root = root == static_cast<unsigned int>(cell);
    c.finished_at_continuous_azimuth_angle = is_root ? cc_ord2d(p.tfinish[cell]) : 0.0;
    c.tree_root_gcol = root == CC_NONE ? -1 : p.slot_gcol[root / R];
    c.tree_root_row = root == CC_NONE ? 0 : static_cast<int>(root % R);
    const unsigned int fp = p.tfirst[cell];
    const bool has_parent = fp != CC_NONE && fp != static_cast<unsigned int>(cell) && root != CC_NONE;
    c.first_parent_gcol = has_parent ? p.slot_gcol[fp / R] : -1;
    c.first_parent_row = has_parent ? static_cast<int>(fp % R) : 0;
    c.tree_num_points = is_root ? p.tnpoints[cell] : 0u;
    c.cluster_width = is_root ? static_cast<unsigned int>(p.tmaxcol[cell] - gcol + 1) : 0u;
    c.number_of_visited_neighbors = p.visited[cell];
    c.pad0_ = 0;
    c.intensity = l.w;
    c.ground_point_label = l.x;
    c.debug_ground_point_label = l.y;
    c.is_ignored = l.z;
    c.belongs_to_finished_cluster = is_root && p.tstate[cell] != 0u ? 1 : 0;
    for (int i = 0; i < 7; i++)
        c.pad1_[i] = 0;
    return c;
}

__global__ void k_export_cells(CcDevCfg cfg, CcDevPtrs p, long long from, int ncols, CcCell* out)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    const long long total = static_cast<long long>(ncols) * cfg.R;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const long long ci = i / cfg.R;
        const int row = static_cast<int>(i - ci * cfg.R);
        union // one full 128-byte line per thread, as eight 16-byte stores
        {
            CcCell c;
            uint4 v[8];
        } u;
        u.c = cc_gather_cell(cfg, p, from + ci, row);
        uint4* dst = reinterpret_cast<uint4*>(out + i);
#pragma unroll
        for (int w = 0; w < 8; w++)
            dst[w] = u.v[w];
    }
}

// =====================================================================================================
// K6 (continued)
//     k_pack_cloud     the sensor_msgs/PointCloud2 payload the ROS node publishes for a range of columns or for a finished
//                      cluster (ros_utils.cpp:11-77: columnToPointCloud / clusterToPointCloud; point layout of
//                      ros_utils.cpp:108-243, field values of addPointToMessage ros_utils.cpp:245-298), byte for byte:
//                      fields packed without padding, 76 bytes per point up to the ground-segmentation stage, 116 with
//                      the clustering fields; column clouds are row-major images (height = rows, width = columns).
//                      A warp assembles 32 consecutive points in shared memory and writes them out as 16-byte vectors
//                      straight into page-locked host memory.
//     k_child_counts   child_points.size() of the cells of a column range (children name their parent, tfirst).
// =====================================================================================================
#define CC_CLOUD_STEP_GROUND 76
#define CC_CLOUD_STEP_CLUSTER 116

__global__ void k_child_counts(CcDevCfg cfg, CcDevPtrs p, long long from, int ncols, int ahead, unsigned int* counts)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    // children sit at most max_steps_in_row columns ahead of their parent (cpp:704-705)
    const long long total = static_cast<long long>(ncols + ahead) * cfg.R;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
    {
        const long long gcol = from + i / cfg.R;
        const int row = static_cast<int>(i % cfg.R);
        const int local = cc_local_col(gcol, cfg.ringcols);
        if (p.slot_gcol[local] != gcol)
            continue;
        const size_t cell = static_cast<size_t>(local) * cfg.R + row;
        const unsigned int fp = p.tfirst[cell];
        if (fp == CC_NONE || fp == static_cast<unsigned int>(cell) || p.tparent[cell] == CC_NONE)
            continue;
        const long long pg = p.slot_gcol[fp / cfg.R];
        if (pg >= from && pg < from + ncols)
            atomicAdd(counts + (pg - from) * cfg.R + fp % cfg.R, 1u);
    }
}

// the words of one point record (stage 2: 19 words, stage 3: 29 words)
CC_DEV void cc_cloud_record(const CcDevCfg& cfg, const CcDevPtrs& p, long long gcol, int row, unsigned int nchild, int nwords,
                            unsigned int* w, unsigned long long* stamp_out)
{
    const CcCell c = cc_gather_cell(cfg, p, gcol, row);
    *stamp_out = c.stamp;
    auto lo = [](double d) { return static_cast<unsigned int>(static_cast<unsigned long long>(__double_as_longlong(d))); };
    auto hi = [](double d) { return static_cast<unsigned int>(static_cast<unsigned long long>(__double_as_longlong(d)) >> 32); };
    const int local = cc_local_col(gcol, cfg.ringcols);
    const int local_column_index = c.global_column_index >= 0 ? local : -1;        // cpp:348-350 / cleared value
    const int row_index = cc_isnan(c.distance) ? -1 : row;                           // set by insertion only (cpp:235)
    w[0] = ccm::f2u(c.x);
    w[1] = ccm::f2u(c.y);
    w[2] = ccm::f2u(c.z);
    const double fi = static_cast<double>(c.firing_index); // (*) UINT64 fields travel as FLOAT64 (ros_utils.cpp:124-126)
    w[3] = lo(fi);
    w[4] = hi(fi);
    // bytes 21 .. 75 are one byte off the word grid (the UINT8 intensity at byte 20): assembled as a stream of words ...
    unsigned int sw[14];
    const double gp = static_cast<double>(c.globally_unique_point_index);
    sw[0] = lo(gp);
    sw[1] = hi(gp);
    sw[2] = static_cast<unsigned int>(c.stamp / 1000000000ull); // ros::Time::fromNSec
    sw[3] = static_cast<unsigned int>(c.stamp % 1000000000ull);
    sw[4] = ccm::f2u(c.distance);
    sw[5] = ccm::f2u(c.azimuth_angle);
    sw[6] = ccm::f2u(c.inclination_angle);
    sw[7] = lo(c.continuous_azimuth_angle);
    sw[8] = hi(c.continuous_azimuth_angle);
    const double gc = static_cast<double>(c.global_column_index);
    sw[9] = lo(gc);
    sw[10] = hi(gc);
    sw[11] = (static_cast<unsigned int>(local_column_index) & 0xffffu) | (static_cast<unsigned int>(row_index) << 16);
    const unsigned int hog = 0x7fc00000u; // height_over_ground is never computed by the reference (only cleared, cpp:1126)
    sw[12] = c.ground_point_label | (static_cast<unsigned int>(c.debug_ground_point_label) << 8) | (hog << 16);
    sw[13] = (hog >> 16) | ((c.is_ignored ? 9u /* BLUE */ : 105u /* ORANGE */) << 16); // ros_utils.cpp:285
    // ... and shifted into place
    unsigned int carry = c.intensity;
#pragma unroll
    for (int k = 0; k < 14; k++)
    {
        w[5 + k] = (sw[k] << 8) | carry;
        carry = sw[k] >> 24;
    }
    if (nwords > 19)
    {
        w[19] = lo(c.finished_at_continuous_azimuth_angle);
        w[20] = hi(c.finished_at_continuous_azimuth_angle);
        w[21] = (nchild & 0xffffu) | (static_cast<unsigned int>(c.tree_root_row) << 16);
        // Point::tree_root_ holds the LOCAL ring column (cpp:661, 814), -1 = not associated
        const double rc = c.tree_root_gcol >= 0 ? static_cast<double>(c.tree_root_gcol % cfg.ringcols) : -1.0;
        w[22] = lo(rc);
        w[23] = hi(rc);
        w[24] = c.number_of_visited_neighbors;
        const double tid = c.tree_root_gcol >= 0 ? static_cast<double>(static_cast<unsigned long long>(c.tree_root_gcol) * cfg.R +
                                                                       static_cast<unsigned long long>(c.tree_root_row))
                                                 : 0.0; // cpp:662, 815
        w[25] = lo(tid);
        w[26] = hi(tid);
        const double id = static_cast<double>(c.id);
        w[27] = lo(id);
        w[28] = hi(id);
    }
}

// mode 0: columns [from, from + ncols): message point o = row * ncols + column (ros_utils.cpp:62); mode 1: the points of
// `list` in order (clusterToPointCloud). `counts`: child counts of the cells of columns [cfrom, ...) (k_child_counts) or
// null. `min_stamp`: smallest non-zero point stamp (the column message's header stamp, ros_utils.cpp:66-74).
__global__ void __launch_bounds__(128) k_pack_cloud(CcDevCfg cfg, CcDevPtrs p, int mode, long long from, int ncols,
                                                    const CcClusterPoint* list, int npoints, int nwords, long long cfrom,
                                                    const unsigned int* counts, unsigned char* out, unsigned long long* min_stamp)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    CC_SMEM(smem);
    const int lane = threadIdx.x % CC_WARP, warp = threadIdx.x / CC_WARP;
    const int nwarps = (blockDim.x + CC_WARP - 1) / CC_WARP;
    unsigned int* tile = reinterpret_cast<unsigned int*>(smem) + static_cast<size_t>(warp) * CC_WARP * 29;
    const long long total = mode == 0 ? static_cast<long long>(ncols) * cfg.R : npoints;
    unsigned long long smin = ~0ull;
    for (long long o0 = (static_cast<long long>(blockIdx.x) * nwarps + warp) * CC_WARP; o0 < total;
         o0 += static_cast<long long>(gridDim.x) * nwarps * CC_WARP)
    {
        const long long o = o0 + lane;
        if (o < total)
        {
            long long gcol;
            int row;
            if (mode == 0)
            {
                row = static_cast<int>(o / ncols);
                gcol = from + o % ncols;
            }
            else
            {
                gcol = list[o].gcol;
                row = list[o].row;
            }
            const unsigned int nchild = counts ? counts[(gcol - cfrom) * cfg.R + row] : 0u;
            unsigned int w[29];
            unsigned long long stamp;
            cc_cloud_record(cfg, p, gcol, row, nchild, nwords, w, &stamp);
            if (stamp != 0ull && stamp < smin)
                smin = stamp;
            for (int k = 0; k < nwords; k++)
                tile[lane * nwords + k] = w[k]; // stride 19 / 29 words: conflict free
        }
        __syncwarp();
        // 32 records are 32 * 76 = 2432 or 32 * 116 = 3712 contiguous bytes, both multiples of 16
        const long long cnt = total - o0 < CC_WARP ? total - o0 : CC_WARP;
        const int nbytes = static_cast<int>(cnt) * nwords * 4;
        unsigned char* dst = out + o0 * nwords * 4;
        const int nvec = nbytes / 16;
        const uint4* src4 = reinterpret_cast<const uint4*>(tile);
        for (int v = lane; v < nvec; v += CC_WARP)
            reinterpret_cast<uint4*>(dst)[v] = src4[v];
        for (int b = nvec * 16 + lane * 4; b < nbytes; b += CC_WARP * 4) // ragged tail of the last warp
            *reinterpret_cast<unsigned int*>(dst + b) = tile[b / 4];
        __syncwarp();
    }
    if (min_stamp && smin != ~0ull)
        atomicMin(min_stamp, smin);
}

// Every message of a push in ONE launch (the node publishes a message per finished-column callback and per finished
// cluster: dozens to hundreds per push): a request table names, per message, what to pack and where its payload starts.
struct CcPackRequest // == cc_pack_request_t + the layout the host computed
{
    int kind;             // 0 ground-stage columns (76-byte points), 1 clustered columns (116), 2 cluster (116)
    int npoints;          // points of the message
    long long from;       // first column (kinds 0 / 1)
    int ncols;            // columns (kinds 0 / 1)
    int list_offset;      // first member-list entry (kind 2)
    long long out_offset; // byte offset of the payload in the output buffer (16-byte aligned)
    int first_task, pad_; // first 32-point task of the request
};

__global__ void __launch_bounds__(128) k_pack_requests(CcDevCfg cfg, CcDevPtrs p, const CcPackRequest* req, int nreq, int ntasks,
                                                       const CcClusterPoint* list, long long cfrom, const unsigned int* counts,
                                                       unsigned char* out, unsigned long long* min_stamps)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    CC_SMEM(smem);
    const int lane = threadIdx.x % CC_WARP, warp = threadIdx.x / CC_WARP;
    const int nwarps = (blockDim.x + CC_WARP - 1) / CC_WARP;
    unsigned int* tile = reinterpret_cast<unsigned int*>(smem) + static_cast<size_t>(warp) * CC_WARP * 29;
    for (int task = blockIdx.x * nwarps + warp; task < ntasks; task += gridDim.x * nwarps)
    {
        int a = 0, b = nreq - 1; // last request with first_task <= task
        while (a < b)
        {
            const int mid = (a + b + 1) >> 1;
            if (req[mid].first_task <= task)
                a = mid;
            else
                b = mid - 1;
        }
        const CcPackRequest rq = req[a];
        const int nwords = rq.kind == 0 ? CC_CLOUD_STEP_GROUND / 4 : CC_CLOUD_STEP_CLUSTER / 4;
        const long long o0 = static_cast<long long>(task - rq.first_task) * CC_WARP;
        const long long o = o0 + lane;
        unsigned long long smin = ~0ull;
        if (o < rq.npoints)
        {
            long long gcol;
            int row;
            if (rq.kind != 2)
            {
                row = static_cast<int>(o / rq.ncols);
                gcol = rq.from + o % rq.ncols;
            }
            else
            {
                gcol = list[rq.list_offset + o].gcol;
                row = list[rq.list_offset + o].row;
            }
            const unsigned int nchild = rq.kind != 0 && counts ? counts[(gcol - cfrom) * cfg.R + row] : 0u;
            unsigned int w[29];
            unsigned long long stamp;
            cc_cloud_record(cfg, p, gcol, row, nchild, nwords, w, &stamp);
            if (stamp != 0ull)
                smin = stamp;
            for (int k = 0; k < nwords; k++)
                tile[lane * nwords + k] = w[k];
        }
        __syncwarp();
        const long long cnt = rq.npoints - o0 < CC_WARP ? rq.npoints - o0 : CC_WARP;
        const int nbytes = static_cast<int>(cnt) * nwords * 4;
        unsigned char* dst = out + rq.out_offset + o0 * nwords * 4;
        const int nvec = nbytes / 16;
        const uint4* src4 = reinterpret_cast<const uint4*>(tile);
        for (int v = lane; v < nvec; v += CC_WARP)
            reinterpret_cast<uint4*>(dst)[v] = src4[v];
        for (int bb = nvec * 16 + lane * 4; bb < nbytes; bb += CC_WARP * 4)
            *reinterpret_cast<unsigned int*>(dst + bb) = tile[bb / 4];
        __syncwarp();
        // smallest non-zero point stamp of the message
        for (int off = CC_WARP / 2; off > 0; off >>= 1)
        {
            const unsigned long long other = __shfl_xor_sync(CC_FULL_MASK, smin, off);
            smin = other < smin ? other : smin;
        }
        if (lane == 0 && smin != ~0ull)
            atomicMin(min_stamps + a, smin);
    }
}

// =====================================================================================================
// K4' the list phases of a whole-push finish pass over ONE thread-block cluster (16 CTAs, cluster barriers) instead of
//     one CTA with block barriers: aggregate | decide + mark | compacted list back (CTA 1) beside the per-column
//     first-unpublished columns and the end-of-push bookkeeping (CTA 0). Same device functions as the fused kernel's tail.
// =====================================================================================================
__global__ void __launch_bounds__(512, 1) k_fin_cluster(CcDevCfg cfg, CcDevPtrs p, unsigned int seq, int last, int smem_bytes,
                                                        CcDevState* snap)
{
    CC_PDL_ENTER();
    const CcGrid g = cc_grid_cluster();
    CcTraceScope cc_trace_scope(p.trace, CC_KID_fin_all, g.bid);
    if (p.st->halted) // an earlier push in flight could not be committed speculatively (see k_halt)
    {
        if (snap && g.bid == 0)
            d_state_snapshot(p, snap);
        return;
    }
    {
        CcTraceScope tr(p.trace, CC_KID_fin_init, g.bid);
        d_fin_init(g, cfg, p, 0, -1, 1);
    }
    cc_cluster_sync();
    {
        CcTraceScope tr(p.trace, CC_KID_fin_agg, g.bid);
        d_fin_agg(g, cfg, p, 1);
    }
    cc_cluster_sync();
    {
        CcTraceScope tr(p.trace, CC_KID_fin_decide, g.bid);
        d_fin_decide_mark(g, cfg, p, seq);
    }
    cc_cluster_sync();
    if (g.bid == (g.nb >= 2 ? 1 : 0))
    {
        CcTraceScope tr(p.trace, CC_KID_fin_copyback, g.bid);
        CcGrid g1;
        g1.bid = 0;
        g1.nb = 1;
        d_fin_copyback(g1, p, 1);
    }
    {
        // every CTA answers its share of the columns (each with its own copy of the prefix maxima in shared memory)
        CcTraceScope tr(p.trace, CC_KID_fin_columns, g.bid);
        __syncthreads();
        d_fin_columns(g, cfg, p, 1, (smem_bytes - static_cast<int>(blockDim.x) * 8) / 4, g.bid, g.nb);
    }
    if (g.bid == 0)
    {
        __syncthreads();
        if (last && threadIdx.x == 0)
            d_push_done(p, 1);
        if (snap)
        {
            __syncthreads();
            d_state_snapshot(p, snap);
        }
    }
}

// ---- device math self-test (bit equality with the host libm) ----
__global__ void k_selftest_math(int op, int n, const float* a, const float* b, float* out)
{
    CC_PDL_ENTER();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        if (op == 0)
            out[i] = ccm::atan2f_glibc(a[i], b[i]);
        else if (op == 1)
            out[i] = ccm::asinf_glibc(a[i]);
        else
            out[i] = ccm::atanf_glibc(a[i]);
    }
}

#endif
