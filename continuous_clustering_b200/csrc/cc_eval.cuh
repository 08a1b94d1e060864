// cc_eval.cuh -- SURVEY 8f-4: the per-frame evaluation metrics of the reference's kitti_evaluation.cpp on the device.
//   ground segmentation  true / false positives / negatives against the SemanticKITTI ground classes (cpp: kitti_evaluation.cpp:44-84)
//   clustering           over- / under-segmentation entropy between the ground-truth clusters and the detections
//                        (kitti_evaluation.cpp:86-146): OSE = - sum over gt clusters g, over detection labels d of its points
//                        (0 = "no detection" included) of f log f with f = n(g, d) / n(g); USE the same with the roles
//                        swapped, skipping detections without any ground-truth point.
// The contingency table n(g, d) lives in an open-addressing hash table in HBM (keys claimed with atomicCAS, counts with
// atomicAdd), the marginals n(g), n(d) in a second one; one pass over the points fills both, one pass over the table
// slots adds up the entropy terms. HBM bound: 11 B read per point; a frame of ~120 k points is a few microseconds.
#ifndef CC_EVAL_CUH
#define CC_EVAL_CUH

#include "cc_kernels.cuh"

#define CC_EVAL_EMPTY 0xffffffffffffffffull

struct CcEvalPtrs
{
    const unsigned short* semantic_label;
    const unsigned char* is_ground;
    const unsigned int* gt_label;  // euclidean_clustering_label, 0 = none
    const unsigned int* det_label; // detection_label (Point::id), 0 = none
    int n;
    unsigned long long* pair_keys; // (gt << 32) | det
    unsigned int* pair_cnt;
    int pair_cap; // power of two
    unsigned long long* marg_keys; // (side << 32) | label, side 0 = ground truth, 1 = detection
    unsigned int* marg_cnt;
    int marg_cap;
    unsigned long long* counts; // tp, fn, fp, tn
    double* entropy;            // over-segmentation, under-segmentation
};

// SemanticKITTI label ids (kitti_loader.cpp:566-603; data constants of the dataset)
#define CC_SK_UNLABELED 0
CC_DEV bool cc_sk_is_ground(unsigned short l)
{
    return l == 60 /* lane-marking */ || l == 40 /* road */ || l == 44 /* parking */ || l == 48 /* sidewalk */ ||
           l == 49 /* other-ground */ || l == 72 /* terrain */;
}

CC_DEV unsigned int cc_hash64(unsigned long long x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return static_cast<unsigned int>(x);
}

// claims (or finds) the slot of `key`; returns its index
CC_DEV int cc_table_slot(unsigned long long* keys, int cap, unsigned long long key)
{
    unsigned int s = cc_hash64(key) & static_cast<unsigned int>(cap - 1);
    while (true)
    {
        const unsigned long long cur = keys[s];
        if (cur == key)
            return static_cast<int>(s);
        if (cur == CC_EVAL_EMPTY)
        {
            const unsigned long long old = atomicCAS(keys + s, CC_EVAL_EMPTY, key);
            if (old == CC_EVAL_EMPTY || old == key)
                return static_cast<int>(s);
        }
        s = (s + 1) & static_cast<unsigned int>(cap - 1);
    }
}
CC_DEV unsigned int cc_table_get(const unsigned long long* keys, const unsigned int* cnt, int cap, unsigned long long key)
{
    unsigned int s = cc_hash64(key) & static_cast<unsigned int>(cap - 1);
    while (true)
    {
        const unsigned long long cur = keys[s];
        if (cur == key)
            return cnt[s];
        if (cur == CC_EVAL_EMPTY)
            return 0u;
        s = (s + 1) & static_cast<unsigned int>(cap - 1);
    }
}

__global__ void k_eval_clear(CcEvalPtrs e)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = t; i < e.pair_cap; i += nt)
    {
        e.pair_keys[i] = CC_EVAL_EMPTY;
        e.pair_cnt[i] = 0u;
    }
    for (int i = t; i < e.marg_cap; i += nt)
    {
        e.marg_keys[i] = CC_EVAL_EMPTY;
        e.marg_cnt[i] = 0u;
    }
    if (t < 4)
        e.counts[t] = 0ull;
    if (t < 2)
        e.entropy[t] = 0.0;
}

__global__ void k_eval_count(CcEvalPtrs e)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    const int lane = threadIdx.x % CC_WARP;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < e.n; i0 += gridDim.x * blockDim.x)
    {
        const int i = i0 + lane;
        int cls = -1; // 0 tp, 1 fn, 2 fp, 3 tn
        if (i < e.n)
        {
            const unsigned short sem = e.semantic_label[i];
            if (sem != CC_SK_UNLABELED) // kitti_evaluation.cpp:49-50
            {
                const bool gt_ground = cc_sk_is_ground(sem), seg_ground = e.is_ground[i] != 0;
                cls = gt_ground ? (seg_ground ? 0 : 1) : (seg_ground ? 2 : 3);
            }
            const unsigned int g = e.gt_label[i], d = e.det_label[i];
            if (g != 0u || d != 0u)
                atomicAdd(e.pair_cnt + cc_table_slot(e.pair_keys, e.pair_cap, (static_cast<unsigned long long>(g) << 32) | d), 1u);
            if (g != 0u)
                atomicAdd(e.marg_cnt + cc_table_slot(e.marg_keys, e.marg_cap, static_cast<unsigned long long>(g)), 1u);
            if (d != 0u)
                atomicAdd(e.marg_cnt + cc_table_slot(e.marg_keys, e.marg_cap, (1ull << 32) | d), 1u);
        }
        for (int c = 0; c < 4; c++)
        {
            const unsigned int m = __ballot_sync(CC_FULL_MASK, cls == c);
            if (m && lane == 0)
                atomicAdd(e.counts + c, static_cast<unsigned long long>(__popc(m)));
        }
    }
}

__global__ void k_eval_entropy(CcEvalPtrs e)
{
    CC_PDL_ENTER(); // launched with programmatic stream serialization (cc_platform.h): wait for the kernel before
    double ose = 0.0, use = 0.0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < e.pair_cap; s += gridDim.x * blockDim.x)
    {
        const unsigned long long key = e.pair_keys[s];
        if (key == CC_EVAL_EMPTY)
            continue;
        const unsigned int g = static_cast<unsigned int>(key >> 32), d = static_cast<unsigned int>(key);
        const double c = static_cast<double>(e.pair_cnt[s]);
        if (g != 0u) // kitti_evaluation.cpp:101-118
        {
            const double frac = c / static_cast<double>(cc_table_get(e.marg_keys, e.marg_cnt, e.marg_cap, static_cast<unsigned long long>(g)));
            ose -= frac * log(frac);
        }
        if (d != 0u) // kitti_evaluation.cpp:121-144
        {
            const unsigned int nd = cc_table_get(e.marg_keys, e.marg_cnt, e.marg_cap, (1ull << 32) | d);
            const unsigned int n0 = cc_table_get(e.pair_keys, e.pair_cnt, e.pair_cap, static_cast<unsigned long long>(d));
            if (n0 != nd) // a detection without any ground-truth point is ignored (cpp:129-131)
            {
                const double frac = c / static_cast<double>(nd);
                use -= frac * log(frac);
            }
        }
    }
    // block reduction, then one atomic per block
    __shared__ double sh[2][32];
    const int lane = threadIdx.x % CC_WARP, warp = threadIdx.x / CC_WARP;
    for (int o = CC_WARP / 2; o > 0; o >>= 1)
    {
        ose += __shfl_xor_sync(CC_FULL_MASK, ose, o);
        use += __shfl_xor_sync(CC_FULL_MASK, use, o);
    }
    if (lane == 0)
    {
        sh[0][warp] = ose;
        sh[1][warp] = use;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const int nw = (blockDim.x + CC_WARP - 1) / CC_WARP;
        double a = 0.0, b = 0.0;
        for (int w = 0; w < nw; w++)
        {
            a += sh[0][w];
            b += sh[1][w];
        }
        if (a != 0.0)
            atomicAdd(e.entropy + 0, a);
        if (b != 0.0)
            atomicAdd(e.entropy + 1, b);
    }
}

#endif
