// cc_kitti.cuh -- SURVEY 8f-2: the KITTI replay front-end of the reference on the device: one frame of the dataset
// (an unorganised list of x, y, z, intensity points in file order) becomes the 2200 pseudo firings kitti_demo feeds into
// addFiring (kitti_demo.cpp:369-403), already resident in HBM in the RawPoint layout, ready for cc_submit_firings_device:
//   k_kitti_flags / k_kitti_rows      KittiLoader::recoverLaserIndices   kitti_loader.cpp:47-99   (row = number of azimuth
//                                     wraps up to the point: flag per point, prefix sum over the frame in segments)
//   k_kitti_undo_ego                  KittiLoader::undoEgoMotionCorrection kitti_loader.cpp:176-210 (per point: azimuth ->
//                                     1 ms time bin -> rigid transform of that bin, f64 in the oracle's evaluation order)
//                                     + the column of generateRangeImage (kitti_loader.cpp:120-128)
//   k_kitti_range_image               KittiLoader::generateRangeImage    kitti_loader.cpp:101-174 (cell collisions: try the
//                                     right neighbour, then the left one, else overwrite -- in file order, which only
//                                     couples points of the same row: one CTA per row, the row's occupancy in shared memory)
//   k_kitti_firings                   KittiDemo::makePseudoFiringFromRangeImageColumn kitti_demo.cpp:123-159
// The per-bin transforms and the per-column poses (KittiLoader::interpolate, kitti_loader.cpp:297-328: slerp) are a few
// thousand f64 operations per frame and are prepared on the host (cc_api.cu).
// HBM bound: 16 B read + 12 B written per point, 48 B written per range-image cell.
// Every kernel works for any block size (the CPU emulation build of the tests runs one-thread blocks).
#ifndef CC_KITTI_CUH
#define CC_KITTI_CUH

#include "cc_kernels.cuh"

#define CC_KITTI_W 2200 /* KittiLoader::RANGE_IMAGE_WIDTH  kitti_loader.hpp:86 */
#define CC_KITTI_H 64   /* KittiLoader::RANGE_IMAGE_HEIGHT kitti_loader.hpp:85 */
#define CC_KITTI_SEG 1024 /* points per scan segment */

struct CcKittiPtrs
{
    const float4* xyzi; // the frame as stored in the .bin file
    int n;
    unsigned char* wrap_flag; // 1 where the azimuth jumps back by more than 0.7 rad: a new row starts (cpp:62-71)
    int* seg_sums;            // flags per segment of CC_KITTI_SEG points
    unsigned char* laser_index;
    float* uncorrected;       // x, y, z after undoEgoMotionCorrection (3 floats per point)
    int* column;              // column of generateRangeImage before collision handling
    int* row_start;           // [CC_KITTI_H + 2] first point of every row (n = the frame has no such row); [H] = first
                              // point after the last row (these keep laser index 0, kitti_loader.cpp:74-76), [H + 1] = n
    int* cell_point;          // [H * W] original index of the point that holds the cell, -1 = empty
    const double* bin_tf;     // [n_bins][12] velodyne_from_velodyne of every 1 ms bin
    int n_bins;
    unsigned long long stamp_start, stamp_end;
};

// std::atan2(y, x) of two floats is atan2f (kitti_loader.cpp:61); made monotonic over 0 .. 2 pi in double (cpp:64-65)
CC_DEV double cc_kitti_monotonic_azimuth(const float4 q)
{
    const double a = static_cast<double>(ccm::atan2f_glibc(q.y, q.x));
    return a < 0 ? a + (2 * M_PI) : a;
}

// wrap flag of every point and their number per segment
__global__ void k_kitti_flags(CcKittiPtrs k)
{
    CC_PDL_ENTER();
    __shared__ int sh[32];
    const int T = blockDim.x, t = threadIdx.x;
    const int nseg = (k.n + CC_KITTI_SEG - 1) / CC_KITTI_SEG;
    if (blockIdx.x == 0)
        for (int r = t; r < CC_KITTI_H + 2; r += T)
            k.row_start[r] = r == 0 ? 0 : k.n;
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x)
    {
        int cnt = 0;
        for (int j = t; j < CC_KITTI_SEG; j += T)
        {
            const int i = seg * CC_KITTI_SEG + j;
            if (i >= k.n)
                break;
            int flag = 0;
            if (i > 0) // (prev_azimuth_monotonic >= 0 holds from the second point on; a NaN never compares below)
            {
                const double cur = cc_kitti_monotonic_azimuth(k.xyzi[i]), prev = cc_kitti_monotonic_azimuth(k.xyzi[i - 1]);
                flag = cur - prev < -0.7 ? 1 : 0;
            }
            k.wrap_flag[i] = static_cast<unsigned char>(flag);
            cnt += flag;
        }
        const int excl = cc_block_exclusive_scan(sh, cnt, 0, CcOpAddI32());
        if (t == T - 1)
            k.seg_sums[seg] = excl + cnt;
    }
}

// row of every point = wraps up to and including it (rows beyond the last one: the reference stops counting and the
// points keep laser index 0, kitti_loader.cpp:74-76); first point of every row
__global__ void k_kitti_rows(CcKittiPtrs k)
{
    CC_PDL_ENTER();
    __shared__ int sh[32];
    __shared__ int sh_base;
    const int T = blockDim.x, t = threadIdx.x;
    const int nseg = (k.n + CC_KITTI_SEG - 1) / CC_KITTI_SEG;
    const int per = (CC_KITTI_SEG + T - 1) / T; // consecutive points per thread
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x)
    {
        // wraps before this segment (a frame has ~120 segments: every CTA adds them up itself)
        int part = 0;
        for (int s = t; s < seg; s += T)
            part += k.seg_sums[s];
        const int before_me = cc_block_exclusive_scan(sh, part, 0, CcOpAddI32());
        if (t == T - 1)
            sh_base = before_me + part;
        __syncthreads();
        const int base = sh_base;
        const int a = seg * CC_KITTI_SEG + t * per;
        int b = a + per;
        b = b > (seg + 1) * CC_KITTI_SEG ? (seg + 1) * CC_KITTI_SEG : b;
        b = b > k.n ? k.n : b;
        int cnt = 0;
        for (int i = a; i < b; i++)
            cnt += k.wrap_flag[i];
        int row = base + cc_block_exclusive_scan(sh, cnt, 0, CcOpAddI32());
        for (int i = a; i < b; i++)
        {
            if (k.wrap_flag[i])
            {
                row++;
                if (row <= CC_KITTI_H)
                    k.row_start[row] = i;
            }
            k.laser_index[i] = row < CC_KITTI_H ? static_cast<unsigned char>(row) : 0;
        }
        __syncthreads();
    }
}

__global__ void k_kitti_undo_ego(CcKittiPtrs k)
{
    CC_PDL_ENTER();
    const double duration = static_cast<double>(k.stamp_end - k.stamp_start);
    const double column_width = (2 * M_PI) / CC_KITTI_W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k.n; i += gridDim.x * blockDim.x)
    {
        const float4 q = k.xyzi[i];
        // kitti_loader.cpp:201-209 (the azimuth of the point as stored)
        const double frac = (M_PI - static_cast<double>(ccm::atan2f_glibc(q.y, q.x))) / (2.0 * M_PI);
        int bin = static_cast<int>((frac * duration) / 1000000.0);
        bin = bin < 0 ? 0 : (bin >= k.n_bins ? k.n_bins - 1 : bin); // (the reference indexes its table unchecked)
        double o[3];
        cc_iso_apply(k.bin_tf + 12 * bin, static_cast<double>(q.x), static_cast<double>(q.y), static_cast<double>(q.z), o);
        const float x = static_cast<float>(o[0]), y = static_cast<float>(o[1]), z = static_cast<float>(o[2]);
        k.uncorrected[3 * i + 0] = x;
        k.uncorrected[3 * i + 1] = y;
        k.uncorrected[3 * i + 2] = z;
        // kitti_loader.cpp:120-128 (the azimuth of the UNCORRECTED point)
        const double az = static_cast<double>(ccm::atan2f_glibc(y, x));
        int col = static_cast<int>((M_PI - az) / column_width);
        if (col == CC_KITTI_W)
            col--;
        col = col < 0 ? 0 : (col >= CC_KITTI_W ? CC_KITTI_W - 1 : col); // (NaN coordinates: out of bounds in the reference)
        k.column[i] = col;
    }
}

// one CTA per row: thread 0 applies the collision rule to the row's points in file order on the occupancy array in
// shared memory (kitti_loader.cpp:130-163); all threads stage the columns and write the row out
#define CC_KITTI_STAGE 1024
__global__ void __launch_bounds__(256) k_kitti_range_image(CcKittiPtrs k)
{
    CC_PDL_ENTER();
    __shared__ int occ[CC_KITTI_W];
    __shared__ int cols[CC_KITTI_STAGE];
    const int row = blockIdx.x;
    if (row >= CC_KITTI_H)
        return;
    for (int c = threadIdx.x; c < CC_KITTI_W; c += blockDim.x)
        occ[c] = -1;
    // the row's points: [row_start[row], row_start[row + 1]); row 0 also gets the points behind the last row, afterwards
    for (int part = 0; part < (row == 0 ? 2 : 1); part++)
    {
        const int a = part == 0 ? k.row_start[row] : k.row_start[CC_KITTI_H];
        const int b = part == 0 ? k.row_start[row + 1] : k.n;
        for (int c0 = a; c0 < b; c0 += CC_KITTI_STAGE)
        {
            const int cnt = b - c0 < CC_KITTI_STAGE ? b - c0 : CC_KITTI_STAGE;
            __syncthreads();
            for (int j = threadIdx.x; j < cnt; j += blockDim.x)
                cols[j] = k.column[c0 + j];
            __syncthreads();
            if (threadIdx.x == 0)
                for (int j = 0; j < cnt; j++)
                {
                    int col = cols[j];
                    if (occ[col] >= 0)
                    {
                        if (col + 1 < CC_KITTI_W && occ[col + 1] < 0)
                            col = col + 1;
                        else if (col - 1 >= 0 && occ[col - 1] < 0)
                            col = col - 1;
                    }
                    occ[col] = c0 + j; // (an occupied cell is overwritten when both neighbours are taken too)
                }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < CC_KITTI_W; c += blockDim.x)
        k.cell_point[row * CC_KITTI_W + c] = occ[c];
}

// kitti_demo.cpp:123-159: firing = column; every row's cell becomes a RawPoint (48 bytes)
__global__ void k_kitti_firings(CcKittiPtrs k, CcRawPoint* firings, int sequence_index, int frame_index)
{
    CC_PDL_ENTER();
    const int total = CC_KITTI_W * CC_KITTI_H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const int col = i / CC_KITTI_H, row = i - col * CC_KITTI_H;
        const int pt = k.cell_point[row * CC_KITTI_W + col];
        const double elapsed_ratio = static_cast<double>(col) / (CC_KITTI_W - 1);
        const double elapsed_time = static_cast<double>(k.stamp_end - k.stamp_start) * elapsed_ratio;
        CcRawPoint r;
        r.pad0 = 0;
        for (int b = 0; b < 7; b++)
            r.pad1[b] = 0;
        r.stamp = k.stamp_start + static_cast<unsigned long long>(elapsed_time);
        r.firing_index = static_cast<unsigned long long>(col);
        if (pt >= 0)
        {
            r.x = k.uncorrected[3 * pt + 0];
            r.y = k.uncorrected[3 * pt + 1];
            r.z = k.uncorrected[3 * pt + 2];
            r.intensity = static_cast<unsigned char>(static_cast<int>(k.xyzi[pt].w * 255)); // uint8_t(i * 255), cpp:148
            r.guid = (static_cast<unsigned long long>(sequence_index) << 48) | (static_cast<unsigned long long>(frame_index) << 32) |
                     static_cast<unsigned long long>(pt);
        }
        else
        {
            // an empty cell is a default KittiPoint: NaN coordinates, original_kitti_index -1 (kitti_loader.hpp:30-43)
            r.x = r.y = r.z = cc_nanf();
            r.intensity = 0;
            r.guid = ~0ull;
        }
        firings[i] = r;
    }
}

#endif
