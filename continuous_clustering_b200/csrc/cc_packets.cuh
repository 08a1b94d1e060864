// cc_packets.cuh -- SURVEY 8f-3: the wire-format step before addFiring on the device. A batch of Ouster lidar UDP packets
// becomes RawPoint firings resident in HBM (ready for cc_submit_firings_device), the way OusterInput::onRawDataArrived
// does it one measurement block at a time (include/continuous_clustering/ros/ouster_input.hpp:105-181):
//   k_ouster_index    which measurement blocks ("columns") are valid (status bit 0, :120-124) and which firing each one
//                     becomes: invalid blocks are dropped, every valid one is one firing (:171-174)
//   k_ouster_decode   per pixel: range / signal fields cut out of the packet with the profile's offsets and masks
//                     (ouster::sensor::packet_format::col_field), x, y, z = range * direction + offset from the sensor's
//                     lookup table at the block's measurement id (ouster::cartesianT, :133-136), NaN for range 0, intensity
//                     scaled from 0..1000 to 0..255 (:150-151), firing index and receive stamp (:160-162)
// The packet layout is DATA (cc_ouster_format_t): the host binding fills it from the SDK's packet_format, or
// cc_ouster_format_legacy for the LEGACY profile of the reference's sensors (calibrations/touareg_os32_*.json).
// Third-party code this restates: ouster-ros / ouster_client (dependencies.repos:10-13, branch master, un-vendored;
// the calibration files were written by ouster_client 0.7.1) -- absent from the image: PARITY UNPINNED for this row.
// HBM bound: 12 B read per pixel (+ 24 B of lookup table, L2 resident), 48 B written.
#ifndef CC_PACKETS_CUH
#define CC_PACKETS_CUH

#include "cc_kernels.cuh"

struct CcOusterFormat // == cc_ouster_format_t
{
    int columns_per_packet, pixels_per_column, columns_per_frame;
    int packet_header_size, col_header_size, col_footer_size, pixel_bytes;
    int col_measurement_id_offset; // u16, from the start of the measurement block
    int col_status_offset;         // from the start of the measurement block
    int col_status_bytes;          // 2 or 4
    int range_offset, range_bytes;
    unsigned int range_mask;
    int range_shift;
    int signal_offset, signal_bytes;
    unsigned int signal_mask;
    int signal_shift;
    int offset_from_direction_table; // 1 = ouster_input.hpp:134 [sic]: the offset block is cut out of the DIRECTION table
};

struct CcOusterPtrs
{
    const unsigned char* packets; // [n_packets][packet_size]
    int n_packets, packet_size;
    const unsigned long long* receive_stamp; // [n_packets] ros::Time::now() of the packet (:111)
    const float* direction;                  // [W * H][3] column-in-frame major, row minor (:84-95)
    const float* offset;
    int* firing_of_column; // [n_packets * columns_per_packet] firing the block becomes, -1 = dropped
    int* n_firings;
    unsigned long long first_firing_index;
    CcRawPoint* firings; // [n_firings][H]
    unsigned long long* firing_stamp; // [n_firings] RawPoints::stamp = min + (max - min) / 2 (sensor_input.hpp:31)
};

CC_DEV unsigned int cc_read_le(const unsigned char* p, int bytes)
{
    unsigned int v = 0;
    for (int b = 0; b < bytes; b++)
        v |= static_cast<unsigned int>(p[b]) << (8 * b);
    return v;
}

// one CTA: valid flag of every measurement block, exclusive prefix sum = firing number
__global__ void k_ouster_index(CcOusterFormat f, CcOusterPtrs o)
{
    CC_PDL_ENTER();
    __shared__ int sh[32];
    const int T = blockDim.x, t = threadIdx.x;
    const int ncol = o.n_packets * f.columns_per_packet;
    const int per = (ncol + T - 1) / T;
    const int col_size = f.col_header_size + f.pixels_per_column * f.pixel_bytes + f.col_footer_size;
    const int a = t * per, b = a + per < ncol ? a + per : ncol;
    int cnt = 0;
    for (int c = a; c < b; c++)
    {
        const unsigned char* col = o.packets + static_cast<size_t>(c / f.columns_per_packet) * o.packet_size + f.packet_header_size +
                                   (c % f.columns_per_packet) * col_size;
        cnt += static_cast<int>(cc_read_le(col + f.col_status_offset, f.col_status_bytes) & 1u);
    }
    int run = cc_block_exclusive_scan(sh, cnt, 0, CcOpAddI32());
    for (int c = a; c < b; c++)
    {
        const unsigned char* col = o.packets + static_cast<size_t>(c / f.columns_per_packet) * o.packet_size + f.packet_header_size +
                                   (c % f.columns_per_packet) * col_size;
        const bool valid = (cc_read_le(col + f.col_status_offset, f.col_status_bytes) & 1u) != 0;
        o.firing_of_column[c] = valid ? run : -1;
        if (valid)
        {
            o.firing_stamp[run] = o.receive_stamp[c / f.columns_per_packet]; // every point carries it: min == max
            run++;
        }
    }
    if (t == T - 1)
        *o.n_firings = run;
}

__global__ void k_ouster_decode(CcOusterFormat f, CcOusterPtrs o)
{
    CC_PDL_ENTER();
    const int H = f.pixels_per_column;
    const int total = o.n_packets * f.columns_per_packet * H;
    const int col_size = f.col_header_size + H * f.pixel_bytes + f.col_footer_size;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const int c = i / H, ring = i - c * H;
        const int firing = o.firing_of_column[c];
        if (firing < 0)
            continue;
        const int packet = c / f.columns_per_packet;
        const unsigned char* col = o.packets + static_cast<size_t>(packet) * o.packet_size + f.packet_header_size +
                                   (c % f.columns_per_packet) * col_size;
        const unsigned int m_id = cc_read_le(col + f.col_measurement_id_offset, 2);
        const unsigned char* px = col + f.col_header_size + ring * f.pixel_bytes;
        unsigned int range = cc_read_le(px + f.range_offset, f.range_bytes);
        if (f.range_mask)
            range &= f.range_mask;
        range = f.range_shift > 0 ? range >> f.range_shift : range << (-f.range_shift);
        unsigned int signal = cc_read_le(px + f.signal_offset, f.signal_bytes);
        if (f.signal_mask)
            signal &= f.signal_mask;
        signal = f.signal_shift > 0 ? signal >> f.signal_shift : signal << (-f.signal_shift);
        CcRawPoint r;
        r.pad0 = 0;
        for (int b = 0; b < 7; b++)
            r.pad1[b] = 0;
        if (range > 0 && m_id < static_cast<unsigned int>(f.columns_per_frame))
        {
            const size_t l = (static_cast<size_t>(m_id) * H + ring) * 3;
            const float* ofs = f.offset_from_direction_table ? o.direction : o.offset;
            const float rr = static_cast<float>(range);
            r.x = rr * o.direction[l + 0] + ofs[l + 0]; // (no contraction: the library is built with -fmad=false)
            r.y = rr * o.direction[l + 1] + ofs[l + 1];
            r.z = rr * o.direction[l + 2] + ofs[l + 2];
            const float s = ccm::div_rn(static_cast<float>(signal), 1000.f);
            r.intensity = static_cast<unsigned char>(static_cast<int>((s < 1.f ? s : 1.f) * 255));
        }
        else
        {
            r.x = r.y = r.z = cc_nanf();
            r.intensity = 0;
        }
        r.firing_index = o.first_firing_index + static_cast<unsigned long long>(firing);
        r.stamp = o.receive_stamp[packet];
        r.guid = 0; // not set by the sensor inputs (RawPoint default)
        o.firings[static_cast<size_t>(firing) * H + ring] = r;
    }
}

#endif
