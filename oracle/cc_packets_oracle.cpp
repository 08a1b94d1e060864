// cc_packets_oracle.cpp -- TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it).
// CPU restatement of the reference's Ouster sensor input, one packet at a time, the way the ROS node runs it:
//   OusterInput::onRawDataArrived            include/continuous_clustering/ros/ouster_input.hpp:105-181
//   SensorInput::reset / publishCurrent... / prepareNewFiring / keepTrackOfMinAndMaxStamp
//                                            include/continuous_clustering/ros/sensor_input.hpp:15-56
// PARITY UNPINNED: the calls it makes into the ouster SDK (packet_format::nth_col / col_measurement_id / col_status /
// col_field, cartesianT) belong to ouster-ros / ouster_client (.github/workflows/dependencies.repos:10-13, branch master,
// un-vendored, not in the image; the reference's calibration files name ouster_client 0.7.1). They are restated here from
// the SDK's published behaviour: little-endian fields at fixed offsets of the measurement block, value = (word & mask)
// >> shift, cartesianT = range * direction + offset in float with NaN for range 0. The reference has no tests or golden
// packets for this path, so this restatement is what the device decoder is compared with.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace
{
struct RawPoint // continuous_clustering::RawPoint, point_types.hpp:10-19
{
    float x{};
    float y{};
    float z{};
    uint64_t firing_index{};
    uint8_t intensity{};
    uint64_t stamp{};
    uint64_t globally_unique_point_index{};
};
static_assert(sizeof(RawPoint) == 48, "RawPoint layout");

struct Format // cc_ouster_format_t (include/cc_b200.h)
{
    int32_t columns_per_packet, pixels_per_column, columns_per_frame;
    int32_t packet_header_size, col_header_size, col_footer_size, pixel_bytes;
    int32_t col_measurement_id_offset, col_status_offset, col_status_bytes;
    int32_t range_offset, range_bytes;
    uint32_t range_mask;
    int32_t range_shift;
    int32_t signal_offset, signal_bytes;
    uint32_t signal_mask;
    int32_t signal_shift;
    int32_t offset_from_direction_table;
};

uint32_t read_le(const uint8_t* p, int bytes)
{
    uint32_t v = 0;
    std::memcpy(&v, p, static_cast<size_t>(bytes)); // x86: little endian
    return v;
}
uint32_t field(const uint8_t* px, int offset, int bytes, uint32_t mask, int shift)
{
    uint32_t v = read_le(px + offset, bytes);
    if (mask)
        v &= mask;
    if (shift > 0)
        v >>= shift;
    else if (shift < 0)
        v <<= -shift;
    return v;
}

struct OusterInputOracle
{
    Format f{};
    const float* lut_direction{nullptr}; // [W * H][3], column major as reordered by ouster_input.hpp:84-95
    const float* lut_offset{nullptr};
    // SensorInput members (sensor_input.hpp:58-63)
    std::vector<RawPoint> current_firing;
    uint64_t firing_index{0};
    uint64_t min_stamp{0}, max_stamp{0};
    bool interrupt_message{false};
    // what the callback received
    std::vector<RawPoint>* out_points{nullptr};
    std::vector<uint64_t>* out_stamps{nullptr};

    void prepareNewFiring() // sensor_input.hpp:37-45
    {
        min_stamp = std::numeric_limits<uint64_t>::max();
        max_stamp = 0;
        current_firing.assign(static_cast<size_t>(f.pixels_per_column), RawPoint());
    }
    void publishCurrentFiringAndPrepareNewFiring() // sensor_input.hpp:27-36
    {
        out_stamps->push_back(min_stamp + (max_stamp - min_stamp) / 2);
        out_points->insert(out_points->end(), current_firing.begin(), current_firing.end());
        firing_index++;
        prepareNewFiring();
    }
    void reset() // sensor_input.hpp:15-19 + ouster_input.hpp:97-101
    {
        firing_index = 0;
        prepareNewFiring();
        interrupt_message = true;
    }
    void onRawDataArrived(const uint8_t* packet_buf, uint64_t packet_receive_time) // ouster_input.hpp:105-181
    {
        const int H = f.pixels_per_column;
        const int col_size = f.col_header_size + H * f.pixel_bytes + f.col_footer_size;
        for (int icol = 0; icol < f.columns_per_packet; icol++)
        {
            const uint8_t* col_buf = packet_buf + f.packet_header_size + icol * col_size; // pf->nth_col
            const uint16_t m_id = static_cast<uint16_t>(read_le(col_buf + f.col_measurement_id_offset, 2));
            const uint32_t status = read_le(col_buf + f.col_status_offset, f.col_status_bytes);
            const bool valid = (status & 0x01);
            if (!valid)
                continue;
            for (int ring = 0; ring < H; ring++)
            {
                const uint8_t* px = col_buf + f.col_header_size + ring * f.pixel_bytes;
                const uint32_t range = field(px, f.range_offset, f.range_bytes, f.range_mask, f.range_shift);
                const uint32_t intensity = field(px, f.signal_offset, f.signal_bytes, f.signal_mask, f.signal_shift);
                if (range > 0)
                {
                    // cartesianT on the m_id-th block of the lookup tables (:132-135; the offset block is taken from
                    // lut_direction there)
                    const float* dir = lut_direction + (static_cast<size_t>(m_id) * H + ring) * 3;
                    const float* ofs = (f.offset_from_direction_table ? lut_direction : lut_offset) + (static_cast<size_t>(m_id) * H + ring) * 3;
                    const float r = static_cast<float>(range);
                    current_firing[ring].x = r * dir[0] + ofs[0];
                    current_firing[ring].y = r * dir[1] + ofs[1];
                    current_firing[ring].z = r * dir[2] + ofs[2];
                    current_firing[ring].intensity =
                        static_cast<uint8_t>(std::min(1.f, static_cast<float>(intensity) / 1000.f) * 255);
                }
                else
                {
                    current_firing[ring].x = nanf("");
                    current_firing[ring].y = nanf("");
                    current_firing[ring].z = nanf("");
                    current_firing[ring].intensity = 0;
                }
                current_firing[ring].firing_index = firing_index;
                current_firing[ring].stamp = packet_receive_time;
                if (packet_receive_time < min_stamp) // keepTrackOfMinAndMaxStamp, sensor_input.hpp:47-53
                    min_stamp = packet_receive_time;
                if (packet_receive_time > max_stamp)
                    max_stamp = packet_receive_time;
            }
            if (!interrupt_message)
                publishCurrentFiringAndPrepareNewFiring();
            else
            {
                prepareNewFiring();
                break;
            }
        }
        interrupt_message = false;
    }
};
} // namespace

// Runs `n_packets` packets through a freshly reset input (or one continuing at `first_firing_index` without a pending
// interrupt). firings: room for n_packets * columns_per_packet * H records; returns the number of firings published.
ORC_API int orc_ouster_decode(const void* format, const float* lut_direction, const float* lut_offset, int n_packets,
                              const uint8_t* packets, int packet_size, const uint64_t* receive_stamps, int after_reset,
                              uint64_t first_firing_index, void* firings, uint64_t* firing_stamps)
{
    OusterInputOracle in;
    std::memcpy(&in.f, format, sizeof(Format));
    in.lut_direction = lut_direction;
    in.lut_offset = lut_offset;
    std::vector<RawPoint> pts;
    std::vector<uint64_t> stamps;
    in.out_points = &pts;
    in.out_stamps = &stamps;
    in.prepareNewFiring();
    if (after_reset)
        in.reset();
    else
        in.firing_index = first_firing_index;
    for (int p = 0; p < n_packets; p++)
        in.onRawDataArrived(packets + static_cast<size_t>(p) * packet_size, receive_stamps[p]);
    if (!pts.empty())
        std::memcpy(firings, static_cast<const void*>(pts.data()), pts.size() * sizeof(RawPoint));
    for (size_t i = 0; i < stamps.size(); i++)
        firing_stamps[i] = stamps[i];
    return static_cast<int>(stamps.size());
}
