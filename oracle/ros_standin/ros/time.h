// TEST INFRASTRUCTURE: minimal stand-in for ros::Time (ROS is not installed in this image). Only what the reference's
// message-conversion code uses (ros_utils.cpp:17, 74, 245-246): fromNSec and the sec / nsec members.
#ifndef CC_STANDIN_ROS_TIME_H
#define CC_STANDIN_ROS_TIME_H
#include <cstdint>
namespace ros
{
struct Time
{
    uint32_t sec{0}, nsec{0};
    Time() = default;
    Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
    Time& fromNSec(uint64_t t)
    {
        sec = static_cast<uint32_t>(t / 1000000000ull);
        nsec = static_cast<uint32_t>(t % 1000000000ull);
        return *this;
    }
    uint64_t toNSec() const { return static_cast<uint64_t>(sec) * 1000000000ull + nsec; }
};
} // namespace ros
#endif
