// TEST INFRASTRUCTURE: functional stand-in for sensor_msgs/PointCloud2 (+ PointField), enough for the reference's
// ros_utils.cpp to build and fill a message: same member names, same packed field layout rules as the ROS type
// (offsets accumulate without padding, point_step = sum of field sizes, row-major data of height x width points).
#ifndef CC_STANDIN_SENSOR_MSGS_POINTCLOUD2_H
#define CC_STANDIN_SENSOR_MSGS_POINTCLOUD2_H
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <ros/time.h>

namespace sensor_msgs
{
struct PointField
{
    enum
    {
        INT8 = 1,
        UINT8 = 2,
        INT16 = 3,
        UINT16 = 4,
        INT32 = 5,
        UINT32 = 6,
        FLOAT32 = 7,
        FLOAT64 = 8
    };
    std::string name;
    uint32_t offset{0};
    uint8_t datatype{0};
    uint32_t count{0};
};

struct PointCloud2
{
    struct Header
    {
        uint32_t seq{0};
        ros::Time stamp;
        std::string frame_id;
    } header;
    uint32_t height{0}, width{0};
    std::vector<PointField> fields;
    bool is_bigendian{false};
    uint32_t point_step{0}, row_step{0};
    std::vector<uint8_t> data;
    bool is_dense{false};
    typedef std::shared_ptr<PointCloud2> Ptr;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
} // namespace sensor_msgs
#endif
