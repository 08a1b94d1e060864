// TEST INFRASTRUCTURE: declaration-level stand-in for <pcl/point_types.h> (PCL is not installed). The reference's
// kitti_evaluation.hpp names pcl::PointXYZINormal in the signature of a private helper of the ground-truth label
// generator (PCL's conditional Euclidean clustering), which is outside the hot path and is never compiled here.
#ifndef CC_STANDIN_PCL_POINT_TYPES_H
#define CC_STANDIN_PCL_POINT_TYPES_H
namespace pcl
{
struct PointXYZINormal
{
    float x, y, z, intensity, normal_x, normal_y, normal_z, curvature;
};
} // namespace pcl
#endif
