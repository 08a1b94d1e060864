// Header-compatible stand-in for the reference's clustering/general.hpp, written for the B200 facade.
// Only what callers of the ContinuousClustering class touch: the float point types carried inside `Point` and the
// label constants. The label VALUES are part of the data contract (they are published in PointCloud2 fields and
// compared by kitti_demo.cpp:211): they index QColor::colorNames() with "transparent" skipped, as in the reference
// (general.hpp:208-357); only the entries the pipeline can emit are named here.
#ifndef CONTINUOUS_CLUSTERING_GENERAL_HPP
#define CONTINUOUS_CLUSTERING_GENERAL_HPP

#include <cmath>

namespace continuous_clustering
{

struct Point2D
{
    Point2D() = default;
    Point2D(float x_, float y_) : x(x_), y(y_) {}
    float x{0.f}, y{0.f};
    float lengthSquared() const { return x * x + y * y; }
    float length() const { return std::sqrt(lengthSquared()); }
};

struct Point3D
{
    Point3D() = default;
    Point3D(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float x{0.f}, y{0.f}, z{0.f};
    float lengthSquared() const { return x * x + y * y + z * z; }
    float length() const { return std::sqrt(lengthSquared()); }
    Point2D xy() const { return {x, y}; }
};

inline Point3D operator-(const Point3D& a, const Point3D& b)
{
    return {a.x - b.x, a.y - b.y, a.z - b.z};
}
inline Point3D operator+(const Point3D& a, const Point3D& b)
{
    return {a.x + b.x, a.y + b.y, a.z + b.z};
}

enum PointCloudColors
{
    BLACK = 7,
    BURLYWOOD = 12,
    CYAN = 20,
    DARKRED = 32,
    GRAY = 53,
    GREEN = 54,
    LIGHTGRAY = 71,
    MAGENTA = 85,
    ORANGE = 105,
    RED = 119,
    VIOLET = 141,
    WHITE = 143,
    YELLOW = 145,
    YELLOWGREEN = 146
};

} // namespace continuous_clustering
#endif
