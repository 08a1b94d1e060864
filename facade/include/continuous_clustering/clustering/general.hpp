// Header-compatible stand-in for the reference's clustering/general.hpp, written for the B200 facade.
// Same names and semantics as general.hpp:7-357 so that callers of the ContinuousClustering class (ros_utils.cpp,
// kitti_demo.cpp, the ROS node) compile unchanged: the float vector types carried inside `Point`, their helpers,
// and the colour table whose VALUES are a data contract (they are published in PointCloud2 fields, e.g.
// ros_utils.cpp:285 `is_ignored ? BLUE : ORANGE`, and compared by kitti_demo.cpp:211). The table is
// QColor::colorNames() with "transparent" skipped, i.e. the SVG colour keywords in alphabetical order.
#ifndef CONTINUOUS_CLUSTERING_GENERAL_HPP
#define CONTINUOUS_CLUSTERING_GENERAL_HPP

#include <cmath>

namespace continuous_clustering
{

struct Point2D
{
    float x{0.f}, y{0.f};
    Point2D() = default;
    Point2D(float x_, float y_) : x(x_), y(y_) {}

    float lengthSquared() const { return x * x + y * y; } // evaluation order of general.hpp:20-23
    float length() const { return std::sqrt(lengthSquared()); }
    // component-wise minimum / maximum with another point (general.hpp:25-43)
    Point2D min(Point2D& o) const { return {x < o.x ? x : o.x, y < o.y ? y : o.y}; }
    Point2D max(Point2D& o) const { return {x > o.x ? x : o.x, y > o.y ? y : o.y}; }
    Point2D normed() const
    {
        const float inv = 1.f / length();
        return {x * inv, y * inv};
    }
};

struct Point3D
{
    float x{0.f}, y{0.f}, z{0.f};
    Point3D() = default;
    Point3D(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}

    float lengthSquared() const { return x * x + y * y + z * z; } // evaluation order of general.hpp:96-99
    float length() const { return std::sqrt(lengthSquared()); }
    Point2D xy() const { return {x, y}; }
    Point2D xz() const { return {x, z}; }
    Point2D yz() const { return {y, z}; }
    float lengthXY() const { return xy().length(); }
    float lengthXZ() const { return xz().length(); }
    float lengthYZ() const { return yz().length(); }
    Point3D min(Point3D& o) const { return {x < o.x ? x : o.x, y < o.y ? y : o.y, z < o.z ? z : o.z}; }
    Point3D max(Point3D& o) const { return {x > o.x ? x : o.x, y > o.y ? y : o.y, z > o.z ? z : o.z}; }
    Point3D normed() const
    {
        const float inv = 1.f / length();
        return {x * inv, y * inv, z * inv};
    }
};

// vector arithmetic (general.hpp:52-80, 166-194): difference, sum, dot product, scaling
inline Point2D operator-(const Point2D& a, const Point2D& b) { return {a.x - b.x, a.y - b.y}; }
inline Point2D operator+(const Point2D& a, const Point2D& b) { return {a.x + b.x, a.y + b.y}; }
inline float operator*(const Point2D& a, const Point2D& b) { return a.x * b.x + a.y * b.y; }
inline Point2D operator*(const Point2D& a, float s) { return {a.x * s, a.y * s}; }
inline Point2D operator*(float s, const Point2D& a) { return a * s; }
inline Point2D operator/(const Point2D& a, float s) { return {a.x / s, a.y / s}; }
inline Point3D operator-(const Point3D& a, const Point3D& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Point3D operator+(const Point3D& a, const Point3D& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float operator*(const Point3D& a, const Point3D& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Point3D operator*(const Point3D& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Point3D operator*(float s, const Point3D& a) { return a * s; }
inline Point3D operator/(const Point3D& a, float s) { return {a.x / s, a.y / s, a.z / s}; }

// general.hpp:196-204
template<typename PointXD>
inline PointXD get_line_direction(PointXD line_start, PointXD line_end)
{
    return (line_end - line_start).normed();
}
template<typename PointXD>
inline float distance_point_from_line(PointXD point, PointXD line_start, PointXD line_direction)
{
    const PointXD rel = point - line_start;
    return (rel - (rel * line_direction) * line_direction).length(); // rejection of rel from the line direction
}

// general.hpp:208-357 -- values are the data contract
enum PointCloudColors
{
    ALICEBLUE = 0, ANTIQUEWHITE = 1, AQUA = 2, AQUAMARINE = 3, AZURE = 4, BEIGE = 5, BISQUE = 6, BLACK = 7,
    BLANCHEDALMOND = 8, BLUE = 9, BLUEVIOLET = 10, BROWN = 11, BURLYWOOD = 12, CADETBLUE = 13, CHARTREUSE = 14,
    CHOCOLATE = 15, CORAL = 16, CORNFLOWERBLUE = 17, CORNSILK = 18, CRIMSON = 19, CYAN = 20, DARKBLUE = 21,
    DARKCYAN = 22, DARKGOLDENROD = 23, DARKGRAY = 24, DARKGREEN = 25, DARKGREY = 26, DARKKHAKI = 27,
    DARKMAGENTA = 28, DARKOLIVEGREEN = 29, DARKORANGE = 30, DARKORCHID = 31, DARKRED = 32, DARKSALMON = 33,
    DARKSEAGREEN = 34, DARKSLATEBLUE = 35, DARKSLATEGRAY = 36, DARKSLATEGREY = 37, DARKTURQUOISE = 38,
    DARKVIOLET = 39, DEEPPINK = 40, DEEPSKYBLUE = 41, DIMGRAY = 42, DIMGREY = 43, DODGERBLUE = 44, FIREBRICK = 45,
    FLORALWHITE = 46, FORESTGREEN = 47, FUCHSIA = 48, GAINSBORO = 49, GHOSTWHITE = 50, GOLD = 51, GOLDENROD = 52,
    GRAY = 53, GREEN = 54, GREENYELLOW = 55, GREY = 56, HONEYDEW = 57, HOTPINK = 58, INDIANRED = 59, INDIGO = 60,
    IVORY = 61, KHAKI = 62, LAVENDER = 63, LAVENDERBLUSH = 64, LAWNGREEN = 65, LEMONCHIFFON = 66, LIGHTBLUE = 67,
    LIGHTCORAL = 68, LIGHTCYAN = 69, LIGHTGOLDENRODYELLOW = 70, LIGHTGRAY = 71, LIGHTGREEN = 72, LIGHTGREY = 73,
    LIGHTPINK = 74, LIGHTSALMON = 75, LIGHTSEAGREEN = 76, LIGHTSKYBLUE = 77, LIGHTSLATEGRAY = 78,
    LIGHTSLATEGREY = 79, LIGHTSTEELBLUE = 80, LIGHTYELLOW = 81, LIME = 82, LIMEGREEN = 83, LINEN = 84, MAGENTA = 85,
    MAROON = 86, MEDIUMAQUAMARINE = 87, MEDIUMBLUE = 88, MEDIUMORCHID = 89, MEDIUMPURPLE = 90, MEDIUMSEAGREEN = 91,
    MEDIUMSLATEBLUE = 92, MEDIUMSPRINGGREEN = 93, MEDIUMTURQUOISE = 94, MEDIUMVIOLETRED = 95, MIDNIGHTBLUE = 96,
    MINTCREAM = 97, MISTYROSE = 98, MOCCASIN = 99, NAVAJOWHITE = 100, NAVY = 101, OLDLACE = 102, OLIVE = 103,
    OLIVEDRAB = 104, ORANGE = 105, ORANGERED = 106, ORCHID = 107, PALEGOLDENROD = 108, PALEGREEN = 109,
    PALETURQUOISE = 110, PALEVIOLETRED = 111, PAPAYAWHIP = 112, PEACHPUFF = 113, PERU = 114, PINK = 115, PLUM = 116,
    POWDERBLUE = 117, PURPLE = 118, RED = 119, ROSYBROWN = 120, ROYALBLUE = 121, SADDLEBROWN = 122, SALMON = 123,
    SANDYBROWN = 124, SEAGREEN = 125, SEASHELL = 126, SIENNA = 127, SILVER = 128, SKYBLUE = 129, SLATEBLUE = 130,
    SLATEGRAY = 131, SLATEGREY = 132, SNOW = 133, SPRINGGREEN = 134, STEELBLUE = 135, TAN = 136, TEAL = 137,
    THISTLE = 138, TOMATO = 139, TURQUOISE = 140, VIOLET = 141, WHEAT = 142, WHITE = 143, WHITESMOKE = 144,
    YELLOW = 145, YELLOWGREEN = 146
};

} // namespace continuous_clustering
#endif
