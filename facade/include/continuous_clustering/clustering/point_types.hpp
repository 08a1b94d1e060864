// Header-compatible stand-in for the reference's clustering/point_types.hpp (point_types.hpp:10-28): the input
// record of addFiring. Layout-identical to cc_raw_point_t (include/cc_b200.h), checked by static_assert in the facade.
#ifndef CONTINUOUS_CLUSTERING_POINT_TYPES_HPP
#define CONTINUOUS_CLUSTERING_POINT_TYPES_HPP

#include <cstdint>
#include <memory>
#include <vector>

namespace continuous_clustering
{

struct RawPoint
{
    float x{}, y{}, z{};
    uint64_t firing_index{};
    uint8_t intensity{};
    uint64_t stamp{};
    uint64_t globally_unique_point_index{};
};

struct RawPoints
{
    uint64_t stamp;
    std::vector<RawPoint> points;
    typedef std::shared_ptr<RawPoints> Ptr;
    typedef std::shared_ptr<RawPoints const> ConstPtr;
};

} // namespace continuous_clustering
#endif
