// TEST INFRASTRUCTURE: functional stand-in for sensor_msgs/point_cloud2_iterator.h: PointCloud2Modifier::
// setPointCloud2Fields (variadic name / count / datatype triples, offsets accumulated without padding, data resized to
// height * width * point_step) and PointCloud2Iterator<T> (operator*, operator+ in units of points).
#ifndef CC_STANDIN_SENSOR_MSGS_POINT_CLOUD2_ITERATOR_H
#define CC_STANDIN_SENSOR_MSGS_POINT_CLOUD2_ITERATOR_H
#include <cstdarg>
#include <stdexcept>

#include <sensor_msgs/PointCloud2.h>

namespace sensor_msgs
{
inline int sizeOfPointField(int datatype)
{
    switch (datatype)
    {
        case PointField::INT8:
        case PointField::UINT8:
            return 1;
        case PointField::INT16:
        case PointField::UINT16:
            return 2;
        case PointField::INT32:
        case PointField::UINT32:
        case PointField::FLOAT32:
            return 4;
        case PointField::FLOAT64:
            return 8;
    }
    throw std::runtime_error("PointField of unknown type");
}

class PointCloud2Modifier
{
  public:
    PointCloud2Modifier(PointCloud2& msg) : msg_(msg) {}
    void setPointCloud2Fields(int n_fields, ...)
    {
        msg_.fields.clear();
        msg_.fields.reserve(n_fields);
        va_list vl;
        va_start(vl, n_fields);
        uint32_t offset = 0;
        for (int i = 0; i < n_fields; i++)
        {
            PointField f;
            f.name = va_arg(vl, char*);
            f.count = static_cast<uint32_t>(va_arg(vl, int));
            f.datatype = static_cast<uint8_t>(va_arg(vl, int));
            f.offset = offset;
            offset += f.count * sizeOfPointField(f.datatype);
            msg_.fields.push_back(f);
        }
        va_end(vl);
        msg_.point_step = offset;
        msg_.row_step = msg_.width * msg_.point_step;
        msg_.data.resize(static_cast<size_t>(msg_.height) * msg_.row_step);
    }

  private:
    PointCloud2& msg_;
};

template<typename T>
class PointCloud2Iterator
{
  public:
    PointCloud2Iterator(PointCloud2& msg, const std::string& field_name) : step_(msg.point_step)
    {
        for (const PointField& f : msg.fields)
            if (f.name == field_name)
            {
                p_ = msg.data.data() + f.offset;
                return;
            }
        throw std::runtime_error("Field " + field_name + " does not exist");
    }
    T& operator*() const { return *reinterpret_cast<T*>(p_); } // unaligned on purpose, like the ROS type
    PointCloud2Iterator operator+(int i) const
    {
        PointCloud2Iterator r(*this);
        r.p_ += static_cast<ptrdiff_t>(i) * step_;
        return r;
    }
    PointCloud2Iterator& operator++()
    {
        p_ += step_;
        return *this;
    }

  private:
    uint8_t* p_{nullptr};
    uint32_t step_{0};
};
} // namespace sensor_msgs
#endif
