// TEST INFRASTRUCTURE: pcl::PointXYZRGBA and pcl::PointCloud as the reference's mapper uses them.
#ifndef SSM_REFSTUB_PCL_POINT_TYPES
#define SSM_REFSTUB_PCL_POINT_TYPES
#include <boost/timer.hpp>   // boost::shared_ptr / make_shared stand-ins
#include <cstdint>
#include <vector>
namespace pcl {
struct PointXYZRGBA {
    float x, y, z, pad_;
    union {
        struct { uint8_t b, g, r, a; };
        uint32_t rgba;
    };
    uint32_t pad2_[3];
    PointXYZRGBA() : x(0), y(0), z(0), pad_(1.0f), rgba(0) { pad2_[0] = pad2_[1] = pad2_[2] = 0; }   // PCL 1.7: r = g = b = a = 0
};
template <typename PointT> class PointCloud {
public:
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    void clear() { points.clear(); width = height = 0; }
    PointCloud& operator+=(const PointCloud& o)
    {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (uint32_t)points.size(); height = 1;
        if (!o.is_dense) is_dense = false;
        return *this;
    }
    void swap(PointCloud& o)
    {
        points.swap(o.points);
        std::swap(width, o.width); std::swap(height, o.height); std::swap(is_dense, o.is_dense);
    }
};
}  // namespace pcl
#endif
