// TEST INFRASTRUCTURE: pcl::transformPointCloud(in, out, Eigen::Matrix4d) -- the WRITTEN definition of SURVEY App. B-1
// (PCL 1.7 evaluates the affine map in double, left to right, and rounds to float; colour copied).  Not the real library.
#ifndef SSM_REFSTUB_PCL_TRANSFORMS
#define SSM_REFSTUB_PCL_TRANSFORMS
#include <Eigen/Core>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT> inline void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4d& T)
{
    out.points.resize(in.points.size());
    out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
    for (size_t i = 0; i < in.points.size(); ++i) {
        const PointT& p = in.points[i];
        PointT q = p;
        const double x = p.x, y = p.y, z = p.z;
        q.x = (float)(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
        q.y = (float)(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
        q.z = (float)(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
        out.points[i] = q;
    }
}
}  // namespace pcl
#endif
