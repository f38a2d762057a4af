// TEST INFRASTRUCTURE: the viewer window is out of scope; showCloud keeps the last cloud it was handed (src/mapper.cpp:159).
#ifndef SSM_REFSTUB_PCL_CLOUD_VIEWER
#define SSM_REFSTUB_PCL_CLOUD_VIEWER
#include <pcl/point_types.h>
#include <string>
namespace pcl { namespace visualization {
class CloudViewer {
public:
    explicit CloudViewer(const std::string&) {}
    template <typename CloudPtr> void showCloud(const CloudPtr&) {}
    bool wasStopped() const { return false; }
};
} }  // namespace pcl::visualization
#endif
