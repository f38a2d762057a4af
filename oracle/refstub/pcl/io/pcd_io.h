// TEST INFRASTRUCTURE: pcl::PCDWriter::write is reached only at shutdown with poseGraph.shutDownFlag set (src/mapper.cpp:165-170).
#ifndef SSM_REFSTUB_PCL_PCD_IO
#define SSM_REFSTUB_PCL_PCD_IO
#include <pcl/point_types.h>
#include <string>
namespace pcl {
class PCDWriter {
public:
    template <typename CloudT> int write(const std::string&, const CloudT&) { return 0; }
};
}  // namespace pcl
#endif
