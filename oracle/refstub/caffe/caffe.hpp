// TEST INFRASTRUCTURE: include/segnet.h declares a Classifier with caffe member types; none of it runs on the mapper path.
#ifndef SSM_REFSTUB_CAFFE
#define SSM_REFSTUB_CAFFE
#include <memory>
namespace caffe {
using std::shared_ptr;
template <typename T> class Net;
}  // namespace caffe
#endif
