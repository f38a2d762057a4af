// prelude.hpp -- TEST INFRASTRUCTURE.  Force-included (-include) in front of the REFERENCE's own src/mapper.cpp,
// src/rgbdframe.cpp and src/parameter_reader.cpp, which oracle/Makefile compiles from where they lie under /root/reference
// into oracle/_ref/libref_mapper.so.  Those files drag in headers of libraries that are not installed in this image
// (OpenCV 2.4 C++, PCL 1.7, Eigen 3, boost, g2o, DBoW2, Caffe).  Two mechanisms make them compile untouched:
//   * oracle/cvstub + oracle/refstub sit on the include path in front of the reference's include/ and provide our own
//     stand-ins, written from the public API documentation of those libraries (nothing is copied): containers and the few
//     functions on the executed path are implemented, everything else is declared and aborts if called;
//   * the include guards of two reference headers are pre-defined on the command line (-DCOMMON_HEADERS_H, -DPOSE_GRAPH_H),
//     so their own text is skipped: common_headers.h is an include bundle (this file includes what the executed code needs
//     of it), pose_graph.h needs g2o / DBoW2 / ORB_SLAM2 -- the Mapper only reads PoseGraph::keyframes and ::shutDownFlag,
//     which the stand-in below carries.
// What is the reference's own compiled arithmetic in that library: FrameReader::next's disparity -> depth loop
// (src/rgbdframe.cpp:85-116), RGBDFrame::project2dTo3d (include/rgbdframe.h:63-75), Mapper::generatePointCloud's filters,
// order and colour tagging (src/mapper.cpp:12-94) and Mapper::semantic_motion_fuse (:189-216).  What is the stand-ins'
// written definition (and therefore NOT pinned to the real libraries): cv::dilate with the reference's uninitialised
// structuring element (all ones), pcl::transformPointCloud and pcl::VoxelGrid.
#ifndef SSM_REFSTUB_PRELUDE_HPP
#define SSM_REFSTUB_PRELUDE_HPP

// what common_headers.h would have brought in (include/common_headers.h:8-16)
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <time.h>
#include <unistd.h>
#include <vector>
using namespace std;

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <opencv2/core/core.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <boost/format.hpp>
#include <boost/timer.hpp>
#include <boost/lexical_cast.hpp>

// stand-in for include/pose_graph.h:34-189: what Mapper reads of it (src/mapper.cpp:114-136, :165)
namespace rgbd_tutor {
class RGBDFrame;
class PoseGraph {
public:
    std::vector<std::shared_ptr<RGBDFrame> > keyframes;
    bool shutDownFlag = false;
};
}  // namespace rgbd_tutor

#endif
