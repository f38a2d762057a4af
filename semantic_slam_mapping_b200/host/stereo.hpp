// stereo.hpp -- host-side mirror of the reference's dense stereo entry point
//   void calDisparity_SGBM(const cv::Mat& img_L, const cv::Mat& img_R, cv::Mat& disp)
//   (/root/reference include/stereo.h:15, src/stereo.cpp:11-38)
// Same name, argument order and meaning: rectified 8-bit grey left/right in, CV_16SC1 disparity x16 out
// (invalid = -16), synchronous, no return code; failures throw (the reference surfaces cv::Exception from
// OpenCV's asserts).  The SGBM parameters are the reference's hard-coded ones (stereo.cpp:16-28) unless a
// StereoConfig is installed first.  All arithmetic runs in libssm.so on the GPU; there is no CPU path.
#pragma once

#include "frame.hpp"

namespace ssm_host {

// process-wide context used by the free function, like the function-local cv::StereoSGBM of the reference
struct StereoConfig {
    int num_disparities = 80;   // stereo.cpp:18
    int device = 0;
    int max_width = 2048, max_height = 1024;
    Camera camera;              // parameters.txt:37-41,50-54,63 (used by disparityToDepth)
};
void setStereoConfig(const StereoConfig& cfg);   // optional; call before the first calDisparity_SGBM
ssm_ctx* stereoContext();                        // the lazily created context (owned by the library)
void releaseStereoContext();

void calDisparity_SGBM(const ImageU8& img_L, const ImageU8& img_R, ImageS16& disp);

// FrameReader::next glue (src/rgbdframe.cpp:85-116): disparity -> depth in camera.scale units
void disparityToDepth(const ImageS16& disp, ImageU16& depth);

}  // namespace ssm_host

#ifdef SSM_WITH_OPENCV
#include <opencv2/core/core.hpp>
// drop-in overload with the reference's exact signature
void calDisparity_SGBM(const cv::Mat& img_L, const cv::Mat& img_R, cv::Mat& disp);
#endif
