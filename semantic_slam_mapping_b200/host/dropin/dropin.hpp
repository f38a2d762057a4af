// dropin.hpp -- what the drop-in replacements of the reference's src/stereo.cpp and src/mapper.cpp share.
#pragma once
#include "ssm.h"

namespace cv { class Mat; }

namespace ssm_dropin {
void check(int rc);                              // throws std::runtime_error with ssm_last_error() on rc != SSM_OK
ssm_ctx* stereo_context(int w, int h);           // the stereo entry points' process-wide context (80 disparities, src/stereo.cpp:16-28)
void release_stereo_context();
// bodies for UVDisparity::calVDisparity / calUDisparity (include/uvdisparity.hpp:88,91): the maintainer's members forward here with
// their own v_dis_int / v_dis_ / u_dis_int / u_dis_ matrices
void calVDisparity(const cv::Mat& img_dis, cv::Mat& xyz, cv::Mat& v_dis_int, cv::Mat& v_dis);
void calUDisparity(const cv::Mat& img_dis, cv::Mat& xyz, cv::Mat& roi_mask, cv::Mat& ground_mask, cv::Mat& u_dis_int, cv::Mat& u_dis);
}  // namespace ssm_dropin
