// dropin/stereo.cpp -- drop-in replacement for the reference's src/stereo.cpp.
//
// Compiled INSIDE the reference's tree in place of src/stereo.cpp: it includes the reference's own include/stereo.h and defines
// the four functions that header declares, with the identical signatures,
//   calDisparity_SGBM   include/stereo.h:15   (reference body: src/stereo.cpp:11-38)
//   triangulate10D      include/stereo.h:25   (src/stereo.cpp:41-118)
//   correct3DPoints     include/stereo.h:36   (src/stereo.cpp:127-181)
//   setImageROI         include/stereo.h:43   (src/stereo.cpp:183-192)
// plus the two UVDisparity histogram passes as free functions a maintainer calls from UVDisparity::calUDisparity /
// calVDisparity (include/uvdisparity.hpp:88,91; those members live in the middle of a 1000-line file that stays).
// Every call forwards to libssm.so through the C ABI of include/ssm.h; there is no CPU path.  Errors surface as
// cv::Exception-style failures: std::runtime_error carrying ssm_last_error().
//
// Type-checked in this repository against stand-ins for the OpenCV names (tests/test_dropin_signatures.py compiles and links
// it with the reference's headers from /root/reference/include); see INTEGRATION.md section 2.
#include "stereo.h"   // the reference's header

#include <mutex>
#include <stdexcept>
#include <string>

#include "ssm.h"
#include "dropin.hpp"

namespace ssm_dropin {

static std::mutex g_mutex;
static ssm_ctx* g_ctx = nullptr;
static int g_w = 0, g_h = 0;

void check(int rc)
{
    if (rc != SSM_OK) throw std::runtime_error(std::string("libssm: ") + ssm_last_error());
}

// one process-wide context for the stereo entry points, like the function-local cv::StereoSGBM of src/stereo.cpp:13; it is
// re-created when a larger frame arrives.  The SGBM parameter block is the reference's hard-coded one (src/stereo.cpp:16-28),
// which ssm_default_params reproduces.
ssm_ctx* stereo_context(int w, int h)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_ctx && w <= g_w && h <= g_h) return g_ctx;
    if (g_ctx) { ssm_destroy(g_ctx); g_ctx = nullptr; }
    ssm_params p;
    ssm_default_params(&p);
    p.max_width = w > 1241 ? w : 1241;
    p.max_height = h > 376 ? h : 376;
    p.max_batch = 1;
    p.map_capacity = 1024;   // this context never maps
    check(ssm_create(&p, 0, &g_ctx));
    g_w = p.max_width; g_h = p.max_height;
    return g_ctx;
}

void release_stereo_context()
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (g_ctx) { ssm_destroy(g_ctx); g_ctx = nullptr; }
    g_w = g_h = 0;
}

static void need(bool ok, const char* what)
{
    if (!ok) throw std::runtime_error(std::string(what) + ": bad matrix type, size or layout");
}

// UVDisparity::calVDisparity(img_dis, xyz) body (src/uvdisparity.cpp:277-366): fills v_dis_int (CV_32SC1) and v_dis (CV_8UC1)
// and channel 8 of xyz
void calVDisparity(const cv::Mat& img_dis, cv::Mat& xyz, cv::Mat& v_dis_int, cv::Mat& v_dis)
{
    need(img_dis.type() == CV_16SC1 && xyz.rows == img_dis.rows && xyz.cols == img_dis.cols && xyz.type() == CV_MAKETYPE(CV_32F, 10) &&
             xyz.step == (size_t)xyz.cols * 40, "calVDisparity");
    ssm_ctx* c = stereo_context(img_dis.cols, img_dis.rows);
    int v_cols = 0;
    check(ssm_v_disparity(c, (const int16_t*)img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, nullptr, nullptr, nullptr, 0, &v_cols));
    v_dis_int = cv::Mat::zeros(img_dis.rows, v_cols, CV_32SC1);
    v_dis = cv::Mat::zeros(img_dis.rows, v_cols, CV_8UC1);
    check(ssm_v_disparity(c, (const int16_t*)img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, (float*)xyz.data,
                          v_cols ? (int32_t*)v_dis_int.data : nullptr, v_cols ? v_dis.data : nullptr, v_cols, &v_cols));
}

// UVDisparity::calUDisparity(img_dis, xyz, roi_mask, ground_mask) body (src/uvdisparity.cpp:195-274)
void calUDisparity(const cv::Mat& img_dis, cv::Mat& xyz, cv::Mat& roi_mask, cv::Mat& ground_mask, cv::Mat& u_dis_int, cv::Mat& u_dis)
{
    need(img_dis.type() == CV_16SC1 && xyz.rows == img_dis.rows && xyz.cols == img_dis.cols && xyz.type() == CV_MAKETYPE(CV_32F, 10) &&
             xyz.step == (size_t)xyz.cols * 40 && roi_mask.type() == CV_8UC1 && ground_mask.type() == CV_8UC1 &&
             roi_mask.step == (size_t)img_dis.cols && ground_mask.step == (size_t)img_dis.cols,
         "calUDisparity");
    ssm_ctx* c = stereo_context(img_dis.cols, img_dis.rows);
    int u_rows = 0;
    check(ssm_u_disparity(c, (const int16_t*)img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, nullptr, roi_mask.data, ground_mask.data,
                          nullptr, nullptr, 0, &u_rows));
    u_dis_int = cv::Mat::zeros(u_rows, img_dis.cols, CV_32SC1);
    u_dis = cv::Mat::zeros(u_rows, img_dis.cols, CV_8UC1);
    check(ssm_u_disparity(c, (const int16_t*)img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, (float*)xyz.data, roi_mask.data,
                          ground_mask.data, (int32_t*)u_dis_int.data, u_dis.data, u_rows, &u_rows));
}

}  // namespace ssm_dropin

using namespace ssm_dropin;

void calDisparity_SGBM(const cv::Mat& img_L, const cv::Mat& img_R, cv::Mat& disp)
{
    need(img_L.type() == CV_8UC1 && img_R.type() == CV_8UC1 && img_L.rows == img_R.rows && img_L.cols == img_R.cols && img_L.step == img_R.step &&
             !img_L.empty(),
         "calDisparity_SGBM");
    disp.create(img_L.size(), CV_16SC1);
    check(ssm_sgbm(stereo_context(img_L.cols, img_L.rows), img_L.data, img_R.data, img_L.cols, img_L.rows, img_L.step, (int16_t*)disp.data,
                   disp.step));
}

void triangulate10D(const cv::Mat& img, const cv::Mat& disp, cv::Mat& xyz, const double f, const double cx, const double cy, const double b,
                    ROI3D roi)
{
    need(img.type() == CV_8UC1 && disp.type() == CV_16SC1 && img.rows == disp.rows && img.cols == disp.cols && !disp.empty(), "triangulate10D");
    xyz.create(disp.size(), CV_MAKETYPE(CV_32F, 10));   // src/stereo.cpp:48
    check(ssm_triangulate10d(stereo_context(disp.cols, disp.rows), img.data, img.step, (const int16_t*)disp.data, disp.step, disp.cols, disp.rows, f,
                             cx, cy, b, roi.x_max, roi.y_max, roi.z_max, (float*)xyz.data));
}

void correct3DPoints(cv::Mat& xyz, ROI3D& roi_, const double& pitch1, const double& pitch2)
{
    need(xyz.type() == CV_MAKETYPE(CV_32F, 10) && !xyz.empty() && xyz.step == (size_t)xyz.cols * 40, "correct3DPoints");
    check(ssm_correct_3d_points(stereo_context(xyz.cols, xyz.rows), (float*)xyz.data, xyz.cols, xyz.rows, roi_.x_max, roi_.y_max, roi_.z_max, pitch1,
                                pitch2));
}

void setImageROI(cv::Mat& xyz, cv::Mat& roi_mask)
{
    need(xyz.type() == CV_MAKETYPE(CV_32F, 10) && !xyz.empty() && xyz.step == (size_t)xyz.cols * 40, "setImageROI");
    roi_mask.create(xyz.size(), CV_8UC1);
    check(ssm_set_image_roi(stereo_context(xyz.cols, xyz.rows), (const float*)xyz.data, xyz.cols, xyz.rows, roi_mask.data, roi_mask.step));
}
