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

// ---- dense motion cues: the tracker thread's consumers of the disparity map (src/track.cpp:67-79) ----------------
// include/basicStructure.hpp:15-38
struct ROI3D {
    double x_max = 30000, y_max = -1000, z_max = 30000;
    ROI3D() = default;
    ROI3D(double x, double y, double z) : x_max(x), y_max(y), z_max(z) {}
};
using ImageXYZ10 = Image<float, 10>;   // CV_32FC(10): X, Y, Z, u, v, disparity, intensity, I_u, I_v, motion mark
using ImageS32 = Image<int32_t, 1>;

// include/stereo.h:25, :36, :43 -- same names, argument order and meaning
void triangulate10D(const ImageU8& img, const ImageS16& disp, ImageXYZ10& xyz, const double f, const double cx, const double cy,
                    const double b, ROI3D roi);
void correct3DPoints(ImageXYZ10& xyz, ROI3D& roi_, const double& pitch1, const double& pitch2);
void setImageROI(ImageXYZ10& xyz, ImageU8& roi_mask);

// the two histogram members of UVDisparity (include/uvdisparity.hpp:88,91; src/uvdisparity.cpp:195-366) with the state
// they leave behind (u_dis_int / u_dis_ / v_dis_int / v_dis_)
class UVDisparity {
public:
    void calUDisparity(const ImageS16& img_dis, ImageXYZ10& xyz, ImageU8& roi_mask, ImageU8& ground_mask);
    void calVDisparity(const ImageS16& img_dis, ImageXYZ10& xyz);
    ImageS32 u_dis_int, v_dis_int;
    ImageU8 u_dis_, v_dis_;
};

}  // namespace ssm_host

#ifdef SSM_WITH_OPENCV
#include <opencv2/core/core.hpp>
// drop-in overload with the reference's exact signature
void calDisparity_SGBM(const cv::Mat& img_L, const cv::Mat& img_R, cv::Mat& disp);
#endif
