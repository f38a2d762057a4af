// ref_stereo_wrap.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's own functions in
// /root/reference/src/stereo.cpp, which oracle/Makefile compiles from where it lies (against oracle/cvstub, a stand-in
// for the OpenCV container types) into oracle/_ref/libref_stereo.so.  Used by tests/ to pin the C oracle's restatement
// of triangulate10D / correct3DPoints / setImageROI, and by tests/golden/make_golden_stereo.py to produce fixtures.
#include "stereo.h"   // the reference's header (include/stereo.h)

extern "C" {

// the oracle's SGBM restatement can be installed behind calDisparity_SGBM (src/stereo.cpp:11-38)
typedef void (*ref_sgbm_fn)(const unsigned char* l, const unsigned char* r, int w, int h, int num_disp, int block, int p1, int p2,
                            int d12, int cap, int uniq, int spw, int spr, short* disp);
static ref_sgbm_fn g_sgbm = 0;
static void sgbm_hook(const cv::StereoSGBM& s, const cv::Mat& l, const cv::Mat& r, cv::Mat& disp)
{
    disp.create(l.rows, l.cols, CV_16SC1);
    g_sgbm(l.data, r.data, l.cols, l.rows, s.numberOfDisparities, s.SADWindowSize, s.P1, s.P2, s.disp12MaxDiff, s.preFilterCap,
           s.uniquenessRatio, s.speckleWindowSize, s.speckleRange, disp.ptr<short>(0));
}
void ref_set_sgbm(ref_sgbm_fn f)
{
    g_sgbm = f;
    cv::StereoSGBM::hook() = f ? sgbm_hook : 0;
}
// calDisparity_SGBM(img_L, img_R, disp): densely packed w x h buffers
void ref_calDisparity_SGBM(const unsigned char* l, const unsigned char* r, int w, int h, short* disp)
{
    cv::Mat L(h, w, CV_8UC1, (void*)l), R(h, w, CV_8UC1, (void*)r), D;
    calDisparity_SGBM(L, R, D);
    for (int i = 0; i < h; ++i) std::memcpy(disp + (size_t)i * w, D.ptr<short>(i), sizeof(short) * w);
}

// triangulate10D(img, disp, xyz, f, cx, cy, b, roi): xyz receives h*w*10 floats
void ref_triangulate10D(const unsigned char* img, const short* disp, int w, int h, double f, double cx, double cy, double b,
                        double roi_x, double roi_y, double roi_z, float* xyz)
{
    cv::Mat I(h, w, CV_8UC1, (void*)img), D(h, w, CV_16SC1, (void*)disp), X;
    triangulate10D(I, D, X, f, cx, cy, b, ROI3D(roi_x, roi_y, roi_z));
    for (int i = 0; i < h; ++i) std::memcpy(xyz + (size_t)i * w * 10, X.ptr<float>(i), sizeof(float) * w * 10);
}

// correct3DPoints(xyz, roi, pitch1, pitch2): in place on h*w*10 floats
void ref_correct3DPoints(float* xyz, int w, int h, double roi_x, double roi_y, double roi_z, double pitch1, double pitch2)
{
    cv::Mat X(h, w, CV_MAKETYPE(CV_32F, 10), xyz);
    ROI3D roi(roi_x, roi_y, roi_z);
    correct3DPoints(X, roi, pitch1, pitch2);
}

// setImageROI(xyz, roi_mask): roi_mask receives h*w bytes
void ref_setImageROI(const float* xyz, int w, int h, unsigned char* roi_mask)
{
    cv::Mat X(h, w, CV_MAKETYPE(CV_32F, 10), (void*)xyz), M;
    setImageROI(X, M);
    for (int i = 0; i < h; ++i) std::memcpy(roi_mask + (size_t)i * w, M.ptr<uchar>(i), w);
}

}  // extern "C"
