// ref_uv_wrap.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's own
// UVDisparity::calVDisparity / calUDisparity (/root/reference/src/uvdisparity.cpp:277-366, 195-274), which oracle/Makefile
// compiles from where the file lies (against oracle/cvstub) into oracle/_ref/libref_stereo.so.  The two member functions
// and the maps they leave behind are private; this translation unit reads the reference's header with `private` opened
// (the reference's own .cpp is compiled untouched).  Used by tests/ to pin the C oracle's restatement of the U/V-disparity
// histograms, and by tests/golden/make_golden_stereo.py to produce fixtures.
#define private public
#include "uvdisparity.hpp"   // the reference's header (include/uvdisparity.hpp)
#undef private

extern "C" {

// calVDisparity(img_dis, xyz): xyz (h*w*10 floats) gets channel 8; v_int / v_u8 receive v_dis_int / v_dis_ (h rows of
// *v_cols entries, densely packed).  Returns -1 when the map is wider than cap_cols.
int ref_calVDisparity(const short* disp, int w, int h, float* xyz, int* v_int, unsigned char* v_u8, int cap_cols)
{
    UVDisparity uv;
    cv::Mat D(h, w, CV_16SC1, (void*)disp), X(h, w, CV_MAKETYPE(CV_32F, 10), xyz);
    uv.calVDisparity(D, X);
    const int vc = uv.v_dis_int.cols, vr = uv.v_dis_int.rows;
    if (vc > cap_cols) return -1;
    for (int i = 0; i < vr; ++i) {
        if (vc > 0) std::memcpy(v_int + (size_t)i * vc, uv.v_dis_int.ptr<int>(i), sizeof(int) * vc);
        if (vc > 0) std::memcpy(v_u8 + (size_t)i * vc, uv.v_dis_.ptr<uchar>(i), vc);
    }
    return vc;
}

// calUDisparity(img_dis, xyz, roi_mask, ground_mask): xyz gets channel 7; u_int / u_u8 receive u_dis_int / u_dis_
// (*u_rows rows of w entries).  Returns -1 when the map is taller than cap_rows.
int ref_calUDisparity(const short* disp, int w, int h, float* xyz, const unsigned char* roi_mask, const unsigned char* ground_mask,
                      int* u_int, unsigned char* u_u8, int cap_rows)
{
    UVDisparity uv;
    cv::Mat D(h, w, CV_16SC1, (void*)disp), X(h, w, CV_MAKETYPE(CV_32F, 10), xyz);
    cv::Mat R(h, w, CV_8UC1, (void*)roi_mask), G(h, w, CV_8UC1, (void*)ground_mask);
    uv.calUDisparity(D, X, R, G);
    const int ur = uv.u_dis_int.rows, uc = uv.u_dis_int.cols;
    if (ur > cap_rows) return -1;
    for (int i = 0; i < ur; ++i) {
        std::memcpy(u_int + (size_t)i * uc, uv.u_dis_int.ptr<int>(i), sizeof(int) * uc);
        std::memcpy(u_u8 + (size_t)i * uc, uv.u_dis_.ptr<uchar>(i), uc);
    }
    return ur;
}

}  // extern "C"
