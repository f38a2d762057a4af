// cvstub.hpp -- TEST INFRASTRUCTURE.  A minimal stand-in for the handful of OpenCV 2.4 C++ names that
// /root/reference/src/stereo.cpp touches, so that the reference's OWN source file can be compiled in this image
// (no OpenCV C++ headers or libraries exist here) and executed as the parity anchor of
//   triangulate10D (src/stereo.cpp:41-118), correct3DPoints (:127-181), setImageROI (:183-192).
// Only the container (cv::Mat), cv::minMaxIdx, cv::split, cv::convertScaleAbs and cvRound are provided; the
// arithmetic under test is the reference's.  cv::StereoSGBM is a field-compatible functor whose call operator forwards
// to a hook (the un-vendored OpenCV implementation is not part of the reference tree).
// This is our own code written against the public OpenCV 2.4 API documentation; nothing is copied from OpenCV.
#ifndef SSM_CVSTUB_HPP
#define SSM_CVSTUB_HPP

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

// round half to even, as the SSE2 / lrint implementation of OpenCV's cvRound
inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline int cvFloor(double v) { return (int)std::floor(v); }

namespace cv {

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
inline bool operator==(const Size& a, const Size& b) { return a.width == b.width && a.height == b.height; }
inline bool operator!=(const Size& a, const Size& b) { return !(a == b); }

struct Scalar;
struct Range;
struct Rect;

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;   // bytes per row
    Mat() : rows(0), cols(0), data(0), step(0), type_(0), uninit_(false) {}
    // cv::Mat(rows, cols, type) leaves its pixels uninitialised in OpenCV; the stand-in zero-fills them and remembers that
    // nobody has defined them (cv::dilate below reads the flag: SURVEY App. C-3)
    Mat(int r, int c, int t) : rows(0), cols(0), data(0), step(0), type_(0), uninit_(false) { create(r, c, t); uninit_ = true; }
    Mat(int r, int c, int t, const Scalar& v);      // filled (cvstub_more.hpp)
    Mat(Size s, int t, const Scalar& v);
    // header over caller-owned memory (no copy), as cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int t, void* d, size_t s = 0) : rows(r), cols(c), data((uchar*)d), step(s), type_(t), uninit_(false)
    {
        if (step == 0) step = (size_t)cols * elemSize();
    }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize1() const
    {
        static const size_t sz[8] = {1, 1, 2, 2, 4, 4, 8, 0};
        return sz[depth()];
    }
    size_t elemSize() const { return elemSize1() * channels(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == 0 || rows * cols == 0; }
    void create(int r, int c, int t)
    {
        if (data && r == rows && c == cols && t == type_) return;
        rows = r; cols = c; type_ = t;
        step = (size_t)cols * elemSize();
        // Zeroed slack before and after the pixels.  The reference indexes a little outside its matrices in
        // UVDisparity::calVDisparity / calUDisparity (row -1 of the 8-bit U map, int reads from the 8-bit V map, histogram
        // bin v_cols); with the slack those accesses have the meaning DESIGN.md section 2 gives them (reads see zeros,
        // writes past the last row are dropped) instead of touching foreign memory.
        const size_t lead = step + 64, trail = 4 * step + 64;
        own_.reset(new std::vector<uchar>(lead + step * (size_t)rows + trail, 0));
        data = own_->data() + lead;
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + step * (size_t)i); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + step * (size_t)i); }
    template <typename T> T& at(int i, int j) { return ((T*)(data + (ptrdiff_t)step * i))[j]; }
    template <typename T> const T& at(int i, int j) const { return ((const T*)(data + (ptrdiff_t)step * i))[j]; }
    template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (ptrdiff_t)step * i); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (ptrdiff_t)step * i); }
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); m.uninit_ = false; return m; }   // create() zero-fills
    static Mat zeros(Size s, int t) { return zeros(s.height, s.width, t); }
    bool uninitialised() const { return uninit_; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        m.uninit_ = uninit_;
        for (int i = 0; i < rows; ++i) std::memcpy(m.data + m.step * (size_t)i, data + step * (size_t)i, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& m) const { m = clone(); }
    // declared for compilation only (cvstub_more.hpp): sub-matrix views, masked copy, fill
    Mat operator()(const Range& rows, const Range& cols) const;
    Mat operator()(const Rect& roi) const;
    void copyTo(Mat& m, const Mat& mask) const;
    Mat& operator=(const Scalar& s);
    void push_back(const Mat& m);

private:
    int type_;
    bool uninit_;
    std::shared_ptr<std::vector<uchar> > own_;
};

template <typename T> static inline void minmax_t(const Mat& m, double* mn, double* mx)
{
    double lo = DBL_MAX, hi = -DBL_MAX;
    const int n = m.cols * m.channels();
    for (int i = 0; i < m.rows; ++i) {
        const T* p = m.ptr<T>(i);
        for (int j = 0; j < n; ++j) {
            if (p[j] < lo) lo = p[j];
            if (p[j] > hi) hi = p[j];
        }
    }
    if (mn) *mn = lo;
    if (mx) *mx = hi;
}
// global minimum / maximum of a single-channel array (index outputs unsupported: the reference passes 0)
inline void minMaxIdx(const Mat& src, double* minVal, double* maxVal = 0, int* = 0, int* = 0)
{
    switch (src.depth()) {
        case CV_8U: minmax_t<uchar>(src, minVal, maxVal); break;
        case CV_16S: minmax_t<short>(src, minVal, maxVal); break;
        case CV_16U: minmax_t<ushort>(src, minVal, maxVal); break;
        case CV_32S: minmax_t<int>(src, minVal, maxVal); break;
        case CV_32F: minmax_t<float>(src, minVal, maxVal); break;
        default: minmax_t<double>(src, minVal, maxVal); break;
    }
}

// planes of a multi-channel array; mv is resized to the channel count
inline void split(const Mat& m, std::vector<Mat>& mv)
{
    const int cn = m.channels();
    const size_t e1 = m.elemSize1();
    mv.resize(cn);
    for (int c = 0; c < cn; ++c) {
        mv[c].create(m.rows, m.cols, CV_MAKETYPE(m.depth(), 1));
        for (int i = 0; i < m.rows; ++i) {
            const uchar* s = m.data + m.step * (size_t)i + e1 * c;
            uchar* d = mv[c].data + mv[c].step * (size_t)i;
            for (int j = 0; j < m.cols; ++j) std::memcpy(d + e1 * j, s + e1 * cn * j, e1);
        }
    }
}

// dst = saturate_cast<uchar>(|src * alpha + beta|), rounding to nearest even
inline void convertScaleAbs(const Mat& src, Mat& dst, double alpha = 1, double beta = 0)
{
    dst.create(src.rows, src.cols, CV_MAKETYPE(CV_8U, src.channels()));
    const int n = src.cols * src.channels();
    for (int i = 0; i < src.rows; ++i) {
        uchar* d = dst.ptr<uchar>(i);
        for (int j = 0; j < n; ++j) {
            double v;
            switch (src.depth()) {
                case CV_8U: v = src.ptr<uchar>(i)[j]; break;
                case CV_16S: v = src.ptr<short>(i)[j]; break;
                case CV_32S: v = src.ptr<int>(i)[j]; break;
                case CV_32F: v = src.ptr<float>(i)[j]; break;
                default: v = src.ptr<double>(i)[j]; break;
            }
            // for CV_32F sources OpenCV evaluates in float
            const int r = src.depth() == CV_32F ? cvRound(std::fabs((float)v * (float)alpha + (float)beta)) : cvRound(std::fabs(v * alpha + beta));
            d[j] = (uchar)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
    }
}

// field-compatible with the OpenCV 2.4 functor that src/stereo.cpp:13-30 configures; the implementation is un-vendored
// third-party code, so the call forwards to a hook (the C oracle's SGBM restatement installs itself there).
class StereoSGBM {
public:
    int minDisparity, numberOfDisparities, SADWindowSize, preFilterCap, uniquenessRatio, P1, P2, speckleWindowSize, speckleRange,
        disp12MaxDiff;
    bool fullDP;
    StereoSGBM()
        : minDisparity(0), numberOfDisparities(0), SADWindowSize(0), preFilterCap(0), uniquenessRatio(0), P1(0), P2(0),
          speckleWindowSize(0), speckleRange(0), disp12MaxDiff(0), fullDP(false) {}
    typedef void (*Hook)(const StereoSGBM&, const Mat&, const Mat&, Mat&);
    static Hook& hook() { static Hook h = 0; return h; }
    void operator()(const Mat& l, const Mat& r, Mat& disp) const
    {
        if (hook()) hook()(*this, l, r, disp);
        else { std::cerr << "cvstub: StereoSGBM has no implementation hook\n"; std::abort(); }
    }
};

}  // namespace cv
#include "cvstub_more.hpp"
#endif
