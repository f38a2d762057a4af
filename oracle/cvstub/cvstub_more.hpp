// cvstub_more.hpp -- TEST INFRASTRUCTURE.  The further OpenCV 2.4 names that /root/reference/src/uvdisparity.cpp (and the
// headers it pulls in: vo_stereo.hpp, vo.hpp, quadmatcher.hpp) mention, so that the reference's OWN file compiles here and
// UVDisparity::calUDisparity / calVDisparity (src/uvdisparity.cpp:195-274, 277-366) can be EXECUTED as the parity anchor of
// the U/V-disparity restatement.  Two kinds of names:
//   * used by those two functions: implemented (Mat::zeros, Mat::size, create(Size), cvtColor GRAY2BGR, cvIsInf / cvIsNaN);
//   * used only by the rest of the file (pitch estimation, Kalman filters, flood-fill segmentation, drawing, feature
//     matching members of the included classes): declared so that the file parses and links; calling one aborts.
// Written against the public OpenCV 2.4 API documentation; nothing is copied from OpenCV.
#ifndef SSM_CVSTUB_MORE_HPP
#define SSM_CVSTUB_MORE_HPP

#include <map>

#define CV_GRAY2BGR 8
#define CV_BGR2GRAY 6
#define CV_DIST_L2 2
#define CV_DIST_HUBER 7
#define CV_AA 16
#define CV_FILLED -1

inline int cvIsNaN(double v) { return std::isnan(v) ? 1 : 0; }
inline int cvIsInf(double v) { return std::isinf(v) ? 1 : 0; }

namespace cv {

[[noreturn]] inline void stub_unreachable(const char* what)
{
    std::cerr << "cvstub: " << what << " is declared for compilation only and must not be executed\n";
    std::abort();
}

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {}
};
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int a, int b) : start(a), end(b) {}
    static Range all() { return Range(-2147483647 - 1, 2147483647); }
};
template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;

template <typename T> struct Ptr {
    T* p;
    Ptr() : p(0) {}
    Ptr(T* q) : p(q) {}
    T* operator->() const { return p; }
    T& operator*() const { return *p; }
    bool empty() const { return p == 0; }
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
};
class Algorithm {};
class FeatureDetector : public Algorithm {};
class DescriptorExtractor : public Algorithm {};
class DescriptorMatcher : public Algorithm {};
struct TermCriteria {
    int type, maxCount;
    double epsilon;
    TermCriteria(int t = 0, int m = 0, double e = 0) : type(t), maxCount(m), epsilon(e) {}
};

// Mat_<T>(r, c) << a, b, ... : the comma initialiser; operator* on it yields the matrix (as OpenCV's does)
template <typename T> struct MatDepth;
template <> struct MatDepth<uchar> { enum { value = CV_8U }; };
template <> struct MatDepth<short> { enum { value = CV_16S }; };
template <> struct MatDepth<int> { enum { value = CV_32S }; };
template <> struct MatDepth<float> { enum { value = CV_32F }; };
template <> struct MatDepth<double> { enum { value = CV_64F }; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(MatDepth<T>::value, 1)) {}
    T& operator()(int i, int j) { return this->template at<T>(i, j); }
};
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<T>* m) : m_(m), n_(0) {}
    template <typename V> MatCommaInitializer_<T>& operator,(V v)
    {
        m_->template at<T>(n_ / m_->cols, n_ % m_->cols) = (T)v;
        ++n_;
        return *this;
    }
    Mat_<T> operator*() const { return *m_; }
    operator Mat_<T>() const { return *m_; }
    Mat_<T>* m_;
    int n_;
};
template <typename T, typename V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v)
{
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    return (ci, v);
}

class KalmanFilter {
public:
    KalmanFilter() {}
    KalmanFilter(int dynamParams, int measureParams, int controlParams = 0, int type = CV_32F) { init(dynamParams, measureParams, controlParams, type); }
    void init(int dp, int mp, int cp = 0, int type = CV_32F)
    {
        statePre.create(dp, 1, type); statePost.create(dp, 1, type); transitionMatrix.create(dp, dp, type);
        processNoiseCov.create(dp, dp, type); measurementMatrix.create(mp, dp, type); measurementNoiseCov.create(mp, mp, type);
        errorCovPre.create(dp, dp, type); errorCovPost.create(dp, dp, type); gain.create(dp, mp, type);
        (void)cp;
    }
    const Mat& predict(const Mat& = Mat()) { stub_unreachable("KalmanFilter::predict"); }
    const Mat& correct(const Mat&) { stub_unreachable("KalmanFilter::correct"); }
    Mat statePre, statePost, transitionMatrix, controlMatrix, measurementMatrix, processNoiseCov, measurementNoiseCov, errorCovPre, gain,
        errorCovPost;
};

// ---- implemented: what calUDisparity / calVDisparity execute -------------------------------------------------------------
inline void setIdentity(Mat& m, const Scalar& s = Scalar(1))
{
    for (int i = 0; i < m.rows; ++i)
        for (int j = 0; j < m.cols; ++j) {
            const double v = i == j ? s.val[0] : 0.0;
            if (m.depth() == CV_32F) m.at<float>(i, j) = (float)v;
            else if (m.depth() == CV_64F) m.at<double>(i, j) = v;
        }
}
// GRAY -> BGR replicates the channel (the only code the reference passes on the executed path, for a display image)
inline void cvtColor(const Mat& src, Mat& dst, int code, int = 0)
{
    if (code != CV_GRAY2BGR || src.type() != CV_8UC1) stub_unreachable("cvtColor (other than 8-bit GRAY2BGR)");
    dst.create(src.rows, src.cols, CV_8UC3);
    for (int i = 0; i < src.rows; ++i)
        for (int j = 0; j < src.cols; ++j) {
            const uchar v = src.ptr<uchar>(i)[j];
            uchar* d = dst.ptr<uchar>(i) + 3 * j;
            d[0] = v; d[1] = v; d[2] = v;
        }
}

// ---- implemented: what FrameReader::next and the Mapper execute (oracle/refstub/prelude.hpp) ---------------------------------------
#define CV_LOAD_IMAGE_UNCHANGED -1
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_LOAD_IMAGE_COLOR 1
inline void fill_scalar(Mat& m, const Scalar& v)
{
    const int cn = m.channels();
    for (int i = 0; i < m.rows; ++i)
        for (int j = 0; j < m.cols; ++j)
            for (int c = 0; c < cn; ++c) {
                const double x = v.val[c < 4 ? c : 3];
                const int k = j * cn + c;
                switch (m.depth()) {
                    case CV_8U: m.ptr<uchar>(i)[k] = (uchar)x; break;
                    case CV_16U: m.ptr<ushort>(i)[k] = (ushort)x; break;
                    case CV_16S: m.ptr<short>(i)[k] = (short)x; break;
                    case CV_32S: m.ptr<int>(i)[k] = (int)x; break;
                    case CV_32F: m.ptr<float>(i)[k] = (float)x; break;
                    default: m.ptr<double>(i)[k] = x; break;
                }
            }
}
inline Mat::Mat(int r, int c, int t, const Scalar& v) : rows(0), cols(0), data(0), step(0), type_(0), uninit_(false) { create(r, c, t); fill_scalar(*this, v); }
inline Mat::Mat(Size s, int t, const Scalar& v) : rows(0), cols(0), data(0), step(0), type_(0), uninit_(false) { create(s.height, s.width, t); fill_scalar(*this, v); }

// cv::imread served from memory: the test harness registers the images under the paths the reference will ask for
// (FrameReader::next reads the same file once with the default flag -- colour -- and once with flag 0 -- grey).  An
// unregistered path is a missing file: an empty matrix, as cv::imread returns.
struct ImreadRegistry {
    std::map<std::string, Mat> colour, grey;
    static ImreadRegistry& get() { static ImreadRegistry r; return r; }
};
inline Mat imread(const std::string& path, int flags = 1)
{
    ImreadRegistry& r = ImreadRegistry::get();
    std::map<std::string, Mat>& m = flags == 0 ? r.grey : r.colour;
    std::map<std::string, Mat>::const_iterator it = m.find(path);
    if (it == m.end() && flags < 0) { it = r.grey.find(path); if (it == r.grey.end()) return Mat(); return it->second.clone(); }
    return it == m.end() ? Mat() : it->second.clone();
}

// cv::dilate on 8UC1 with the default border (out-of-image pixels never win) and the anchor at the kernel centre.  A kernel
// whose pixels were never defined -- the reference passes cv::Mat(3,3,CV_8UC1) straight from the constructor,
// src/mapper.cpp:214 -- acts as all ones: the canonical choice of SURVEY App. C-3 (heap garbage is non-zero).
inline void dilate(const Mat& src, Mat& dst, const Mat& kernel, Point anchor = Point(-1, -1), int iterations = 1)
{
    if (src.type() != CV_8UC1) stub_unreachable("dilate (other than 8UC1)");
    const int kr = kernel.rows, kc = kernel.cols;
    const int ay = anchor.y < 0 ? kr / 2 : anchor.y, ax = anchor.x < 0 ? kc / 2 : anchor.x;
    Mat cur = src.clone();
    for (int it = 0; it < iterations; ++it) {
        Mat out = Mat::zeros(cur.rows, cur.cols, CV_8UC1);
        for (int y = 0; y < cur.rows; ++y)
            for (int x = 0; x < cur.cols; ++x) {
                int best = 0;
                for (int i = 0; i < kr; ++i)
                    for (int j = 0; j < kc; ++j) {
                        if (!kernel.uninitialised() && kernel.ptr<uchar>(i)[j] == 0) continue;
                        const int yy = y + i - ay, xx = x + j - ax;
                        if (yy < 0 || yy >= cur.rows || xx < 0 || xx >= cur.cols) continue;
                        const int v = cur.ptr<uchar>(yy)[xx];
                        if (v > best) best = v;
                    }
                out.ptr<uchar>(y)[x] = (uchar)best;
            }
        cur = out;
    }
    dst = cur;
}

// ---- declared only ---------------------------------------------------------------------------------------------------------
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3, THRESH_OTSU = 8 };
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { FLOODFILL_FIXED_RANGE = 1 << 16, FLOODFILL_MASK_ONLY = 1 << 17 };
inline double threshold(const Mat&, Mat&, double, double, int) { stub_unreachable("threshold"); }
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { stub_unreachable("line"); }
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) { stub_unreachable("circle"); }
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { stub_unreachable("rectangle"); }
inline void imshow(const std::string&, const Mat&) { stub_unreachable("imshow"); }
inline int waitKey(int = 0) { stub_unreachable("waitKey"); }
inline bool imwrite(const std::string&, const Mat&) { stub_unreachable("imwrite"); }
inline Mat getStructuringElement(int, Size, Point = Point(-1, -1)) { stub_unreachable("getStructuringElement"); }
inline void erode(const Mat&, Mat&, const Mat&, Point = Point(-1, -1), int = 1) { stub_unreachable("erode"); }
inline void bitwise_or(const Mat&, const Mat&, Mat&) { stub_unreachable("bitwise_or"); }
inline void bitwise_and(const Mat&, const Mat&, Mat&) { stub_unreachable("bitwise_and"); }
inline int floodFill(Mat&, Point, Scalar, Rect* = 0, Scalar = Scalar(), Scalar = Scalar(), int = 4) { stub_unreachable("floodFill"); }
inline int floodFill(Mat&, Mat&, Point, Scalar, Rect* = 0, Scalar = Scalar(), Scalar = Scalar(), int = 4) { stub_unreachable("floodFill"); }
inline void GaussianBlur(const Mat&, Mat&, Size, double, double = 0, int = 4) { stub_unreachable("GaussianBlur"); }
inline Mat Mat::operator()(const Range&, const Range&) const { stub_unreachable("Mat::operator()(Range, Range)"); }
inline Mat Mat::operator()(const Rect&) const { stub_unreachable("Mat::operator()(Rect)"); }
inline void Mat::copyTo(Mat&, const Mat&) const { stub_unreachable("Mat::copyTo(dst, mask)"); }
inline Mat& Mat::operator=(const Scalar&) { stub_unreachable("Mat::operator=(Scalar)"); }
inline void Mat::push_back(const Mat&) { stub_unreachable("Mat::push_back"); }
template <typename P> inline void fitLine(const std::vector<P>&, Vec4f&, int, double, double, double) { stub_unreachable("fitLine"); }
inline void fitLine(const Mat&, Vec4f&, int, double, double, double) { stub_unreachable("fitLine"); }

}  // namespace cv
#endif
