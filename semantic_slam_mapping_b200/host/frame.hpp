// frame.hpp -- the data that crosses between the two halves of the path, shaped like the reference's
// RGBDFrame (/root/reference include/rgbdframe.h:26-121) but without OpenCV / PCL / Eigen types, so this
// host layer builds wherever libssm.so does.  With -DSSM_WITH_OPENCV the cv::Mat overloads at the bottom
// let reference code pass its own matrices unchanged.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ssm.h"

namespace ssm_host {

// thrown where the reference would surface a cv::Exception from an OpenCV assert (stereo.cpp:30)
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc)
{
    if (rc != SSM_OK) throw Error(rc, std::string("libssm: ") + ssm_last_error());
}

// minimal cv::Mat stand-in: an owning, row-major image with a byte step
template <typename T, int CH = 1>
struct Image {
    int rows = 0, cols = 0;
    size_t step = 0;   // bytes between rows
    std::vector<T> store;
    T* data = nullptr;

    Image() = default;
    Image(int r, int c) { create(r, c); }
    void create(int r, int c)
    {
        if (r == rows && c == cols && data) return;
        rows = r; cols = c; step = sizeof(T) * CH * (size_t)c;
        store.assign((size_t)r * c * CH, T());
        data = store.data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    T* ptr(int r) { return reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(data) + step * r); }
    const T* ptr(int r) const { return reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(data) + step * r); }
};
using ImageU8 = Image<uint8_t, 1>;
using ImageBGR = Image<uint8_t, 3>;
using ImageS16 = Image<int16_t, 1>;
using ImageU16 = Image<uint16_t, 1>;

// include/utils.h CAMERA_INTRINSIC_PARAMETERS + the ROI / baseline keys read at src/rgbdframe.cpp:87-94
struct Camera {
    double cx = 607.1928, cy = 185.2157, fx = 718.8560, fy = 718.8560, scale = 1000.0;
    double baseline = 0.532331858, roix = 20, roiy = 5, roiz = 40;
};

struct Point3f { float x = 0, y = 0, z = 0; };

// pcl::PointXYZRGBA as the Mapper uses it: xyz + packed 0x00RRGGBB; label is the north_star addition
struct PointXYZRGBL {
    float x, y, z;
    uint32_t rgba;
    uint8_t label;
};
using PointCloud = std::vector<PointXYZRGBL>;

struct Frame {
    using Ptr = std::shared_ptr<Frame>;
    int id = -1;
    ImageBGR rgb, semantic;
    ImageU16 depth;
    ImageS16 disparity;
    Camera camera;
    std::array<double, 16> T_f_w{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};   // camera -> world, row-major
    std::mutex mutexT;
    int cloud_id = -1;   // frame->pointcloud (rgbdframe.h:58): id of the cached camera-space cloud on the Mapper's device, -1 = not built

    void setTransform(const std::array<double, 16>& T)
    {
        std::unique_lock<std::mutex> lck(mutexT);
        T_f_w = T;
    }
    std::array<double, 16> getTransform()
    {
        std::unique_lock<std::mutex> lck(mutexT);
        return T_f_w;
    }
    // rgbdframe.h:63-75 (kept on the host for sparse callers such as OrbFeature, orb.h:50; the dense
    // path evaluates the same expression on the GPU)
    Point3f project2dTo3d(int u, int v) const
    {
        Point3f p;
        if (depth.empty()) return p;
        const uint16_t d = depth.ptr(v)[u];
        if (d == 0) return p;
        p.z = (float)(double(d) / camera.scale);
        p.x = (float)((u - camera.cx) * p.z / camera.fx);
        p.y = (float)((v - camera.cy) * p.z / camera.fy);
        return p;
    }
};

// what the Mapper needs from the reference's PoseGraph (include/pose_graph.h:150-158): the keyframe list
// under its mutex, and the shutdown flag
struct KeyframeSource {
    std::vector<Frame::Ptr> keyframes;
    std::mutex keyframes_mutex;
    bool shutDownFlag = false;
};

}  // namespace ssm_host
