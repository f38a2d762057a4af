// host.cpp -- C++ host layer above the C ABI: calDisparity_SGBM and rgbd_tutor::Mapper mirrors.
// Links against libssm.so only (no CUDA headers, no OpenCV/PCL).  See stereo.hpp / mapper.hpp.
#include <chrono>
#include <cstdio>
#include <cstring>

#include "mapper.hpp"
#include "stereo.hpp"

namespace ssm_host {

// ---------------------------------------------------------------------------------------------
// stereo.h
// ---------------------------------------------------------------------------------------------
static StereoConfig g_cfg;
static ssm_ctx* g_stereo = nullptr;
static std::mutex g_stereo_mutex;

void setStereoConfig(const StereoConfig& cfg)
{
    std::lock_guard<std::mutex> lk(g_stereo_mutex);
    if (g_stereo) { ssm_destroy(g_stereo); g_stereo = nullptr; }
    g_cfg = cfg;
}

static void fill_camera(ssm_params& p, const Camera& c)
{
    p.cx = c.cx; p.cy = c.cy; p.fx = c.fx; p.fy = c.fy; p.scale = c.scale;
    p.baseline = c.baseline; p.roix = c.roix; p.roiy = c.roiy; p.roiz = c.roiz;
}

ssm_ctx* stereoContext()
{
    std::lock_guard<std::mutex> lk(g_stereo_mutex);
    if (!g_stereo) {
        ssm_params p;
        ssm_default_params(&p);   // stereo.cpp:16-28
        p.num_disparities = g_cfg.num_disparities;
        p.max_width = g_cfg.max_width; p.max_height = g_cfg.max_height; p.max_batch = 1;
        p.map_capacity = 1024;    // this context never maps
        fill_camera(p, g_cfg.camera);
        check(ssm_create(&p, g_cfg.device, &g_stereo));
    }
    return g_stereo;
}

void releaseStereoContext()
{
    std::lock_guard<std::mutex> lk(g_stereo_mutex);
    if (g_stereo) { ssm_destroy(g_stereo); g_stereo = nullptr; }
}

void calDisparity_SGBM(const ImageU8& img_L, const ImageU8& img_R, ImageS16& disp)
{
    if (img_L.empty() || img_R.empty() || img_L.rows != img_R.rows || img_L.cols != img_R.cols || img_L.step != img_R.step)
        throw Error(SSM_ERR_INVALID_ARGUMENT, "calDisparity_SGBM: left/right must be non-empty 8UC1 images of equal size");
    disp.create(img_L.rows, img_L.cols);
    check(ssm_sgbm(stereoContext(), img_L.data, img_R.data, img_L.cols, img_L.rows, img_L.step, disp.data, disp.step));
}

void disparityToDepth(const ImageS16& disp, ImageU16& depth)
{
    if (disp.empty()) throw Error(SSM_ERR_INVALID_ARGUMENT, "disparityToDepth: empty disparity");
    depth.create(disp.rows, disp.cols);
    check(ssm_disparity_to_depth(stereoContext(), disp.data, disp.cols, disp.rows, disp.step, depth.data, depth.step));
}

// ---------------------------------------------------------------------------------------------
// mapper.h
// ---------------------------------------------------------------------------------------------
Mapper::Mapper(const MapperConfig& para, KeyframeSource& graph, bool start_thread) : config(para), poseGraph(graph)
{
    ssm_params p;
    ssm_default_params(&p);
    p.resolution = para.mapper_resolution;      // mapper.h:24
    p.max_distance = para.mapper_max_distance;  // mapper.h:25
    p.num_labels = para.num_labels;
    p.max_width = para.max_width; p.max_height = para.max_height; p.max_batch = 1;
    p.num_disparities = 16;                     // the mapping context never runs SGBM: keep its cost buffers minimal
    p.map_capacity = para.map_capacity;
    fill_camera(p, para.camera);
    check(ssm_create(&p, para.device, &ctx));
    if (start_thread) viewerThread = std::make_shared<std::thread>(std::bind(&Mapper::viewer, this));
}

Mapper::~Mapper()
{
    shutdown();
    if (ctx) ssm_destroy(ctx);
}

void Mapper::shutdown()
{
    shutdownFlag = true;
    if (viewerThread != nullptr && viewerThread->joinable()) viewerThread->join();
}

void Mapper::semantic_motion_fuse(const Frame::Ptr& frame)
{
    const ImageBGR& s = frame->semantic;
    moving_mask.create(s.rows, s.cols);
    check(ssm_semantic_motion_fuse(ctx, s.data, s.cols, s.rows, s.step, moving_mask.data, moving_mask.step));
}

std::shared_ptr<Mapper::PointCloud> Mapper::generatePointCloud(const Frame::Ptr& frame)
{
    const int w = frame->depth.cols, h = frame->depth.rows;
    if (frame->depth.empty() || frame->semantic.rows != h || frame->semantic.cols != w || frame->rgb.rows != h || frame->rgb.cols != w)
        throw Error(SSM_ERR_INVALID_ARGUMENT, "generatePointCloud: depth / semantic / rgb must have equal size");
    if (frame->depth.step != (size_t)w * 2 || frame->semantic.step != (size_t)w * 3 || frame->rgb.step != (size_t)w * 3)
        throw Error(SSM_ERR_INVALID_ARGUMENT, "generatePointCloud: images must be densely packed");
    const std::array<double, 16> T = frame->getTransform();   // under frame->mutexT (rgbdframe.h:116-120)
    std::vector<float> xyz((size_t)w * h * 3);
    std::vector<uint32_t> rgba((size_t)w * h);
    std::vector<uint8_t> label((size_t)w * h);
    int n = 0;
    check(ssm_generate_point_cloud(ctx, frame->depth.data, frame->semantic.data, frame->rgb.data, w, h, T.data(), xyz.data(),
                                   rgba.data(), label.data(), w * h, &n));
    auto cloud = std::make_shared<PointCloud>((size_t)n);
    for (int i = 0; i < n; ++i) (*cloud)[i] = PointT{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgba[i], label[i]};
    return cloud;
}

void Mapper::integrate(const Frame::Ptr& frame)
{
    const int w = frame->depth.cols, h = frame->depth.rows;
    const std::array<double, 16> T = frame->getTransform();
    check(ssm_map_integrate_frame(ctx, frame->depth.data, frame->semantic.data, frame->rgb.data, w, h, T.data()));
}

void Mapper::viewer()
{
    while (!shutdownFlag) {
        std::vector<Frame::Ptr> kfs;
        {
            std::lock_guard<std::mutex> lk(poseGraph.keyframes_mutex);   // the reference reads this list unlocked (App. C-11)
            if (poseGraph.keyframes.size() > keyframe_size) kfs = poseGraph.keyframes;
        }
        if (kfs.empty()) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));   // usleep(1000), mapper.cpp:116
            continue;
        }
        const auto t0 = std::chrono::steady_clock::now();
        if (config.redraw_every > 0 && cntGlobalUpdate % config.redraw_every == 0) {
            check(ssm_map_clear(ctx));                                   // globalMap->clear(), mapper.cpp:125
            for (const Frame::Ptr& f : kfs) integrate(f);
        } else {
            for (size_t i = keyframe_size; i < kfs.size(); ++i) integrate(kfs[i]);
        }
        cntGlobalUpdate++;
        keyframe_size = kfs.size();
        uint64_t n = 0;
        check(ssm_map_size(ctx, &n));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::printf("points in global map: %llu\nMapping cost time: %.3fms\n", (unsigned long long)n, ms);   // mapper.cpp:161-162
    }
    if (poseGraph.shutDownFlag && !config.save_path.empty()) {          // mapper.cpp:165-170
        SaveMap();
        std::printf("Map saved!\n");
    }
}

void Mapper::SaveMap()
{
    if (config.save_path.empty()) throw Error(SSM_ERR_INVALID_ARGUMENT, "SaveMap: MapperConfig.save_path is empty");
    check(ssm_map_save_pcd(ctx, config.save_path.c_str()));
}

uint64_t Mapper::mapSize()
{
    uint64_t n = 0;
    check(ssm_map_size(ctx, &n));
    return n;
}

VoxelCloud Mapper::exportMap(bool pcl_order)
{
    VoxelCloud v;
    const uint64_t n = mapSize();
    v.ijk.resize(n * 3); v.xyz.resize(n * 3); v.rgba.resize(n); v.count.resize(n); v.label.resize(n);
    v.votes.resize(n * (size_t)config.num_labels);
    ssm_voxel_export e{v.ijk.data(), v.xyz.data(), v.rgba.data(), v.label.data(), v.count.data(), v.votes.data()};
    uint64_t got = 0;
    check(ssm_map_export(ctx, &e, n, pcl_order ? 1 : 0, &got));
    return v;
}

}  // namespace ssm_host

#ifdef SSM_WITH_OPENCV
void calDisparity_SGBM(const cv::Mat& img_L, const cv::Mat& img_R, cv::Mat& disp)
{
    CV_Assert(img_L.type() == CV_8UC1 && img_R.type() == CV_8UC1 && img_L.size() == img_R.size() && img_L.step == img_R.step);
    disp.create(img_L.size(), CV_16SC1);
    ssm_host::check(ssm_sgbm(ssm_host::stereoContext(), img_L.data, img_R.data, img_L.cols, img_L.rows, img_L.step,
                             (int16_t*)disp.data, disp.step));
}
#endif
