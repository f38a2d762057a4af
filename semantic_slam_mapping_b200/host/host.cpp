// host.cpp -- C++ host layer above the C ABI: calDisparity_SGBM and rgbd_tutor::Mapper mirrors.
// Links against libssm.so only (no CUDA headers, no OpenCV/PCL).  See stereo.hpp / mapper.hpp.
#include <algorithm>
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
// stereo.h / uvdisparity.hpp: dense motion cues
// ---------------------------------------------------------------------------------------------
static void need_dense(bool ok, const char* what)
{
    if (!ok) throw Error(SSM_ERR_INVALID_ARGUMENT, std::string(what) + ": images must be non-empty, equally sized and densely packed");
}

void triangulate10D(const ImageU8& img, const ImageS16& disp, ImageXYZ10& xyz, const double f, const double cx, const double cy,
                    const double b, ROI3D roi)
{
    need_dense(!img.empty() && !disp.empty() && img.rows == disp.rows && img.cols == disp.cols, "triangulate10D");
    xyz.create(disp.rows, disp.cols);
    check(ssm_triangulate10d(stereoContext(), img.data, img.step, disp.data, disp.step, disp.cols, disp.rows, f, cx, cy, b, roi.x_max,
                             roi.y_max, roi.z_max, xyz.data));
}

void correct3DPoints(ImageXYZ10& xyz, ROI3D& roi_, const double& pitch1, const double& pitch2)
{
    need_dense(!xyz.empty(), "correct3DPoints");
    check(ssm_correct_3d_points(stereoContext(), xyz.data, xyz.cols, xyz.rows, roi_.x_max, roi_.y_max, roi_.z_max, pitch1, pitch2));
}

void setImageROI(ImageXYZ10& xyz, ImageU8& roi_mask)
{
    need_dense(!xyz.empty(), "setImageROI");
    roi_mask.create(xyz.rows, xyz.cols);
    check(ssm_set_image_roi(stereoContext(), xyz.data, xyz.cols, xyz.rows, roi_mask.data, roi_mask.step));
}

void UVDisparity::calVDisparity(const ImageS16& img_dis, ImageXYZ10& xyz)
{
    need_dense(!img_dis.empty() && xyz.rows == img_dis.rows && xyz.cols == img_dis.cols, "calVDisparity");
    int v_cols = 0;
    ssm_ctx* c = stereoContext();
    check(ssm_v_disparity(c, img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, nullptr, nullptr, nullptr, 0, &v_cols));   // size
    v_dis_int = ImageS32(); v_dis_ = ImageU8();
    v_dis_int.create(img_dis.rows, v_cols);
    v_dis_.create(img_dis.rows, v_cols);
    check(ssm_v_disparity(c, img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, xyz.data, v_cols ? v_dis_int.data : nullptr,
                          v_cols ? v_dis_.data : nullptr, v_cols, &v_cols));
}

void UVDisparity::calUDisparity(const ImageS16& img_dis, ImageXYZ10& xyz, ImageU8& roi_mask, ImageU8& ground_mask)
{
    need_dense(!img_dis.empty() && xyz.rows == img_dis.rows && xyz.cols == img_dis.cols && roi_mask.rows == img_dis.rows &&
                   roi_mask.cols == img_dis.cols && ground_mask.rows == img_dis.rows && ground_mask.cols == img_dis.cols &&
                   roi_mask.step == (size_t)img_dis.cols && ground_mask.step == (size_t)img_dis.cols,
               "calUDisparity");
    int u_rows = 0;
    ssm_ctx* c = stereoContext();
    check(ssm_u_disparity(c, img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, nullptr, roi_mask.data, ground_mask.data, nullptr, nullptr, 0,
                          &u_rows));
    u_dis_int = ImageS32(); u_dis_ = ImageU8();
    u_dis_int.create(u_rows, img_dis.cols);
    u_dis_.create(u_rows, img_dis.cols);
    check(ssm_u_disparity(c, img_dis.data, img_dis.step, img_dis.cols, img_dis.rows, xyz.data, roi_mask.data, ground_mask.data, u_dis_int.data,
                          u_dis_.data, u_rows, &u_rows));
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

// frame->pointcloud of the reference (mapper.cpp:17-20): the keyframe's cloud is built once, kept on the device in camera
// coordinates and re-transformed by the frame's current pose whenever the map is (re)drawn
int Mapper::cloudOf(const Frame::Ptr& frame)
{
    if (frame->cloud_id < 0) {
        const int w = frame->depth.cols, h = frame->depth.rows;
        const std::array<double, 16> T = frame->getTransform();
        check(ssm_keyframe_add(ctx, frame->depth.data, frame->semantic.data, frame->rgb.data, w, h, T.data(), &frame->cloud_id));
    }
    return frame->cloud_id;
}

void Mapper::viewer()
{
    // The body runs on its own std::thread: an exception that left it would end the process through std::terminate (the
    // reference's thread has the same property for cv::Exception / PCL exceptions).  A library error -- the voxel hash out
    // of device memory, a CUDA failure -- is recorded instead and stops the updates; failed() / lastError() report it and
    // shutdown() still joins cleanly.
    try {
        viewerLoop();
    } catch (const std::exception& e) {
        {
            std::lock_guard<std::mutex> lk(errorMutex);
            errorText = e.what();
        }
        failedFlag = true;
        std::fprintf(stderr, "Mapper::viewer stopped: %s\n", e.what());
    }
}

bool Mapper::failed() const { return failedFlag; }
std::string Mapper::lastError() const
{
    std::lock_guard<std::mutex> lk(errorMutex);
    return errorText;
}

void Mapper::viewerLoop()
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
        std::vector<int> ids;
        if (config.redraw_every > 0 && cntGlobalUpdate % config.redraw_every == 0) {
            // periodic full redraw (mapper.cpp:121-131): poses may have been rewritten by the pose graph since the last one
            for (size_t i = 0; i < kfs.size(); i += (size_t)std::max(1, config.redraw_stride)) {
                const int id = cloudOf(kfs[i]);
                const std::array<double, 16> T = kfs[i]->getTransform();
                check(ssm_keyframe_set_pose(ctx, id, T.data()));
                ids.push_back(id);
            }
            check(ssm_map_redraw(ctx, ids.data(), (int)ids.size()));     // globalMap->clear(); += every listed keyframe
        } else {
            for (size_t i = keyframe_size; i < kfs.size(); ++i) {
                const int id = cloudOf(kfs[i]);
                const std::array<double, 16> T = kfs[i]->getTransform();
                check(ssm_keyframe_set_pose(ctx, id, T.data()));
                ids.push_back(id);
            }
            check(ssm_map_integrate_keyframes(ctx, ids.data(), (int)ids.size()));
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
