// ref_mapper_wrap.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points around the REFERENCE's own compiled
//   FrameReader::next            (src/rgbdframe.cpp:34-191: image reads, calDisparity_SGBM, the disparity -> depth loop :85-116),
//   RGBDFrame::project2dTo3d     (include/rgbdframe.h:63-75),
//   Mapper::semantic_motion_fuse (src/mapper.cpp:189-216),
//   Mapper::generatePointCloud   (src/mapper.cpp:12-94)
// which oracle/Makefile compiles from where they lie (against oracle/cvstub + oracle/refstub, see refstub/prelude.hpp) into
// oracle/_ref/libref_mapper.so.  Used by tests/ to pin the C oracle's restatement of that glue and by
// tests/golden/make_golden_mapper.py to produce fixtures.  cv::StereoSGBM's un-vendored implementation is supplied through the
// hook of ref_stereo_wrap.cpp (the C oracle's SGBM, itself pinned to cv2 4.13).
#define protected public   // Mapper::generatePointCloud / semantic_motion_fuse / moving_mask are protected (include/mapper.h:47-65)
#include "mapper.h"        // the reference's header
#undef protected
#include "stereo.h"

#include <sys/stat.h>

using namespace rgbd_tutor;

namespace {

struct Cam { double cx, cy, fx, fy, baseline, scale, roix, roiy, roiz; };

std::string num(double v)
{
    char b[64];
    snprintf(b, sizeof(b), "%.17g", v);
    return b;
}

// the keys FrameReader / Mapper / ParameterReader::getCamera read (parameters.txt)
void fill_params(ParameterReader& pr, const Cam& c, const std::string& data_source, int n_files, double resolution, double max_distance)
{
    std::map<std::string, std::string>& d = pr.data;
    d["data_source"] = data_source;
    d["rgb_dir"] = "image_2/";
    d["depth_dir"] = "image_3/";
    d["start_index"] = "0";
    d["end_index"] = std::to_string(n_files);
    d["camera.cx"] = num(c.cx); d["camera.cy"] = num(c.cy); d["camera.fx"] = num(c.fx); d["camera.fy"] = num(c.fy);
    d["camera.baseline"] = num(c.baseline); d["camera.scale"] = num(c.scale);
    d["camera.roix"] = num(c.roix); d["camera.roiy"] = num(c.roiy); d["camera.roiz"] = num(c.roiz);
    d["camera.d0"] = d["camera.d1"] = d["camera.d2"] = d["camera.d3"] = d["camera.d4"] = "0";
    d["mapper_resolution"] = num(resolution);
    d["mapper_max_distance"] = num(max_distance);
    d["motion_area_thres"] = "1000";
    d["motion_overlay_portion_thres"] = "0.143";
}

cv::Mat wrap(int h, int w, int type, const void* p) { return cv::Mat(h, w, type, const_cast<void*>(p)).clone(); }

// an empty parameter file, so that the ParameterReader constructor finds a file and stays quiet
std::string empty_params_file(const std::string& dir)
{
    const std::string path = dir + "/parameters.txt";
    FILE* f = fopen(path.c_str(), "w");
    if (f) { fputs("# filled in memory\n", f); fclose(f); }
    return path;
}

}  // namespace

extern "C" {

// FrameReader::next on one stereo frame.  `tmpdir` must exist and be writable: init_kitti counts the files of
// <tmpdir>/image_2/ (src/rgbdframe.cpp:229-247), so two empty files are created there (next() reads file index + 1 and index).
// left / right: 8UC1 w x h; rgb / semantic: 8UC3.  Outputs: depth (16UC1) and disparity (16SC1) of the returned frame.
// Returns 0, or -1 when next() returned no frame.
int ref_frame_next(const char* tmpdir, const unsigned char* left, const unsigned char* right, const unsigned char* rgb, const unsigned char* semantic,
                   int w, int h, const double* cam9, unsigned short* depth_out, short* disp_out)
{
    const Cam c = {cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8]};
    const std::string root = std::string(tmpdir) + "/";
    mkdir((root + "image_2").c_str(), 0700);
    for (int i = 0; i < 2; ++i) {
        char name[64];
        snprintf(name, sizeof(name), "image_2/%06d.png", i);
        FILE* f = fopen((root + name).c_str(), "w");
        if (f) fclose(f);
    }
    cv::ImreadRegistry& reg = cv::ImreadRegistry::get();
    reg.colour.clear(); reg.grey.clear();
    const cv::Mat L = wrap(h, w, CV_8UC1, left), R = wrap(h, w, CV_8UC1, right), C = wrap(h, w, CV_8UC3, rgb), S = wrap(h, w, CV_8UC3, semantic);
    for (int i = 0; i < 2; ++i) {
        char name[64];
        snprintf(name, sizeof(name), "%06d.png", i);
        reg.colour[root + "image_2/" + name] = C; reg.grey[root + "image_2/" + name] = L;
        reg.colour[root + "image_3/" + name] = C; reg.grey[root + "image_3/" + name] = R;
        reg.colour[root + "segnet_0/" + name] = S; reg.colour[root + "result_0/" + name] = S; reg.colour[root + "segnet_1/" + name] = S;
    }
    ParameterReader pr(empty_params_file(tmpdir));
    fill_params(pr, c, root, 1, 0.1, 40.0);
    FrameReader reader(pr, FrameReader::KITTI);
    RGBDFrame::Ptr frame = reader.next();
    reg.colour.clear(); reg.grey.clear();
    if (!frame || frame->depth.empty()) return -1;
    for (int i = 0; i < h; ++i) {
        std::memcpy(depth_out + (size_t)i * w, frame->depth.ptr<ushort>(i), sizeof(ushort) * w);
        std::memcpy(disp_out + (size_t)i * w, frame->disparity.ptr<short>(i), sizeof(short) * w);
    }
    return 0;
}

// Mapper::semantic_motion_fuse + Mapper::generatePointCloud on one frame.  Outputs: the moving mask (8UC1), the frame's cached
// camera-space cloud (frame->pointcloud: the reference's own push_back loop, no library arithmetic) as xyz_cam [n][3] / rgba [n]
// (0x00RRGGBB as PointXYZRGBA packs b, g, r, a), and the cloud generatePointCloud returns, transformed by T (row-major 4x4,
// through the stand-in's pcl::transformPointCloud), as xyz_world.  Returns the number of points (outputs hold max_points).
int ref_mapper_cloud(const char* tmpdir, const unsigned short* depth, const unsigned char* semantic, const unsigned char* rgb, int w, int h,
                     const double* cam9, double max_distance, const double* T16, unsigned char* mask_out, float* xyz_cam, float* xyz_world,
                     unsigned int* rgba, int max_points)
{
    const Cam c = {cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8]};
    ParameterReader pr(empty_params_file(tmpdir));
    fill_params(pr, c, std::string(tmpdir) + "/", 0, 0.1, max_distance);
    PoseGraph graph;
    Mapper mapper(pr, graph);          // spawns the viewer thread, which idles: the stand-in PoseGraph has no keyframes
    RGBDFrame::Ptr frame(new RGBDFrame);
    frame->id = 0;
    frame->depth = wrap(h, w, CV_16UC1, depth);
    frame->semantic = wrap(h, w, CV_8UC3, semantic);
    frame->rgb = wrap(h, w, CV_8UC3, rgb);
    frame->result = frame->rgb.clone();          // read (and dropped) by generatePointCloud, src/mapper.cpp:62-68
    frame->camera = pr.getCamera();
    // first under the identity pose: the returned cloud is the camera-space cloud (1 * x + 0 * y + 0 * z + 0 is exact); the
    // reference also leaves it in frame->pointcloud (its own push_back loop, no library arithmetic), which is preferred when set
    frame->setTransform(Eigen::Isometry3d::Identity());
    Mapper::PointCloud::Ptr cam = mapper.generatePointCloud(frame);
    if (frame->pointcloud != nullptr) cam = frame->pointcloud;
    Eigen::Isometry3d T;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T(i, j) = T16[4 * i + j];
    frame->setTransform(T);
    Mapper::PointCloud::Ptr world = mapper.generatePointCloud(frame);
    for (int i = 0; i < h; ++i) std::memcpy(mask_out + (size_t)i * w, mapper.moving_mask.ptr<uchar>(i), w);
    const int n = (int)cam->points.size();
    if ((int)world->points.size() != n) { mapper.shutdown(); return -1; }
    for (int i = 0; i < n && i < max_points; ++i) {
        const Mapper::PointT& p = cam->points[i];
        const Mapper::PointT& q = world->points[i];
        xyz_cam[3 * i] = p.x; xyz_cam[3 * i + 1] = p.y; xyz_cam[3 * i + 2] = p.z;
        xyz_world[3 * i] = q.x; xyz_world[3 * i + 1] = q.y; xyz_world[3 * i + 2] = q.z;
        rgba[i] = ((unsigned)p.r << 16) | ((unsigned)p.g << 8) | (unsigned)p.b;
    }
    mapper.shutdown();
    return n;
}

}  // extern "C"
