// dropin/mapper.cpp -- drop-in replacement for the reference's src/mapper.cpp.
//
// Compiled INSIDE the reference's tree in place of src/mapper.cpp: it includes the reference's own include/mapper.h (class
// rgbd_tutor::Mapper with its inline constructor Mapper(const ParameterReader&, PoseGraph&), include/mapper.h:21-30) and defines
// the member functions that header declares,
//   Mapper::generatePointCloud     reference body src/mapper.cpp:12-94
//   Mapper::viewer                 src/mapper.cpp:96-178
//   Mapper::SaveMap                src/mapper.cpp:179-187 (empty in the reference)
//   Mapper::semantic_motion_fuse   src/mapper.cpp:189-216
// on top of libssm.so.  The class layout is the reference's, so the per-Mapper state the GPU path needs (its ssm_ctx, the ids of
// the keyframes' device-resident clouds) lives in a side table keyed by the Mapper's address.
//
// Semantics kept (src/mapper.cpp:109-163): poll the keyframe list, never block the producer; every 15th update redraw the whole
// map from the keyframes' CURRENT poses (the pose graph rewrites them, src/pose_graph.cpp:253-260), otherwise fuse what arrived;
// show the fused cloud on every update; write the PCD when the pose graph has shut down.  Canonical differences:
// each keyframe contributes exactly once and fusion is one VoxelGrid pass over the union -- on the GPU voxel hash, not by
// re-filtering a growing pcl cloud.
#include "mapper.h"   // the reference's header

#include <pcl/filters/voxel_grid.h>
#include <pcl/io/pcd_io.h>
#include <pcl/visualization/cloud_viewer.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "ssm.h"
#include "dropin.hpp"

using namespace rgbd_tutor;

namespace {

struct MapperState {
    ssm_ctx* ctx = nullptr;
    int w = 0, h = 0;
    std::map<const RGBDFrame*, int> cloud_id;   // frame->pointcloud of the reference (src/mapper.cpp:17-20), kept on the device
};
std::mutex g_mutex;
std::map<const void*, MapperState> g_state;

// the Mapper's context: created on first use from the first frame's camera and size
MapperState& state_of(const void* self, const RGBDFrame::Ptr& frame, double resolution, double max_distance)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    MapperState& st = g_state[self];
    const int w = frame->depth.cols, h = frame->depth.rows;
    if (st.ctx && (w > st.w || h > st.h)) throw std::runtime_error("Mapper: frame larger than the first keyframe");
    if (!st.ctx) {
        ssm_params p;
        ssm_default_params(&p);
        p.cx = frame->camera.cx; p.cy = frame->camera.cy; p.fx = frame->camera.fx; p.fy = frame->camera.fy; p.scale = frame->camera.scale;
        p.resolution = resolution;        // mapper_resolution (include/mapper.h:24)
        p.max_distance = max_distance;    // mapper_max_distance (:25)
        p.roiz = 65535.0 / p.scale;       // the mapping context never converts disparities: only the 16-bit depth bound applies
        p.num_disparities = 16;           // ... nor runs SGBM: keep its cost buffers minimal
        p.max_width = w > 32 ? w : 32; p.max_height = h; p.max_batch = 1;
        ssm_dropin::check(ssm_create(&p, 0, &st.ctx));
        st.w = p.max_width; st.h = h;
    }
    return st;
}

void pose_of(const RGBDFrame::Ptr& frame, double T[16])
{
    const Eigen::Isometry3d iso = frame->getTransform();   // under frame->mutexT (include/rgbdframe.h:116-120)
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T[4 * i + j] = iso(i, j);
}

void need_dense(const RGBDFrame::Ptr& f)
{
    const int w = f->depth.cols, h = f->depth.rows;
    if (f->depth.empty() || f->depth.type() != CV_16UC1 || f->semantic.type() != CV_8UC3 || f->rgb.type() != CV_8UC3 || f->semantic.rows != h ||
        f->semantic.cols != w || f->rgb.rows != h || f->rgb.cols != w || f->depth.step != (size_t)w * 2 || f->semantic.step != (size_t)w * 3 ||
        f->rgb.step != (size_t)w * 3)
        throw std::runtime_error("Mapper: depth (16UC1), semantic and rgb (8UC3) must be continuous and of equal size");
}

// the fused map as the pcl cloud the viewer / PCD writer take
Mapper::PointCloud::Ptr export_map(ssm_ctx* ctx)
{
    uint64_t n = 0;
    ssm_dropin::check(ssm_map_size(ctx, &n));
    std::vector<float> xyz(3 * n);
    std::vector<uint32_t> rgba(n);
    ssm_voxel_export e = {};
    e.xyz = xyz.data();
    e.rgba = rgba.data();
    uint64_t got = 0;
    ssm_dropin::check(ssm_map_export(ctx, &e, n, /*sorted: pcl::VoxelGrid's output order*/ 1, &got));
    Mapper::PointCloud::Ptr cloud(new Mapper::PointCloud());
    cloud->points.resize(got);
    for (uint64_t i = 0; i < got; ++i) {
        Mapper::PointT& p = cloud->points[i];
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        p.rgba = rgba[i];   // 0x00RRGGBB: alpha 0 as pcl::VoxelGrid leaves it
    }
    cloud->width = (uint32_t)got; cloud->height = 1; cloud->is_dense = true;
    return cloud;
}

}  // namespace

void Mapper::semantic_motion_fuse(const RGBDFrame::Ptr& frame)
{
    need_dense(frame);
    MapperState& st = state_of(this, frame, resolution, max_distance);
    moving_mask.create(frame->semantic.size(), CV_8UC1);
    ssm_dropin::check(ssm_semantic_motion_fuse(st.ctx, frame->semantic.data, frame->semantic.cols, frame->semantic.rows, frame->semantic.step,
                                               moving_mask.data, moving_mask.step));
}

Mapper::PointCloud::Ptr Mapper::generatePointCloud(const RGBDFrame::Ptr& frame)
{
    semantic_motion_fuse(frame);   // src/mapper.cpp:14 (the mask is a member the reference keeps up to date)
    MapperState& st = state_of(this, frame, resolution, max_distance);
    const int w = frame->depth.cols, h = frame->depth.rows;
    double T[16];
    pose_of(frame, T);
    std::vector<float> xyz((size_t)w * h * 3);
    std::vector<uint32_t> rgba((size_t)w * h);
    int n = 0;
    ssm_dropin::check(ssm_generate_point_cloud(st.ctx, (const uint16_t*)frame->depth.data, frame->semantic.data, frame->rgb.data, w, h, T, xyz.data(),
                                               rgba.data(), nullptr, w * h, &n));
    PointCloud::Ptr tmp(new PointCloud());
    tmp->points.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        PointT& p = tmp->points[i];
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        p.rgba = rgba[i];
    }
    tmp->width = (uint32_t)n; tmp->height = 1;
    tmp->is_dense = false;         // src/mapper.cpp:92
    return tmp;
}

void Mapper::viewer()
{
    pcl::visualization::CloudViewer viewer("viewer");
    int cntGlobalUpdate = 0;
    MapperState* st = nullptr;
    PointCloud::Ptr globalMap(new PointCloud);
    try {
        while (shutdownFlag == false) {
            boost::timer timer;
            const size_t n_kf = poseGraph.keyframes.size();
            if (n_kf <= (size_t)this->keyframe_size) {
                usleep(1000);      // src/mapper.cpp:116
                continue;
            }
            std::vector<RGBDFrame::Ptr> kfs(poseGraph.keyframes.begin(), poseGraph.keyframes.begin() + n_kf);
            st = &state_of(this, kfs[0], resolution, max_distance);
            auto cloud_of = [&](const RGBDFrame::Ptr& f) {
                need_dense(f);
                double T[16];
                pose_of(f, T);
                auto it = st->cloud_id.find(f.get());
                int id;
                if (it == st->cloud_id.end()) {
                    ssm_dropin::check(ssm_keyframe_add(st->ctx, (const uint16_t*)f->depth.data, f->semantic.data, f->rgb.data, f->depth.cols,
                                                       f->depth.rows, T, &id));
                    st->cloud_id[f.get()] = id;
                } else {
                    id = it->second;
                    ssm_dropin::check(ssm_keyframe_set_pose(st->ctx, id, T));   // the pose graph may have rewritten it
                }
                return id;
            };
            std::vector<int> ids;
            if (cntGlobalUpdate % 15 == 0) {      // src/mapper.cpp:121-131: redraw everything from the current poses
                cout << "redrawing frames" << endl;
                for (size_t i = 0; i < kfs.size(); ++i) ids.push_back(cloud_of(kfs[i]));
                ssm_dropin::check(ssm_map_redraw(st->ctx, ids.data(), (int)ids.size()));
            } else {                              // :132-149: the keyframes that arrived since the last update
                for (size_t i = (size_t)this->keyframe_size; i < kfs.size(); ++i) ids.push_back(cloud_of(kfs[i]));
                ssm_dropin::check(ssm_map_integrate_keyframes(st->ctx, ids.data(), (int)ids.size()));
            }
            cntGlobalUpdate++;
            keyframe_size = (int)kfs.size();
            globalMap = export_map(st->ctx);      // voxel.filter(*tmp); globalMap->swap(*tmp)   (:154-158)
            viewer.showCloud(globalMap);          // :159
            cout << "points in global map: " << globalMap->points.size() << endl;
            cout << "Mapping cost time: " << timer.elapsed() * 1000.0 << "ms" << endl;
        }
        if (poseGraph.shutDownFlag == true && st) {   // src/mapper.cpp:165-170 (the reference hard-codes its own path there)
            ssm_dropin::check(ssm_map_save_pcd(st->ctx, "map.pcd"));
            cout << "Map saved!" << endl;
        }
    } catch (const std::exception& e) {
        // the body runs on its own std::thread: an escaping exception would end the process through std::terminate
        cerr << "Mapper::viewer stopped: " << e.what() << endl;
    }
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_state.find(this);
    if (it != g_state.end()) {
        if (it->second.ctx) ssm_destroy(it->second.ctx);
        g_state.erase(it);
    }
}

void Mapper::SaveMap()
{
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_state.find(this);
    if (it != g_state.end() && it->second.ctx) ssm_dropin::check(ssm_map_save_pcd(it->second.ctx, "map.pcd"));
}
