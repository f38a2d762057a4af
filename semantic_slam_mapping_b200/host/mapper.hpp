// mapper.hpp -- host-side mirror of rgbd_tutor::Mapper (/root/reference include/mapper.h:15-70,
// src/mapper.cpp).  Same public surface -- constructor from the configuration + the keyframe owner,
// shutdown(), viewer() (the thread body), SaveMap() -- and the same protected workers
// generatePointCloud() / semantic_motion_fuse().  What differs is where the arithmetic runs: every per-pixel
// and per-point operation is a libssm.so call (GPU); the global map is the device-resident voxel hash instead
// of a pcl::PointCloud that is re-filtered on every update.
//
// Semantics kept from the reference loop (mapper.cpp:109-163):
//   * waits until the keyframe list grew (poll, 1 ms sleep), never blocks the producer;
//   * every 15th update redraws the whole map from the keyframes' CURRENT poses (the pose graph rewrites
//     them, pose_graph.cpp:253-260) -- here: ssm_map_clear + re-integration of every keyframe;
//   * otherwise integrates the keyframes that arrived since the last update.
// Canonical differences: each keyframe contributes exactly once (the reference
// re-adds the last five keyframes on every update and redraws every second keyframe only); fusion is one
// VoxelGrid pass over the union of clouds.  The PCL viewer window is out of scope.
#pragma once

#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

#include "frame.hpp"

namespace ssm_host {

struct MapperConfig {
    double mapper_resolution = 0.1;      // parameters.txt:97
    double mapper_max_distance = 40.0;   // parameters.txt:98
    Camera camera;
    int num_labels = 12;                 // 12-class SegNet palette of ssm_default_params
    int device = 0;
    int max_width = 1241, max_height = 376;
    uint64_t map_capacity = 1ull << 22;
    int redraw_every = 15;               // mapper.cpp:121
    int redraw_stride = 1;               // keyframes taken on a full redraw: every one (the reference's loop takes every 2nd, mapper.cpp:126)
    std::string save_path;               // mapper.cpp:168 hard-codes a path; empty = do not save at shutdown
};

struct VoxelCloud {                      // what viewer.showCloud() / PCDWriter received in the reference
    std::vector<int32_t> ijk;
    std::vector<float> xyz;
    std::vector<uint32_t> rgba, count, votes;
    std::vector<uint8_t> label;
    size_t size() const { return rgba.size(); }
};

class Mapper {
public:
    typedef PointXYZRGBL PointT;
    typedef ssm_host::PointCloud PointCloud;

    Mapper(const MapperConfig& para, KeyframeSource& graph, bool start_thread = true);
    ~Mapper();

    void shutdown();
    void viewer();       // thread body; library errors stop the updates and are reported by failed() / lastError()
    bool failed() const;
    std::string lastError() const;
    void SaveMap();      // writes the fused map as binary PCD to config.save_path (reference body is empty)

    // inspection (the reference prints "points in global map", mapper.cpp:161)
    uint64_t mapSize();
    VoxelCloud exportMap(bool pcl_order = true);
    int updates() const { return cntGlobalUpdate; }
    size_t consumed() const { return keyframe_size; }   // keyframes fused so far
    ssm_ctx* context() { return ctx; }

    // exposed for tests; protected in the reference
    std::shared_ptr<PointCloud> generatePointCloud(const Frame::Ptr& frame);
    void semantic_motion_fuse(const Frame::Ptr& frame);
    ImageU8 moving_mask;

protected:
    int cloudOf(const Frame::Ptr& frame);   // device-resident camera-space cloud of a keyframe (built on first use)
    void viewerLoop();
    std::atomic<bool> failedFlag{false};
    mutable std::mutex errorMutex;
    std::string errorText;

    std::shared_ptr<std::thread> viewerThread;
    MapperConfig config;
    KeyframeSource& poseGraph;
    ssm_ctx* ctx = nullptr;
    std::atomic<size_t> keyframe_size{0};
    std::atomic<int> cntGlobalUpdate{0};
    std::atomic<bool> shutdownFlag{false};
};

}  // namespace ssm_host
