// host_driver.cpp -- exercises the C++ host layer (stereo.hpp / mapper.hpp) the way the reference's
// exp_mapping.cpp drives stereo.h / mapper.h: frames in, disparity per frame, keyframes pushed to the
// keyframe list while the Mapper's viewer thread fuses them.  Reads / writes flat binary files so that
// tests/test_gpu_host_cpp.py can compare against the CPU oracle.
//   usage: host_driver <in.bin> <out.bin> [pcd path]
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../../semantic_slam_mapping_b200/host/mapper.hpp"
#include "../../semantic_slam_mapping_b200/host/stereo.hpp"

using namespace ssm_host;

template <typename T>
static void rd(FILE* f, T* p, size_t n)
{
    if (fread(p, sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    FILE* fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    int hdr[4];
    double leaf;
    rd(fi, hdr, 4);
    rd(fi, &leaf, 1);
    const int n = hdr[0], H = hdr[1], W = hdr[2], D = hdr[3];
    try {
        StereoConfig sc;
        sc.num_disparities = D; sc.max_width = W; sc.max_height = H;
        setStereoConfig(sc);
        MapperConfig mc;
        mc.mapper_resolution = leaf; mc.max_width = W; mc.max_height = H; mc.map_capacity = 1 << 18;
        mc.redraw_every = 3;   // exercise the redraw branch a few times
        if (argc > 3) mc.save_path = argv[3];
        KeyframeSource graph;
        Mapper mapper(mc, graph);   // spawns the viewer thread, like the reference constructor
        for (int i = 0; i < n; ++i) {
            auto fr = std::make_shared<Frame>();
            fr->id = i;
            ImageU8 L(H, W), R(H, W);
            fr->semantic.create(H, W);
            fr->rgb.create(H, W);
            rd(fi, L.data, (size_t)H * W);
            rd(fi, R.data, (size_t)H * W);
            rd(fi, fr->semantic.data, (size_t)H * W * 3);
            rd(fi, fr->rgb.data, (size_t)H * W * 3);
            std::array<double, 16> T;
            rd(fi, T.data(), 16);
            calDisparity_SGBM(L, R, fr->disparity);        // rgbdframe.cpp:82
            disparityToDepth(fr->disparity, fr->depth);    // rgbdframe.cpp:85-116
            fr->setTransform(T);                           // track.cpp:128
            fwrite(fr->disparity.data, 2, (size_t)H * W, fo);
            {
                std::lock_guard<std::mutex> lk(graph.keyframes_mutex);
                graph.keyframes.push_back(fr);             // pose_graph.cpp:11-77
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(3));
        }
        // wait until the viewer thread has consumed every keyframe, then stop it
        for (int spin = 0; spin < 20000 && mapper.consumed() < (size_t)n; ++spin)
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        if (mapper.consumed() != (size_t)n) { std::fprintf(stderr, "viewer thread did not consume all keyframes\n"); return 4; }
        graph.shutDownFlag = true;
        mapper.shutdown();
        // one more cloud through the protected worker, for the per-point parity check
        auto pc = mapper.generatePointCloud(graph.keyframes[0]);
        VoxelCloud v = mapper.exportMap(true);
        const long long nv = (long long)v.size(), np = (long long)pc->size();
        fwrite(&nv, 8, 1, fo);
        fwrite(v.ijk.data(), 4, v.ijk.size(), fo);
        fwrite(v.xyz.data(), 4, v.xyz.size(), fo);
        fwrite(v.rgba.data(), 4, v.rgba.size(), fo);
        fwrite(v.count.data(), 4, v.count.size(), fo);
        fwrite(v.votes.data(), 4, v.votes.size(), fo);
        fwrite(v.label.data(), 1, v.label.size(), fo);
        fwrite(&np, 8, 1, fo);
        for (const auto& p : *pc) { fwrite(&p.x, 4, 3, fo); fwrite(&p.rgba, 4, 1, fo); }
        releaseStereoContext();
        // error behaviour: mismatched sizes throw, like the reference's cv::Exception
        bool threw = false;
        try { ImageU8 a(4, 40), b(5, 40); ImageS16 d; calDisparity_SGBM(a, b, d); } catch (const Error&) { threw = true; }
        if (!threw) { std::fprintf(stderr, "no exception on mismatched sizes\n"); return 3; }
    } catch (const Error& e) {
        std::fprintf(stderr, "ssm_host::Error %d: %s\n", e.code, e.what());
        return 1;
    }
    fclose(fi);
    fclose(fo);
    return 0;
}
