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
        // the tracker thread's dense cues on keyframe 0, in the order of UVDisparity::Process (uvdisparity.cpp:842-903)
        {
            const Frame::Ptr& f0 = graph.keyframes[0];
            ImageU8 grey(H, W), roi_mask, ground(H, W);
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) { grey.ptr(y)[x] = f0->rgb.ptr(y)[3 * x]; ground.ptr(y)[x] = y >= H / 2 ? 255 : 0; }
            ImageXYZ10 xyz;
            ROI3D roi(15.0, 1.5, 35.0);
            const Camera cam;
            triangulate10D(grey, f0->disparity, xyz, cam.fx, cam.cx, cam.cy, cam.baseline, roi);   // track.cpp:67-71
            UVDisparity uv;
            uv.calVDisparity(f0->disparity, xyz);
            const double pitch1 = 0.02, pitch2 = 0.0;
            correct3DPoints(xyz, roi, pitch1, pitch2);
            setImageROI(xyz, roi_mask);
            uv.calUDisparity(f0->disparity, xyz, roi_mask, ground);
            const int dims[2] = {uv.v_dis_int.cols, uv.u_dis_int.rows};
            fwrite(dims, 4, 2, fo);
            fwrite(xyz.data, 4, (size_t)H * W * 10, fo);
            fwrite(roi_mask.data, 1, (size_t)H * W, fo);
            fwrite(uv.v_dis_int.data, 4, (size_t)H * dims[0], fo);
            fwrite(uv.u_dis_.data, 1, (size_t)dims[1] * W, fo);
        }
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
