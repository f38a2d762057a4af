// sgbm_aggregate.cu -- SGM path aggregation, sm_100a.
//
// Replaces the 5-direction single-pass aggregation inside cv::StereoSGBM (called from /root/reference
// src/stereo.cpp:30).  One warp walks one path; the D-wide state L_r(p, .) lives in registers as packed
// u16x2 lanes (NR words = 2*NR disparities per lane), the d+-1 neighbours come from warp shuffles, the
// min over the disparity range from one CREDUX (__reduce_min_sync), and the recurrence itself is
// VIADDMNMX.U16x2 / VIMNMX.U16x2:
//      L(p,d) = C(p,d) + min(L(p-r,d), L(p-r,d-1)+P1, L(p-r,d+1)+P1, m+P2) - m,   m = min_k L(p-r,k)
// Out-of-image predecessors are the all-zero state; d = -1 and d = D read "+inf" (kBig).  Lanes beyond D
// carry a self-maintaining pad value (>= kBig - P2), so no per-step select is needed.
// S = min(sum_r L_r, 32767) is accumulated with saturating u16x2 adds (all L >= 0, so order is irrelevant).
#include "ssm_internal.cuh"
#include "sgbm_path.cuh"

namespace ssm {

struct AggrArgs {
    const int16_t* C;
    uint16_t* S;
    int W1, H, D, Dl;   // D valid disparities, Dl disparities per column in the layout
    int P1, P2;
    int dir;       // 0: ->  1: down-right  2: down  3: down-left  4: <-
    int first;     // 1: S = L, 0: S = sat(S + L)
    int npaths;    // per frame
    int total;     // npaths * batch
    uint32_t one;  // 1 (opaque to the compiler, see PathLane)
};

template <int NR>
__global__ void __launch_bounds__(256) k_aggr_path(AggrArgs a)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= a.total) return;
    const int b = gw / a.npaths, path = gw - b * a.npaths;
    int x, y, dx, dy, len;
    switch (a.dir) {
        case 0: x = 0; y = path; dx = 1; dy = 0; len = a.W1; break;
        case 4: x = a.W1 - 1; y = path; dx = -1; dy = 0; len = a.W1; break;
        case 2: x = path; y = 0; dx = 0; dy = 1; len = a.H; break;
        case 1: {
            const int c = path - (a.H - 1);
            x = c >= 0 ? c : 0; y = c >= 0 ? 0 : -c; dx = 1; dy = 1;
            len = min(a.W1 - x, a.H - y);
            break;
        }
        default: {
            const int c = path;
            x = c <= a.W1 - 1 ? c : a.W1 - 1; y = c <= a.W1 - 1 ? 0 : c - (a.W1 - 1); dx = -1; dy = 1;
            len = min(x + 1, a.H - y);
            break;
        }
    }
    const int d0 = lane * 2 * NR;
    const bool active = d0 < a.D;
    const PathLane pl = make_path_lane(lane, a.one, (uint32_t)a.P1);
    const uint32_t P1w = (uint32_t)a.P1 * 0x10001u, P2w = (uint32_t)a.P2 * 0x10001u;
    const uint32_t padC = (kBig - (uint32_t)a.P2) * 0x10001u;
    const size_t frame = (size_t)b * a.H * a.W1;
    const ptrdiff_t stride = ((ptrdiff_t)dy * a.W1 + dx) * a.Dl;
    size_t off = (frame + (size_t)y * a.W1 + x) * a.Dl + d0;

    uint32_t L[NR], Cw[NR], Cn[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { L[r] = 0u; Cw[r] = padC; Cn[r] = padC; }
    uint32_t m = 0u;
    if (active) load_words<NR>(a.C + off, Cw);
    for (int t = 0; t < len; ++t) {
        if (active && t + 1 < len) load_words<NR>(a.C + off + stride, Cn);   // prefetch the next pixel of the path
        m = path_step<NR>(L, Cw, m, P1w, P2w, pl);
        if (active) {
            if (a.first) {
                store_words<NR>(a.S + off, L);
            } else {
                uint32_t Sw[NR];
                load_words<NR>(a.S + off, Sw);
#pragma unroll
                for (int r = 0; r < NR; ++r) Sw[r] = __viaddmin_u16x2(Sw[r], L[r], kSatW);
                store_words<NR>(a.S + off, Sw);
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) Cw[r] = Cn[r];
        off += stride;
    }
}

int launch_aggregate_horizontal(ssm_ctx* c, int B, cudaStream_t s)
{
    return hsweep2_supported(c) ? launch_hsweep2(c, B, s) : launch_hsweep(c, B, s);
}

int launch_aggregate_vertical(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    AggrArgs a;
    a.C = c->d_C; a.S = c->d_S; a.W1 = p.W1; a.H = p.H; a.D = p.D; a.Dl = p.Dl; a.P1 = p.P1; a.P2 = p.P2; a.one = 1u;
    const int nr = p.Dl <= 64 ? 1 : (p.Dl <= 128 ? 2 : (p.Dl <= 256 ? 4 : 8));
    // preferred: all three top-down directions in one cluster launch (sgbm_vertical.cu)
    bool done = false;
    int rc = launch_vertical(c, B, s, &done);
    if (rc) return rc;
    if (done) return SSM_OK;
    // fallback: the three top-down directions walk one warp per path; the horizontal pair + WTA is k_hsweep
    const int order[3] = {2, 1, 3};
    for (int i = 0; i < 3; ++i) {
        a.dir = order[i];
        a.first = i == 0;
        a.npaths = (a.dir == 0 || a.dir == 4) ? p.H : (a.dir == 2 ? p.W1 : p.W1 + p.H - 1);
        a.total = a.npaths * B;
        const int wpb = 8;
        const unsigned grid = (unsigned)((a.total + wpb - 1) / wpb);
        switch (nr) {
            case 1: k_aggr_path<1><<<grid, wpb * 32, 0, s>>>(a); break;
            case 2: k_aggr_path<2><<<grid, wpb * 32, 0, s>>>(a); break;
            case 4: k_aggr_path<4><<<grid, wpb * 32, 0, s>>>(a); break;
            default: k_aggr_path<8><<<grid, wpb * 32, 0, s>>>(a); break;
        }
        SSM_LAUNCH_CHECK(c);
    }
    return SSM_OK;
}

}  // namespace ssm
