// sgbm_path.cuh -- the SGM path recurrence on packed u16x2 lanes, shared by the aggregation kernels.
#pragma once

#include "ssm_internal.cuh"

namespace ssm {

template <int NR> struct WordVec;
template <> struct WordVec<1> { using T = uint32_t; };
template <> struct WordVec<2> { using T = uint2; };
template <> struct WordVec<4> { using T = uint4; };

template <int NR>
__device__ __forceinline__ void load_words(const void* p, uint32_t (&w)[NR])
{
    if constexpr (NR == 8) {
        const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
        const typename WordVec<NR>::T v = *reinterpret_cast<const typename WordVec<NR>::T*>(p);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int i = 0; i < NR; ++i) w[i] = s[i];
    }
}
template <int NR>
__device__ __forceinline__ void store_words(void* p, const uint32_t (&w)[NR])
{
    if constexpr (NR == 8) {
        reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
        typename WordVec<NR>::T v;
        uint32_t* s = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int i = 0; i < NR; ++i) s[i] = w[i];
        *reinterpret_cast<typename WordVec<NR>::T*>(p) = v;
    }
}

// One step of the recurrence for this lane's 2*NR disparities.  Returns the new warp-wide minimum.
template <int NR>
__device__ __forceinline__ uint32_t path_step(uint32_t (&L)[NR], const uint32_t (&Cw)[NR], uint32_t m, uint32_t P1w,
                                              uint32_t P2, int lane)
{
    uint32_t up = __shfl_up_sync(0xffffffffu, L[NR - 1], 1);
    uint32_t dn = __shfl_down_sync(0xffffffffu, L[0], 1);
    if (lane == 0) up = kBigW;
    if (lane == 31) dn = kBigW;
    const uint32_t mw = m * 0x10001u;
    const uint32_t mP2w = (m + P2) * 0x10001u;
    uint32_t Ln[NR];
    uint32_t mn = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const uint32_t prev = r == 0 ? up : L[r - 1];
        const uint32_t next = r == NR - 1 ? dn : L[r + 1];
        const uint32_t lm1 = __funnelshift_l(prev, L[r], 16);   // (L[d-1], L[d]) for the pair (d, d+1)
        const uint32_t lp1 = __funnelshift_r(L[r], next, 16);   // (L[d+1], L[d+2])
        uint32_t t = __viaddmin_u16x2(lm1, P1w, L[r]);
        t = __viaddmin_u16x2(lp1, P1w, t);
        t = __vminu2(t, mP2w);
        Ln[r] = t - mw + Cw[r];                                 // every lane of t >= m: plain 32-bit arithmetic is exact
        mn = r == 0 ? Ln[r] : __vminu2(mn, Ln[r]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) L[r] = Ln[r];
    const uint32_t lane_min = min(mn & 0xffffu, mn >> 16);
    return __reduce_min_sync(0xffffffffu, lane_min);
}

}  // namespace ssm
