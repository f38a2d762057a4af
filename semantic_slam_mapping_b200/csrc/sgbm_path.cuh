// sgbm_path.cuh -- the SGM path recurrence on packed u16x2 lanes, shared by the aggregation kernels.
#pragma once

#include "ssm_internal.cuh"

namespace ssm {

// TMA-issued L2 prefetch of a contiguous global range (16-byte aligned, size a multiple of 16): one instruction per
// range, no registers held while the data travels.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

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

// the same with the evict-first ("streaming") cache hint: data that is read or written exactly once by the kernel and is far too
// large to be found in L2 by the next kernel either
template <int NR>
__device__ __forceinline__ void load_words_cs(const void* p, uint32_t (&w)[NR])
{
    if constexpr (NR == 1) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(w[0]) : "l"(p));
    else if constexpr (NR == 2) asm volatile("ld.global.cs.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
    else {
#pragma unroll
        for (int q = 0; q < NR / 4; ++q)
            asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * q]), "=r"(w[4 * q + 1]), "=r"(w[4 * q + 2]), "=r"(w[4 * q + 3])
                         : "l"(static_cast<const char*>(p) + 16 * q));
    }
}
template <int NR>
__device__ __forceinline__ void store_words_cs(void* p, const uint32_t (&w)[NR])
{
    if constexpr (NR == 1) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
    else if constexpr (NR == 2) asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
    else {
#pragma unroll
        for (int q = 0; q < NR / 4; ++q)
            asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(static_cast<char*>(p) + 16 * q), "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]),
                         "r"(w[4 * q + 3]) : "memory");
    }
}

// shared-memory accesses through 32-bit shared-window addresses (no generic-to-shared conversion per access)
template <int NR>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t (&w)[NR])
{
    if constexpr (NR == 1) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(addr));
    else if constexpr (NR == 2) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
    else {
#pragma unroll
        for (int q = 0; q < NR / 4; ++q)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * q]), "=r"(w[4 * q + 1]), "=r"(w[4 * q + 2]), "=r"(w[4 * q + 3]) : "r"(addr + 16 * q));
    }
}
template <int NR>
__device__ __forceinline__ void sts_words(uint32_t addr, const uint32_t (&w)[NR])
{
    if constexpr (NR == 1) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(w[0]) : "memory");
    else if constexpr (NR == 2) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(w[0]), "r"(w[1]) : "memory");
    else {
#pragma unroll
        for (int q = 0; q < NR / 4; ++q)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 16 * q), "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]) : "memory");
    }
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// make a loop-invariant value opaque so that it stays in its register instead of being re-derived every iteration
__device__ __forceinline__ void keep(uint32_t& x) { asm volatile("" : "+r"(x)); }

// Per-lane constants of path_step.  `one` is the integer 1 passed in as a kernel argument: the compiler cannot
// fold a multiplication by it, so `a * one + b` is emitted as IMAD, which issues on the FMA pipe.  The aggregation
// kernels are bound by the ALU pipe (IADD3 / LOP3 / SHF / SEL / VIMNMX / VIADDMNMX issue one warp instruction per
// two cycles per scheduler); every add, select and shift that can be phrased as a multiply-add moves there.
struct PathLane {
    uint32_t one;      // 1
    uint32_t sh16;     // 65536
    uint32_t keep0;    // lane 0: 0, else 1          -- d = -1 reads "+inf": up = shfl * keep0 + add0
    uint32_t add0;     // lane 0: kBig | P1 << 16 (the upper half carries the +P1 the shuffled word would have brought), else 0
    uint32_t mul31;    // lane 31: 0, else 65536     -- d past the last lane reads "+inf": s = dn * mul31 + (hi + add31)
    uint32_t add31;    // lane 31: kBig << 16, else 0
    uint32_t seg_lo;   // LANES == 16 only: all-ones in lanes 16..31 (masks this lane out of the lower segment's REDUX)
    uint32_t seg_hi;   //                   all-ones in lanes 0..15
};
// LANES = lanes that share one pixel's disparity range: 32 (a warp per pixel) or 16 (two pixels per warp, one per
// half-warp: the per-step overhead -- shuffles, border fix-ups, reduction, addressing -- is paid once for both)
template <int LANES = 32>
__device__ __forceinline__ PathLane make_path_lane(int lane, uint32_t one, uint32_t P1)
{
    const int sl = lane & (LANES - 1);
    PathLane p;
    p.one = one;
    p.sh16 = one << 16;
    p.keep0 = sl == 0 ? 0u : one;
    p.add0 = sl == 0 ? (kBig | (P1 << 16)) : 0u;
    p.mul31 = sl == LANES - 1 ? 0u : one << 16;
    p.add31 = sl == LANES - 1 ? kBig << 16 : 0u;
    p.seg_lo = lane >= 16 ? 0xffffffffu : 0u;
    p.seg_hi = lane >= 16 ? 0u : 0xffffffffu;
    keep(p.sh16); keep(p.keep0); keep(p.add0); keep(p.mul31); keep(p.add31);
    if (LANES == 16) { keep(p.seg_lo); keep(p.seg_hi); }
    return p;
}

// One step of the recurrence for this lane's 2*NR disparities.
//   mw   packed (m, m): the minimum of the incoming state over the pixel's disparities, in both 16-bit halves
//   P1w, P2w   packed penalties
// Returns the packed minimum of the new state.  The NR+1 distinct "shifted by one disparity, plus P1" words
// s[k] = (w[k-1].hi, w[k].lo) + P1 are built once each: hp = (w >> 16) + P1w carries P1 in BOTH halves (one LEA.HI), so
// s[k] = w[k] * 65536 + hp[k-1] is one IMAD; then min(L[d], m + P2, L[d-1] + P1) is one VIMNMX3 and L[d+1] + P1 one
// VIMNMX.  The lane minimum stays packed (PRMT + VIMNMX.U16x2) so the REDUX result is directly the next step's mw.
template <int NR, int LANES = 32>
__device__ __forceinline__ uint32_t path_step(uint32_t (&L)[NR], const uint32_t (&Cw)[NR], uint32_t mw, uint32_t P1w,
                                              uint32_t P2w, const PathLane& pl)
{
    uint32_t hp[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) hp[r] = (L[r] >> 16) + P1w;   // (IMAD.HI would move this to the FMA pipe, but it issues at a lower rate: measured slower)
    const uint32_t up_hp = __shfl_up_sync(0xffffffffu, hp[NR - 1], 1, LANES) * pl.keep0 + pl.add0;
    const uint32_t dn = __shfl_down_sync(0xffffffffu, L[0], 1, LANES);
    uint32_t sft[NR + 1];
    sft[0] = L[0] * pl.sh16 + up_hp;
#pragma unroll
    for (int r = 1; r < NR; ++r) sft[r] = L[r] * pl.sh16 + hp[r - 1];
    sft[NR] = dn * pl.mul31 + (hp[NR - 1] * pl.one + pl.add31);
    const uint32_t mP2w = mw * pl.one + P2w;
    const uint32_t nmw = 0u - mw;
    uint32_t mn = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const uint32_t cm = Cw[r] * pl.one + nmw;                // C - m, off the critical path
        uint32_t t = __vimin3_u16x2(L[r], mP2w, sft[r]);         // min(L[d], m + P2, L[d-1] + P1)
        t = __vminu2(t, sft[r + 1]);                             // min(., L[d+1] + P1)
        L[r] = t * pl.one + cm;                                  // every lane of t >= m: plain 32-bit arithmetic is exact
        mn = r == 0 ? L[r] : __vminu2(mn, L[r]);
    }
    mn = __vminu2(mn, __byte_perm(mn, 0, 0x1032));               // both halves = the lane minimum
    if constexpr (LANES == 32) {
        return __reduce_min_sync(0xffffffffu, mn);
    } else {
        // one REDUX per half-warp segment: the other segment's lanes contribute all-ones
        const uint32_t a = __reduce_min_sync(0xffffffffu, mn | pl.seg_lo);
        const uint32_t b = __reduce_min_sync(0xffffffffu, mn | pl.seg_hi);
        return (a & pl.seg_hi) | (b & pl.seg_lo);
    }
}

}  // namespace ssm
