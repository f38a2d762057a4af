// sgbm_wta.cuh -- decoding of the 16-byte winner-take-all records written by k_hrev (sgbm_hsweep2.cu): shared by the
// stand-alone finalize kernel and the fused selection kernel (sgbm_select.cu).
#pragma once

#include "ssm_internal.cuh"

namespace ssm {

// C-truncating num / den for den > 0 and a quotient of a few bits (here |num / den| <= 9): reciprocal estimate, which is
// within one of the answer, plus an integer fix-up -- instead of the generic 32-bit division sequence (~35 instructions)
__device__ __forceinline__ int div_trunc_small(int num, int den)
{
    int q = __float2int_rz(__fdividef((float)num, (float)den));
    const int r = num - q * den;
    if (num >= 0) q += r < 0 ? -1 : (r >= den ? 1 : 0);
    else q += r > 0 ? 1 : (r <= -den ? -1 : 0);
    return q;
}

// One record -> the pixel's raw disparity (x16, sub-pixel refined; kInvalidDisp when the uniqueness test rejected it),
// SURVEY.md App. A-5.  Record: x = minS | best << 16 | reject << 31, then the winner lane's packed costs with the two values
// across its lane borders: half-words {up, v0, .., v(2NR-1), down}, the winner is element q + 1.
//   NR <= 2: 16 bytes  {x, v0|v1, v2|v3 (NR = 2), up|down}
//   NR == 4: 32 bytes  {x, v0|v1, v2|v3, v4|v5} {v6|v7, up|down, -, -}
template <int NR> struct WtaRec { static constexpr int kQuads = NR <= 2 ? 1 : 2; };   // uint4s per record
template <int NR>
__device__ __forceinline__ int wta2_decode(const uint4& r, const uint4& r2, int D, int& minS, int& best, bool& valid)
{
    minS = (int)(r.x & 0xffffu);
    best = (int)((r.x >> 16) & 0x7fffu);
    valid = !(r.x >> 31);
    if (!valid) return kInvalidDisp;
    int d16 = best * kDispScale;
    if (best > 0 && best < D - 1) {
        uint32_t w[5];
        if (NR == 4) {
            w[0] = (r2.y & 0xffffu) | (r.y << 16);
            w[1] = (r.y >> 16) | (r.z << 16); w[2] = (r.z >> 16) | (r.w << 16); w[3] = (r.w >> 16) | (r2.x << 16);
            w[4] = (r2.x >> 16) | (r2.y & 0xffff0000u);
        } else {
            w[0] = (r.w & 0xffffu) | (r.y << 16);
            if (NR == 2) { w[1] = (r.y >> 16) | (r.z << 16); w[2] = (r.z >> 16) | (r.w & 0xffff0000u); }
            else { w[1] = (r.y >> 16) | (r.w & 0xffff0000u); w[2] = 0u; }
            w[3] = 0u; w[4] = 0u;
        }
        const int q = best & (2 * NR - 1);
        auto elem = [&](int i) {
            const uint32_t v = i < 2 ? w[0] : (i < 4 ? w[1] : (i < 6 ? w[2] : (i < 8 ? w[3] : w[4])));
            return (int)((i & 1) ? (v >> 16) : (v & 0xffffu));
        };
        const int sm = elem(q), sp = elem(q + 2);
        const int denom2 = max(sm + sp - 2 * minS, 1);
        d16 += div_trunc_small((sm - sp) * kDispScale + denom2, denom2 * 2);   // sm, sp >= minS: |quotient| <= 8
    }
    return d16;
}

}  // namespace ssm
