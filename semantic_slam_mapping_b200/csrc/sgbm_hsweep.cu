// sgbm_hsweep.cu -- the two horizontal SGM paths fused with the winner-take-all, sm_100a.
//
// cv::StereoSGBM (called from /root/reference src/stereo.cpp:30) finishes every image row with a left-to-right
// path, then a right-to-left path during which it forms S = sat(sum of the five L_r), picks the first minimum,
// applies the uniqueness test and records what the sub-pixel fit and the right-image disparity need.  Here one
// warp owns one row of one frame:
//   forward sweep  x' = 0 .. W1-1 : L0 in registers, S_f = sat(S_v + L0) written in place over S_v
//   reverse sweep  x' = W1-1 .. 0 : L0r in registers, S = sat(S_f + L0r) formed in registers only, then
//                  argmin (one CREDUX on (S<<16 | d) keys), uniqueness (masked second minimum against a
//                  precomputed threshold table), S[best-1], S[best+1] by shuffle -> one 64-bit record per pixel,
//                  flushed 32 pixels at a time as a coalesced 256-byte store.
// A per-pixel kernel (k_wta_finalize) then does the serial-unfriendly tail in parallel: sub-pixel interpolation
// with C-truncating division and the disp2 atomicMin.  Cost rows stream through registers with an 8-deep
// (2 x 4 steps) software prefetch; every load is a coalesced 2*D-byte row segment.
#include "sgbm_path.cuh"

namespace ssm {

constexpr int kPF = 4;   // steps per prefetch group

template <int NR, bool REVERSE>
__device__ __forceinline__ void load_group(const uint16_t* __restrict__ Crow, const uint16_t* Srow, int g, int W1,
                                           int D, bool active, uint32_t padC, uint32_t (&Cg)[kPF][NR], uint32_t (&Sg)[kPF][NR])
{
#pragma unroll
    for (int j = 0; j < kPF; ++j) {
        const int t = g * kPF + j;
        const int x = REVERSE ? W1 - 1 - t : t;
        if (active && t < W1) {
            load_words<NR>(Crow + (size_t)x * D, Cg[j]);
            load_words<NR>(Srow + (size_t)x * D, Sg[j]);
        } else {
#pragma unroll
            for (int r = 0; r < NR; ++r) { Cg[j][r] = padC; Sg[j][r] = 0xffffffffu; }
        }
    }
}

// L2 prefetch (TMA bulk) of pixels [x_lo, x_hi) of this row of C and S: one lane, two instructions
constexpr int kAhead = 8;   // prefetch distance in groups of kPF pixels
__device__ __forceinline__ void prefetch_span(const uint16_t* Cbase, const uint16_t* Sbase, int x_lo, int x_hi, int D)
{
    const uint32_t bytes = (uint32_t)(x_hi - x_lo) * D * 2;
    l2_prefetch_bulk(Cbase + (size_t)x_lo * D, bytes);
    l2_prefetch_bulk(Sbase + (size_t)x_lo * D, bytes);
}

template <int NR>
__device__ __forceinline__ uint32_t pick(const uint32_t (&w)[NR], int idx)
{   // 16-bit element idx (0 .. 2*NR-1) of this lane's packed words
    uint32_t v = w[0];
#pragma unroll
    for (int r = 1; r < NR; ++r)
        if ((idx >> 1) == r) v = w[r];
    return (idx & 1) ? (v >> 16) : (v & 0xffffu);
}

template <int NR>
__global__ void __launch_bounds__(128) k_hsweep(const int16_t* __restrict__ C, uint16_t* S /* read and written: no restrict */, uint2* __restrict__ rec,
                                                const uint32_t* __restrict__ uniq_thr, int W1, int D, int P1, int P2, int nrows, uint32_t one, int pf)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    constexpr int LOG_DPL = NR == 1 ? 1 : (NR == 2 ? 2 : (NR == 4 ? 3 : 4));
    const int d0 = lane * 2 * NR;
    const bool active = d0 < D;
    const PathLane pl = make_path_lane(lane, one, (uint32_t)P1);
    const uint32_t P1w = (uint32_t)P1 * 0x10001u, P2w = (uint32_t)P2 * 0x10001u;
    const uint32_t padC = (kBig - (uint32_t)P2) * 0x10001u;
    const uint16_t* Crow = reinterpret_cast<const uint16_t*>(C) + (size_t)row * W1 * D + d0;
    uint16_t* Srow = S + (size_t)row * W1 * D + d0;
    const int ngroups = (W1 + kPF - 1) / kPF;
    const uint16_t* Cbase = Crow - d0;
    const uint16_t* Sbase = Srow - d0;

    uint32_t L[NR], Ca[kPF][NR], Sa[kPF][NR], Cb[kPF][NR], Sb[kPF][NR];

    // ---------------- forward: S_f = sat(S_v + L0) ----------------
    {
#pragma unroll
        for (int r = 0; r < NR; ++r) L[r] = 0u;
        uint32_t m = 0u;
        load_group<NR, false>(Crow, Srow, 0, W1, D, active, padC, Ca, Sa);
        if (lane == 0 && pf) prefetch_span(Cbase, Sbase, 0, min(kAhead * kPF, W1), D);
        for (int g = 0; g < ngroups; g += 2) {
            if (lane == 0 && pf) {   // the 8 pixels that will be loaded kAhead groups from now, into L2
                const int x0 = (g + kAhead) * kPF;
                if (x0 < W1) prefetch_span(Cbase, Sbase, x0, min(x0 + 2 * kPF, W1), D);
            }
            load_group<NR, false>(Crow, Srow, g + 1, W1, D, active, padC, Cb, Sb);
#pragma unroll
            for (int j = 0; j < kPF; ++j) {
                const int x = g * kPF + j;
                if (x < W1) {
                    m = path_step<NR>(L, Ca[j], m, P1w, P2w, pl);
                    if (active) {
                        uint32_t o[NR];
#pragma unroll
                        for (int r = 0; r < NR; ++r) o[r] = __viaddmin_u16x2(Sa[j][r], L[r], kSatW);
                        store_words<NR>(Srow + (size_t)x * D, o);
                    }
                }
            }
            load_group<NR, false>(Crow, Srow, g + 2, W1, D, active, padC, Ca, Sa);
#pragma unroll
            for (int j = 0; j < kPF; ++j) {
                const int x = (g + 1) * kPF + j;
                if (x < W1) {
                    m = path_step<NR>(L, Cb[j], m, P1w, P2w, pl);
                    if (active) {
                        uint32_t o[NR];
#pragma unroll
                        for (int r = 0; r < NR; ++r) o[r] = __viaddmin_u16x2(Sb[j][r], L[r], kSatW);
                        store_words<NR>(Srow + (size_t)x * D, o);
                    }
                }
            }
        }
    }
    // The reverse sweep re-reads S_f written above by this same warp; a warp-level fence orders the accesses.
    __syncwarp();

    // ---------------- reverse: S = sat(S_f + L0r), winner-take-all records ----------------
    uint2* rrow = rec + (size_t)row * W1;
    uint32_t rlo = 0u, rhi = 0u;
    auto wta = [&](const uint32_t (&Sf)[NR], int x) {
        uint32_t Sw[NR];
        uint32_t kmin = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            Sw[r] = active ? __viaddmin_u16x2(Sf[r], L[r], kSatW) : 0xffffffffu;
            const uint32_t klo = (Sw[r] << 16) | (uint32_t)(d0 + 2 * r);
            const uint32_t khi = (Sw[r] & 0xffff0000u) | (uint32_t)(d0 + 2 * r + 1);
            kmin = __vimin3_u32(kmin, klo, khi);   // smaller d wins ties: first minimum
        }
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        const int minS = (int)(kmin >> 16), best = (int)(kmin & 0xffffu);
        const uint32_t thr = __ldg(uniq_thr + minS);
        // smallest S outside the window [best-1, best+1]
        uint32_t mm = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int t = best - (d0 + 2 * r);
            const uint32_t mask = ((unsigned)(t + 1) <= 2u ? 0xffffu : 0u) | ((unsigned)t <= 2u ? 0xffff0000u : 0u);
            mm = __vminu2(mm, Sw[r] | mask);
        }
        uint32_t min2 = min(mm & 0xffffu, mm >> 16);
        min2 = __reduce_min_sync(0xffffffffu, min2);
        const uint32_t reject = min2 < thr ? 1u : 0u;
        const int im = best - 1, ip = best + 1;
        const uint32_t sm1 = __shfl_sync(0xffffffffu, pick<NR>(Sw, im & (2 * NR - 1)), (im >> LOG_DPL) & 31);
        const uint32_t sp1 = __shfl_sync(0xffffffffu, pick<NR>(Sw, ip & (2 * NR - 1)), (ip >> LOG_DPL) & 31);
        if (lane == (x & 31)) {
            rlo = (uint32_t)minS | (sm1 << 16);
            rhi = (sp1 & 0xffffu) | ((uint32_t)best << 16) | (reject << 31);
        }
        if ((x & 31) == 0) {
            const int xo = x + lane;
            if (xo < W1) rrow[xo] = make_uint2(rlo, rhi);
        }
    };
    {
#pragma unroll
        for (int r = 0; r < NR; ++r) L[r] = 0u;
        uint32_t m = 0u;
        load_group<NR, true>(Crow, Srow, 0, W1, D, active, padC, Ca, Sa);
        if (lane == 0 && pf) prefetch_span(Cbase, Sbase, max(W1 - kAhead * kPF, 0), W1, D);
        for (int g = 0; g < ngroups; g += 2) {
            if (lane == 0 && pf) {
                const int x1 = W1 - (g + kAhead) * kPF;        // pixels [x1 - 8, x1) are loaded kAhead groups from now
                if (x1 > 0) prefetch_span(Cbase, Sbase, max(x1 - 2 * kPF, 0), x1, D);
            }
            load_group<NR, true>(Crow, Srow, g + 1, W1, D, active, padC, Cb, Sb);
#pragma unroll
            for (int j = 0; j < kPF; ++j) {
                const int t = g * kPF + j;
                if (t < W1) {
                    m = path_step<NR>(L, Ca[j], m, P1w, P2w, pl);
                    wta(Sa[j], W1 - 1 - t);
                }
            }
            load_group<NR, true>(Crow, Srow, g + 2, W1, D, active, padC, Ca, Sa);
#pragma unroll
            for (int j = 0; j < kPF; ++j) {
                const int t = (g + 1) * kPF + j;
                if (t < W1) {
                    m = path_step<NR>(L, Cb[j], m, P1w, P2w, pl);
                    wta(Sb[j], W1 - 1 - t);
                }
            }
        }
    }
}

// sub-pixel fit, disp2 candidates, raw disparity: one thread per pixel of the valid region
__global__ void __launch_bounds__(256) k_wta_finalize(const uint2* __restrict__ rec, int16_t* __restrict__ disp_raw,
                                                      uint32_t* __restrict__ disp2key, int W, int D, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int W1 = W - D;
    const int xp = (int)(idx % W1);
    const size_t row = idx / W1;
    const uint2 r = rec[idx];
    const int minS = (int)(r.x & 0xffffu), sm = (int)(r.x >> 16), sp = (int)(r.y & 0xffffu);
    const int best = (int)((r.y >> 16) & 0x7fffu);
    const int x = xp + D;
    int out = kInvalidDisp;
    if (!(r.y >> 31)) {
        atomicMin(&disp2key[row * W + (x - best)], ((uint32_t)minS << 16) | (uint32_t)(0xffff - x));
        int d16 = best * kDispScale;
        if (best > 0 && best < D - 1) {
            const int denom2 = max(sm + sp - 2 * minS, 1);
            d16 += ((sm - sp) * kDispScale + denom2) / (denom2 * 2);
        }
        out = d16;
    }
    disp_raw[row * W + x] = (int16_t)out;
}

int launch_hsweep(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const int nrows = B * p.H;
    const int nr = p.D <= 64 ? 1 : (p.D <= 128 ? 2 : (p.D <= 256 ? 4 : 8));
    const int wpb = 4;
    const unsigned grid = (unsigned)((nrows + wpb - 1) / wpb);
    uint2* rec = reinterpret_cast<uint2*>(c->d_wta_rec);
    switch (nr) {
        case 1: k_hsweep<1><<<grid, wpb * 32, 0, s>>>(c->d_C, c->d_S, rec, c->d_uniq_thr, p.W1, p.D, p.P1, p.P2, nrows, 1u, c->tune[1]); break;
        case 2: k_hsweep<2><<<grid, wpb * 32, 0, s>>>(c->d_C, c->d_S, rec, c->d_uniq_thr, p.W1, p.D, p.P1, p.P2, nrows, 1u, c->tune[1]); break;
        case 4: k_hsweep<4><<<grid, wpb * 32, 0, s>>>(c->d_C, c->d_S, rec, c->d_uniq_thr, p.W1, p.D, p.P1, p.P2, nrows, 1u, c->tune[1]); break;
        default: k_hsweep<8><<<grid, wpb * 32, 0, s>>>(c->d_C, c->d_S, rec, c->d_uniq_thr, p.W1, p.D, p.P1, p.P2, nrows, 1u, c->tune[1]); break;
    }
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_wta_finalize(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t total = (size_t)B * p.H * p.W1;
    k_wta_finalize<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint2*>(c->d_wta_rec), c->d_disp_raw,
                                                                    c->d_disp2key, p.W, p.D, total);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

}  // namespace ssm
