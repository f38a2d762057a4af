// sgbm_cost.cu -- matching-cost stage of the SGBM chain, sm_100a.
//
// Replaces the calcPixelCostBT + block-sum part of cv::StereoSGBM that /root/reference
// src/stereo.cpp:13-30 calls.  Three generations, all bit-exact with the oracle and all still covered by parity tests:
//   r2  k_prefilter_tab + k_cost_tma     128- / 256-disparity layouts at the reference's block size: the right image is written in
//                                        table format and pulled into shared memory by TMA bulk copies (mbarrier, two rows ahead)
//   r1  k_prefilter8 + k_cost_fused      every power-of-two layout: one launch, tables built per row inside the kernel
//   r0  k_prefilter + k_pix_hsum + k_vsum  any multiple of 16 disparities (SSM_LEGACY_COST=1): horizontal sums through HBM
//   k_prefilter   image -> per-pixel record {v,-v,lo,-hi} for the Sobel-x and the raw channel (A-1, A-2 intervals)
//   k_pix_hsum    records -> Birchfield-Tomasi pixel cost in packed s16x2 lanes (VIADDMNMX/VIMNMX), staged as a
//                 shared-memory tile, then the bs-wide horizontal window sum hs[y][x'][d] (A-2, A-3 first half)
//   k_vsum        bs-tall running window sum down the rows -> C[y][x'][d] int16 (A-3 second half)
#include <algorithm>

#include "ssm_internal.cuh"
#include "ssm_tma.cuh"

namespace ssm {

// ------------------------------------------------------------------------------------------------
// K0: prefilter + BT half-sample intervals.  One thread per pixel, both images in one launch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sobel_clip(const uint8_t* __restrict__ img, int W, int H, int y, int x, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero;  // first/last column forced to ftzero (A-1)
    const uint8_t* cur = img + (size_t)y * W;
    const uint8_t* up = img + (size_t)(y > 0 ? y - 1 : y) * W;
    const uint8_t* dn = img + (size_t)(y < H - 1 ? y + 1 : y) * W;
    int s = ((int)cur[x + 1] - (int)cur[x - 1]) * 2 + ((int)up[x + 1] - (int)up[x - 1]) + ((int)dn[x + 1] - (int)dn[x - 1]);
    s = max(-ftzero, min(ftzero, s));
    return s + ftzero;
}
__device__ __forceinline__ int raw_val(const uint8_t* __restrict__ img, int W, int y, int x, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero;  // the raw channel's border columns are ftzero too (A-1)
    return img[(size_t)y * W + x];
}
__device__ __forceinline__ uint2 bt_record(int pm, int p, int pp, bool has_left, bool has_right)
{
    const int a = has_left ? (p + pm) >> 1 : p;
    const int b = has_right ? (p + pp) >> 1 : p;
    const int lo = min(min(a, b), p), hi = max(max(a, b), p);
    uint2 r;
    r.x = (uint32_t)(p & 0xffff) | ((uint32_t)((-p) & 0xffff) << 16);
    r.y = (uint32_t)(lo & 0xffff) | ((uint32_t)((-hi) & 0xffff) << 16);
    return r;
}

__global__ void __launch_bounds__(256) k_prefilter(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right,
                                                   uint4* __restrict__ recL, uint4* __restrict__ recR, int W, int H,
                                                   int ftzero, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int which = blockIdx.y;  // 0 = left, 1 = right
    const uint8_t* img_all = which ? right : left;
    uint4* rec = which ? recR : recL;
    const int x = (int)(idx % W);
    const size_t row = idx / W;
    const int y = (int)(row % H);
    const uint8_t* img = img_all + (row / H) * (size_t)W * H;
    const int g0 = sobel_clip(img, W, H, y, x - 1, ftzero), g1 = sobel_clip(img, W, H, y, x, ftzero),
              g2 = sobel_clip(img, W, H, y, x + 1, ftzero);
    const int r0 = raw_val(img, W, y, x - 1, ftzero), r1 = raw_val(img, W, y, x, ftzero),
              r2 = raw_val(img, W, y, x + 1, ftzero);
    const uint2 a = bt_record(g0, g1, g2, x > 0, x < W - 1);
    const uint2 b = bt_record(r0, r1, r2, x > 0, x < W - 1);
    rec[idx] = make_uint4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------
// K1a: pixel cost + horizontal window sum.  CTA = (x' tile of TX columns, row y, batch item b).
// ------------------------------------------------------------------------------------------------
constexpr int kTX = 64;     // output columns per CTA
constexpr int kRL = 4;      // consecutive outputs per thread in the window-sum phase
constexpr int kMaxR = 5;    // block_size <= 11

// BT cost of one channel for two disparities at once (s16x2 lanes):
//   c0 = max(0, u - hiR, loR - u), c1 = max(0, v - hiL, loL - v), cost = min(c0, c1)
__device__ __forceinline__ uint32_t bt2(uint32_t u, uint32_t nu, uint32_t loL, uint32_t nhiL, uint32_t v, uint32_t nv,
                                        uint32_t loR, uint32_t nhiR)
{
    const uint32_t t0 = __viaddmax_s16x2(loR, nu, 0u);
    const uint32_t c0 = __viaddmax_s16x2(u, nhiR, t0);
    const uint32_t t1 = __viaddmax_s16x2(nv, loL, 0u);
    const uint32_t c1 = __viaddmax_s16x2(v, nhiL, t1);
    return __vmins2(c0, c1);
}

__global__ void __launch_bounds__(256) k_pix_hsum(const uint4* __restrict__ recL, const uint4* __restrict__ recR,
                                                  uint16_t* __restrict__ hs, int W, int H, int D, int radius)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int W1 = W - D;
    const int octets = D >> 3;
    const int wpp = D >> 1;                          // 32-bit words per pixel in the cost tile
    const int t0 = blockIdx.x * kTX, y = blockIdx.y, b = blockIdx.z;
    const int e_lo = max(t0 - radius, 0), e_hi = min(t0 + kTX - 1 + radius, W1 - 1);
    const int n_e = e_hi - e_lo + 1;
    const int n_r = n_e + D - 1;                     // right-image pixels [e_lo + 1, e_hi + D], reversed
    const int tw = (kTX + 2 * kMaxR + D) / 2 + 4;    // words per (quantity, copy) table

    uint32_t* Rt = smem;                             // [8 quantities][2 copies][tw]
    uint32_t* Lt = Rt + 16 * tw;                     // [kTX + 2*kMaxR][8] pre-duplicated left record
    uint32_t* pix = Lt + (kTX + 2 * kMaxR) * 8;      // [kTX + 2*kMaxR][wpp]

    const size_t rowbase = ((size_t)b * H + y) * W;
    // right tables: element i <-> pixel xr = e_hi + D - i; copy 1 is shifted down by one element
    for (int i = threadIdx.x; i < n_r; i += blockDim.x) {
        const uint4 r = recR[rowbase + (e_hi + D - i)];
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint16_t val = (uint16_t)(w[q >> 1] >> ((q & 1) * 16));
            uint16_t* c0 = reinterpret_cast<uint16_t*>(Rt + (q * 2 + 0) * tw);
            uint16_t* c1 = reinterpret_cast<uint16_t*>(Rt + (q * 2 + 1) * tw);
            c0[i] = val;
            if (i > 0) c1[i - 1] = val;
        }
    }
    for (int i = threadIdx.x; i < n_e; i += blockDim.x) {
        const uint4 r = recL[rowbase + (e_lo + i + D)];
        uint4* dst = reinterpret_cast<uint4*>(Lt + i * 8);
        dst[0] = make_uint4(__byte_perm(r.x, 0, 0x1010), __byte_perm(r.x, 0, 0x3232), __byte_perm(r.y, 0, 0x1010),
                            __byte_perm(r.y, 0, 0x3232));
        dst[1] = make_uint4(__byte_perm(r.z, 0, 0x1010), __byte_perm(r.z, 0, 0x3232), __byte_perm(r.w, 0, 0x1010),
                            __byte_perm(r.w, 0, 0x3232));
    }
    __syncthreads();

    // phase 1: pixel costs for (e, octet) items
    for (int it = threadIdx.x; it < n_e * octets; it += blockDim.x) {
        const int el = it / octets, o = it - el * octets;
        const int i0 = (n_e - 1 - el) + 8 * o;       // reversed index of d = 8*o at pixel e = e_lo + el
        const int par = i0 & 1, wb = i0 >> 1;
        const uint4 la = reinterpret_cast<const uint4*>(Lt + el * 8)[0];   // u, -u, lo, -hi (gradient), duplicated
        const uint4 lb = reinterpret_cast<const uint4*>(Lt + el * 8)[1];   // same for the raw channel
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t* base = Rt + par * tw + wb + k;
            const uint32_t cg = bt2(la.x, la.y, la.z, la.w, base[0 * 2 * tw], base[1 * 2 * tw], base[2 * 2 * tw],
                                    base[3 * 2 * tw]);
            const uint32_t cr = bt2(lb.x, lb.y, lb.z, lb.w, base[4 * 2 * tw], base[5 * 2 * tw], base[6 * 2 * tw],
                                    base[7 * 2 * tw]);
            out[k] = cg + ((cr >> 2) & 0x3fff3fffu);
        }
        reinterpret_cast<uint4*>(pix + el * wpp)[o] = make_uint4(out[0], out[1], out[2], out[3]);
    }
    __syncthreads();

    // phase 2: horizontal window sums, kRL consecutive columns per thread (sliding)
    const int runs = kTX / kRL;
    for (int it = threadIdx.x; it < runs * octets; it += blockDim.x) {
        const int run = it / octets, o = it - run * octets;
        const int x0 = t0 + run * kRL;
        if (x0 >= W1) continue;
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int k = -radius; k <= radius; ++k) {
            const int e = min(max(x0 + k, 0), W1 - 1) - e_lo;
            const uint4 v = reinterpret_cast<const uint4*>(pix + e * wpp)[o];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;   // lanes stay < 2^15: no carry between halves
        }
        uint16_t* dst = hs + (((size_t)b * H + y) * W1 + x0) * D + 8 * o;
#pragma unroll
        for (int t = 0; t < kRL; ++t) {
            const int xp = x0 + t;
            if (xp >= W1) break;
            *reinterpret_cast<uint4*>(dst + (size_t)t * D) = acc;
            if (t == kRL - 1) break;                 // the next column belongs to another thread's run
            const int ein = min(max(xp + 1 + radius, 0), W1 - 1) - e_lo;
            const int eout = min(max(xp - radius, 0), W1 - 1) - e_lo;
            const uint4 vi = reinterpret_cast<const uint4*>(pix + ein * wpp)[o];
            const uint4 vo = reinterpret_cast<const uint4*>(pix + eout * wpp)[o];
            acc.x = acc.x + vi.x - vo.x; acc.y = acc.y + vi.y - vo.y;
            acc.z = acc.z + vi.z - vo.z; acc.w = acc.w + vi.w - vo.w;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1b: vertical window sum.  One thread per (band, x', octet) marching down its band of rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vsum(const uint16_t* __restrict__ hs, int16_t* __restrict__ C, int H,
                                              int rowwords4 /* W1*D/8 uint4 per row */, int radius, int band_rows,
                                              int nbands, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int col = (int)(idx % rowwords4);
    const size_t rest = idx / rowwords4;
    const int band = (int)(rest % nbands);
    const size_t b = rest / nbands;
    const int y0 = band * band_rows, y1 = min(H, y0 + band_rows);
    const uint4* src = reinterpret_cast<const uint4*>(hs) + b * (size_t)H * rowwords4 + col;
    uint4* dst = reinterpret_cast<uint4*>(C) + b * (size_t)H * rowwords4 + col;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int k = -radius; k <= radius; ++k) {
        const uint4 v = src[(size_t)min(max(y0 + k, 0), H - 1) * rowwords4];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int y = y0; y < y1; ++y) {
        dst[(size_t)y * rowwords4] = acc;
        const uint4 vi = src[(size_t)min(y + 1 + radius, H - 1) * rowwords4];
        const uint4 vo = src[(size_t)max(y - radius, 0) * rowwords4];
        acc.x = acc.x + vi.x - vo.x; acc.y = acc.y + vi.y - vo.y;
        acc.z = acc.z + vi.z - vo.z; acc.w = acc.w + vi.w - vo.w;
    }
}

// ------------------------------------------------------------------------------------------------
// K0 (compact): prefilter + BT half-sample intervals as 8-byte records {gv, glo, ghi, rv, rlo, rhi, 0, 0} (all
// values fit a byte: Sobel channel in [0, 126], raw channel in [0, 255]).  One thread makes 4 consecutive pixels
// from one 8 x 3 pixel neighbourhood; both images in one launch.
// ------------------------------------------------------------------------------------------------
// 8 consecutive bytes starting at byte address `a` of a 4-byte aligned array, from three aligned 32-bit loads
__device__ __forceinline__ uint2 load8_unaligned(const uint8_t* __restrict__ base, size_t a)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (a >> 2);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
__device__ __forceinline__ int byte_of(const uint2& v, int i) { return (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xffu); }

__global__ void __launch_bounds__(256) k_prefilter8(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right,
                                                    uint2* __restrict__ recL, uint2* __restrict__ recR, int W, int H, int ftzero,
                                                    int quads_per_row, size_t total_quads, int aligned4 /* both images 4-byte aligned */)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total_quads) return;
    const int which = blockIdx.y;                    // 0 = left, 1 = right
    const uint8_t* img_all = which ? right : left;
    uint2* rec = which ? recR : recL;
    const int x0 = (int)(q % quads_per_row) * 4;
    const size_t row = q / quads_per_row;            // frame * H + y
    const int y = (int)(row % H);
    const size_t cur_o = row * (size_t)W;
    const size_t up_o = y > 0 ? cur_o - W : cur_o, dn_o = y < H - 1 ? cur_o + W : cur_o;
    int g[6], r[6];
    // columns x0 - 2 .. x0 + 5 of the three rows
    if (aligned4 && x0 >= 8 && x0 + 10 <= W) {
        // interior quad: every byte (and the aligned words around them) lies inside the frame -- three aligned 32-bit loads and
        // two funnel shifts per row instead of eight byte loads; no border column among x0 - 1 .. x0 + 4
        const uint2 a = load8_unaligned(img_all, up_o + x0 - 2), b = load8_unaligned(img_all, cur_o + x0 - 2),
                    c = load8_unaligned(img_all, dn_o + x0 - 2);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int sx = (byte_of(b, i + 2) - byte_of(b, i)) * 2 + (byte_of(a, i + 2) - byte_of(a, i)) + (byte_of(c, i + 2) - byte_of(c, i));
            g[i] = max(-ftzero, min(ftzero, sx)) + ftzero;
            r[i] = byte_of(b, i + 1);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ga = (g[i + 1] + g[i]) >> 1, gb = (g[i + 1] + g[i + 2]) >> 1;
            const int ra = (r[i + 1] + r[i]) >> 1, rb = (r[i + 1] + r[i + 2]) >> 1;
            const int glo = min(min(ga, gb), g[i + 1]), ghi = max(max(ga, gb), g[i + 1]);
            const int rlo = min(min(ra, rb), r[i + 1]), rhi = max(max(ra, rb), r[i + 1]);
            rec[cur_o + x0 + i] = make_uint2((uint32_t)g[i + 1] | ((uint32_t)glo << 8) | ((uint32_t)ghi << 16) | ((uint32_t)r[i + 1] << 24),
                                             (uint32_t)rlo | ((uint32_t)rhi << 8));
        }
        return;
    }
    const uint8_t* cur = img_all + cur_o;
    const uint8_t* up = img_all + up_o;
    const uint8_t* dn = img_all + dn_o;
    // border quads: clamped byte loads; out-of-range values are never used unclamped
    int a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int x = min(max(x0 - 2 + i, 0), W - 1);
        a[i] = up[x]; b[i] = cur[x]; c[i] = dn[x];
    }
    // g[i], r[i] for columns x0 - 1 .. x0 + 4
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int x = x0 - 1 + i;
        const bool border = x <= 0 || x >= W - 1;    // first / last column (and beyond) are ftzero in both channels (A-1)
        int sx = (b[i + 2] - b[i]) * 2 + (a[i + 2] - a[i]) + (c[i + 2] - c[i]);
        sx = max(-ftzero, min(ftzero, sx)) + ftzero;
        g[i] = border ? ftzero : sx;
        r[i] = border ? ftzero : b[i + 1];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + i;
        if (x >= W) break;
        const bool hl = x > 0, hr = x < W - 1;
        const int ga = hl ? (g[i + 1] + g[i]) >> 1 : g[i + 1], gb = hr ? (g[i + 1] + g[i + 2]) >> 1 : g[i + 1];
        const int ra = hl ? (r[i + 1] + r[i]) >> 1 : r[i + 1], rb = hr ? (r[i + 1] + r[i + 2]) >> 1 : r[i + 1];
        const int glo = min(min(ga, gb), g[i + 1]), ghi = max(max(ga, gb), g[i + 1]);
        const int rlo = min(min(ra, rb), r[i + 1]), rhi = max(max(ra, rb), r[i + 1]);
        rec[row * (size_t)W + x] = make_uint2((uint32_t)g[i + 1] | ((uint32_t)glo << 8) | ((uint32_t)ghi << 16) | ((uint32_t)r[i + 1] << 24),
                                              (uint32_t)rlo | ((uint32_t)rhi << 8));
    }
}

// ------------------------------------------------------------------------------------------------
// K1 fused: pixel cost + horizontal window sum + vertical window sum -> C, one launch, no intermediate volume.
//
// CTA = (tile of TX output columns, band of rows, frame); TX * D/2 = 2048 packed words, 256 threads.  The CTA
// marches down its band (plus `radius` rows above and below); per image row:
//   phase 0  the row's prefilter records -> shared tables: right image {lo, hi, v} x {Sobel, raw} as s16x2 words in
//            two copies (element-aligned and shifted by one, so that the pair (d, d+1) <-> right pixels
//            (x-d, x-d-1) is ONE aligned 32-bit word for every x and d), left image 8 pre-combined words per pixel
//   phase 1  Birchfield-Tomasi cost of (pixel, word) items, 16 lanes per pixel on consecutive words (bank-conflict
//            free: the two copies sit 16 banks apart), 4 VIADDMNMX.S16x2 + 1 VIMNMX per channel and word;
//            negated operands are formed as K - x by IMAD (FMA pipe) instead of being loaded  -> pix tile (shared)
//   phase 2  thread = (8 consecutive columns, one word): 18 loads feed eight bs-wide window sums held in registers;
//            vertical running sum C_run += hs - hs[bs rows ago] against a bs-row ring of hs in shared memory;
//            one coalesced store of C per column once the window is full
// Clamping is relative to the valid region [0, W1) x [0, H); rows beyond the image edge re-use
// the previous row's window sums.  All lanes stay < 2^15, so 32-bit adds never carry between the halves.
// ------------------------------------------------------------------------------------------------
constexpr int kKb = 0x4000;                 // bias that keeps K - x positive in both halves
constexpr uint32_t kKw = 0x40004000u;
constexpr uint32_t kUnbias = 0u - 0x50005000u;   // -(K + K/4) in both halves

// compile-time geometry of one tile width (TX * D / 2 = 2048 packed words per CTA row)
template <int TX>
struct CostGeom {
    static constexpr int D = 4096 / TX;
    static constexpr int WPP = D / 2;                        // packed words per pixel
    static constexpr int NE = TX + 2 * kMaxR;                // pixels whose cost a CTA computes per row (at most)
    static constexpr int NR_MAX = NE + D - 1;                // right-image pixels it needs
    static constexpr int TW = ((NE + D) / 2 + 2 + 31) / 32 * 32 + 16;   // words per (quantity, copy) table, = 16 mod 32
    static constexpr int PS = WPP + 16;                      // pix row stride in words
    static constexpr int LPP = WPP < 16 ? WPP : 16;          // lanes per pixel in phase 1
    static constexpr int WPL = WPP / LPP;                    // words per lane
    static constexpr int RPT = (NR_MAX + 255) / 256;         // right records per thread
    static constexpr int LPT = (NE + 255) / 256;             // left records per thread
    static constexpr int SL = ((NE + 1) / 2 + 6) / 7;        // diagonal sweep: pixels per segment (7 segments per pixel parity)
    static constexpr bool DIAG = WPP >= 64 && SL <= 3;       // phase-1 diagonal sweep (D >= 128)
    static constexpr size_t smem_words(int bs) { return (size_t)12 * TW + NE * 8 + (size_t)NE * PS + (size_t)bs * TX * WPP; }
};

template <int TX, int RAD /* block_size / 2, or -1: run-time radius (slow generic window sums) */,
          bool PAD = false /* padded layout: cells at d >= Dv are written as `padw`, the aggregation kernels' "+inf" cost */>
__global__ void __launch_bounds__(256, 2) k_cost_fused(const uint2* __restrict__ recL, const uint2* __restrict__ recR,
                                                       int16_t* __restrict__ C, int W, int H, int Dv /* valid disparities <= D: the image geometry */,
                                                       int radius, int band_rows,
                                                       uint32_t mone /* 0xffffffff, opaque: x * mone + K is one IMAD */,
                                                       uint32_t padw = 0u)
{
    using G = CostGeom<TX>;
    constexpr int D = G::D, WPP = G::WPP, TW = G::TW, PS = G::PS, LPP = G::LPP, WPL = G::WPL;
    extern __shared__ __align__(16) uint32_t smem[];
    const int W1 = W - Dv;                           // D is the layout (words per column); disparities >= Dv are computed from
                                                     // whatever lies left of the image (zeros) and never read by the aggregation kernels
    const int t0 = blockIdx.x * TX, b = blockIdx.z;
    const int y0 = blockIdx.y * band_rows, y1 = min(H, y0 + band_rows);
    const int e_lo = max(t0 - radius, 0), e_hi = min(t0 + TX - 1 + radius, W1 - 1);
    const int n_e = e_hi - e_lo + 1;
    const int n_r = n_e + D - 1;                     // right-image pixels [e_lo + 1, e_hi + D], stored reversed
    const int bs = 2 * radius + 1;
    uint32_t one = 0u - mone;                        // 1, opaque to the compiler: x * one + y stays an IMAD
    asm volatile("" : "+r"(one));

    uint32_t* Rt = smem;                             // [6 quantities][2 copies][TW]
    uint32_t* Lt = Rt + 12 * TW;                     // [NE][8]
    uint32_t* pix = Lt + G::NE * 8;                  // [NE][PS]
    uint32_t* ring = pix + G::NE * PS;               // [bs][TX][WPP]
    for (int i = threadIdx.x; i < bs * TX * WPP; i += 256) ring[i] = 0u;

    // phase-2 identity: 8 consecutive columns x one word; (TX / 8) groups x WPP words = 256 threads exactly
    const int cgp = threadIdx.x / WPP;
    const int w2 = threadIdx.x - cgp * WPP;
    const int c0 = t0 + cgp * 8;                     // first of my 8 output columns
    const bool interior = RAD >= 0 && c0 - RAD >= 0 && c0 + 7 + RAD <= W1 - 1;
    uint32_t crun[8], hs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { crun[i] = 0u; hs[i] = 0u; }

    // records of the next fresh row, prefetched one row ahead
    uint2 qr[G::RPT], ql[G::LPT];
    auto fetch = [&](int r) {
        const size_t rowbase = ((size_t)b * H + r) * W;
#pragma unroll
        for (int k = 0; k < G::RPT; ++k) {
            const int i = threadIdx.x + k * 256;
            if (i < n_r) qr[k] = e_hi + Dv - i >= 0 ? recR[rowbase + (e_hi + Dv - i)] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < G::LPT; ++k) {
            const int i = threadIdx.x + k * 256;
            if (i < n_e) ql[k] = recL[rowbase + (e_lo + i + Dv)];
        }
    };
    const int j_begin = y0 - radius, j_end = y1 + radius;
    fetch(min(max(j_begin, 0), H - 1));

    int slot = 0, prev_r = -1;
    for (int j = j_begin; j < j_end; ++j) {
        const int r = min(max(j, 0), H - 1);
        const bool fresh = r != prev_r;              // CTA-uniform
        prev_r = r;
        if (fresh) {
            // ---- phase 0: tables from the prefetched records.  element i of the reversed right tables <-> pixel e_hi + D - i
#pragma unroll
            for (int k = 0; k < G::RPT; ++k) {
                const int i = threadIdx.x + k * 256;
                if (i < n_r) {
                    const uint2 q = qr[k];
                    // record bytes: gv, glo, ghi, rv | rlo, rhi
                    const uint16_t val[6] = {(uint16_t)((q.x >> 8) & 0xffu), (uint16_t)((q.x >> 16) & 0xffu), (uint16_t)(q.x & 0xffu),
                                             (uint16_t)(q.y & 0xffu), (uint16_t)((q.y >> 8) & 0xffu), (uint16_t)(q.x >> 24)};
#pragma unroll
                    for (int t = 0; t < 6; ++t) {
                        uint16_t* a0 = reinterpret_cast<uint16_t*>(Rt + (2 * t) * TW);
                        uint16_t* a1 = reinterpret_cast<uint16_t*>(Rt + (2 * t + 1) * TW);
                        a0[i] = val[t];
                        if (i > 0) a1[i - 1] = val[t];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < G::LPT; ++k) {
                const int i = threadIdx.x + k * 256;
                if (i < n_e) {
                    const uint2 q = ql[k];
                    // per channel: dup(u + K), dup(K - u), dup(K - hi), dup(lo + K)  (K keeps every packed difference positive)
                    const uint32_t ug = q.x & 0xffu, log_ = (q.x >> 8) & 0xffu, hig = (q.x >> 16) & 0xffu;
                    const uint32_t ur = q.x >> 24, lor = q.y & 0xffu, hir = (q.y >> 8) & 0xffu;
                    uint4* dst = reinterpret_cast<uint4*>(Lt + i * 8);
                    dst[0] = make_uint4((ug + kKb) * 0x10001u, (kKb - ug) * 0x10001u, (kKb - hig) * 0x10001u, (log_ + kKb) * 0x10001u);
                    dst[1] = make_uint4((ur + kKb) * 0x10001u, (kKb - ur) * 0x10001u, (kKb - hir) * 0x10001u, (lor + kKb) * 0x10001u);
                }
            }
            // the records of the next distinct row travel while this row is processed
            {
                const int rn = min(r + 1, H - 1);
                if (rn != r && j + 1 < j_end) fetch(rn);
            }
            __syncthreads();
            // ---- phase 1: pixel costs.  bt_word() is the Birchfield-Tomasi cost of one packed word, biased operands:
            //   c0 = max(u - hiR, loR - u, 0), c1 = max(v - hiL, loL - v, 0), cost = min(c0, c1), all + K per half.
            // The four differences are IMADs (FMA pipe); one VIMNMX3 per max, one VIMNMX per min (ALU pipe).
            auto bt_word = [&](const uint4& la, const uint4& lb, uint32_t loG, uint32_t hiG, uint32_t vG, uint32_t loR, uint32_t hiR,
                               uint32_t vR) -> uint32_t {
                const uint32_t g0 = __vimax3_s16x2(hiG * mone + la.x, loG * one + la.y, kKw);
                const uint32_t g1 = __vimax3_s16x2(vG * one + la.z, vG * mone + la.w, kKw);
                const uint32_t r0 = __vimax3_s16x2(hiR * mone + lb.x, loR * one + lb.y, kKw);
                const uint32_t r1 = __vimax3_s16x2(vR * one + lb.z, vR * mone + lb.w, kKw);
                const uint32_t cgv = __vmins2(g0, g1), crv = __vmins2(r0, r1);
                // (crv >> 2) per half = (cost_raw >> 2) + K/4; the sum carries K + K/4 per half
                return cgv + ((crv >> 2) & 0x3fff3fffu) + kUnbias;
            };
            if constexpr (G::DIAG) {
                // Diagonal sweep: word w of pixel e and word w + 1 of pixel e + 2 read the SAME right-image pair, so a
                // thread that walks pixels e, e + 2, e + 4, ... while its words move up by one keeps its right-image
                // operands in registers: 6 * WPL table loads per SL pixels instead of per pixel.  16 lanes x WPL words
                // cover a pixel; the words that enter at the bottom of the disparity range on the way (word < step) are
                // left to the spare warp below.
                constexpr int SL = G::SL;
                const int sg = threadIdx.x >> 4, jl = threadIdx.x & 15;
                const int el0 = (sg >> 1) * 2 * SL + (sg & 1);
                if (threadIdx.x < 224) {
                    if (el0 < n_e) {
                        const int i0 = n_e - 1 - el0;
                        const uint32_t* q = Rt + (i0 & 1) * TW + (i0 >> 1) + jl;
                        uint32_t loG[WPL], hiG[WPL], vG[WPL], loR[WPL], hiR[WPL], vR[WPL];
#pragma unroll
                        for (int k = 0; k < WPL; ++k) {
                            loG[k] = q[k * LPP]; hiG[k] = q[k * LPP + 2 * TW]; vG[k] = q[k * LPP + 4 * TW];
                            loR[k] = q[k * LPP + 6 * TW]; hiR[k] = q[k * LPP + 8 * TW]; vR[k] = q[k * LPP + 10 * TW];
                        }
#pragma unroll
                        for (int t = 0; t < SL; ++t) {
                            const int el = el0 + 2 * t;
                            if (el < n_e) {
                                const uint4 la = reinterpret_cast<const uint4*>(Lt + el * 8)[0];
                                const uint4 lb = reinterpret_cast<const uint4*>(Lt + el * 8)[1];
                                uint32_t* out = pix + el * PS + jl + t;
#pragma unroll
                                for (int k = 0; k < WPL; ++k) {
                                    const uint32_t v = bt_word(la, lb, loG[k], hiG[k], vG[k], loR[k], hiR[k], vR[k]);
                                    if (k < WPL - 1 || jl + t < LPP) out[k * LPP] = v;
                                }
                            }
                        }
                    }
                } else {
                    // spare warp: per group of 2 * SL pixels the words (step t, word w < t) of both pixel parities
                    constexpr int IPG = SL * (SL - 1);                 // items per group
                    const int n_items = ((n_e + 2 * SL - 1) / (2 * SL)) * IPG;
                    for (int c = threadIdx.x - 224; c < n_items; c += 32) {
                        const int g = c / IPG, r = c - g * IPG;
                        const int qd = r >> 1;                         // (t, w): (1, 0), (2, 0), (2, 1)
                        const int t = qd == 0 ? 1 : 2, w = qd == 2 ? 1 : 0;
                        const int el = g * 2 * SL + 2 * t + (r & 1);
                        if (el < n_e) {
                            const uint4 la = reinterpret_cast<const uint4*>(Lt + el * 8)[0];
                            const uint4 lb = reinterpret_cast<const uint4*>(Lt + el * 8)[1];
                            const int i0 = n_e - 1 - el;
                            const uint32_t* q = Rt + (i0 & 1) * TW + (i0 >> 1) + w;
                            pix[el * PS + w] = bt_word(la, lb, q[0], q[2 * TW], q[4 * TW], q[6 * TW], q[8 * TW], q[10 * TW]);
                        }
                    }
                }
            } else {
                // (pixel, lane-in-pixel) items; lane jl takes words jl, jl + LPP, ...
                const int items = n_e * LPP;
                for (int it = threadIdx.x; it < items; it += 256) {
                    const int el = it / LPP, jl = it % LPP;
                    const uint4 la = reinterpret_cast<const uint4*>(Lt + el * 8)[0];
                    const uint4 lb = reinterpret_cast<const uint4*>(Lt + el * 8)[1];
                    const int i0 = n_e - 1 - el;             // reversed index of d = 0 at pixel e_lo + el
                    const uint32_t* q = Rt + (i0 & 1) * TW + (i0 >> 1) + jl;
                    uint32_t* out = pix + el * PS + jl;
#pragma unroll
                    for (int k = 0; k < WPL; ++k)
                        out[k * LPP] = bt_word(la, lb, q[k * LPP], q[k * LPP + 2 * TW], q[k * LPP + 4 * TW], q[k * LPP + 6 * TW],
                                               q[k * LPP + 8 * TW], q[k * LPP + 10 * TW]);
                }
            }
            __syncthreads();
            // ---- phase 2a: horizontal window sums of my 8 columns (registers)
            if (c0 < W1) {
                if constexpr (RAD >= 0) {
                    uint32_t v[8 + 2 * RAD];
                    if (interior) {
                        const uint32_t* src = pix + (c0 - RAD - e_lo) * PS + w2;
#pragma unroll
                        for (int i = 0; i < 8 + 2 * RAD; ++i) v[i] = src[i * PS];
                    } else {
#pragma unroll
                        for (int i = 0; i < 8 + 2 * RAD; ++i) {
                            const int e = min(max(c0 + i - RAD, 0), W1 - 1) - e_lo;
                            v[i] = pix[e * PS + w2];
                        }
                    }
                    uint32_t acc = 0u;
#pragma unroll
                    for (int i = 0; i < 2 * RAD + 1; ++i) acc += v[i];
                    hs[0] = acc;
#pragma unroll
                    for (int c = 1; c < 8; ++c) {
                        acc = acc + v[c + 2 * RAD] - v[c - 1];
                        hs[c] = acc;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) hs[c] = 0u;
                    for (int off = -radius; off <= 7 + radius; ++off) {
                        const int e = min(max(c0 + off, 0), W1 - 1) - e_lo;
                        const uint32_t val = pix[e * PS + w2];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (off >= c - radius && off <= c + radius) hs[c] += val;
                    }
                }
            }
        }
        // ---- phase 2b: vertical running sum against the ring, output
        if (c0 < W1) {
            uint32_t* rg = ring + (slot * TX + cgp * 8) * WPP + w2;
            const int yout = j - radius;
            int16_t* dst = C + (((size_t)b * H + max(yout, 0)) * W1 + c0) * D + 2 * w2;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t old = rg[c * WPP];
                rg[c * WPP] = hs[c];
                crun[c] = crun[c] + hs[c] - old;
                if (yout >= y0 && c0 + c < W1) *reinterpret_cast<uint32_t*>(dst + (size_t)c * D) = (PAD && 2 * w2 >= Dv) ? padw : crun[c];
            }
        }
        slot = slot + 1 == bs ? 0 : slot + 1;
        // hazards: the next row's tables are written after every thread has left phase 1 (second barrier); the pix
        // tile is re-written only after the next row's first barrier, which every thread reaches after its phase 2
    }
}

// ------------------------------------------------------------------------------------------------
// K1 fused, r2: the same three phases with the per-row table building taken off the instruction stream.
//
//   * k_prefilter_tab writes the RIGHT image's six Birchfield-Tomasi quantities directly in table format: one 32-bit word
//     P_q[x] = val_q[x] | val_q[x-1] << 16 per pixel and quantity (array [frame][row][q][pitch], zero margin on the left for
//     padded layouts).  The packed pair (d, d+1) = 2w, 2w+1 of pixel x is then the single word P_q[x - 2w]: sixteen lanes on
//     consecutive words read sixteen banks of one parity, the other half-warp (the adjacent pixel) the other parity.
//   * The cost kernel pulls the six table rows of an image row with SIX TMA 1-D BULK COPIES tracked by one mbarrier, two
//     rows ahead (double-buffered): phase 0 of k_cost_fused (record extraction + 12 STS.U16 per right pixel, 13 % of its
//     instructions) and one of its two per-row fetch/convert steps are gone.
//   * The left image's 8 pre-combined words per pixel are built for the NEXT row by the spare warp of phase 1, which has
//     two thirds of a main warp's work to spare.
//   * Phase 2b keeps each thread's 8-column slice of the ring contiguous ([slot][group][half][word][4 columns]): two
//     LDS.128 + two STS.128 instead of 8 + 8 32-bit accesses; the output pointer is advanced, not recomputed.
// Shapes: D >= 128 layouts with the diagonal sweep and two table buffers inside the 113 KB of a CTA pair (D = 128).
// ------------------------------------------------------------------------------------------------
// 12 consecutive bytes starting at byte address `a` of a 4-byte aligned array, from four aligned 32-bit loads
__device__ __forceinline__ uint3 load12_unaligned(const uint8_t* __restrict__ base, size_t a)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (a >> 2);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    const uint32_t sh = (uint32_t)(a & 3) * 8u;
    return make_uint3(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh));
}
__device__ __forceinline__ int byte12(const uint3& v, int i)
{
    return (int)(((i < 4 ? v.x : (i < 8 ? v.y : v.z)) >> (8 * (i & 3))) & 0xffu);
}

// One thread makes 4 consecutive pixels x0 .. x0 + 3 (and the values of pixel x0 - 1, the upper half of P[x0]).
// blockIdx.y = 0: left image -> 8-byte records (as k_prefilter8); 1: right image -> table words.
__global__ void __launch_bounds__(256) k_prefilter_tab(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right,
                                                       uint2* __restrict__ recL, uint32_t* __restrict__ ptab, int W, int H, int ftzero,
                                                       int quads_per_row, size_t total_quads, int aligned4, int pitch, int margin)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total_quads) return;
    const int which = blockIdx.y;
    const uint8_t* img_all = which ? right : left;
    const int x0 = (int)(q % quads_per_row) * 4;
    const size_t row = q / quads_per_row;            // frame * H + y
    const int y = (int)(row % H);
    const size_t cur_o = row * (size_t)W;
    const size_t up_o = y > 0 ? cur_o - W : cur_o, dn_o = y < H - 1 ? cur_o + W : cur_o;
    int g[7], r[7];                                  // columns x0 - 2 .. x0 + 4
    if (aligned4 && x0 >= 8 && x0 + 16 <= W) {
        const uint3 a = load12_unaligned(img_all, up_o + x0 - 3), b = load12_unaligned(img_all, cur_o + x0 - 3),
                    c = load12_unaligned(img_all, dn_o + x0 - 3);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int sx = (byte12(b, i + 2) - byte12(b, i)) * 2 + (byte12(a, i + 2) - byte12(a, i)) + (byte12(c, i + 2) - byte12(c, i));
            g[i] = max(-ftzero, min(ftzero, sx)) + ftzero;
            r[i] = byte12(b, i + 1);
        }
    } else {
        const uint8_t* cur = img_all + cur_o;
        const uint8_t* up = img_all + up_o;
        const uint8_t* dn = img_all + dn_o;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int x = x0 - 2 + i;
            const int xm = min(max(x - 1, 0), W - 1), xc = min(max(x, 0), W - 1), xp = min(max(x + 1, 0), W - 1);
            const bool border = x <= 0 || x >= W - 1;    // first / last column (and beyond) are ftzero in both channels (A-1)
            int sx = ((int)cur[xp] - (int)cur[xm]) * 2 + ((int)up[xp] - (int)up[xm]) + ((int)dn[xp] - (int)dn[xm]);
            sx = max(-ftzero, min(ftzero, sx)) + ftzero;
            g[i] = border ? ftzero : sx;
            r[i] = border ? ftzero : (int)cur[xc];
        }
    }
    // values of pixels x0 - 1 .. x0 + 3 (index p = 0 .. 4 <-> g[p + 1]); zero outside the image
    uint32_t val[6][5];                              // loG, hiG, vG, loR, hiR, vR
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const int x = x0 - 1 + p;
        const bool inside = x >= 0 && x < W;
        const bool hl = x > 0, hr = x < W - 1;
        const int gc = g[p + 1], rc = r[p + 1];
        const int ga = hl ? (gc + g[p]) >> 1 : gc, gb = hr ? (gc + g[p + 2]) >> 1 : gc;
        const int ra = hl ? (rc + r[p]) >> 1 : rc, rb = hr ? (rc + r[p + 2]) >> 1 : rc;
        val[0][p] = inside ? (uint32_t)min(min(ga, gb), gc) : 0u;
        val[1][p] = inside ? (uint32_t)max(max(ga, gb), gc) : 0u;
        val[2][p] = inside ? (uint32_t)gc : 0u;
        val[3][p] = inside ? (uint32_t)min(min(ra, rb), rc) : 0u;
        val[4][p] = inside ? (uint32_t)max(max(ra, rb), rc) : 0u;
        val[5][p] = inside ? (uint32_t)rc : 0u;
    }
    if (which == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (x0 + i >= W) break;
            const int p = i + 1;
            recL[cur_o + x0 + i] = make_uint2(val[2][p] | (val[0][p] << 8) | (val[1][p] << 16) | (val[5][p] << 24), val[3][p] | (val[4][p] << 8));
        }
        return;
    }
    uint32_t* trow = ptab + row * 6 * (size_t)pitch;
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        uint32_t* dst = trow + (size_t)t * pitch;
        *reinterpret_cast<uint4*>(dst + margin + x0) = make_uint4(val[t][1] | (val[t][0] << 16), val[t][2] | (val[t][1] << 16),
                                                                  val[t][3] | (val[t][2] << 16), val[t][4] | (val[t][3] << 16));
        if (x0 == 0)
            for (int i = 0; i < margin; i += 4) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0u, 0u, 0u, 0u);
        if (x0 + 4 >= W)
            for (int i = margin + x0 + 4; i < pitch; i += 4) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int TX>
struct CostGeom2 : CostGeom<TX> {
    using G = CostGeom<TX>;
    static constexpr int PTW = (G::D + G::NE + 2 + 3) / 4 * 4;          // words per table row in shared memory
    static constexpr int NG = TX / 8;                                   // phase-2 column groups
    static constexpr size_t smem_bytes(int bs)
    {
        return sizeof(uint32_t) * ((size_t)2 * 6 * PTW + (size_t)2 * G::NE * 8 + (size_t)G::NE * G::PS + (size_t)bs * TX * G::WPP) + 16;
    }
};

template <int TX, int RAD, bool PAD, int NKK = CostGeom<TX>::WPL /* 16-word groups of a pixel that phase 1 computes: the groups above hold
          only cells at d >= Dv of a padded layout (written as `padw`, never read from the tile) */,
          int NSPLIT = 1 /* the layout has NSPLIT * D disparities per column: blockIdx.x = tile * NSPLIT + part, part = which D of them this CTA computes */>
__global__ void __launch_bounds__(256, 2) k_cost_tma(const uint2* __restrict__ recL, const uint32_t* __restrict__ ptab,
                                                     int16_t* __restrict__ C, int W, int H, int Dv, int band_rows, int pitch, int margin,
                                                     uint32_t mone, uint32_t padw)
{
    using G = CostGeom2<TX>;
    constexpr int D = G::D, WPP = G::WPP, PS = G::PS, LPP = G::LPP, WPL = G::WPL, PTW = G::PTW, NG = G::NG, SL = G::SL;
    constexpr int radius = RAD, bs = 2 * RAD + 1;
    static_assert(G::DIAG && RAD >= 0, "diagonal-sweep shapes with a compile-time window only");
    extern __shared__ __align__(16) uint32_t smem[];
    const int W1 = W - Dv;
    const int part = (int)(blockIdx.x % (unsigned)NSPLIT), dbase = part * D;   // this CTA's disparities: dbase .. dbase + D - 1
    const int t0 = (int)(blockIdx.x / (unsigned)NSPLIT) * TX, b = blockIdx.z;
    const int y0 = blockIdx.y * band_rows, y1 = min(H, y0 + band_rows);
    const int e_lo = max(t0 - radius, 0), e_hi = min(t0 + TX - 1 + radius, W1 - 1);
    const int n_e = e_hi - e_lo + 1;
    const int xs = (e_lo + Dv - dbase - (D - 2)) & ~3;   // image pixel of table word 0 (may be negative: the zero margin)
    const int xoff = e_lo + Dv - dbase - xs;         // table index of (pixel e_lo, word 0); index(el, w) = xoff + el - 2w
    const uint32_t copy_bytes = (uint32_t)((xoff + n_e + 3) & ~3) * 4u;
    uint32_t one = 0u - mone;
    asm volatile("" : "+r"(one));

    uint32_t* Pt = smem;                             // [2 buffers][6 quantities][PTW]
    uint32_t* Lt = Pt + 2 * 6 * PTW;                 // [2 buffers][NE][8]
    uint32_t* pix = Lt + 2 * G::NE * 8;              // [NE][PS]
    uint32_t* ring = pix + G::NE * PS;               // [bs][NG][2][WPP][4]
    const uint32_t bar0 = smem_addr(ring + bs * TX * WPP);
    {
        uint4* r4 = reinterpret_cast<uint4*>(ring);
        for (int i = threadIdx.x; i < bs * TX * WPP / 4; i += 256) r4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int j_begin = y0 - radius, j_end = y1 + radius;
    const int r_first = max(j_begin, 0), r_last = min(j_end - 1, H - 1);
    const bool spare = threadIdx.x >= 224;
    const int sl = threadIdx.x - 224;                // lane of the spare warp

    auto issue = [&](int r, int buf) {               // one thread: the six table rows of image row r
        const uint32_t bar = bar0 + 8 * buf;
        mbar_expect(bar, 6 * copy_bytes);
        const uint32_t* src = ptab + ((size_t)b * H + r) * 6 * (size_t)pitch + margin + xs;
#pragma unroll
        for (int t = 0; t < 6; ++t) tma_load_1d(smem_addr(Pt + (buf * 6 + t) * PTW), src + (size_t)t * pitch, copy_bytes, bar);
    };
    if (threadIdx.x == 224) {
        mbar_init1(bar0);
        mbar_init1(bar0 + 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue(r_first, 0);
        if (r_first + 1 <= r_last) issue(r_first + 1, 1);
    }
    // left records: spare warp, two per lane, one row ahead
    uint2 ql[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
    auto fetch_left = [&](int r) {
        const uint2* src = recL + ((size_t)b * H + r) * W + e_lo + Dv;
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (sl + 32 * k < n_e) ql[k] = src[sl + 32 * k];
    };
    auto build_left = [&](int buf) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = sl + 32 * k;
            if (i < n_e) {
                const uint2 q = ql[k];
                const uint32_t ug = q.x & 0xffu, log_ = (q.x >> 8) & 0xffu, hig = (q.x >> 16) & 0xffu;
                const uint32_t ur = q.x >> 24, lor = q.y & 0xffu, hir = (q.y >> 8) & 0xffu;
                uint4* dst = reinterpret_cast<uint4*>(Lt + (buf * G::NE + i) * 8);
                dst[0] = make_uint4((ug + kKb) * 0x10001u, (kKb - ug) * 0x10001u, (kKb - hig) * 0x10001u, (log_ + kKb) * 0x10001u);
                dst[1] = make_uint4((ur + kKb) * 0x10001u, (kKb - ur) * 0x10001u, (kKb - hir) * 0x10001u, (lor + kKb) * 0x10001u);
            }
        }
    };
    if (spare) {
        fetch_left(r_first);
        build_left(0);
        if (r_first + 1 <= r_last) fetch_left(r_first + 1);
    }

    // phase-2 identity: 8 consecutive columns x one word
    const int cgp = threadIdx.x / WPP;
    const int w2 = threadIdx.x - cgp * WPP;
    const int c0 = t0 + cgp * 8;
    const bool interior = c0 - RAD >= 0 && c0 + 7 + RAD <= W1 - 1;
    const int nvalid = min(8, W1 - c0);              // <= 0: this thread has no columns
    const bool padded_word = PAD && dbase + 2 * w2 >= Dv;
    uint32_t crun[8], hs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { crun[i] = 0u; hs[i] = 0u; }
    constexpr int colw = WPP * NSPLIT;               // words per column of the layout
    uint32_t* dst = reinterpret_cast<uint32_t*>(C) + (((size_t)b * H + y0) * W1 + c0) * colw + part * WPP + w2;
    const size_t rowstep = (size_t)W1 * colw;
    uint4* const ring_t = reinterpret_cast<uint4*>(ring) + (size_t)cgp * 2 * WPP + w2;   // this thread's cell of slot 0
    __syncthreads();                                 // ring zeroed, Lt[0] built, mbarriers initialised

    auto bt_word = [&](const uint4& la, const uint4& lb, uint32_t loG, uint32_t hiG, uint32_t vG, uint32_t loR, uint32_t hiR,
                       uint32_t vR) -> uint32_t {
        const uint32_t g0 = __vimax3_s16x2(hiG * mone + la.x, loG * one + la.y, kKw);
        const uint32_t g1 = __vimax3_s16x2(vG * one + la.z, vG * mone + la.w, kKw);
        const uint32_t r0 = __vimax3_s16x2(hiR * mone + lb.x, loR * one + lb.y, kKw);
        const uint32_t r1 = __vimax3_s16x2(vR * one + lb.z, vR * mone + lb.w, kKw);
        const uint32_t cgv = __vmins2(g0, g1), crv = __vmins2(r0, r1);
        return cgv + ((crv >> 2) & 0x3fff3fffu) + kUnbias;
    };

    auto phase2b = [&](int jj, int slot_) {          // vertical running sum against the ring, output of row jj - radius
        if (nvalid > 0) {
            uint4* rg = ring_t + (size_t)slot_ * (NG * 2 * WPP);
            const uint4 o0 = rg[0], o1 = rg[WPP];
            rg[0] = make_uint4(hs[0], hs[1], hs[2], hs[3]);
            rg[WPP] = make_uint4(hs[4], hs[5], hs[6], hs[7]);
            crun[0] = crun[0] + hs[0] - o0.x; crun[1] = crun[1] + hs[1] - o0.y; crun[2] = crun[2] + hs[2] - o0.z; crun[3] = crun[3] + hs[3] - o0.w;
            crun[4] = crun[4] + hs[4] - o1.x; crun[5] = crun[5] + hs[5] - o1.y; crun[6] = crun[6] + hs[6] - o1.z; crun[7] = crun[7] + hs[7] - o1.w;
            if (jj - radius >= y0) {
                if (nvalid == 8) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) dst[c * colw] = padded_word ? padw : crun[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < nvalid) dst[c * colw] = padded_word ? padw : crun[c];
                }
                dst += rowstep;
            }
        }
    };

    int slot = 0, prev_r = -1;
    for (int j = j_begin; j < j_end; ++j) {
        const int r = min(max(j, 0), H - 1);
        const bool fresh = r != prev_r;              // CTA-uniform
        prev_r = r;
        const int k = r - r_first, buf = k & 1;
        if (fresh) {
            mbar_wait(bar0 + 8 * buf, (uint32_t)(k >> 1) & 1u);
            const uint32_t* P = Pt + buf * 6 * PTW + xoff;
            const uint32_t* L = Lt + buf * G::NE * 8;
            // ---- phase 1: pixel costs.  Diagonal sweep as in k_cost_fused (word w of pixel e and word w + 1 of pixel e + 2 read the
            // same right-image words), closed into a ring: the lane whose top word runs off the disparity range at step t takes the
            // word that enters at the bottom (w = jl + t - LPP) -- six fresh table loads for that lane, then the sweep property
            // holds for it as well.  No left-over items: the eighth warp only prepares the next row's left tables.
            if (!spare) {
                const int sg = threadIdx.x >> 4, jl = threadIdx.x & 15;
                const int el0 = (sg >> 1) * 2 * SL + (sg & 1);
                const bool work = el0 < n_e;
                uint32_t loG[NKK], hiG[NKK], vG[NKK], loR[NKK], hiR[NKK], vR[NKK];
                if (work) {
                    const uint32_t* q = P + el0 - 2 * jl;
#pragma unroll
                    for (int kk = 0; kk < NKK; ++kk) {
                        loG[kk] = q[0 * PTW - 2 * kk * LPP]; hiG[kk] = q[1 * PTW - 2 * kk * LPP]; vG[kk] = q[2 * PTW - 2 * kk * LPP];
                        loR[kk] = q[3 * PTW - 2 * kk * LPP]; hiR[kk] = q[4 * PTW - 2 * kk * LPP]; vR[kk] = q[5 * PTW - 2 * kk * LPP];
                    }
#pragma unroll
                    for (int t = 0; t < SL; ++t) {
                        const int el = el0 + 2 * t;
                        if (el < n_e) {
                            if (t > 0 && jl + t == LPP) {
                                const uint32_t* q2 = P + el;         // word 0 of pixel el
                                loG[NKK - 1] = q2[0]; hiG[NKK - 1] = q2[PTW]; vG[NKK - 1] = q2[2 * PTW];
                                loR[NKK - 1] = q2[3 * PTW]; hiR[NKK - 1] = q2[4 * PTW]; vR[NKK - 1] = q2[5 * PTW];
                            }
                            const uint4 la = reinterpret_cast<const uint4*>(L + el * 8)[0];
                            const uint4 lb = reinterpret_cast<const uint4*>(L + el * 8)[1];
                            uint32_t* out = pix + el * PS;
#pragma unroll
                            for (int kk = 0; kk < NKK; ++kk) {
                                // the ring of NKK * LPP words: the top group's lanes that run off it re-enter at the bottom
                                const int wi = (kk == NKK - 1 && jl + t >= LPP) ? jl + t - LPP : kk * LPP + jl + t;
                                out[wi] = bt_word(la, lb, loG[kk], hiG[kk], vG[kk], loR[kk], hiR[kk], vR[kk]);
                            }
                        }
                    }
                }
            } else if (r + 1 <= r_last) {
                build_left(buf ^ 1);
                if (r + 2 <= r_last) fetch_left(r + 2);
            }
            __syncthreads();                         // tile (and the next row's left table) complete; table buffer `buf` is free
            // ---- phase 2a: horizontal window sums of my 8 columns (registers): the middle one as a tree, the others outwards from it
            if (nvalid > 0) {
                uint32_t v[8 + 2 * RAD];
                if (interior) {
                    const uint32_t* src = pix + (c0 - RAD - e_lo) * PS + w2;
#pragma unroll
                    for (int i = 0; i < 8 + 2 * RAD; ++i) v[i] = src[i * PS];
                } else {
#pragma unroll
                    for (int i = 0; i < 8 + 2 * RAD; ++i) {
                        const int e = min(max(c0 + i - RAD, 0), W1 - 1) - e_lo;
                        v[i] = pix[e * PS + w2];
                    }
                }
                uint32_t part[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < 2 * RAD + 1; ++i) part[i & 3] += v[3 + i];
                hs[3] = (part[0] + part[1]) + (part[2] + part[3]);
#pragma unroll
                for (int c = 4; c < 8; ++c) hs[c] = hs[c - 1] + v[c + 2 * RAD] - v[c - 1];
#pragma unroll
                for (int c = 2; c >= 0; --c) hs[c] = hs[c + 1] + v[c] - v[c + 2 * RAD + 1];
            }
            __syncthreads();                         // tile consumed
            if (threadIdx.x == 224 && r + 2 <= r_last) issue(r + 2, buf);   // behind the barrier: the eighth warp has the slack
        }
        phase2b(j, slot);                            // flows into the next row's phase 1 without a barrier
        slot = slot + 1 == bs ? 0 : slot + 1;
    }
}

// ------------------------------------------------------------------------------------------------
static bool use_fused_cost(const ssm_ctx* c)
{
    const int D = c->dp.Dl;
    return !c->force_legacy_cost && (D == 16 || D == 32 || D == 64 || D == 128 || D == 256 || D == 512);
}

// k_cost_tma: 128- and 256-disparity layouts (D = 128 / 256, or 80 ... 112 / 144 ... 240 padded), the reference's block size
static bool use_cost_tma(const ssm_ctx* c)
{
    return use_fused_cost(c) && !c->no_cost_tma && c->d_ptab && (c->dp.Dl == 128 || c->dp.Dl == 256) && c->dp.bs == 11;
}

int launch_prefilter(ssm_ctx* c, int B, const uint8_t* dL, const uint8_t* dR, cudaStream_t s)
{
    const DevParams& p = c->dp;
    if (use_cost_tma(c)) {
        const int quads_per_row = (p.W + 3) / 4;
        const size_t total_quads = (size_t)B * p.H * quads_per_row;
        dim3 grid((unsigned)((total_quads + 255) / 256), 2);
        k_prefilter_tab<<<grid, 256, 0, s>>>(dL, dR, reinterpret_cast<uint2*>(c->d_recL), c->d_ptab, p.W, p.H, p.ftzero, quads_per_row, total_quads,
                                             (int)(((reinterpret_cast<uintptr_t>(dL) | reinterpret_cast<uintptr_t>(dR)) & 3) == 0), c->ptab_pitch,
                                             c->ptab_margin);
        SSM_LAUNCH_CHECK(c);
        return SSM_OK;
    }
    if (use_fused_cost(c)) {
        const int quads_per_row = (p.W + 3) / 4;
        const size_t total_quads = (size_t)B * p.H * quads_per_row;
        dim3 grid((unsigned)((total_quads + 255) / 256), 2);
        k_prefilter8<<<grid, 256, 0, s>>>(dL, dR, reinterpret_cast<uint2*>(c->d_recL), reinterpret_cast<uint2*>(c->d_recR), p.W, p.H,
                                          p.ftzero, quads_per_row, total_quads,
                                          (int)(((reinterpret_cast<uintptr_t>(dL) | reinterpret_cast<uintptr_t>(dR)) & 3) == 0));
        SSM_LAUNCH_CHECK(c);
        return SSM_OK;
    }
    const size_t total = (size_t)B * p.H * p.W;
    dim3 grid((unsigned)((total + 255) / 256), 2);
    k_prefilter<<<grid, 256, 0, s>>>(dL, dR, c->d_recL, c->d_recR, p.W, p.H, p.ftzero, total);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

template <int TX, int RAD, bool PAD = false>
static int launch_cost_fused_t(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const int radius = p.bs / 2;
    const size_t smem = sizeof(uint32_t) * CostGeom<TX>::smem_words(p.bs);
    SSM_CUDA(cudaFuncSetAttribute(k_cost_fused<TX, RAD, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // bands: enough CTAs to fill the machine a few times over, but tall enough to amortise the 2*radius halo rows
    const int tiles = (p.W1 + TX - 1) / TX;
    int bands = std::max(1, std::min(p.H / 32, (c->sm_count * 8 + tiles * B - 1) / (tiles * B)));
    const int band_rows = (p.H + bands - 1) / bands;
    bands = (p.H + band_rows - 1) / band_rows;
    dim3 grid(tiles, bands, B);
    k_cost_fused<TX, RAD, PAD><<<grid, 256, smem, s>>>(reinterpret_cast<const uint2*>(c->d_recL), reinterpret_cast<const uint2*>(c->d_recR), c->d_C, p.W, p.H, p.D, radius, band_rows, 0xffffffffu,
                                                       (kBig - (uint32_t)p.P2) * 0x10001u);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

template <int TX, int RAD, bool PAD, int NKK = CostGeom<TX>::WPL, int NSPLIT = 1>
static int launch_cost_tma_t(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t smem = CostGeom2<TX>::smem_bytes(p.bs);
    SSM_CUDA(cudaFuncSetAttribute(k_cost_tma<TX, RAD, PAD, NKK, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (NSPLIT = 2: 256-disparity layouts, two CTAs per tile with 128 disparities each -- halo 1.31 x instead of the 1.62 x of a 16-column tile)
    const int tiles = (p.W1 + TX - 1) / TX * NSPLIT;
    // bands of rows per tile: every band re-computes 2 * radius rows of pixel costs, so as few as still give every SM four CTAs
    // (33 KITTI frames: one band, 1155 CTAs, 2.037 ms -- two bands 2.072 ms, three 2.110 ms)
    int bands = std::max(1, std::min(p.H / 32, (c->sm_count * 4 + tiles * B - 1) / (tiles * B)));
    static const int force_bands = [] { const char* e = getenv("SSM_COST_BANDS"); return e ? atoi(e) : 0; }();
    if (force_bands > 0) bands = std::min(force_bands, std::max(1, p.H / 16));
    const int band_rows = (p.H + bands - 1) / bands;
    bands = (p.H + band_rows - 1) / band_rows;
    dim3 grid(tiles, bands, B);
    k_cost_tma<TX, RAD, PAD, NKK, NSPLIT><<<grid, 256, smem, s>>>(reinterpret_cast<const uint2*>(c->d_recL), c->d_ptab, c->d_C, p.W, p.H, p.D, band_rows,
                                                     c->ptab_pitch, c->ptab_margin, 0xffffffffu, (kBig - (uint32_t)p.P2) * 0x10001u);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_cost_volume(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    if (use_cost_tma(c)) {
        if (p.Dl == 256) return p.Dl == p.D ? launch_cost_tma_t<32, 5, false, 4, 2>(c, B, s) : launch_cost_tma_t<32, 5, true, 4, 2>(c, B, s);
        if (p.Dl == p.D) return launch_cost_tma_t<32, 5, false>(c, B, s);
        // padded layouts: 80 and 96 disparities (the reference's default is 80) fill three of the four 16-word groups of a pixel
        return p.D <= 96 ? launch_cost_tma_t<32, 5, true, 3>(c, B, s) : launch_cost_tma_t<32, 5, true>(c, B, s);
    }
    if (use_fused_cost(c)) {
        // TX * D/2 = 2048 words per CTA row
        const bool r5 = p.bs == 11;   // the reference's block size gets the compile-time window; others the generic one
        if (p.Dl != p.D) {            // padded layouts (only formed for block size 11, api.cu: layout_disparities)
            switch (p.Dl) {
                case 64: return launch_cost_fused_t<64, 5, true>(c, B, s);
                case 128: return launch_cost_fused_t<32, 5, true>(c, B, s);
                case 256: return launch_cost_fused_t<16, 5, true>(c, B, s);
                default: break;
            }
        }
        switch (p.Dl) {
            case 16: return r5 ? launch_cost_fused_t<256, 5>(c, B, s) : launch_cost_fused_t<256, -1>(c, B, s);
            case 32: return r5 ? launch_cost_fused_t<128, 5>(c, B, s) : launch_cost_fused_t<128, -1>(c, B, s);
            case 64: return r5 ? launch_cost_fused_t<64, 5>(c, B, s) : launch_cost_fused_t<64, -1>(c, B, s);
            case 128: return r5 ? launch_cost_fused_t<32, 5>(c, B, s) : launch_cost_fused_t<32, -1>(c, B, s);
            case 256: return r5 ? launch_cost_fused_t<16, 5>(c, B, s) : launch_cost_fused_t<16, -1>(c, B, s);
            case 512: return r5 ? launch_cost_fused_t<8, 5>(c, B, s) : launch_cost_fused_t<8, -1>(c, B, s);
            default: break;   // other multiples of 16: the two-kernel path below
        }
    }
    const int radius = p.bs / 2;
    const int tw = (kTX + 2 * kMaxR + p.D) / 2 + 4;
    const size_t smem = sizeof(uint32_t) * ((size_t)16 * tw + (kTX + 2 * kMaxR) * 8 + (size_t)(kTX + 2 * kMaxR) * (p.D / 2));
    static bool attr_done = false;
    if (!attr_done) {
        SSM_CUDA(cudaFuncSetAttribute(k_pix_hsum, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    dim3 grid((p.W1 + kTX - 1) / kTX, p.H, B);
    k_pix_hsum<<<grid, 256, smem, s>>>(c->d_recL, c->d_recR, c->d_hs, p.W, p.H, p.D, radius);
    SSM_LAUNCH_CHECK(c);

    const int band_rows = 64;
    const int nbands = (p.H + band_rows - 1) / band_rows;
    const int rowwords4 = p.W1 * p.D / 8;
    const size_t total = (size_t)B * nbands * rowwords4;
    k_vsum<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->d_hs, c->d_C, p.H, rowwords4, radius, band_rows, nbands,
                                                            total);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

}  // namespace ssm
