// sgbm_cost.cu -- matching-cost stage of the SGBM chain, sm_100a.
//
// Replaces the calcPixelCostBT + block-sum part of cv::StereoSGBM that /root/reference
// src/stereo.cpp:13-30 calls.  Three kernels:
//   k_prefilter   image -> per-pixel record {v,-v,lo,-hi} for the Sobel-x and the raw channel (A-1, A-2 intervals)
//   k_pix_hsum    records -> Birchfield-Tomasi pixel cost in packed s16x2 lanes (VIADDMNMX/VIMNMX), staged as a
//                 shared-memory tile, then the bs-wide horizontal window sum hs[y][x'][d] (A-2, A-3 first half)
//   k_vsum        bs-tall running window sum down the rows -> C[y][x'][d] int16 (A-3 second half)
// All arithmetic is integer and bit-exact with the oracle.
#include "ssm_internal.cuh"

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
int launch_prefilter(ssm_ctx* c, int B, const uint8_t* dL, const uint8_t* dR, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t total = (size_t)B * p.H * p.W;
    dim3 grid((unsigned)((total + 255) / 256), 2);
    k_prefilter<<<grid, 256, 0, s>>>(dL, dR, c->d_recL, c->d_recR, p.W, p.H, p.ftzero, total);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_cost_volume(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
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
