// sgbm_select.cu -- disparity selection and post-filters, sm_100a.
//
// Replaces the tail of cv::StereoSGBM (called from /root/reference src/stereo.cpp:30); winner-take-all itself is fused
// into the horizontal sweep (sgbm_hsweep2.cu / sgbm_hsweep.cu), which leaves one record per pixel.
//   k_select_fused      records -> sub-pixel disparity + disp2 -> L-R check (A-6) -> cv::medianBlur(3) -> speckle components
//                       inside a band of rows, all in shared memory (checkpointed-sweep record format, D <= 128)
//   k_cc_merge_bands, k_cc_count_roots, k_cc_apply_bands
//                       the rest of cv::filterSpeckles: join components across bands, total their sizes, filter
//   k_lrcheck, k_median3, k_cc_init / merge / count / apply
//                       the same steps as separate kernels through HBM (D > 128, SSM_LEGACY_SELECT=1); union-find with atomicMin
#include "sgbm_wta.cuh"

namespace ssm {

__global__ void __launch_bounds__(256) k_lrcheck(const int16_t* __restrict__ disp_raw, const uint32_t* __restrict__ disp2key,
                                                 int16_t* __restrict__ disp_lr, int W, int D, int d12, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const size_t rowbase = idx - x;
    int out = kInvalidDisp;
    if (x >= D) {
        const int d1 = disp_raw[idx];
        out = d1;
        if (d1 != kInvalidDisp) {
            const int dlo = d1 >> 4, dhi = (d1 + kDispScale - 1) >> 4;
            const int xlo = x - dlo, xhi = x - dhi;
            bool bad_lo = false, bad_hi = false;
            if (xlo >= 0 && xlo < W) {
                const uint32_t k = disp2key[rowbase + xlo];
                if (k != 0xffffffffu) bad_lo = abs((0xffff - (int)(k & 0xffffu)) - xlo - dlo) > d12;
            }
            if (xhi >= 0 && xhi < W) {
                const uint32_t k = disp2key[rowbase + xhi];
                if (k != 0xffffffffu) bad_hi = abs((0xffff - (int)(k & 0xffffu)) - xhi - dhi) > d12;
            }
            if (bad_lo && bad_hi) out = kInvalidDisp;
        }
    }
    disp_lr[idx] = (int16_t)out;
}

// ------------------------------------------------------------------------------------------------
// 3x3 median, replicate border (9-element sorting network, min/max only)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(int& a, int& b)
{
    const int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
__global__ void __launch_bounds__(256) k_median3(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int W, int H,
                                                 size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const size_t r = idx / W;
    const int y = (int)(r % H);
    const int16_t* img = src + (r / H) * (size_t)W * H;
    const int xm = max(x - 1, 0), xq = min(x + 1, W - 1);
    const int16_t* r0 = img + (size_t)max(y - 1, 0) * W;
    const int16_t* r1 = img + (size_t)y * W;
    const int16_t* r2 = img + (size_t)min(y + 1, H - 1) * W;
    int p0 = r0[xm], p1 = r0[x], p2 = r0[xq], p3 = r1[xm], p4 = r1[x], p5 = r1[xq], p6 = r2[xm], p7 = r2[x], p8 = r2[xq];
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    dst[idx] = (int16_t)p4;
}

// ------------------------------------------------------------------------------------------------
// speckle filter: label = smallest linear index of the component (lock-free union-find)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cc_find(int* label, int x)
{
    int p = __ldcg(label + x);   // L2 reads: parents change under concurrent atomicMin hooks
    while (p != x) {
        x = p;
        p = __ldcg(label + x);
    }
    return x;
}
__device__ __forceinline__ void cc_union(int* label, int a, int b)
{
    while (true) {
        a = cc_find(label, a);
        b = cc_find(label, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // a > b: hook the larger root under the smaller
        const int old = atomicMin(&label[a], b);
        if (old == a) return;
        a = old;
    }
}
__device__ __forceinline__ bool cc_conn(int a, int b, int max_diff)
{
    return a != kInvalidDisp && b != kInvalidDisp && abs(a - b) <= max_diff;
}
// initial label = start of the pixel's horizontal run inside its 32-pixel segment (one warp per segment)
__global__ void __launch_bounds__(256) k_cc_init(const int16_t* __restrict__ img, int* __restrict__ label,
                                                 int* __restrict__ size, int W, int max_diff, int segs_per_row,
                                                 size_t total_segs)
{
    const int lane = threadIdx.x & 31;
    const size_t seg = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= total_segs) return;
    const size_t row = seg / segs_per_row;
    const int x0 = (int)(seg % segs_per_row) * 32, x = x0 + lane;
    const size_t idx = row * W + x;
    const int v = x < W ? (int)img[idx] : kInvalidDisp;
    const int vl = __shfl_up_sync(0xffffffffu, v, 1);
    const bool head = v != kInvalidDisp && !(lane > 0 && cc_conn(v, vl, max_diff));
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (x < W) {
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        label[idx] = v == kInvalidDisp ? -1 : (int)(row * W + x0 + start);
        size[idx] = 0;
    }
}
// unions: across segment boundaries in a row, and between rows -- skipping a vertical edge whenever the
// left neighbour's vertical edge plus the two horizontal edges already connect the same pair
__global__ void __launch_bounds__(256) k_cc_merge(const int16_t* __restrict__ img, int* __restrict__ label, int W, int H,
                                                  int max_diff, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int v = img[idx];
    if (v == kInvalidDisp) return;
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int vl = x > 0 ? (int)img[idx - 1] : kInvalidDisp;
    const bool cl = cc_conn(v, vl, max_diff);
    if (cl && (x & 31) == 0) cc_union(label, (int)idx, (int)idx - 1);
    if (y > 0) {
        const int vu = img[idx - W];
        if (cc_conn(v, vu, max_diff)) {
            bool skip = false;
            if (cl) {
                const int vul = img[idx - W - 1];
                skip = cc_conn(vl, vul, max_diff) && cc_conn(vul, vu, max_diff);
            }
            if (!skip) cc_union(label, (int)idx, (int)idx - W);
        }
    }
}
// component sizes, only as far as the decision "size <= max_size" needs them: one atomic per run of equal roots
// inside a warp, and none once the counter is already past the threshold
__global__ void __launch_bounds__(256) k_cc_count(int* __restrict__ label, int* __restrict__ size, int max_size, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int root = -1;
    if (idx < total && label[idx] >= 0) {
        root = cc_find(label, (int)idx);
        label[idx] = root;   // path compression; roots are fixed points, so concurrent finds stay correct
    }
    const int rp = __shfl_up_sync(0xffffffffu, root, 1);
    const bool head = lane == 0 || root != rp;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (root >= 0 && head) {
        const uint32_t after = lane == 31 ? 0u : (heads >> (lane + 1));
        const int run = after ? __ffs(after) : 32 - lane;
        if (__ldcg(size + root) <= max_size) atomicAdd(&size[root], run);
    }
}
__global__ void __launch_bounds__(256) k_cc_apply(const int16_t* __restrict__ img, const int* __restrict__ label,
                                                  const int* __restrict__ size, int16_t* __restrict__ out, int max_size,
                                                  size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int v = img[idx];
    const int l = label[idx];
    if (l >= 0 && size[cc_find(const_cast<int*>(label), l)] <= max_size) v = kInvalidDisp;
    out[idx] = (int16_t)v;
}

// ------------------------------------------------------------------------------------------------
// Fused selection (checkpointed-sweep record format, D <= 128): one CTA per band of R image rows does, entirely in
// shared memory, what the five kernels above do through HBM:
//   records -> raw disparity + disp2 candidates (shared-memory atomicMin; no 4-byte key image, no memset)
//   -> L-R check in place -> 3x3 median, two pixels per thread on packed s16x2 min / max
//   -> speckle components INSIDE the band: run-start labels, unions with shared-memory atomics, flattening, sizes.
// The rows above and below the band are recomputed (R + 2 rows of records per R rows of output).  What leaves the
// CTA: disp_lr (debug tap), disp_med, label = global index of the pixel's band-local root (-1: invalid pixel),
// size = band-local component size at roots, 0 elsewhere.  Components are completed across band borders by
// k_cc_merge_bands (one row of vertical edges per band) and k_cc_count_roots (one atomic per hooked band root).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kNoLabel = 0xffffffffu;

__device__ __forceinline__ uint32_t s_find(const volatile uint32_t* lab, uint32_t x)
{
    uint32_t p = lab[x];
    while (p != x) {
        x = p;
        p = lab[x];
    }
    return x;
}
__device__ __forceinline__ void s_union(uint32_t* lab, uint32_t a, uint32_t b)
{
    while (true) {
        a = s_find(lab, a);
        b = s_find(lab, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}
__device__ __forceinline__ void cswap2(uint32_t& a, uint32_t& b)
{
    const uint32_t lo = __vmins2(a, b), hi = __vmaxs2(a, b);
    a = lo; b = hi;
}

struct SelArgs {
    const uint4* rec;
    int16_t *disp_lr, *disp_med;
    int *label, *size;
    int W, H, D, d12, R, WS, max_diff, do_cc;
    uint32_t mW1, mW, mHp;    // ceil(2^32 / d) for d = W1, W, (W + 1) / 2: i / d == umulhi(i, m) for every index the kernel forms
};
__device__ __forceinline__ void divmod_magic(int i, uint32_t m, int d, int& q, int& r)
{
    q = (int)__umulhi((uint32_t)i, m);
    r = i - q * d;
}

template <int NR>
__global__ void __launch_bounds__(512) k_select_fused(SelArgs a)
{
    extern __shared__ __align__(16) uint32_t sel_smem[];
    const int W = a.W, H = a.H, D = a.D, W1 = W - D, R = a.R, WS = a.WS;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int y0 = blockIdx.x * R, b = blockIdx.y;
    const int nr = min(R, H - y0), ns = nr + 2;                               // band rows; slots (slot j <-> image row y0 - 1 + j, clamped)
    uint32_t* keys = sel_smem;                                                // [R + 2][W]   disp2 candidates (minS << 16 | 0xffff - x)
    int16_t* dsp = reinterpret_cast<int16_t*>(keys + (size_t)(R + 2) * W);    // [R + 2][WS]  element x + 2 <-> pixel x, replicated borders
    int16_t* med = dsp + (size_t)(R + 2) * WS;                                // [R][W]
    const size_t frame_row0 = (size_t)b * H;

    for (int i = tid; i < ns * W; i += nt) keys[i] = 0xffffffffu;
    __syncthreads();
    // ---- records -> raw disparity, disp2 candidates; two records in flight per thread
    {
        constexpr int RQ = WtaRec<NR>::kQuads;
        const uint4* recf = a.rec + frame_row0 * W1 * RQ;
        const int n1 = ns * W1;
        auto item = [&](const uint4& r, const uint4& rb, int j, int xp) {
            int minS, best;
            bool valid;
            const int out = wta2_decode<NR>(r, rb, D, minS, best, valid);
            const int x = xp + D;
            if (valid) atomicMin(&keys[j * W + (x - best)], ((uint32_t)minS << 16) | (uint32_t)(0xffff - x));
            dsp[j * WS + x + 2] = (int16_t)out;
        };
        for (int i = tid; i < n1; i += 2 * nt) {
            const int i2 = i + nt;
            const bool two = i2 < n1;
            int j, xp, j2, xp2;
            divmod_magic(i, a.mW1, W1, j, xp);
            divmod_magic(two ? i2 : i, a.mW1, W1, j2, xp2);
            const int ia = (min(max(y0 - 1 + j, 0), H - 1) * W1 + xp) * RQ, ib = (min(max(y0 - 1 + j2, 0), H - 1) * W1 + xp2) * RQ;
            const uint4 r = recf[ia], r2 = recf[ib];
            const uint4 rb = NR == 4 ? recf[ia + 1] : r, r2b = NR == 4 ? recf[ib + 1] : r2;
            item(r, rb, j, xp);
            if (two) item(r2, r2b, j2, xp2);
        }
    }
    __syncthreads();
    // ---- L-R check in place (A-6); columns [0, D) are always invalid; border replication for the median
    {
        int16_t* lrg = a.disp_lr + (frame_row0 + y0) * W;
        const int n2 = ns * W;
        for (int i = tid; i < n2; i += nt) {
            int j, x;
            divmod_magic(i, a.mW, W, j, x);
            int16_t* row = dsp + j * WS;
            int out = kInvalidDisp;
            if (x >= D) {
                const int d1 = row[x + 2];
                out = d1;
                if (d1 != kInvalidDisp) {
                    const uint32_t* krow = keys + j * W;
                    const int dlo = d1 >> 4, dhi = (d1 + kDispScale - 1) >> 4;
                    const int xlo = x - dlo, xhi = x - dhi;   // 0 < x - (D - 1) <= xhi <= xlo <= x: always inside the row
                    const uint32_t klo = krow[xlo], khi = krow[xhi];
                    const bool bad_lo = klo != 0xffffffffu && abs((0xffff - (int)(klo & 0xffffu)) - xlo - dlo) > a.d12;
                    const bool bad_hi = khi != 0xffffffffu && abs((0xffff - (int)(khi & 0xffffu)) - xhi - dhi) > a.d12;
                    if (bad_lo && bad_hi) out = kInvalidDisp;
                }
            }
            row[x + 2] = (int16_t)out;
            if (x == 0) { row[0] = (int16_t)out; row[1] = (int16_t)out; }
            if (x == W - 1)
                for (int e = W + 2; e < WS; ++e) row[e] = (int16_t)out;
            if (j >= 1 && j <= nr) lrg[(j - 1) * W + x] = (int16_t)out;
        }
    }
    __syncthreads();
    // ---- 3x3 median (cv::medianBlur(3), replicate border), pixels (x, x + 1) in the two halves of packed words
    {
        const int hp = (W + 1) >> 1;
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(dsp);
        int16_t* mg = a.disp_med + (frame_row0 + y0) * W;
        for (int i = tid; i < nr * hp; i += nt) {
            int r, k;
            divmod_magic(i, a.mHp, hp, r, k);
            const int x = 2 * k;
            uint32_t p[9];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const uint32_t* q = dw + (((r + t) * WS + x) >> 1);
                const uint32_t A0 = q[0], A1 = q[1], A2 = q[2];
                p[3 * t] = __byte_perm(A0, A1, 0x5432);        // pixels (x - 1, x)
                p[3 * t + 1] = A1;                             // pixels (x, x + 1)
                p[3 * t + 2] = __byte_perm(A1, A2, 0x5432);    // pixels (x + 1, x + 2)
            }
            cswap2(p[1], p[2]); cswap2(p[4], p[5]); cswap2(p[7], p[8]); cswap2(p[0], p[1]); cswap2(p[3], p[4]); cswap2(p[6], p[7]);
            cswap2(p[1], p[2]); cswap2(p[4], p[5]); cswap2(p[7], p[8]); cswap2(p[0], p[3]); cswap2(p[5], p[8]); cswap2(p[4], p[7]);
            cswap2(p[3], p[6]); cswap2(p[1], p[4]); cswap2(p[2], p[5]); cswap2(p[4], p[7]); cswap2(p[4], p[2]); cswap2(p[6], p[4]);
            cswap2(p[4], p[2]);
            const int16_t m0 = (int16_t)(p[4] & 0xffffu), m1 = (int16_t)(p[4] >> 16);
            const int li = r * W + x;
            med[li] = m0;
            mg[li] = m0;
            if (x + 1 < W) { med[li + 1] = m1; mg[li + 1] = m1; }
        }
    }
    if (!a.do_cc) return;
    __syncthreads();
    // ---- speckle components inside the band.  label = smallest linear index (r * W + x) of the component's pixels
    uint32_t* lab = keys;                                       // [nr * W], over the key rows
    uint32_t* szw = reinterpret_cast<uint32_t*>(dsp);           // band-local sizes as packed u16 pairs, over the disparity rows
    const int npx = nr * W;
    const int md = a.max_diff;
    for (int i = tid; i < (npx + 1) / 2; i += nt) szw[i] = 0u;
    // initial label = start of the pixel's horizontal run: one warp walks a row, the open run is carried across its 32-pixel steps
    for (int r = warp; r < nr; r += nwarps) {
        const int16_t* mrow = med + r * W;
        uint32_t* lrow = lab + r * W;
        int carry_start = 0, carry_v = kInvalidDisp;            // start of the run the previous step ended in, its last value
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            const int v = x < W ? (int)mrow[x] : kInvalidDisp;
            int vl = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) vl = carry_v;
            const bool head = v != kInvalidDisp && !cc_conn(v, vl, md);
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            const uint32_t upto = heads & (0xffffffffu >> (31 - lane));
            const int start = upto ? x0 + 31 - __clz(upto) : carry_start;
            if (x < W) lrow[x] = v == kInvalidDisp ? kNoLabel : (uint32_t)(r * W + start);
            carry_start = __shfl_sync(0xffffffffu, start, 31);
            carry_v = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // vertical edges (rows 1 .. nr - 1), skipping an edge whenever the left neighbour's vertical edge and the two horizontal
    // edges already connect the same pair
    for (int li = W + tid; li < npx; li += nt) {
        const int v = med[li];
        if (v == kInvalidDisp) continue;
        const int vu = med[li - W];
        if (!cc_conn(v, vu, md)) continue;
        int r, x;
        divmod_magic(li, a.mW, W, r, x);
        if (x > 0) {
            const int vl = med[li - 1], vul = med[li - W - 1];
            if (cc_conn(v, vl, md) && cc_conn(vl, vul, md) && cc_conn(vul, vu, md)) continue;
        }
        s_union(lab, (uint32_t)li, (uint32_t)(li - W));
    }
    __syncthreads();
    // run heads look their root up; every other pixel then reaches it in two loads (its run's head, the head's root)
    for (int li = tid; li < npx; li += nt) {
        const int v = med[li];
        if (v == kInvalidDisp) continue;
        int r, x;
        divmod_magic(li, a.mW, W, r, x);
        if (x > 0 && cc_conn(v, (int)med[li - 1], md)) continue;
        lab[li] = s_find(lab, (uint32_t)li);
    }
    __syncthreads();
    {
        int* lg = a.label + (frame_row0 + y0) * W;
        const int gbase = (int)((frame_row0 + y0) * W);
        for (int base = tid - lane; base < npx; base += nt) {   // warp-uniform trip count
            const int li = base + lane;
            uint32_t root = kNoLabel;
            if (li < npx) {
                const uint32_t h = lab[li];
                if (h != kNoLabel) root = lab[h];
                lg[li] = root == kNoLabel ? -1 : gbase + (int)root;
            }
            const uint32_t rp = __shfl_up_sync(0xffffffffu, root, 1);
            const bool head = lane == 0 || root != rp;
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            if (root != kNoLabel && head) {
                const uint32_t after = lane == 31 ? 0u : (heads >> (lane + 1));
                const uint32_t run = after ? (uint32_t)__ffs(after) : (uint32_t)(32 - lane);
                atomicAdd(&szw[root >> 1], run << ((root & 1u) * 16));   // sizes <= R * W < 65536: no carry between the halves
            }
        }
    }
    __syncthreads();
    {
        int* sg = a.size + (frame_row0 + y0) * W;
        const uint16_t* sz16 = reinterpret_cast<const uint16_t*>(szw);
        for (int li = tid; li < npx; li += nt) sg[li] = (int)sz16[li];   // non-zero exactly at the band-local roots
    }
}

// vertical edges between the last row of a band and the first row of the next (same redundant-edge rule as k_cc_merge)
__global__ void __launch_bounds__(256) k_cc_merge_bands(const int16_t* __restrict__ img, int* __restrict__ label, int W, int H, int R,
                                                        int inner /* bands per frame - 1 */, int max_diff, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const size_t t = idx / W;
    const int y = ((int)(t % inner) + 1) * R;
    const size_t p = ((t / inner) * H + y) * W + x;
    const int v = img[p];
    if (v == kInvalidDisp) return;
    const int vu = img[p - W];
    if (!cc_conn(v, vu, max_diff)) return;
    if (x > 0) {
        const int vl = img[p - 1];
        if (cc_conn(v, vl, max_diff)) {
            const int vul = img[p - W - 1];
            if (cc_conn(vl, vul, max_diff) && cc_conn(vul, vu, max_diff)) return;
        }
    }
    cc_union(label, (int)p, (int)(p - W));
}
// band roots hooked under another band's root hand their size to the component's root and point straight at it
__global__ void __launch_bounds__(256) k_cc_count_roots(int* __restrict__ label, int* __restrict__ size, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int s = size[idx];
    if (s <= 0) return;
    const int g = cc_find(label, (int)idx);
    if (g != (int)idx) {
        atomicAdd(&size[g], s);   // only roots of whole components receive additions; a hooked root keeps its band-local size
        label[idx] = g;
    }
}
// a pixel stays unless its component has at most max_size pixels; a band-local size above the bound already settles it
template <bool VEC>
__global__ void __launch_bounds__(256) k_cc_apply_bands(const int16_t* __restrict__ img, const int* __restrict__ label,
                                                        const int* __restrict__ size, int16_t* __restrict__ out, int max_size,
                                                        size_t total)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= total) return;
    int v[4], l[4];
    const int n = (int)min((size_t)4, total - i0);
    if (VEC && n == 4) {
        const uint2 vv = *reinterpret_cast<const uint2*>(img + i0);
        const int4 ll = *reinterpret_cast<const int4*>(label + i0);
        v[0] = (int16_t)(vv.x & 0xffffu); v[1] = (int16_t)(vv.x >> 16); v[2] = (int16_t)(vv.y & 0xffffu); v[3] = (int16_t)(vv.y >> 16);
        l[0] = ll.x; l[1] = ll.y; l[2] = ll.z; l[3] = ll.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = k < n ? (int)img[i0 + k] : kInvalidDisp;
            l[k] = k < n ? label[i0 + k] : -1;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (l[k] < 0) continue;
        int s = size[l[k]];
        if (s <= max_size) {
            const int g = cc_find(const_cast<int*>(label), l[k]);
            if (g != l[k]) s = size[g];
            if (s <= max_size) v[k] = kInvalidDisp;
        }
    }
    if (VEC && n == 4) {
        uint2 o;
        o.x = (uint32_t)(v[0] & 0xffff) | ((uint32_t)(v[1] & 0xffff) << 16);
        o.y = (uint32_t)(v[2] & 0xffff) | ((uint32_t)(v[3] & 0xffff) << 16);
        *reinterpret_cast<uint2*>(out + i0) = o;
    } else {
        for (int k = 0; k < n; ++k) out[i0 + k] = (int16_t)v[k];
    }
}

// rows per band of the fused selection kernel (0: not applicable -> the separate kernels run)
static int select_fused_rows(const ssm_ctx* c, size_t* smem_out)
{
    const DevParams& p = c->dp;
    // index / d as one multiply-high needs d >= 2 and index * d < 2^32 for every index formed (at most (R + 2) * W + 2 * 512)
    if (c->force_legacy_select || !hsweep2_supported(c) || p.W1 < 2 || p.W > 8192) return 0;
    const int WS = (p.W + 4 + 1) & ~1;
    const int first = c->select_rows > 0 ? c->select_rows : 8;
    for (int R = first; R >= 1; R >>= 1) {
        const size_t smem = (size_t)(R + 2) * p.W * 4 + (size_t)(R + 2) * WS * 2 + (size_t)R * p.W * 2 + 16;
        if ((size_t)R * p.W < 65535 && smem <= (R > 1 ? 110u : 220u) * 1024u) {
            if (smem_out) *smem_out = smem;
            return R;
        }
    }
    return 0;
}

template <int NR>
static int launch_select_fused_t(ssm_ctx* c, int B, int R, size_t smem, cudaStream_t s)
{
    const DevParams& p = c->dp;
    SelArgs a;
    a.rec = reinterpret_cast<const uint4*>(c->d_wta_rec);
    a.disp_lr = c->d_disp_lr; a.disp_med = c->d_disp_med; a.label = c->d_cc_label; a.size = c->d_cc_size;
    a.W = p.W; a.H = p.H; a.D = p.D; a.d12 = p.d12; a.R = R; a.WS = (p.W + 4 + 1) & ~1;
    a.max_diff = p.speckle_diff; a.do_cc = p.speckle_win > 0;
    auto magic = [](int d) { return (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
    a.mW1 = magic(p.W1); a.mW = magic(p.W); a.mHp = magic((p.W + 1) / 2);
    SSM_CUDA(cudaFuncSetAttribute(k_select_fused<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.H + R - 1) / R), (unsigned)B);
    k_select_fused<NR><<<grid, 512, smem, s>>>(a);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

// ------------------------------------------------------------------------------------------------
int launch_select(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t npix = (size_t)B * p.H * p.W;
    size_t smem = 0;
    if (const int R = select_fused_rows(c, &smem))   // records -> L-R checked, median-filtered disparity + band-local speckle labels
        return p.Dl <= 64 ? launch_select_fused_t<1>(c, B, R, smem, s)
                          : (p.Dl <= 128 ? launch_select_fused_t<2>(c, B, R, smem, s) : launch_select_fused_t<4>(c, B, R, smem, s));
    SSM_CUDA(cudaMemsetAsync(c->d_disp2key, 0xff, npix * sizeof(uint32_t), s));
    int rc = hsweep2_supported(c) ? launch_wta_finalize2(c, B, s) : launch_wta_finalize(c, B, s);
    if (rc) return rc;
    k_lrcheck<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(c->d_disp_raw, c->d_disp2key, c->d_disp_lr, p.W, p.D, p.d12, npix);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_post(ssm_ctx* c, int B, int16_t* d_out, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t npix = (size_t)B * p.H * p.W;
    const unsigned grid = (unsigned)((npix + 255) / 256);
    const int R = select_fused_rows(c, nullptr);
    if (R && p.speckle_win > 0) {
        // the fused selection kernel left the median image, band-local labels and sizes: join the bands, total the sizes, filter
        const int inner = (p.H + R - 1) / R - 1;
        if (inner > 0) {
            const size_t edges = (size_t)B * inner * p.W;
            k_cc_merge_bands<<<(unsigned)((edges + 255) / 256), 256, 0, s>>>(c->d_disp_med, c->d_cc_label, p.W, p.H, R, inner, p.speckle_diff, edges);
            SSM_LAUNCH_CHECK(c);
            k_cc_count_roots<<<grid, 256, 0, s>>>(c->d_cc_label, c->d_cc_size, npix);
            SSM_LAUNCH_CHECK(c);
        }
        const unsigned grid4 = (unsigned)(((npix + 3) / 4 + 255) / 256);
        const bool vec = (reinterpret_cast<uintptr_t>(c->d_disp_med) % 8 == 0) && (reinterpret_cast<uintptr_t>(d_out) % 8 == 0) &&
                         (reinterpret_cast<uintptr_t>(c->d_cc_label) % 16 == 0);
        if (vec) k_cc_apply_bands<true><<<grid4, 256, 0, s>>>(c->d_disp_med, c->d_cc_label, c->d_cc_size, d_out, p.speckle_win, npix);
        else k_cc_apply_bands<false><<<grid4, 256, 0, s>>>(c->d_disp_med, c->d_cc_label, c->d_cc_size, d_out, p.speckle_win, npix);
        SSM_LAUNCH_CHECK(c);
        return SSM_OK;
    }
    if (!R) {
        k_median3<<<grid, 256, 0, s>>>(c->d_disp_lr, c->d_disp_med, p.W, p.H, npix);
        SSM_LAUNCH_CHECK(c);
    }
    if (p.speckle_win <= 0) {
        SSM_CUDA(cudaMemcpyAsync(d_out, c->d_disp_med, npix * sizeof(int16_t), cudaMemcpyDeviceToDevice, s));
        return SSM_OK;
    }
    // labels are linear indices over the whole batch, but merges never cross a frame (x/y bounds are per frame)
    const int segs_per_row = (p.W + 31) / 32;
    const size_t total_segs = (size_t)B * p.H * segs_per_row;
    k_cc_init<<<(unsigned)((total_segs + 7) / 8), 256, 0, s>>>(c->d_disp_med, c->d_cc_label, c->d_cc_size, p.W, p.speckle_diff,
                                                             segs_per_row, total_segs);
    SSM_LAUNCH_CHECK(c);
    k_cc_merge<<<grid, 256, 0, s>>>(c->d_disp_med, c->d_cc_label, p.W, p.H, p.speckle_diff, npix);
    SSM_LAUNCH_CHECK(c);
    k_cc_count<<<grid, 256, 0, s>>>(c->d_cc_label, c->d_cc_size, p.speckle_win, npix);
    SSM_LAUNCH_CHECK(c);
    k_cc_apply<<<grid, 256, 0, s>>>(c->d_disp_med, c->d_cc_label, c->d_cc_size, d_out, p.speckle_win, npix);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

}  // namespace ssm
