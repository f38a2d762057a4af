// sgbm_select.cu -- disparity selection and post-filters, sm_100a.
//
// Replaces the tail of cv::StereoSGBM (called from /root/reference src/stereo.cpp:30):
//   (winner-take-all itself is fused into the horizontal sweep, sgbm_hsweep.cu)
//   k_lrcheck     left-right consistency (A-6); also writes the always-invalid columns [0, D)
//   k_median3     cv::medianBlur(disp, 3) on int16 with replicate border
//   k_cc_*        cv::filterSpeckles == connected components under |a-b| <= maxDiff; union-find with atomicMin
#include "ssm_internal.cuh"

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
int launch_select(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t npix = (size_t)B * p.H * p.W;
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
    k_median3<<<grid, 256, 0, s>>>(c->d_disp_lr, c->d_disp_med, p.W, p.H, npix);
    SSM_LAUNCH_CHECK(c);
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
