// cues.cu -- dense motion cues from the disparity map, sm_100a.
//
// The other dense consumer of calDisparity_SGBM's output in the reference, run per frame on the tracker thread
// (/root/reference src/track.cpp:67-79):
//   triangulate10D                  src/stereo.cpp:41-118      disparity + grey image -> [H][W][10] fp32 record image
//   UVDisparity::calVDisparity      src/uvdisparity.cpp:277-366  per-row disparity histogram, 8-bit map, channel 8
//   correct3DPoints                 src/stereo.cpp:127-181     pitch rotation of Y/Z, ROI clears the intensity channel
//   setImageROI                     src/stereo.cpp:183-192     intensity channel -> 8-bit ROI mask
//   UVDisparity::calUDisparity      src/uvdisparity.cpp:195-274  per-column disparity histogram, 8-bit map, channel 7
// (Pitch_Classify / the Kalman filters between them are sequential host code and stay with the reference.)
//
// The record image is 40 bytes per pixel; every kernel that touches it moves whole 128-pixel blocks (5120 contiguous
// bytes) through shared memory so that global accesses are coalesced 16-byte vectors, and the two stages below touch
// it once each:
//   stage 1  min/max -> V histogram -> 8-bit V map -> k_tri10d writes all ten channels (channel 8 included)
//   stage 2  k_correct_roi_uhist: rotate, ROI, roi mask, U histogram in one read-modify-write -> 8-bit U map ->
//            k_assign_u writes channel 7
// fp64 expressions are evaluated left to right without contraction (the library is built with --fmad=false), so the
// fp32 results are bit-identical to the reference's.  The reference's out-of-bounds accesses are given the canonical
// meaning stated in DESIGN.md (section 2) and at the kernels below.
#include <algorithm>
#include <cmath>
#include <string>

#include "ssm_internal.cuh"

namespace ssm {

constexpr int kCueBlock = 128;              // pixels per block
constexpr int kRec = 10;                    // floats per pixel record

struct CueWs {
    size_t cap_pix = 0;                     // pixels of one frame the host-call workspace holds
    size_t cap_hist = 0;                    // histogram elements it holds
    uint8_t* d_img = nullptr;
    int16_t* d_disp = nullptr;
    float* d_xyz = nullptr;
    uint8_t *d_roi = nullptr, *d_ground = nullptr;
    int32_t* d_hist = nullptr;              // V: [H][v_cols] flat;  U: [u_rows][W]
    uint8_t* d_hist8 = nullptr;
    int32_t* d_mm = nullptr;                // [max_batch][2] min, max of the disparity map
    int mm_cap = 0;
};

// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_cue_mm_init(int32_t* mm, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { mm[2 * i] = 32767; mm[2 * i + 1] = -32768; }
}

// min / max of each frame's disparity map (cv::minMaxIdx at stereo.cpp:58-59, uvdisparity.cpp:197-198, :279-280)
__global__ void __launch_bounds__(256) k_cue_minmax(const int16_t* __restrict__ disp, size_t npix, int32_t* __restrict__ mm)
{
    const int b = blockIdx.y;
    const int16_t* d = disp + (size_t)b * npix;
    int lo = 32767, hi = -32768;
    // 8 values per 16-byte load where aligned; scalar tail
    const size_t nvec = ((npix * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) ? npix / 8 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(d)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = (int16_t)(w[k] & 0xffffu), c = (int16_t)(w[k] >> 16);
            lo = min(lo, min(a, c)); hi = max(hi, max(a, c));
        }
    }
    for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const int a = d[i];
        lo = min(lo, a); hi = max(hi, a);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * b], lo); atomicMax(&mm[2 * b + 1], hi); }
}

// cvCeil(max / 16) clamped at 0 (a negative size is an empty map)
__device__ __forceinline__ int cols_of_max(int mx)
{
    const int c = (int)ceil((double)mx / 16);
    return max(c, 0);
}
// cvRound((float)d / 16.0f): round half to even
__device__ __forceinline__ int round_disp(int d16) { return __float2int_rn(__fdiv_rn((float)d16, 16.0f)); }

// V histogram: one block per (row, frame); bins in shared memory, flushed to the flat [H][v_cols] matrix.  Bin
// id == v_cols is the flat element after the row's last one (uvdisparity.cpp:307-311), dropped past the matrix end.
__global__ void __launch_bounds__(256) k_cue_vhist(const int16_t* __restrict__ disp, const int32_t* __restrict__ mm, int32_t* __restrict__ vint,
                                                   int W, int H, size_t hist_stride /* elements per frame */, int cap_cols,
                                                   uint32_t* __restrict__ overflow)
{
    extern __shared__ int32_t bins[];
    const int i = blockIdx.x, b = blockIdx.y;
    const int v_cols = cols_of_max(mm[2 * b + 1]);
    if (v_cols > cap_cols) { if (threadIdx.x == 0) atomicOr(overflow, 1u); return; }
    for (int k = threadIdx.x; k <= v_cols; k += blockDim.x) bins[k] = 0;
    __syncthreads();
    const int16_t* row = disp + ((size_t)b * H + i) * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        const int d = row[j];
        if (d > 0) atomicAdd(&bins[max(0, min(v_cols, round_disp(d)))], 1);
    }
    __syncthreads();
    int32_t* out = vint + (size_t)b * hist_stride;
    const size_t n = (size_t)H * v_cols;
    for (int k = threadIdx.x; k <= v_cols; k += blockDim.x) {
        const size_t flat = (size_t)i * v_cols + k;
        if (bins[k] && flat < n) atomicAdd(&out[flat], bins[k]);
    }
}

// 8-bit map = (uchar)(count * scale): truncation toward zero, modulo 256 (uvdisparity.cpp:235-247, :319-333).
// which = 0: V map, n = H * v_cols;  which = 1: U map, n = u_rows * W.
__global__ void __launch_bounds__(256) k_cue_scale(const int32_t* __restrict__ hist, uint8_t* __restrict__ hist8, const int32_t* __restrict__ mm,
                                                   int W, int H, size_t hist_stride, int which, float scale)
{
    const int b = blockIdx.y;
    const int c = cols_of_max(mm[2 * b + 1]);
    const size_t n = which == 0 ? (size_t)H * c : (size_t)(c + 1) * W;
    const int32_t* src = hist + (size_t)b * hist_stride;
    uint8_t* dst = hist8 + (size_t)b * hist_stride;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        dst[k] = (uint8_t)(int)__fmul_rn((float)src[k], scale);
}

// coalesced transfer of one block's records between global and shared memory (kCueBlock * 10 floats; 16-byte vectors
// when the block is full and aligned)
__device__ __forceinline__ void recs_store(float* __restrict__ g, const float* s, int npx)
{
    if (npx == kCueBlock && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        for (int k = threadIdx.x; k < kCueBlock * kRec / 4; k += kCueBlock) reinterpret_cast<float4*>(g)[k] = reinterpret_cast<const float4*>(s)[k];
    } else {
        for (int k = threadIdx.x; k < npx * kRec; k += kCueBlock) g[k] = s[k];
    }
}
__device__ __forceinline__ void recs_load(float* s, const float* __restrict__ g, int npx)
{
    if (npx == kCueBlock && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        for (int k = threadIdx.x; k < kCueBlock * kRec / 4; k += kCueBlock) reinterpret_cast<float4*>(s)[k] = reinterpret_cast<const float4*>(g)[k];
    } else {
        for (int k = threadIdx.x; k < npx * kRec; k += kCueBlock) s[k] = g[k];
    }
}

// triangulate10D (stereo.cpp:41-118) with channel 8 of calVDisparity (uvdisparity.cpp:341-362) when v8 != nullptr.
__global__ void __launch_bounds__(kCueBlock) k_cue_tri10d(const uint8_t* __restrict__ img, const int16_t* __restrict__ disp,
                                                          const int32_t* __restrict__ mm, const uint8_t* __restrict__ v8,
                                                          float* __restrict__ xyz, int W, int H, size_t hist_stride, double f, double cx,
                                                          double cy, double bl)
{
    __shared__ __align__(16) float rec[kCueBlock * kRec];
    const int b = blockIdx.y;
    const size_t npix = (size_t)W * H;
    const size_t p0 = (size_t)blockIdx.x * kCueBlock;
    const int npx = (int)min((size_t)kCueBlock, npix - p0);
    const int t = threadIdx.x;
    if (t < npx) {
        const size_t p = p0 + t;
        const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
        const int d = disp[b * npix + p];
        const double pw = __ddiv_rn(bl, (double)d);
        double px = __dmul_rn(__dmul_rn(__dsub_rn((double)j, cx), pw), 16.0);
        double py = __dmul_rn(__dmul_rn(__dsub_rn((double)i, cy), pw), 16.0);
        double pz = __dmul_rn(__dmul_rn(f, pw), 16.0);
        if (d == mm[2 * b]) px = py = pz = __longlong_as_double(0x7ff0000000000000ll);   // |d - min| <= FLT_EPSILON on integers
        float* o = rec + t * kRec;
        o[0] = __double2float_rn(px); o[1] = __double2float_rn(py); o[2] = __double2float_rn(pz);
        o[3] = (float)j; o[4] = (float)i;
        o[5] = __fdiv_rn((float)d, 16.0f);
        o[6] = (float)(int)img[b * npix + p];
        o[7] = 0.f;
        float c8 = 0.f;
        if (v8) {
            const int dr = round_disp(d);          // cvRound(channel 5); channel 4 is the row
            if (dr > 0) {
                const int v_cols = cols_of_max(mm[2 * b + 1]);
                const size_t n = (size_t)H * v_cols;
                const uint8_t* m = v8 + (size_t)b * hist_stride;
                uint32_t word = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const size_t off = (size_t)i * v_cols + 4 * (size_t)dr + q;
                    if (off < n) word |= (uint32_t)m[off] << (8 * q);
                }
                c8 = (float)(int)word;
            }
        }
        o[8] = c8;
        o[9] = 0.f;
    }
    __syncthreads();
    recs_store(xyz + ((size_t)b * npix + p0) * kRec, rec, npx);
}

// channel 8 alone on an existing record image (the stateless calVDisparity entry point): reads channels 4, 5
__global__ void __launch_bounds__(kCueBlock) k_cue_assign_v(float* __restrict__ xyz, const uint8_t* __restrict__ v8, const int32_t* __restrict__ mm,
                                                            int W, int H, size_t hist_stride)
{
    __shared__ __align__(16) float rec[kCueBlock * kRec];
    const int b = blockIdx.y;
    const size_t npix = (size_t)W * H;
    const size_t p0 = (size_t)blockIdx.x * kCueBlock;
    const int npx = (int)min((size_t)kCueBlock, npix - p0);
    float* g = xyz + ((size_t)b * npix + p0) * kRec;
    recs_load(rec, g, npx);
    __syncthreads();
    if (threadIdx.x < npx) {
        float* o = rec + threadIdx.x * kRec;
        const int v = __float2int_rn(o[4]), dr = __float2int_rn(o[5]);
        float c8 = 0.f;
        if (dr > 0) {
            const int v_cols = cols_of_max(mm[2 * b + 1]);
            const size_t n = (size_t)H * v_cols;
            const uint8_t* m = v8 + (size_t)b * hist_stride;
            uint32_t word = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long off = (long long)v * v_cols + 4 * (long long)dr + q;
                if (off >= 0 && (size_t)off < n) word |= (uint32_t)m[off] << (8 * q);
            }
            c8 = (float)(int)word;
        }
        o[8] = c8;
    }
    __syncthreads();
    recs_store(g, rec, npx);
}

// correct3DPoints (stereo.cpp:127-181) + setImageROI (:183-192) + the counting loop of calUDisparity
// (uvdisparity.cpp:209-231) in one read-modify-write of the record image.  Any of the three parts can be switched off.
__global__ void __launch_bounds__(kCueBlock) k_cue_correct_roi_uhist(float* __restrict__ xyz, const int16_t* __restrict__ disp,
                                                                     const uint8_t* __restrict__ roi_in, const uint8_t* __restrict__ ground,
                                                                     uint8_t* __restrict__ roi_out, int32_t* __restrict__ uint_, const int32_t* __restrict__ mm,
                                                                     int W, int H, size_t hist_stride, int cap_rows, int do_correct,
                                                                     double cos_p1, double sin_p1, double roi_x, double roi_y, double roi_z,
                                                                     uint32_t* __restrict__ overflow)
{
    __shared__ __align__(16) float rec[kCueBlock * kRec];
    const int b = blockIdx.y;
    const size_t npix = (size_t)W * H;
    const size_t p0 = (size_t)blockIdx.x * kCueBlock;
    const int npx = (int)min((size_t)kCueBlock, npix - p0);
    float* g = xyz ? xyz + ((size_t)b * npix + p0) * kRec : nullptr;
    if (g) recs_load(rec, g, npx);
    __syncthreads();
    if (threadIdx.x < npx) {
        const size_t p = p0 + threadIdx.x;
        float* o = rec + threadIdx.x * kRec;
        if (g && do_correct) {
            const float yp = o[1], zp = o[2];
            const int d = __float2int_rn(o[5]);
            if (d > 0 && d < 100) {
                o[1] = __double2float_rn(__dadd_rn(__dmul_rn(cos_p1, (double)yp), __dmul_rn(sin_p1, (double)zp)));
                o[2] = __double2float_rn(__dsub_rn(__dmul_rn(cos_p1, (double)zp), __dmul_rn(sin_p1, (double)yp)));
                if ((double)o[0] > roi_x || (double)o[1] > roi_y || (double)o[2] > roi_z) o[6] = 0.f;
            } else {
                o[6] = 0.f;
            }
        }
        int roi = 0;
        if (g) {
            roi = min(__float2int_rn(fabsf(o[6])), 255);     // convertScaleAbs: saturate_cast<uchar>(round(|x|))
            if (roi_out) roi_out[b * npix + p] = (uint8_t)roi;
        } else if (roi_in) {
            roi = roi_in[b * npix + p];
        }
        if (uint_) {
            const int u_rows = cols_of_max(mm[2 * b + 1]) + 1;
            if (u_rows > cap_rows) {
                atomicOr(overflow, 2u);
            } else {
                const int d16 = disp[b * npix + p];
                const int dis = d16 / 16;                    // cvRound(d/16): integer division first (uvdisparity.cpp:219)
                const bool gm = ground ? ground[b * npix + p] > 0 : true;
                if (d16 > 0 && roi > 0 && gm && dis > 0) atomicAdd(&uint_[(size_t)b * hist_stride + (size_t)dis * W + (p % W)], 1);
            }
        }
    }
    __syncthreads();
    if (g && do_correct) recs_store(g, rec, npx);
}

// channel 7 = u_dis_(round(channel 5), round(channel 3)) (uvdisparity.cpp:257-271); a negative row reads 0
__global__ void __launch_bounds__(kCueBlock) k_cue_assign_u(float* __restrict__ xyz, const uint8_t* __restrict__ u8, const int32_t* __restrict__ mm,
                                                            int W, int H, size_t hist_stride)
{
    __shared__ __align__(16) float rec[kCueBlock * kRec];
    const int b = blockIdx.y;
    const size_t npix = (size_t)W * H;
    const size_t p0 = (size_t)blockIdx.x * kCueBlock;
    const int npx = (int)min((size_t)kCueBlock, npix - p0);
    float* g = xyz + ((size_t)b * npix + p0) * kRec;
    recs_load(rec, g, npx);
    __syncthreads();
    if (threadIdx.x < npx) {
        float* o = rec + threadIdx.x * kRec;
        const int u = __float2int_rn(o[3]), d = __float2int_rn(o[5]);
        const int u_rows = cols_of_max(mm[2 * b + 1]) + 1;
        o[7] = (d >= 0 && d < u_rows && u >= 0 && u < W) ? (float)u8[(size_t)b * hist_stride + (size_t)d * W + u] : 0.f;
    }
    __syncthreads();
    recs_store(g, rec, npx);
}

// ---------------------------------------------------------------------------------------------------------------------
static int cue_fail(int code, const std::string& msg)
{
    set_error(msg);
    return code;
}

static CueWs* cue_ws(ssm_ctx* c)
{
    if (!c->cues_ws) c->cues_ws = new CueWs();
    return static_cast<CueWs*>(c->cues_ws);
}

void cues_free(ssm_ctx* c)
{
    CueWs* w = static_cast<CueWs*>(c->cues_ws);
    if (!w) return;
    void* ptrs[] = {w->d_img, w->d_disp, w->d_xyz, w->d_roi, w->d_ground, w->d_hist, w->d_hist8, w->d_mm};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete w;
    c->cues_ws = nullptr;
}

static int ensure_mm(CueWs* w, int B)
{
    if (B <= w->mm_cap) return SSM_OK;
    if (w->d_mm) cudaFree(w->d_mm);
    w->d_mm = nullptr; w->mm_cap = 0;
    SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_mm), sizeof(int32_t) * 2 * B));
    w->mm_cap = B;
    return SSM_OK;
}

// one-frame workspace of the host entry points; `bins` = histogram columns (V) / rows (U) needed
static int ensure_frame(CueWs* w, int W, int H, int bins)
{
    int rc;
    if ((rc = ensure_mm(w, 1))) return rc;
    const size_t npix = (size_t)W * H;
    if (npix > w->cap_pix) {
        void* ptrs[] = {w->d_img, w->d_disp, w->d_xyz, w->d_roi, w->d_ground};
        for (void* q : ptrs)
            if (q) cudaFree(q);
        w->d_img = nullptr; w->d_disp = nullptr; w->d_xyz = nullptr; w->d_roi = nullptr; w->d_ground = nullptr; w->cap_pix = 0;
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_img), npix));
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_disp), npix * 2));
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_xyz), npix * kRec * sizeof(float)));
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_roi), npix));
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_ground), npix));
        w->cap_pix = npix;
    }
    const size_t need = bins > 0 ? (size_t)(bins + 1) * std::max(W, H) + 8 : 0;   // V: H * bins + 4;  U: bins * W
    if (need > w->cap_hist) {
        if (w->d_hist) cudaFree(w->d_hist);
        if (w->d_hist8) cudaFree(w->d_hist8);
        w->d_hist = nullptr; w->d_hist8 = nullptr; w->cap_hist = 0;
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_hist), sizeof(int32_t) * need));
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_hist8), need));
        w->cap_hist = need;
    }
    return SSM_OK;
}

static int launch_minmax(ssm_ctx* c, CueWs* w, int B, const int16_t* d_disp, size_t npix, cudaStream_t s)
{
    k_cue_mm_init<<<(B + 255) / 256, 256, 0, s>>>(w->d_mm, B);
    SSM_LAUNCH_CHECK(c);
    const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((npix / 8 + 255) / 256 + 1, (size_t)c->sm_count * 4 / B + 1));
    k_cue_minmax<<<dim3(gx, B), 256, 0, s>>>(d_disp, npix, w->d_mm);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

static int check_dims(int w, int h)
{
    if (w < 1 || h < 1 || w > 65535 || h > 65535) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "image dimensions must be in [1, 65535]");
    return SSM_OK;
}

// stage 1 on device buffers: min/max, V histogram + 8-bit map (optional), record image incl. channel 8
static int cues_stage1(ssm_ctx* c, CueWs* w, int B, const uint8_t* d_img, const int16_t* d_disp, int W, int H, double f, double cx, double cy,
                       double bl, float* d_xyz, int32_t* d_vint, uint8_t* d_v8, size_t hist_stride, int cap_cols, cudaStream_t s)
{
    int rc;
    const size_t npix = (size_t)W * H;
    if ((rc = launch_minmax(c, w, B, d_disp, npix, s))) return rc;
    if (d_vint) {
        SSM_CUDA(cudaMemsetAsync(d_vint, 0, sizeof(int32_t) * hist_stride * B, s));
        SSM_CUDA(cudaMemsetAsync(c->d_counters + 4, 0, sizeof(uint32_t), s));
        k_cue_vhist<<<dim3(H, B), 256, sizeof(int32_t) * (cap_cols + 1), s>>>(d_disp, w->d_mm, d_vint, W, H, hist_stride, cap_cols, c->d_counters + 4);
        SSM_LAUNCH_CHECK(c);
        const float scale = 255 * 1.0f / (float)W;
        k_cue_scale<<<dim3((unsigned)std::min<size_t>((hist_stride + 255) / 256, 1024), B), 256, 0, s>>>(d_vint, d_v8, w->d_mm, W, H, hist_stride, 0, scale);
        SSM_LAUNCH_CHECK(c);
    }
    if (d_xyz) {
        k_cue_tri10d<<<dim3((unsigned)((npix + kCueBlock - 1) / kCueBlock), B), kCueBlock, 0, s>>>(d_img, d_disp, w->d_mm, d_vint ? d_v8 : nullptr, d_xyz, W, H,
                                                                                                 hist_stride, f, cx, cy, bl);
        SSM_LAUNCH_CHECK(c);
    }
    return SSM_OK;
}

// stage 2 on device buffers: correction + ROI mask + U histogram (one pass), 8-bit U map, channel 7
static int cues_stage2(ssm_ctx* c, CueWs* w, int B, const int16_t* d_disp, int W, int H, float* d_xyz, int do_correct, double pitch1,
                       const double roi[3], const uint8_t* d_roi_in, const uint8_t* d_ground, uint8_t* d_roi_out, int32_t* d_uint, uint8_t* d_u8,
                       size_t hist_stride, int cap_rows, bool assign, cudaStream_t s)
{
    const size_t npix = (size_t)W * H;
    const dim3 grid((unsigned)((npix + kCueBlock - 1) / kCueBlock), B);
    if (d_uint) {
        SSM_CUDA(cudaMemsetAsync(d_uint, 0, sizeof(int32_t) * hist_stride * B, s));
        SSM_CUDA(cudaMemsetAsync(c->d_counters + 4, 0, sizeof(uint32_t), s));
    }
    k_cue_correct_roi_uhist<<<grid, kCueBlock, 0, s>>>(d_xyz, d_disp, d_roi_in, d_ground, d_roi_out, d_uint, w->d_mm, W, H, hist_stride, cap_rows, do_correct,
                                                       std::cos(pitch1), std::sin(pitch1), roi[0], roi[1], roi[2], c->d_counters + 4);
    SSM_LAUNCH_CHECK(c);
    if (d_uint) {
        const float scale = 255 * 1.0f / (float)H;
        k_cue_scale<<<dim3((unsigned)std::min<size_t>((hist_stride + 255) / 256, 1024), B), 256, 0, s>>>(d_uint, d_u8, w->d_mm, W, H, hist_stride, 1, scale);
        SSM_LAUNCH_CHECK(c);
        if (assign && d_xyz) {
            k_cue_assign_u<<<grid, kCueBlock, 0, s>>>(d_xyz, d_u8, w->d_mm, W, H, hist_stride);
            SSM_LAUNCH_CHECK(c);
        }
    }
    return SSM_OK;
}

static int read_max(ssm_ctx* c, CueWs* w, cudaStream_t s, int* mx)
{
    int32_t mm[2];
    SSM_CUDA(cudaMemcpyAsync(mm, w->d_mm, sizeof(mm), cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    (void)c;
    *mx = mm[1];
    return SSM_OK;
}

}  // namespace ssm

using namespace ssm;

extern "C" {

int ssm_triangulate10d(ssm_ctx* c, const uint8_t* img, size_t img_stride, const int16_t* disp, size_t disp_stride, int w, int h, double f,
                       double cx, double cy, double b, double roi_x, double roi_y, double roi_z, float* xyz)
{
    (void)roi_x; (void)roi_y; (void)roi_z;   // both branches of the reference's ROI test store the same values (stereo.cpp:88-113)
    if (!c || !img || !disp || !xyz) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    CueWs* ws = cue_ws(c);
    const size_t npix = (size_t)w * h;
    if ((rc = ensure_frame(ws, w, h, 0))) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpy2DAsync(ws->d_img, w, img, img_stride, w, h, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpy2DAsync(ws->d_disp, (size_t)w * 2, disp, disp_stride, (size_t)w * 2, h, cudaMemcpyHostToDevice, s));
    if ((rc = cues_stage1(c, ws, 1, ws->d_img, ws->d_disp, w, h, f, cx, cy, b, ws->d_xyz, nullptr, nullptr, 0, 0, s))) return rc;
    SSM_CUDA(cudaMemcpyAsync(xyz, ws->d_xyz, npix * kRec * sizeof(float), cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_correct_3d_points(ssm_ctx* c, float* xyz, int w, int h, double roi_x, double roi_y, double roi_z, double pitch1, double pitch2)
{
    (void)pitch2;                            // accepted and unused, as in the reference (stereo.cpp:127-181)
    if (!c || !xyz) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    CueWs* ws = cue_ws(c);
    const size_t npix = (size_t)w * h;
    if ((rc = ensure_frame(ws, w, h, 0))) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpyAsync(ws->d_xyz, xyz, npix * kRec * sizeof(float), cudaMemcpyHostToDevice, s));
    const double roi[3] = {roi_x, roi_y, roi_z};
    if ((rc = cues_stage2(c, ws, 1, nullptr, w, h, ws->d_xyz, 1, pitch1, roi, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, false, s))) return rc;
    SSM_CUDA(cudaMemcpyAsync(xyz, ws->d_xyz, npix * kRec * sizeof(float), cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_set_image_roi(ssm_ctx* c, const float* xyz, int w, int h, uint8_t* roi_mask, size_t mask_stride)
{
    if (!c || !xyz || !roi_mask) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    CueWs* ws = cue_ws(c);
    const size_t npix = (size_t)w * h;
    if ((rc = ensure_frame(ws, w, h, 0))) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpyAsync(ws->d_xyz, xyz, npix * kRec * sizeof(float), cudaMemcpyHostToDevice, s));
    const double roi[3] = {0, 0, 0};
    if ((rc = cues_stage2(c, ws, 1, nullptr, w, h, ws->d_xyz, 0, 0.0, roi, nullptr, nullptr, ws->d_roi, nullptr, nullptr, 0, 0, false, s))) return rc;
    SSM_CUDA(cudaMemcpy2DAsync(roi_mask, mask_stride, ws->d_roi, w, w, h, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_v_disparity(ssm_ctx* c, const int16_t* disp, size_t disp_stride, int w, int h, float* xyz, int32_t* v_dis_int, uint8_t* v_dis,
                    int cap_cols, int* v_cols_out)
{
    if (!c || !disp) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    CueWs* ws = cue_ws(c);
    const size_t npix = (size_t)w * h;
    if ((rc = ensure_frame(ws, w, h, 0))) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpy2DAsync(ws->d_disp, (size_t)w * 2, disp, disp_stride, (size_t)w * 2, h, cudaMemcpyHostToDevice, s));
    if ((rc = launch_minmax(c, ws, 1, ws->d_disp, npix, s))) return rc;
    int mx;
    if ((rc = read_max(c, ws, s, &mx))) return rc;
    const int v_cols = std::max(0, (int)std::ceil((double)mx / 16));
    if (v_cols_out) *v_cols_out = v_cols;
    if ((v_dis_int || v_dis) && v_cols > cap_cols) return cue_fail(SSM_ERR_CAPACITY, "v-disparity map wider than cap_cols");
    if ((rc = ensure_frame(ws, w, h, v_cols))) return rc;
    const size_t hist_stride = (size_t)h * v_cols + 4;
    if ((rc = cues_stage1(c, ws, 1, nullptr, ws->d_disp, w, h, 0, 0, 0, 0, nullptr, ws->d_hist, ws->d_hist8, hist_stride, v_cols, s))) return rc;
    if (xyz) {
        SSM_CUDA(cudaMemcpyAsync(ws->d_xyz, xyz, npix * kRec * sizeof(float), cudaMemcpyHostToDevice, s));
        k_cue_assign_v<<<dim3((unsigned)((npix + kCueBlock - 1) / kCueBlock), 1), kCueBlock, 0, s>>>(ws->d_xyz, ws->d_hist8, ws->d_mm, w, h, hist_stride);
        SSM_LAUNCH_CHECK(c);
        SSM_CUDA(cudaMemcpyAsync(xyz, ws->d_xyz, npix * kRec * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (v_dis_int && v_cols) SSM_CUDA(cudaMemcpyAsync(v_dis_int, ws->d_hist, sizeof(int32_t) * h * v_cols, cudaMemcpyDeviceToHost, s));
    if (v_dis && v_cols) SSM_CUDA(cudaMemcpyAsync(v_dis, ws->d_hist8, (size_t)h * v_cols, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_u_disparity(ssm_ctx* c, const int16_t* disp, size_t disp_stride, int w, int h, float* xyz, const uint8_t* roi_mask,
                    const uint8_t* ground_mask, int32_t* u_dis_int, uint8_t* u_dis, int cap_rows, int* u_rows_out)
{
    if (!c || !disp || !roi_mask || !ground_mask) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    CueWs* ws = cue_ws(c);
    const size_t npix = (size_t)w * h;
    if ((rc = ensure_frame(ws, w, h, 0))) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpy2DAsync(ws->d_disp, (size_t)w * 2, disp, disp_stride, (size_t)w * 2, h, cudaMemcpyHostToDevice, s));
    if ((rc = launch_minmax(c, ws, 1, ws->d_disp, npix, s))) return rc;
    int mx;
    if ((rc = read_max(c, ws, s, &mx))) return rc;
    const int u_rows = std::max(0, (int)std::ceil((double)mx / 16)) + 1;
    if (u_rows_out) *u_rows_out = u_rows;
    if ((u_dis_int || u_dis) && u_rows > cap_rows) return cue_fail(SSM_ERR_CAPACITY, "u-disparity map taller than cap_rows");
    if ((rc = ensure_frame(ws, w, h, u_rows))) return rc;
    SSM_CUDA(cudaMemcpyAsync(ws->d_roi, roi_mask, npix, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(ws->d_ground, ground_mask, npix, cudaMemcpyHostToDevice, s));
    if (xyz) SSM_CUDA(cudaMemcpyAsync(ws->d_xyz, xyz, npix * kRec * sizeof(float), cudaMemcpyHostToDevice, s));
    const size_t hist_stride = (size_t)u_rows * w;
    const double roi[3] = {0, 0, 0};
    // counting from the given masks (no correction, no ROI rewrite), then the 8-bit map and channel 7
    if ((rc = cues_stage2(c, ws, 1, ws->d_disp, w, h, nullptr, 0, 0.0, roi, ws->d_roi, ws->d_ground, nullptr, ws->d_hist, ws->d_hist8, hist_stride, u_rows,
                          false, s)))
        return rc;
    if (xyz) {
        k_cue_assign_u<<<dim3((unsigned)((npix + kCueBlock - 1) / kCueBlock), 1), kCueBlock, 0, s>>>(ws->d_xyz, ws->d_hist8, ws->d_mm, w, h, hist_stride);
        SSM_LAUNCH_CHECK(c);
        SSM_CUDA(cudaMemcpyAsync(xyz, ws->d_xyz, npix * kRec * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (u_dis_int) SSM_CUDA(cudaMemcpyAsync(u_dis_int, ws->d_hist, sizeof(int32_t) * hist_stride, cudaMemcpyDeviceToHost, s));
    if (u_dis) SSM_CUDA(cudaMemcpyAsync(u_dis, ws->d_hist8, hist_stride, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_motion_cues_stage1_device(ssm_ctx* c, int batch, const uint8_t* d_img, const int16_t* d_disp, int w, int h, double f, double cx,
                                  double cy, double b, float* d_xyz, int32_t* d_v_dis_int, uint8_t* d_v_dis, size_t hist_stride, int cap_cols,
                                  void* stream)
{
    if (!c || !d_img || !d_disp || !d_xyz || batch < 1) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    if ((d_v_dis_int == nullptr) != (d_v_dis == nullptr)) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "v_dis_int and v_dis go together");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    if (d_v_dis_int && (hist_stride < (size_t)h * cap_cols + 4 || cap_cols < 1 || cap_cols > 8192))
        return cue_fail(SSM_ERR_INVALID_ARGUMENT, "hist_stride must be >= h * cap_cols + 4");
    CueWs* ws = cue_ws(c);
    if ((rc = ensure_mm(ws, batch))) return rc;
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return cues_stage1(c, ws, batch, d_img, d_disp, w, h, f, cx, cy, b, d_xyz, d_v_dis_int, d_v_dis, hist_stride, cap_cols, s);
}

int ssm_motion_cues_stage2_device(ssm_ctx* c, int batch, const int16_t* d_disp, int w, int h, float* d_xyz, double roi_x, double roi_y,
                                  double roi_z, double pitch1, const uint8_t* d_ground_mask, uint8_t* d_roi_mask, int32_t* d_u_dis_int,
                                  uint8_t* d_u_dis, size_t hist_stride, int cap_rows, void* stream)
{
    if (!c || !d_disp || !d_xyz || !d_roi_mask || batch < 1) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    if ((d_u_dis_int == nullptr) != (d_u_dis == nullptr)) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "u_dis_int and u_dis go together");
    int rc;
    if ((rc = check_dims(w, h))) return rc;
    if (d_u_dis_int && hist_stride < (size_t)cap_rows * w) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "hist_stride must be >= cap_rows * w");
    CueWs* ws = cue_ws(c);
    if (batch > ws->mm_cap) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "stage 2 follows stage 1 of the same batch");
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    const double roi[3] = {roi_x, roi_y, roi_z};
    return cues_stage2(c, ws, batch, d_disp, w, h, d_xyz, 1, pitch1, roi, nullptr, d_ground_mask, d_roi_mask, d_u_dis_int, d_u_dis, hist_stride, cap_rows, true, s);
}

int ssm_motion_cues_overflow(ssm_ctx* c, int* flags)
{
    if (!c || !flags) return cue_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    uint32_t f = 0;
    SSM_CUDA(cudaMemcpy(&f, c->d_counters + 4, sizeof(f), cudaMemcpyDeviceToHost));
    *flags = (int)f;
    return SSM_OK;
}

}  // extern "C"
