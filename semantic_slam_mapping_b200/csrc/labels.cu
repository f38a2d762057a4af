// labels.cu -- label production in front of the path, sm_100a.
//
// /root/reference experiment/segnet.cpp:121-135 (and the commented online path src/rgbdframe.cpp:118-136) turn SegNet's
// 480x360 argmax index image into the semantic colour image the Mapper reads:
//     cv::resize(index image, frame size)   default INTER_LINEAR on 8-bit data: 11-bit fixed-point weights,
//                                           horizontal pass in int, vertical pass ((b * (r >> 4)) >> 16), + 2 >> 2
//     cv::LUT(., color)                     256-entry BGR table (segnet.cpp:58-84)
//     cvtColor(BGR2GRAY) -> raw_semantic    the resized index itself (all three channels are equal)
// (bilinear interpolation of class INDICES produces in-between ids along class borders; that is the reference's
// behaviour and is reproduced bit for bit -- the result is pinned against cv2.resize / cv2.LUT.)
// The weight tables are built on the host with OpenCV's float expressions; one kernel does both passes and the lookup.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "ssm_internal.cuh"

namespace ssm {

struct LabelWs {
    int sw = 0, sh = 0, dw = 0, dh = 0;
    int4* d_xtab = nullptr;     // [dw] {x0, x1, a0, a1}
    int4* d_ytab = nullptr;     // [dh] {y0, y1, b0, b1}
    uint32_t* d_lut = nullptr;  // [256] 0x00RRGGBB as B | G << 8 | R << 16
    uint8_t* d_index = nullptr; // host-call staging
    uint8_t* d_sem = nullptr;
    uint8_t* d_raw = nullptr;
    size_t cap_src = 0, cap_dst = 0;
    uint8_t lut_host[768];
    bool lut_valid = false;
};

__global__ void __launch_bounds__(256) k_labels_resize_lut(const uint8_t* __restrict__ index, const int4* __restrict__ xtab,
                                                           const int4* __restrict__ ytab, const uint32_t* __restrict__ lut,
                                                           uint8_t* __restrict__ sem, uint8_t* __restrict__ raw, int sw, int sh, int dw, int dh)
{
    __shared__ uint32_t slut[256];
    slut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= dw) return;
    const int4 xt = xtab[x], yt = ytab[y];
    const uint8_t* src = index + (size_t)b * sw * sh;
    const uint8_t* r0 = src + (size_t)yt.x * sw;
    const uint8_t* r1 = src + (size_t)yt.y * sw;
    const int h0 = (int)r0[xt.x] * xt.z + (int)r0[xt.y] * xt.w;     // HResizeLinear: int rows, 11 fractional bits
    const int h1 = (int)r1[xt.x] * xt.z + (int)r1[xt.y] * xt.w;
    const int v = (((yt.z * (h0 >> 4)) >> 16) + ((yt.w * (h1 >> 4)) >> 16) + 2) >> 2;   // VResizeLinear, uchar cast
    const uint32_t id = (uint32_t)v & 255u;
    const size_t o = ((size_t)b * dh + y) * dw + x;
    if (raw) raw[o] = (uint8_t)id;
    const uint32_t c = slut[id];
    sem[o * 3] = (uint8_t)c; sem[o * 3 + 1] = (uint8_t)(c >> 8); sem[o * 3 + 2] = (uint8_t)(c >> 16);
}

static int lab_fail(int code, const std::string& msg)
{
    set_error(msg);
    return code;
}

void labels_free(ssm_ctx* c)
{
    LabelWs* w = static_cast<LabelWs*>(c->labels_ws);
    if (!w) return;
    void* ptrs[] = {w->d_xtab, w->d_ytab, w->d_lut, w->d_index, w->d_sem, w->d_raw};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete w;
    c->labels_ws = nullptr;
}

// cv::resize's tables: fx = (float)((d + 0.5) * scale - 0.5) with scale = 1. / (dst / (double)src); weights are
// saturate_cast<short>(w * 2048) (round half to even).  x clamps at the borders, y keeps its weights and clamps the rows.
static void make_table(int dst, int src, bool clamp, std::vector<int4>& tab)
{
    tab.resize(dst);
    const double inv_scale = (double)dst / src;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (clamp) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= src - 1) { f = 0.f; s = src - 1; }
        }
        const int w0 = (int)std::nearbyint((1.f - f) * 2048.f), w1 = (int)std::nearbyint(f * 2048.f);
        const int i0 = std::min(std::max(s, 0), src - 1), i1 = std::min(std::max(s + 1, 0), src - 1);
        tab[d] = make_int4(i0, i1, w0, w1);
    }
}

static int ensure_tables(LabelWs* w, int sw, int sh, int dw, int dh, const uint8_t* lut_bgr, cudaStream_t s)
{
    if (!w->d_lut) SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_lut), 256 * sizeof(uint32_t)));
    if (!w->lut_valid || std::memcmp(w->lut_host, lut_bgr, 768) != 0) {
        uint32_t lut[256];
        for (int i = 0; i < 256; ++i) lut[i] = (uint32_t)lut_bgr[3 * i] | ((uint32_t)lut_bgr[3 * i + 1] << 8) | ((uint32_t)lut_bgr[3 * i + 2] << 16);
        SSM_CUDA(cudaMemcpyAsync(w->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, s));
        SSM_CUDA(cudaStreamSynchronize(s));       // `lut` lives on this stack frame
        std::memcpy(w->lut_host, lut_bgr, 768);
        w->lut_valid = true;
    }
    if (w->sw == sw && w->sh == sh && w->dw == dw && w->dh == dh && w->d_xtab) return SSM_OK;
    if (w->d_xtab) cudaFree(w->d_xtab);
    if (w->d_ytab) cudaFree(w->d_ytab);
    w->d_xtab = w->d_ytab = nullptr;
    std::vector<int4> xt, yt;
    make_table(dw, sw, true, xt);
    make_table(dh, sh, false, yt);
    SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_xtab), sizeof(int4) * dw));
    SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_ytab), sizeof(int4) * dh));
    SSM_CUDA(cudaMemcpy(w->d_xtab, xt.data(), sizeof(int4) * dw, cudaMemcpyHostToDevice));
    SSM_CUDA(cudaMemcpy(w->d_ytab, yt.data(), sizeof(int4) * dh, cudaMemcpyHostToDevice));
    w->sw = sw; w->sh = sh; w->dw = dw; w->dh = dh;
    return SSM_OK;
}

static int check_args(int sw, int sh, int dw, int dh)
{
    if (sw < 1 || sh < 1 || dw < 1 || dh < 1 || sw > 32768 || sh > 32768 || dw > 65535 || dh > 65535)
        return lab_fail(SSM_ERR_INVALID_ARGUMENT, "label image dimensions out of range");
    return SSM_OK;
}

static int launch_labels(ssm_ctx* c, LabelWs* w, int B, const uint8_t* d_index, uint8_t* d_sem, uint8_t* d_raw, cudaStream_t s)
{
    dim3 grid((w->dw + 255) / 256, w->dh, B);
    k_labels_resize_lut<<<grid, 256, 0, s>>>(d_index, w->d_xtab, w->d_ytab, w->d_lut, d_sem, d_raw, w->sw, w->sh, w->dw, w->dh);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

}  // namespace ssm

using namespace ssm;

extern "C" {

int ssm_labels_from_indices(ssm_ctx* c, const uint8_t* index, size_t index_stride, int sw, int sh, int dw, int dh, const uint8_t* lut_bgr,
                            uint8_t* semantic_bgr, size_t sem_stride, uint8_t* raw, size_t raw_stride)
{
    if (!c || !index || !lut_bgr || !semantic_bgr) return lab_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_args(sw, sh, dw, dh))) return rc;
    if (!c->labels_ws) c->labels_ws = new LabelWs();
    LabelWs* w = static_cast<LabelWs*>(c->labels_ws);
    cudaStream_t s = c->stream;
    if ((rc = ensure_tables(w, sw, sh, dw, dh, lut_bgr, s))) return rc;
    const size_t nsrc = (size_t)sw * sh, ndst = (size_t)dw * dh;
    if (nsrc > w->cap_src) {
        if (w->d_index) cudaFree(w->d_index);
        w->d_index = nullptr; w->cap_src = 0;
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_index), nsrc));
        w->cap_src = nsrc;
    }
    if (ndst > w->cap_dst) {
        if (w->d_sem) cudaFree(w->d_sem);
        if (w->d_raw) cudaFree(w->d_raw);
        w->d_sem = w->d_raw = nullptr; w->cap_dst = 0;
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_sem), ndst * 3));
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&w->d_raw), ndst));
        w->cap_dst = ndst;
    }
    SSM_CUDA(cudaMemcpy2DAsync(w->d_index, sw, index, index_stride, sw, sh, cudaMemcpyHostToDevice, s));
    if ((rc = launch_labels(c, w, 1, w->d_index, w->d_sem, raw ? w->d_raw : nullptr, s))) return rc;
    SSM_CUDA(cudaMemcpy2DAsync(semantic_bgr, sem_stride, w->d_sem, (size_t)dw * 3, (size_t)dw * 3, dh, cudaMemcpyDeviceToHost, s));
    if (raw) SSM_CUDA(cudaMemcpy2DAsync(raw, raw_stride, w->d_raw, dw, dw, dh, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_labels_from_indices_batch_device(ssm_ctx* c, int batch, const uint8_t* d_index, int sw, int sh, int dw, int dh,
                                         const uint8_t* lut_bgr, uint8_t* d_semantic_bgr, uint8_t* d_raw, void* stream)
{
    if (!c || !d_index || !lut_bgr || !d_semantic_bgr || batch < 1) return lab_fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    int rc;
    if ((rc = check_args(sw, sh, dw, dh))) return rc;
    if (batch > 65535) return lab_fail(SSM_ERR_INVALID_ARGUMENT, "batch too large");
    if (!c->labels_ws) c->labels_ws = new LabelWs();
    LabelWs* w = static_cast<LabelWs*>(c->labels_ws);
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    if ((rc = ensure_tables(w, sw, sh, dw, dh, lut_bgr, s))) return rc;
    return launch_labels(c, w, batch, d_index, d_semantic_bgr, d_raw, s);
}

}  // extern "C"
