// ingest.cu -- PNG ingest in front of the path, sm_100a.
//
// Replaces the cv::imread calls of FrameReader::next (/root/reference src/rgbdframe.cpp:45-78, 138-180: the grey stereo
// pair via imread(path, 0), the colour and label images via imread(path)) for batches of frames.  A PNG is a zlib stream
// of filtered scanlines.  The stream is inherently serial, so it is inflated on host threads (one image per task); the
// rest -- PNG un-filtering (None / Sub / Up / Average / Paeth), palette expansion, alpha stripping, RGB -> BGR reordering
// and the colour -> grey conversion -- runs on the GPU and writes straight into the [batch][h][w] / [batch][h][w][3] device
// images the pipeline entry points take.  Bit-exact with cv2 4.13 imread / imdecode: grey from colour is libpng's
// png_set_rgb_to_gray(0.299, 0.587) arithmetic, (9797 R + 19234 G + 3737 B) >> 15, which is what OpenCV asks libpng for --
// and, for files that carry a gAMA chunk outside 1 +- 0.05 or an sRGB chunk, libpng's linear-light variant of it (samples
// through the gamma_to_1 table, weighted sum rounded, back through gamma_from_1; pixels with R == G == B pass unchanged).
// Like libpng, the decoder verifies the CRC of the critical chunks and rejects scanline filter types above 4: cv::imread
// returns an empty image for such files, so they are errors here, never silently different pixels.
// Supported: 8-bit grey, grey + alpha, RGB, RGBA and palette images, non-interlaced (what KITTI and SegNet tools write).
#include <zlib.h>

#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "ssm_internal.cuh"

namespace ssm {

// ---- host: container parsing + inflate ------------------------------------------------------------------------
struct PngInfo {
    int w = 0, h = 0, bpp = 0;        // bytes per pixel of the filtered scanlines (1, 2, 3, 4)
    int colour_type = 0;              // 0 grey, 2 RGB, 3 palette, 4 grey + alpha, 6 RGBA
    uint8_t palette[256 * 3] = {};
    uint32_t file_gamma = 0;          // gAMA value (x 100000) or 45455 for an sRGB chunk; 0 = none / not significant
};

// libpng 1.6 png_build_gamma_table as it runs for OpenCV's grey read of a colour file (no png_set_gamma call, so the screen
// gamma defaults to the reciprocal of the file gamma): to_1 = correct(i, 1 / file), from_1 = correct(i, 1 / screen) with
// correct(i, g) = floor(255 * pow(i / 255, g * 1e-5) + .5), identity when g is within 1 +- 0.05 (png_gamma_significant).
static inline long long png_reciprocal_fixed(long long a) { return (long long)floor(1e10 / (double)a + .5); }
static void png_gamma_table(long long g, uint8_t* t)
{
    const bool significant = g < 95000 || g > 105000;
    for (int i = 0; i < 256; ++i)
        t[i] = (!significant || i == 0 || i == 255) ? (uint8_t)i : (uint8_t)floor(255.0 * pow(i / 255.0, (double)g * .00001) + .5);
}
static void png_gray_tables(uint32_t file_gamma, uint8_t* to1, uint8_t* from1)
{
    const long long screen = png_reciprocal_fixed(file_gamma);
    png_gamma_table(png_reciprocal_fixed(file_gamma), to1);
    png_gamma_table(png_reciprocal_fixed(screen), from1);
}
static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Parses the chunks and inflates the IDAT stream into `out` (h rows of 1 filter byte + w * bpp bytes).  Returns an
// error text or nullptr.
static const char* png_inflate(const uint8_t* png, size_t n, PngInfo& info, uint8_t* out, size_t out_cap, bool header_only)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n < 8 + 25 || memcmp(png, sig, 8) != 0) return "not a PNG file";
    size_t pos = 8;
    bool have_ihdr = false, stream_open = false, done = false, seen_idat = false, have_srgb = false;
    uint32_t gama = 0;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    const char* err = nullptr;
    while (pos + 12 <= n && !done) {
        const uint32_t len = be32(png + pos);
        const uint8_t* type = png + pos + 4;
        const uint8_t* data = png + pos + 8;
        if ((size_t)len > n - pos - 12) { err = "truncated PNG chunk"; break; }
        // libpng treats a CRC mismatch in a critical chunk (upper-case first letter) as an error, in an ancillary one as a warning
        if (!(type[0] & 0x20) && (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, (uInt)(4 + len)) != be32(data + len)) {
            err = "PNG chunk CRC mismatch";
            break;
        }
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; break; }
            info.w = (int)be32(data); info.h = (int)be32(data + 4);
            const int depth = data[8];
            info.colour_type = data[9];
            if (data[10] != 0 || data[11] != 0) { err = "unknown PNG compression / filter method"; break; }
            if (data[12] != 0) { err = "interlaced PNG files are not supported"; break; }
            if (depth != 8) { err = "only 8-bit PNG files are supported"; break; }
            switch (info.colour_type) {
                case 0: info.bpp = 1; break;
                case 2: info.bpp = 3; break;
                case 3: info.bpp = 1; break;
                case 4: info.bpp = 2; break;
                case 6: info.bpp = 4; break;
                default: err = "unknown PNG colour type"; break;
            }
            if (err) break;
            have_ihdr = true;
            if (header_only) return nullptr;
            if (info.w < 1 || info.h < 1 || info.w > (1 << 20) || info.h > (1 << 20)) { err = "PNG dimensions out of range"; break; }
            if ((size_t)info.h * ((size_t)info.w * info.bpp + 1) > out_cap) { err = "PNG larger than the batch's frame size"; break; }
            if ((size_t)info.h * ((size_t)info.w * info.bpp + 1) > 0x7fffffffull) { err = "PNG too large"; break; }
            if (inflateInit(&zs) != Z_OK) { err = "zlib inflateInit failed"; break; }
            stream_open = true;
            zs.next_out = out;
            zs.avail_out = (uInt)((size_t)info.h * ((size_t)info.w * info.bpp + 1));
        } else if (!have_ihdr) {
            err = "PNG does not start with IHDR";
            break;
        } else if (!memcmp(type, "PLTE", 4)) {
            if (len > 768 || len % 3) { err = "bad PLTE"; break; }
            memcpy(info.palette, data, len);
        } else if (!memcmp(type, "gAMA", 4) && !seen_idat) {
            if (len == 4 && be32(data) >= 16 && be32(data) <= 625000000u) gama = be32(data);   // out-of-range values are ignored by libpng
        } else if (!memcmp(type, "sRGB", 4) && !seen_idat) {
            have_srgb = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            seen_idat = true;
            zs.next_in = const_cast<Bytef*>(data);
            zs.avail_in = len;
            const int rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END) { err = "corrupt PNG data stream"; break; }
        } else if (!memcmp(type, "IEND", 4)) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    if (stream_open) {
        if (!err && zs.avail_out != 0) err = "PNG data stream ends early";
        inflateEnd(&zs);
    }
    if (!err && !have_ihdr) err = "PNG without IHDR";
    if (!err) {
        const size_t rb = (size_t)info.w * info.bpp + 1;
        for (int y = 0; y < info.h; ++y)
            if (out[(size_t)y * rb] > 4) { err = "bad PNG scanline filter type"; break; }
        const uint32_t g = have_srgb ? 45455u : gama;   // an sRGB chunk overrides gAMA
        info.file_gamma = (g != 0 && (g < 95000u || g > 105000u)) ? g : 0u;
    }
    return err;
}

// ---- device: un-filter + convert --------------------------------------------------------------------------------
struct PngDesc {                      // one image of the batch
    unsigned long long src_off;       // byte offset of its filtered scanlines in the staging buffer
    int bpp, colour_type;
    int palette_index;                // index into the palette table, or -1
    int gamma_index;                  // index into the gamma table pairs (to_1, from_1), or -1: plain integer weights
};

__device__ __forceinline__ int paeth(int a, int b, int c)
{
    const int p = a + b - c;
    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// One CTA per image walks the rows top to bottom with the previous and the current un-filtered row in shared memory.  Up is
// parallel over the row; Sub is a prefix sum per channel (chunked scan over the CTA); Average and Paeth depend on the byte
// to the left, so one thread per channel walks the row (the running left value stays in a register).  Conversion and the
// store of the finished row are parallel again.  The kernel uses a few warps of one SM per image: a batch of frames keeps
// the machine's other slots free for the path's own kernels.
template <int MODE /* 0: grey [h][w], 1: BGR [h][w][3] */>
__global__ void __launch_bounds__(128) k_png_unfilter(const uint8_t* __restrict__ staged, const PngDesc* __restrict__ desc,
                                                      const uint8_t* __restrict__ palettes, const uint8_t* __restrict__ gammas,
                                                      uint8_t* __restrict__ out, int W, int H)
{
    extern __shared__ __align__(16) uint8_t png_smem[];
    __shared__ int chunk_sum[128 * 4];
    __shared__ uint8_t gam[512];
    const PngDesc d = desc[blockIdx.x];
    const int bpp = d.bpp, rb = W * bpp, tid = threadIdx.x, nt = blockDim.x;
    uint8_t* rowA = png_smem;
    uint8_t* rowB = png_smem + ((rb + 15) & ~15);
    const uint8_t* src = staged + d.src_off;
    const uint8_t* pal = d.palette_index >= 0 ? palettes + (size_t)d.palette_index * 768 : nullptr;
    uint8_t* dst = out + (size_t)blockIdx.x * W * H * (MODE ? 3 : 1);
    for (int i = tid; i < rb; i += nt) rowB[i] = 0;      // the row above the first one is all zeros
    const bool linear = MODE == 0 && d.gamma_index >= 0;
    if (linear)
        for (int i = tid; i < 512; i += nt) gam[i] = gammas[(size_t)d.gamma_index * 512 + i];
    __syncthreads();
    uint8_t* cur = rowA;
    uint8_t* prev = rowB;
    for (int y = 0; y < H; ++y) {
        const uint8_t* line = src + (size_t)y * (rb + 1);
        const int ft = line[0];
        for (int i = tid; i < rb; i += nt) cur[i] = line[1 + i];
        __syncthreads();
        if (ft == 2) {
            for (int i = tid; i < rb; i += nt) cur[i] = (uint8_t)(cur[i] + prev[i]);
        } else if (ft == 1) {
            // prefix sum (mod 256) along each channel: every thread sums a chunk of pixels, the chunk totals are scanned
            // by the first bpp threads, and every thread re-walks its chunk with its offset
            const int per = (W + nt - 1) / nt, x0 = tid * per, x1 = min(W, x0 + per);
            int tot[4] = {0, 0, 0, 0};
            for (int x = x0; x < x1; ++x)
                for (int c = 0; c < bpp; ++c) tot[c] += cur[x * bpp + c];
            // exclusive scan of the chunk totals over the CTA's threads: shuffles inside a warp, the warp totals through shared memory
            const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
            int incl[4];
            for (int c = 0; c < 4; ++c) {
                int v = tot[c];
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += u;
                }
                incl[c] = v;
                if (lane == 31) chunk_sum[c * 128 + wid] = v;
            }
            __syncthreads();
            for (int c = 0; c < 4; ++c) {
                int before = 0;
                for (int q = 0; q < nw; ++q) before += q < wid ? chunk_sum[c * 128 + q] : 0;
                tot[c] = before + incl[c] - tot[c];  // exclusive prefix of my chunk
            }
            int acc[4];
            for (int c = 0; c < 4; ++c) acc[c] = tot[c];
            for (int x = x0; x < x1; ++x)
                for (int c = 0; c < bpp; ++c) {
                    acc[c] += cur[x * bpp + c];
                    cur[x * bpp + c] = (uint8_t)acc[c];
                }
        } else if (ft == 3 || ft == 4) {
            if (tid < bpp) {
                int left = 0, upleft = 0;
                for (int i = tid; i < rb; i += bpp) {
                    const int up = prev[i];
                    const int pred = ft == 3 ? ((left + up) >> 1) : paeth(left, up, upleft);
                    left = (cur[i] + pred) & 255;
                    cur[i] = (uint8_t)left;
                    upleft = up;
                }
            }
        }
        __syncthreads();
        // the finished row -> output pixels
        for (int x = tid; x < W; x += nt) {
            int r, g, b;
            if (d.colour_type == 2 || d.colour_type == 6) { r = cur[x * bpp]; g = cur[x * bpp + 1]; b = cur[x * bpp + 2]; }
            else if (d.colour_type == 3) { const uint8_t* e = pal + 3 * cur[x]; r = e[0]; g = e[1]; b = e[2]; }
            else { r = g = b = cur[x * bpp]; }
            if (MODE == 0) {
                // libpng's rgb_to_gray as OpenCV configures it (png_set_rgb_to_gray(png, 1, 0.299, 0.587)); the weights sum to 2^15
                // (libpng's linear-light variant when the file carries a significant gamma; grey pixels pass unchanged)
                if (linear && (r != g || r != b))
                    dst[(size_t)y * W + x] = gam[256 + ((9797 * gam[r] + 19234 * gam[g] + 3737 * gam[b] + 16384) >> 15)];
                else
                    dst[(size_t)y * W + x] = (uint8_t)((9797 * r + 19234 * g + 3737 * b) >> 15);
            } else {
                uint8_t* q = dst + ((size_t)y * W + x) * 3;
                q[0] = (uint8_t)b; q[1] = (uint8_t)g; q[2] = (uint8_t)r;
            }
        }
        __syncthreads();
        uint8_t* t = cur; cur = prev; prev = t;
    }
}

// Two staging sets alternate between calls, so the host inflates call n + 1 while the GPU still copies and un-filters call n.
struct IngestSet {
    uint8_t* h_staged = nullptr;      // pinned: filtered scanlines of a batch
    uint8_t* d_staged = nullptr;
    size_t staged_cap = 0;
    PngDesc *h_desc = nullptr, *d_desc = nullptr;
    uint8_t *h_pal = nullptr, *d_pal = nullptr;
    uint8_t *h_gam = nullptr, *d_gam = nullptr;   // [images][2][256] gamma_to_1 / gamma_from_1
    int cap_images = 0;
    cudaEvent_t done = nullptr;       // the set's last batch has left the staging buffers
};
struct IngestWs {
    IngestSet set[2];
    unsigned calls = 0;
};

static void ingest_set_free(IngestSet& w)
{
    if (w.h_staged) cudaFreeHost(w.h_staged);
    if (w.d_staged) cudaFree(w.d_staged);
    if (w.h_desc) cudaFreeHost(w.h_desc);
    if (w.d_desc) cudaFree(w.d_desc);
    if (w.h_pal) cudaFreeHost(w.h_pal);
    if (w.d_pal) cudaFree(w.d_pal);
    if (w.h_gam) cudaFreeHost(w.h_gam);
    if (w.d_gam) cudaFree(w.d_gam);
    if (w.done) cudaEventDestroy(w.done);
    w = IngestSet();
}

void ingest_free(ssm_ctx* c)
{
    IngestWs* w = static_cast<IngestWs*>(c->ingest_ws);
    if (!w) return;
    ingest_set_free(w->set[0]);
    ingest_set_free(w->set[1]);
    delete w;
    c->ingest_ws = nullptr;
}

static int ingest_reserve(IngestSet& w, int images, size_t per_image)
{
    if (!w.done) SSM_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    SSM_CUDA(cudaEventSynchronize(w.done));   // (a never-recorded event is complete)
    const size_t need = per_image * (size_t)images;
    if (need > w.staged_cap) {
        if (w.h_staged) cudaFreeHost(w.h_staged);
        if (w.d_staged) cudaFree(w.d_staged);
        w.h_staged = nullptr; w.d_staged = nullptr; w.staged_cap = 0;
        SSM_CUDA(cudaMallocHost(&w.h_staged, need));
        SSM_CUDA(cudaMalloc(&w.d_staged, need));
        w.staged_cap = need;
    }
    if (images > w.cap_images) {
        if (w.h_desc) cudaFreeHost(w.h_desc);
        if (w.d_desc) cudaFree(w.d_desc);
        if (w.h_pal) cudaFreeHost(w.h_pal);
        if (w.d_pal) cudaFree(w.d_pal);
        if (w.h_gam) cudaFreeHost(w.h_gam);
        if (w.d_gam) cudaFree(w.d_gam);
        w.h_desc = nullptr; w.d_desc = nullptr; w.h_pal = nullptr; w.d_pal = nullptr; w.h_gam = nullptr; w.d_gam = nullptr; w.cap_images = 0;
        SSM_CUDA(cudaMallocHost(&w.h_desc, sizeof(PngDesc) * images));
        SSM_CUDA(cudaMalloc(&w.d_desc, sizeof(PngDesc) * images));
        SSM_CUDA(cudaMallocHost(&w.h_pal, (size_t)768 * images));
        SSM_CUDA(cudaMalloc(&w.d_pal, (size_t)768 * images));
        SSM_CUDA(cudaMallocHost(&w.h_gam, (size_t)512 * images));
        SSM_CUDA(cudaMalloc(&w.d_gam, (size_t)512 * images));
        w.cap_images = images;
    }
    return SSM_OK;
}

}  // namespace ssm

using namespace ssm;

extern "C" {

int ssm_png_info(const uint8_t* png, size_t png_bytes, int* w, int* h, int* channels)
{
    if (!png || !w || !h) { set_error("null argument"); return SSM_ERR_INVALID_ARGUMENT; }
    PngInfo info;
    if (const char* e = png_inflate(png, png_bytes, info, nullptr, 0, true)) { set_error(e); return SSM_ERR_INVALID_ARGUMENT; }
    *w = info.w; *h = info.h;
    if (channels) *channels = (info.colour_type == 0 || info.colour_type == 4) ? 1 : 3;
    return SSM_OK;
}

int ssm_png_decode_batch_device(ssm_ctx* c, int batch, const uint8_t* const* png, const size_t* png_bytes, int w, int h, int mode,
                                uint8_t* d_out, int host_threads, void* stream)
{
    if (!c || !png || !png_bytes || !d_out || batch < 1 || w < 1 || h < 1 || (mode != 0 && mode != 1)) {
        set_error("bad argument");
        return SSM_ERR_INVALID_ARGUMENT;
    }
    SSM_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    const size_t per_image = (size_t)h * ((size_t)w * 4 + 1);          // worst case: RGBA
    if (!c->ingest_ws) c->ingest_ws = new IngestWs();
    IngestWs* all = static_cast<IngestWs*>(c->ingest_ws);
    IngestSet* ws = &all->set[all->calls++ & 1u];
    int rc = ingest_reserve(*ws, batch, per_image);   // waits until the set's previous batch has left its buffers
    if (rc) return rc;
    // inflate: one image per task on host threads
    std::atomic<int> next(0);
    std::vector<const char*> errs((size_t)batch, nullptr);
    auto worker = [&]() {
        for (int i = next.fetch_add(1); i < batch; i = next.fetch_add(1)) {
            PngInfo info;
            const char* e = png_inflate(png[i], png_bytes[i], info, ws->h_staged + per_image * i, per_image, false);
            if (!e && (info.w != w || info.h != h)) e = "PNG size differs from the batch's frame size";
            errs[i] = e;
            if (e) continue;
            PngDesc& d = ws->h_desc[i];
            d.src_off = (unsigned long long)(per_image * i);
            d.bpp = info.bpp; d.colour_type = info.colour_type;
            d.palette_index = info.colour_type == 3 ? i : -1;
            if (info.colour_type == 3) memcpy(ws->h_pal + (size_t)768 * i, info.palette, 768);
            const bool colour_file = info.colour_type == 2 || info.colour_type == 3 || info.colour_type == 6;
            d.gamma_index = (mode == 0 && colour_file && info.file_gamma) ? i : -1;
            if (d.gamma_index >= 0) png_gray_tables(info.file_gamma, ws->h_gam + (size_t)512 * i, ws->h_gam + (size_t)512 * i + 256);
        }
    };
    const int nthreads = std::max(1, std::min(host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency(), batch));
    if (nthreads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    for (int i = 0; i < batch; ++i)
        if (errs[i]) { set_error(std::string("image ") + std::to_string(i) + ": " + errs[i]); return SSM_ERR_INVALID_ARGUMENT; }
    int max_bpp = 1;
    for (int i = 0; i < batch; ++i) max_bpp = std::max(max_bpp, ws->h_desc[i].bpp);
    SSM_CUDA(cudaMemcpyAsync(ws->d_staged, ws->h_staged, per_image * batch, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(ws->d_desc, ws->h_desc, sizeof(PngDesc) * batch, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(ws->d_pal, ws->h_pal, (size_t)768 * batch, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(ws->d_gam, ws->h_gam, (size_t)512 * batch, cudaMemcpyHostToDevice, s));
    const size_t smem = 2 * (((size_t)w * max_bpp + 15) & ~(size_t)15);
    if (smem > 200 * 1024) { set_error("PNG rows too wide for the un-filter kernel"); return SSM_ERR_INVALID_ARGUMENT; }
    if (mode == 0) {
        SSM_CUDA(cudaFuncSetAttribute(k_png_unfilter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_png_unfilter<0><<<batch, 128, smem, s>>>(ws->d_staged, ws->d_desc, ws->d_pal, ws->d_gam, d_out, w, h);
    } else {
        SSM_CUDA(cudaFuncSetAttribute(k_png_unfilter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_png_unfilter<1><<<batch, 128, smem, s>>>(ws->d_staged, ws->d_desc, ws->d_pal, ws->d_gam, d_out, w, h);
    }
    SSM_LAUNCH_CHECK(c);
    SSM_CUDA(cudaEventRecord(ws->done, s));
    return SSM_OK;
}

int ssm_png_decode(ssm_ctx* c, const uint8_t* png, size_t png_bytes, int mode, uint8_t* out, size_t out_bytes, int* w, int* h)
{
    if (!c || !png || !out || (mode != 0 && mode != 1)) { set_error("bad argument"); return SSM_ERR_INVALID_ARGUMENT; }
    int W = 0, H = 0;
    int rc = ssm_png_info(png, png_bytes, &W, &H, nullptr);
    if (rc) return rc;
    const size_t need = (size_t)W * H * (mode ? 3 : 1);
    if (out_bytes < need) { set_error("output buffer too small"); return SSM_ERR_INVALID_ARGUMENT; }
    SSM_CUDA(cudaSetDevice(c->device));
    uint8_t* d_out = nullptr;
    SSM_CUDA(cudaMalloc(&d_out, need));
    const uint8_t* one[1] = {png};
    const size_t nb[1] = {png_bytes};
    rc = ssm_png_decode_batch_device(c, 1, one, nb, W, H, mode, d_out, 1, c->stream);
    if (rc == SSM_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, need, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "ssm_png_decode copy");
    }
    cudaFree(d_out);
    if (rc == SSM_OK) { if (w) *w = W; if (h) *h = H; }
    return rc;
}

}  // extern "C"
