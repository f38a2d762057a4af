// api.cu -- the C ABI of include/ssm.h: context, buffers, host/device entry points, the batched path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ssm_internal.cuh"

namespace ssm {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what)
{
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
    return SSM_ERR_CUDA;
}

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

static int validate(const ssm_params& p)
{
    if (p.min_disparity != 0) return fail(SSM_ERR_UNSUPPORTED, "min_disparity must be 0 (src/stereo.cpp:19)");
    if (p.num_disparities < 16 || p.num_disparities % 16 != 0 || p.num_disparities > 512)
        return fail(SSM_ERR_INVALID_ARGUMENT, "num_disparities must be a multiple of 16 in [16, 512]");
    if (p.block_size < 1 || p.block_size % 2 == 0) return fail(SSM_ERR_INVALID_ARGUMENT, "block_size must be odd and >= 1");
    if (p.block_size > 11) return fail(SSM_ERR_UNSUPPORTED, "block_size > 11 is outside the 16-bit cost range");
    const int P1 = p.p1 > 0 ? p.p1 : 2, P2 = std::max(p.p2 > 0 ? p.p2 : 5, P1 + 1);
    if (189 * p.block_size * p.block_size + P2 > 32767 || P1 > 8000)
        return fail(SSM_ERR_UNSUPPORTED, "189*block_size^2 + P2 must stay <= 32767 (16-bit path costs)");
    if (p.max_width <= p.num_disparities || p.max_width > 65535 || p.max_height < 1 || p.max_batch < 1)
        return fail(SSM_ERR_INVALID_ARGUMENT, "max_width must exceed num_disparities (and be <= 65535); max_height, max_batch >= 1");
    if (p.num_labels < 1 || p.num_labels > SSM_MAX_LABELS) return fail(SSM_ERR_INVALID_ARGUMENT, "num_labels must be in [1, 20]");
    if (!(p.resolution > 0) || !(p.scale > 0) || !(p.fx != 0) || !(p.fy != 0))
        return fail(SSM_ERR_INVALID_ARGUMENT, "resolution, scale, fx, fy must be non-zero/positive");
    if (p.uniqueness_ratio > 100) return fail(SSM_ERR_INVALID_ARGUMENT, "uniqueness_ratio must be <= 100");
    if (p.dilate_iterations < 0 || p.dilate_iterations > 8) return fail(SSM_ERR_INVALID_ARGUMENT, "dilate_iterations in [0, 8]");
    if (p.map_capacity < 1024) return fail(SSM_ERR_INVALID_ARGUMENT, "map_capacity must be >= 1024 slots");
    // rgbdframe.cpp:112-113 stores ushort(pz * scale) for pz < roiz: the product has to fit 16 bits, or depths would wrap (and a
    // wrapped value of 0 reads as "no depth")
    if (!(p.roix > 0) || !(p.roiy > 0) || !(p.roiz > 0)) return fail(SSM_ERR_INVALID_ARGUMENT, "roix, roiy, roiz must be positive");
    if (!(p.roiz * p.scale < 65536.0)) return fail(SSM_ERR_INVALID_ARGUMENT, "roiz * scale must stay below 65536 (16-bit depth image)");
    return SSM_OK;
}

// Disparities per column of the cost-volume layout.  D = 80 (the reference's src/stereo.cpp:18), 96, 112 run on the D = 128
// kernels and D = 48 on the D = 64 ones: the lanes above D are switched off, their cells are computed but never read.
static int layout_disparities(const ssm_ctx* c, int D)
{
    if (c->no_pad || c->force_legacy_cost || c->force_legacy_hsweep || c->force_legacy_vertical || c->p.block_size != 11) return D;
    if (D > 128 && D < 256) return 256;
    if (D > 64 && D < 128) return 128;
    if (D > 32 && D < 64) return 64;
    return D;
}

static void fill_dev_params(ssm_ctx* c, int w, int h)
{
    const ssm_params& p = c->p;
    DevParams& d = c->dp;
    d.W = w; d.H = h; d.D = p.num_disparities; d.W1 = w - p.num_disparities;
    d.Dl = layout_disparities(c, d.D);
    d.bs = p.block_size;
    d.P1 = p.p1 > 0 ? p.p1 : 2;
    d.P2 = std::max(p.p2 > 0 ? p.p2 : 5, d.P1 + 1);
    d.uniq = p.uniqueness_ratio >= 0 ? p.uniqueness_ratio : 10;
    d.d12 = p.disp12_max_diff > 0 ? p.disp12_max_diff : 1;
    d.ftzero = std::max(p.pre_filter_cap, 15) | 1;
    d.speckle_win = p.speckle_window_size;
    d.speckle_diff = kDispScale * p.speckle_range;
    d.cx = p.cx; d.cy = p.cy; d.fx = p.fx; d.fy = p.fy; d.baseline = p.baseline; d.scale = p.scale;
    d.roix = p.roix; d.roiy = p.roiy; d.roiz = p.roiz;
    d.max_depth_units = p.max_distance * p.scale;          // mapper.cpp:30
    d.inv_leaf = 1.0f / (float)p.resolution;               // VoxelGrid::setLeafSize -> inverse_leaf_size_ (fp32)
    d.num_labels = p.num_labels;
    for (int i = 0; i < SSM_MAX_LABELS; ++i)
        d.palette[i] = i < p.num_labels ? ((uint32_t)p.palette_bgr[i][0] | ((uint32_t)p.palette_bgr[i][1] << 8) |
                                           ((uint32_t)p.palette_bgr[i][2] << 16))
                                        : 0xffffffffu;
    d.drop_mask = p.drop_mask; d.dynamic_mask = p.dynamic_mask;
    d.dilate_radius = p.dilate_iterations; d.colour_source = p.colour_source;
    c->ptab_margin = d.Dl != d.D ? (d.Dl - d.D + 4 + 3) / 4 * 4 : 0;
    c->ptab_pitch = c->ptab_margin + (w + 3) / 4 * 4 + 4;
}

static int set_shape(ssm_ctx* c, int w, int h, int batch)
{
    if (w <= c->p.num_disparities) return fail(SSM_ERR_INVALID_ARGUMENT, "image width must exceed num_disparities");
    if (w > c->cap_w || h > c->cap_h || h < 1 || (size_t)w * h > (size_t)c->cap_w * c->cap_h)
        return fail(SSM_ERR_CAPACITY, "frame larger than the context's max_width x max_height");
    if (batch < 1 || batch > c->cap_b) return fail(SSM_ERR_CAPACITY, "batch larger than the context's max_batch");
    if (c->dp.W != w || c->dp.H != h) fill_dev_params(c, w, h);
    return SSM_OK;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)); }

static void free_all(ssm_ctx* c)
{
    void* ptrs[] = {c->d_left, c->d_right, c->d_recL, c->d_recR, c->d_ptab, c->d_C, c->d_S, c->d_disp_raw, c->d_disp_lr, c->d_disp_med,
                    c->d_disp, c->d_disp2key, c->d_wta_rec, c->d_ck, c->d_uniq_thr, c->d_cc_label, c->d_cc_size, c->d_depth, c->d_label, c->d_mask, c->d_label_lut, c->d_sem,
                    c->d_rgb, c->d_pose, c->d_min_disp, c->d_points, c->d_blk_count, c->d_counters, c->d_table, c->d_send,
                    c->d_recv, c->d_send_counts};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    for (auto& q : c->d_spill)
        if (q) cudaFree(q);
    if (c->h_mirror) cudaFreeHost(c->h_mirror);
    if (c->export_ws) cudaFree(c->export_ws);
    for (auto& e : c->ev_batch)
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < ssm_ctx::kRetired; ++i) {
        if (c->retired[i]) cudaFree(c->retired[i]);
        if (c->retired_ev[i]) cudaEventDestroy(c->retired_ev[i]);
    }
    for (auto& set : c->ev)
        for (auto& e : set)
            if (e) cudaEventDestroy(e);
    for (auto& st : c->sub_stream)
        if (st) cudaStreamDestroy(st);
    for (auto& e : c->sub_join)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->sub_cost)
        if (e) cudaEventDestroy(e);
    if (c->sub_fork) cudaEventDestroy(c->sub_fork);
    for (int k = 0; k < 2; ++k) {
        void* st[] = {c->stage_left[k], c->stage_right[k], c->stage_sem[k], c->stage_rgb[k], c->stage_pose[k]};
        for (void* q : st)
            if (q) cudaFree(q);
        if (c->ev_copied[k]) cudaEventDestroy(c->ev_copied[k]);
        if (c->ev_consumed[k]) cudaEventDestroy(c->ev_consumed[k]);
    }
    if (c->route_stream) cudaStreamDestroy(c->route_stream);
    if (c->ev_route_done) cudaEventDestroy(c->ev_route_done);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
}

// stage timing helpers
static inline void mark(ssm_ctx* c, int i, cudaStream_t s)
{
    if (c->timing) cudaEventRecord(c->ev[c->ev_set][i], s);
}

// the stereo half on device buffers: prefilter -> cost volume -> 5-path aggregation -> selection -> post-filters
static int run_sgbm(ssm_ctx* c, int B, const uint8_t* dL, const uint8_t* dR, int16_t* d_out, cudaStream_t s)
{
    int rc;
    mark(c, 0, s);
    if ((rc = launch_prefilter(c, B, dL, dR, s))) return rc;
    if ((rc = launch_cost_volume(c, B, s))) return rc;
    if (c->ev_after_cost) SSM_CUDA(cudaEventRecord(c->ev_after_cost, s));
    mark(c, 1, s);
    if ((rc = launch_aggregate_vertical(c, B, s))) return rc;
    mark(c, 2, s);
    if ((rc = launch_aggregate_horizontal(c, B, s))) return rc;
    mark(c, 3, s);
    if ((rc = launch_select(c, B, s))) return rc;
    mark(c, 4, s);
    if ((rc = launch_post(c, B, d_out, s))) return rc;
    mark(c, 5, s);
    return SSM_OK;
}

// the mapper half on device buffers: per-frame preparation (depth image, label ids, moving mask) ...
static int run_map_prepare(ssm_ctx* c, int B, const int16_t* d_disp, const uint8_t* d_sem, cudaStream_t s)
{
    int rc;
    if ((rc = launch_depth(c, B, d_disp, c->d_depth, s))) return rc;
    return launch_labels_mask(c, B, d_sem, s);
}
// ... and point generation + voxel fusion (the only part with shared state; across ranks: routing to the owner)
static int run_map_points(ssm_ctx* c, int B, const uint8_t* d_sem, const uint8_t* d_rgb, const double* d_pose, cudaStream_t s)
{
    int rc;
    if (c->nranks > 1 && c->p2p) {
        mark(c, 6, s);   // points are made, fused or sent to their owner in one kernel; then barrier + inbox fusion
        if ((rc = points_route_p2p(c, B, c->d_depth, d_sem, d_rgb, d_pose, s))) return rc;
    } else if (c->nranks > 1) {
        if ((rc = launch_points(c, B, c->d_depth, d_sem, d_rgb, d_pose, false, s))) return rc;
        mark(c, 6, s);
        if ((rc = route_and_fuse(c, s))) return rc;
    } else {
        mark(c, 6, s);   // single GPU: points are fused as they are generated, one kernel
        if ((rc = launch_points(c, B, c->d_depth, d_sem, d_rgb, d_pose, true, s))) return rc;
    }
    mark(c, 7, s);
    if (c->timing) {
        c->ev_used = std::min(c->ev_used + 1, (int)ssm_ctx::kEvSets);
        c->ev_set = (c->ev_set + 1) % ssm_ctx::kEvSets;
    }
    return SSM_OK;
}
static int run_map(ssm_ctx* c, int B, const int16_t* d_disp, const uint8_t* d_sem, const uint8_t* d_rgb,
                   const double* d_pose, cudaStream_t s)
{
    int rc;
    if ((rc = run_map_prepare(c, B, d_disp, d_sem, s))) return rc;
    return run_map_points(c, B, d_sem, d_rgb, d_pose, s);
}

// Shift every per-frame work buffer of the context by `frames` frames (positive or negative): a sub-batch then runs
// through the unchanged stage launchers on its own slice of the buffers.
static void offset_buffers(ssm_ctx* c, ptrdiff_t frames)
{
    const ptrdiff_t npix = (ptrdiff_t)c->dp.W * c->dp.H * frames;
    const ptrdiff_t cells = (ptrdiff_t)c->dp.W1 * c->dp.H * c->dp.Dl * frames;
    c->d_recL += npix; c->d_recR += npix;
    if (c->d_ptab) c->d_ptab += (ptrdiff_t)c->dp.H * 6 * c->ptab_pitch * frames;
    c->d_C += cells; c->d_S += cells; c->d_hs += cells;
    c->d_disp_raw += npix; c->d_disp_lr += npix; c->d_disp_med += npix; c->d_disp += npix;
    c->d_disp2key += npix; c->d_wta_rec += (c->dp.Dl > 128 ? 4 : 2) * npix; c->d_cc_label += npix;
    if (c->d_ck) c->d_ck += (ptrdiff_t)hsweep2_ck_words(c->dp.W1, c->dp.Dl, c->dp.H, 1) * frames; c->d_cc_size += npix;
    c->d_depth += npix; c->d_label += npix; c->d_mask += npix;
    c->d_min_disp += frames;
}

static int run_map(ssm_ctx* c, int B, const int16_t* d_disp, const uint8_t* d_sem, const uint8_t* d_rgb, const double* d_pose, cudaStream_t s);
static int sync_route(ssm_ctx* c);

// Growth policy of the streaming path.  The host never waits for the device here: it looks at the pinned mirror of the device
// counters (refreshed behind every pipeline call, so stale by the one or two batches still in flight) and keeps the table
// below half full with room for three more batches of the largest increase seen so far.  Whatever still does not fit lands in
// the spill list and is re-inserted by the growth step this triggers.
static int maybe_grow(ssm_ctx* c, cudaStream_t s)
{
    if (!c->auto_grow || !c->h_mirror) return SSM_OK;
    // until one batch's voxel count has been seen there is no growth rate to plan with: the second and third call wait for
    // the batch before them (once per map; the streaming overlap starts right after)
    const uint64_t calls = c->pipeline_calls++;
    if (c->mirror_dmax == 0 && calls >= 1 && calls <= 2) {
        int rr = sync_route(c);
        if (rr) return rr;
        SSM_CUDA(cudaStreamSynchronize(s));
    } else if (calls >= 2 && c->ev_batch[(calls - 2) % ssm_ctx::kBatchRing]) {
        // the host may run ahead of the device by two batches, not more: the mirror then reflects every batch but the two in
        // flight, which (with the one being enqueued) is what the head-room of three increases below covers
        SSM_CUDA(cudaEventSynchronize(c->ev_batch[(calls - 2) % ssm_ctx::kBatchRing]));
    }
    {
        int rp = table_reap(c, false);
        if (rp) return rp;
    }
    const volatile uint32_t* m = c->h_mirror;
    const uint64_t nvox = m[1], parked = m[4];
    if (nvox >= c->mirror_prev) c->mirror_dmax = std::max(c->mirror_dmax, nvox - c->mirror_prev);
    c->mirror_prev = nvox;
    const uint64_t need = nvox + 3 * c->mirror_dmax;
    if (parked == 0 && 2 * need <= c->table_slots) return SSM_OK;
    if (c->route_pending) {   // the exchange stream inserts into the table as well
        SSM_CUDA(cudaStreamWaitEvent(s, c->ev_route_done, 0));
        c->route_pending = false;
    }
    uint64_t want = c->table_slots;
    while (want < 2 * need) want <<= 1;
    if (want == c->table_slots) want <<= 1;
    int rc = table_grow(c, want, s);
    if (rc == SSM_ERR_CAPACITY) {   // no room for a bigger table: keep going on the old one; a full table surfaces at the next blocking call
        c->auto_grow = false;
        return SSM_OK;
    }
    if (rc) return rc;
    // the mirror still shows the parked points until the next refresh: do not grow again for them
    c->h_mirror[4] = 0;
    return SSM_OK;
}
static int refresh_mirror(ssm_ctx* c, cudaStream_t s)
{
    if (!c->h_mirror || !c->auto_grow) return SSM_OK;
    SSM_CUDA(cudaMemcpyAsync(c->h_mirror, c->d_counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    cudaEvent_t& e = c->ev_batch[(c->pipeline_calls - 1) % ssm_ctx::kBatchRing];   // this call's slot (maybe_grow counted it)
    if (!e) SSM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SSM_CUDA(cudaEventRecord(e, s));
    return SSM_OK;
}

// The whole path on device buffers.  On a single GPU a batch of 64 frames or more (or SSM_TUNE3 = n > 1) is cut into
// sub-batches that run on their own streams and meet again on `s`.
static int run_pipeline(ssm_ctx* c, int B, const uint8_t* dL, const uint8_t* dR, const uint8_t* d_sem, const uint8_t* d_rgb,
                        const double* d_pose, int16_t* d_disp, cudaStream_t s)
{
    int rc;
    if ((rc = maybe_grow(c, s))) return rc;
    // sub-batches of at least ~one full wave of the vertical cluster kernel each (33 KITTI frames on a B200): kernels of
    // different sub-batches are bound by different units (shared memory, issue slots, HBM) and fill each other's tails
    // (the wave of the vertical cluster kernel sets the sub-batch size: 33 KITTI frames at 128 disparities, 7 Cityscapes-sized ones at 256)
    const int wave = vertical_wave_frames(c);
    const int want = c->tune[3] >= 0 ? c->tune[3] : (wave > 0 && B >= 2 * wave - wave / 8 ? std::min((B + wave - 1) / wave, (int)ssm_ctx::kMaxSplit) : 1);
    const int nsplit = std::min({want, (int)ssm_ctx::kMaxSplit, B});
    if (nsplit <= 1 || c->timing) {
        if (c->route_pending) {   // an overlapped exchange is still in flight: order this call after it
            SSM_CUDA(cudaStreamWaitEvent(s, c->ev_route_done, 0));
            c->route_pending = false;
        }
        if ((rc = run_sgbm(c, B, dL, dR, d_disp, s))) return rc;
        if ((rc = run_map(c, B, d_disp, d_sem, d_rgb, d_pose, s))) return rc;
        return refresh_mirror(c, s);
    }
    const size_t npix = (size_t)c->dp.W * c->dp.H;
    if (!c->sub_fork) {
        SSM_CUDA(cudaEventCreateWithFlags(&c->sub_fork, cudaEventDisableTiming));
        for (int i = 0; i < ssm_ctx::kMaxSplit; ++i) {
            SSM_CUDA(cudaStreamCreateWithFlags(&c->sub_stream[i], cudaStreamNonBlocking));
            SSM_CUDA(cudaEventCreateWithFlags(&c->sub_join[i], cudaEventDisableTiming));
            SSM_CUDA(cudaEventCreateWithFlags(&c->sub_cost[i], cudaEventDisableTiming));
        }
    }
    const bool overlap = c->route_overlap && c->nranks > 1;
    if (overlap && !c->route_stream) {
        // highest priority: the exchange's small kernels (and the NCCL barrier) take the first free SM slots instead of
        // queueing behind a full wave of SGBM blocks, so the ranks meet at the barrier early
        int prio_lo = 0, prio_hi = 0;
        SSM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        SSM_CUDA(cudaStreamCreateWithPriority(&c->route_stream, cudaStreamNonBlocking, prio_hi));
        SSM_CUDA(cudaEventCreateWithFlags(&c->ev_route_done, cudaEventDisableTiming));
    }
    if (!overlap && c->route_pending) {   // an overlapped exchange is still in flight: order this call after it
        SSM_CUDA(cudaStreamWaitEvent(s, c->ev_route_done, 0));
        c->route_pending = false;
    }
    SSM_CUDA(cudaEventRecord(c->sub_fork, s));
    int first = 0;
    rc = SSM_OK;
    auto cuda_rc = [](cudaError_t e, const char* what) { return e == cudaSuccess ? (int)SSM_OK : cuda_fail(e, what); };
    for (int i = 0; i < nsplit && rc == SSM_OK; ++i) {
        const int n = B / nsplit + (i < B % nsplit ? 1 : 0);
        cudaStream_t ss = c->sub_stream[i];
        SSM_CUDA(cudaStreamWaitEvent(ss, c->sub_fork, 0));
        // no early return between the two offset_buffers calls: the work pointers must always be shifted back
        if (c->stagger && i > 0) SSM_CUDA(cudaStreamWaitEvent(ss, c->sub_cost[i - 1], 0));
        offset_buffers(c, first);
        c->ev_after_cost = c->stagger ? c->sub_cost[i] : nullptr;
        rc = run_sgbm(c, n, dL + first * npix, dR + first * npix, d_disp + first * npix, ss);
        c->ev_after_cost = nullptr;
        if (rc == SSM_OK) {
            // one GPU: the sub-batch fuses its own points; several ranks: routing is one exchange per batch (below)
            if (c->nranks > 1) {
                // the previous batch's exchange may still be reading the depth / label / mask buffers
                if (c->route_pending) rc = cuda_rc(cudaStreamWaitEvent(ss, c->ev_route_done, 0), "cudaStreamWaitEvent(route)");
                if (rc == SSM_OK) rc = run_map_prepare(c, n, d_disp + first * npix, d_sem + first * npix * 3, ss);
            }
            else rc = run_map(c, n, d_disp + first * npix, d_sem + first * npix * 3, d_rgb + first * npix * 3, d_pose + (size_t)first * 16, ss);
        }
        offset_buffers(c, -first);
        if (rc == SSM_OK) {
            SSM_CUDA(cudaEventRecord(c->sub_join[i], ss));
            SSM_CUDA(cudaStreamWaitEvent(s, c->sub_join[i], 0));
            if (overlap) SSM_CUDA(cudaStreamWaitEvent(c->route_stream, c->sub_join[i], 0));
        }
        first += n;
    }
    if (rc == SSM_OK && c->nranks > 1) {
        // The exchange (points to their owners, barrier, inbox fusion) is one step per batch.  With route overlap it runs
        // on its own stream: `s` only covers the stereo half, so the next batch's SGBM starts while this batch's points
        // travel and the ranks' skew at the barrier is absorbed; ssm_synchronize / ssm_map_* wait for it.
        cudaStream_t rs = overlap ? c->route_stream : s;
        rc = run_map_points(c, B, d_sem, d_rgb, d_pose, rs);
        if (rc == SSM_OK && overlap) {
            SSM_CUDA(cudaEventRecord(c->ev_route_done, rs));
            c->route_pending = true;
        }
        if (rc == SSM_OK) rc = refresh_mirror(c, rs);
    } else if (rc == SSM_OK) {
        rc = refresh_mirror(c, s);
    }
    return rc;
}

static int finish_timing(ssm_ctx* c)
{
    // mean per-call stage time over the timed pipeline calls recorded so far
    if (!c->timing || c->ev_used == 0) return SSM_OK;
    double acc[SSM_STAGE_COUNT] = {};
    for (int k = 0; k < c->ev_used; ++k) {
        const int set = (c->ev_set - 1 - k + 2 * ssm_ctx::kEvSets) % ssm_ctx::kEvSets;
        SSM_CUDA(cudaEventSynchronize(c->ev[set][SSM_STAGE_COUNT]));
        for (int i = 0; i < SSM_STAGE_COUNT; ++i) {
            float ms = 0.f;
            SSM_CUDA(cudaEventElapsedTime(&ms, c->ev[set][i], c->ev[set][i + 1]));
            acc[i] += ms;
        }
    }
    for (int i = 0; i < SSM_STAGE_COUNT; ++i) c->stage_ms[i] = (float)(acc[i] / c->ev_used);
    return SSM_OK;
}

static int sync_route(ssm_ctx* c)
{
    if (c->route_stream && c->route_pending) {
        SSM_CUDA(cudaStreamSynchronize(c->route_stream));
        c->route_pending = false;
    }
    return SSM_OK;
}

// Blocking checkpoint of the voxel hash: waits for the work in flight, grows the table while points are parked in the spill
// list (or the table is more than half full), reports sticky errors and the voxel count.
static int check_overflow(ssm_ctx* c, cudaStream_t s, uint64_t* n_voxels)
{
    int rr = sync_route(c);
    if (rr) return rr;
    if (c->user_stream && c->user_stream != s) SSM_CUDA(cudaStreamSynchronize(c->user_stream));
    uint32_t h[8];
    for (int round = 0;; ++round) {
        SSM_CUDA(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, s));
        SSM_CUDA(cudaStreamSynchronize(s));
        if (h[2] & 1u) break;
        const bool crowded = 2ull * h[1] > c->table_slots;
        if (!c->auto_grow || (h[4] == 0 && !crowded) || round >= 8) break;
        int rc = table_grow(c, 4ull * ((uint64_t)h[1] + h[4]), s);
        if (rc == SSM_ERR_CAPACITY && h[4] == 0) { c->auto_grow = false; break; }   // crowded but complete: carry on
        if (rc) return rc;
    }
    if (c->h_mirror) memcpy(c->h_mirror, h, sizeof(h));
    {
        int rp = table_reap(c, true);
        if (rp) return rp;
    }
    if (c->p2p && c->ipc_base) {
        uint32_t flag = 0;
        SSM_CUDA(cudaMemcpy(&flag, static_cast<char*>(c->ipc_base) + 8, sizeof(flag), cudaMemcpyDeviceToHost));
        if (flag) return fail(SSM_ERR_CAPACITY, "peer inbox overflow: more routed points than 2x a local batch (raise max_batch)");
    }
    if ((h[2] & 1u) || h[4] != 0)
        return fail(SSM_ERR_CAPACITY, c->auto_grow ? "voxel hash table and its spill list are full (raise ssm_params.map_capacity)"
                                                   : "voxel hash table is full and cannot grow (device memory, or SSM_NO_GROW)");
    if (h[2] & 2u) return fail(SSM_ERR_CAPACITY, "point outside the 21-bit voxel coordinate range");
    if (h[2] & 4u) return fail(SSM_ERR_COMM, "a peer rank did not reach the exchange step within the time-out (its points are missing from the map)");
    if (n_voxels) *n_voxels = h[1];
    return SSM_OK;
}

}  // namespace ssm

using namespace ssm;

// =================================================================================================
extern "C" {
static void keyframes_free(ssm_ctx* c);

void ssm_default_params(ssm_params* p)
{
    memset(p, 0, sizeof(*p));
    // src/stereo.cpp:16-28
    p->min_disparity = 0; p->num_disparities = 80; p->block_size = 11;
    p->p1 = 4 * 11 * 11; p->p2 = 32 * 11 * 11;
    p->disp12_max_diff = 1; p->pre_filter_cap = 63; p->uniqueness_ratio = 10;
    p->speckle_window_size = 100; p->speckle_range = 32;
    // parameters.txt:37-41,50-54,63
    p->cx = 607.1928; p->cy = 185.2157; p->fx = 718.8560; p->fy = 718.8560; p->baseline = 0.532331858; p->scale = 1000.0;
    p->roix = 20; p->roiy = 5; p->roiz = 40;
    // parameters.txt:97-98
    p->resolution = 0.1; p->max_distance = 40;
    // src/mapper.cpp:42-54
    static const uint8_t pal[12][3] = {{128, 128, 128}, {0, 0, 128}, {128, 192, 192}, {0, 69, 255}, {128, 64, 128}, {222, 40, 60},
                                       {0, 128, 128}, {128, 128, 192}, {128, 64, 64}, {128, 0, 64}, {0, 64, 64}, {192, 128, 0}};
    p->num_labels = 12;
    memcpy(p->palette_bgr, pal, sizeof(pal));
    p->drop_mask = (1u << 0) | (1u << 2) | (1u << 11);
    p->dynamic_mask = (1u << 10) | (1u << 11);
    p->dilate_iterations = 2; p->colour_source = 0;
    p->max_width = 1241; p->max_height = 376; p->max_batch = 1;
    p->map_capacity = 1ull << 22;
}

const char* ssm_last_error(void) { return g_err.c_str(); }
const char* ssm_version(void) { return "semantic_slam_mapping_b200 0.1 (sm_100a)"; }
uint64_t ssm_kernel_launches(const ssm_ctx* c) { return c ? c->launches : 0; }

int ssm_create(const ssm_params* p, int device, ssm_ctx** out)
{
    if (!p || !out) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    int rc = validate(*p);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(SSM_ERR_NO_DEVICE, "no CUDA device available: libssm has no CPU fallback");
    }
    cudaDeviceProp prop;
    SSM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(SSM_ERR_NO_DEVICE, "libssm is built for sm_100a (B200) only");
    SSM_CUDA(cudaSetDevice(device));

    ssm_ctx* c = new ssm_ctx();
    c->p = *p;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_cluster = 16;
    {
        const char* e = getenv("SSM_LEGACY_VERTICAL");
        c->force_legacy_vertical = e && e[0] == '1';
        for (int t = 0; t < 4; ++t) {
            const std::string name = "SSM_TUNE" + std::to_string(t);
            const char* v = getenv(name.c_str());
            static const int defaults[4] = {-1 /* vertical: L2 prefetch distance in rows (automatic: sgbm_vertical.cu) */, 0 /* hsweep: L2 prefetch off */, 0 /* vertical kernel variant */,
                                             -1 /* sub-batch streams: automatic */};
            c->tune[t] = v ? atoi(v) : defaults[t];
        }
        const char* lh = getenv("SSM_LEGACY_HSWEEP");
        c->force_legacy_hsweep = lh && lh[0] == '1';
        const char* lc = getenv("SSM_LEGACY_COST");
        c->force_legacy_cost = lc && lc[0] == '1';
        const char* ls = getenv("SSM_LEGACY_SELECT");
        c->force_legacy_select = ls && ls[0] == '1';
        const char* np = getenv("SSM_NO_PAD");
        c->no_pad = np && np[0] == '1';
        const char* sr = getenv("SSM_SELECT_ROWS");
        if (sr && atoi(sr) > 0) c->select_rows = atoi(sr);
        const char* nt = getenv("SSM_NO_COST_TMA");
        c->no_cost_tma = nt && nt[0] == '1';
        const char* ng = getenv("SSM_NO_GROW");
        c->auto_grow = !(ng && ng[0] == '1');
        const char* sg = getenv("SSM_STAGGER");
        if (sg) c->stagger = atoi(sg);
        const char* m = getenv("SSM_MAX_CLUSTER");
        if (m && atoi(m) > 0) c->max_cluster = atoi(m);
        const char* n = getenv("SSM_MIN_CLUSTER");
        if (n && atoi(n) > 0) c->min_cluster = atoi(n);
    }
    c->cap_w = p->max_width; c->cap_h = p->max_height; c->cap_b = p->max_batch;
    const size_t npix = (size_t)c->cap_w * c->cap_h * c->cap_b;
    const size_t ncell = (size_t)(c->cap_w - p->num_disparities) * c->cap_h * c->cap_b * layout_disparities(c, p->num_disparities);
    uint64_t slots = 1024;
    while (slots < p->map_capacity) slots <<= 1;
    c->table_slots = slots;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& set : c->ev)
        for (auto& ev : set) A(cudaEventCreate(&ev));
    A(dalloc(&c->d_left, npix)); A(dalloc(&c->d_right, npix));
    A(dalloc(&c->d_recL, npix)); A(dalloc(&c->d_recR, npix));
    if (layout_disparities(c, p->num_disparities) == 128 || layout_disparities(c, p->num_disparities) == 256) {   // right image in table format for k_cost_tma (sgbm_cost.cu)
        const int ld = layout_disparities(c, p->num_disparities);
        const size_t margin = ld != p->num_disparities ? (size_t)(ld - p->num_disparities + 4 + 3) / 4 * 4 : 0;
        A(dalloc(&c->d_ptab, (margin + (size_t)(c->cap_w + 3) / 4 * 4 + 4) * 6 * c->cap_h * c->cap_b));
    }
    A(dalloc(&c->d_C, ncell)); A(dalloc(&c->d_S, ncell));
    c->d_hs = c->d_S;   // horizontal sums are dead once C exists; S is written afterwards
    A(dalloc(&c->d_disp_raw, npix)); A(dalloc(&c->d_disp_lr, npix)); A(dalloc(&c->d_disp_med, npix)); A(dalloc(&c->d_disp, npix));
    A(dalloc(&c->d_disp2key, npix)); A(dalloc(&c->d_wta_rec, npix * (layout_disparities(c, p->num_disparities) > 128 ? 4 : 2)));   // 16-byte records, 32-byte ones for 256-disparity layouts
    if (p->num_disparities <= 256)
        A(dalloc(&c->d_ck, hsweep2_ck_words(c->cap_w - p->num_disparities, layout_disparities(c, p->num_disparities), c->cap_h, c->cap_b))); A(dalloc(&c->d_uniq_thr, (size_t)32768)); A(dalloc(&c->d_cc_label, npix)); A(dalloc(&c->d_cc_size, npix));
    A(dalloc(&c->d_depth, npix)); A(dalloc(&c->d_label, npix)); A(dalloc(&c->d_mask, npix)); A(dalloc(&c->d_label_lut, (size_t)1 << 24));
    A(dalloc(&c->d_sem, npix * 3)); A(dalloc(&c->d_rgb, npix * 3));
    A(dalloc(&c->d_pose, (size_t)16 * c->cap_b)); A(dalloc(&c->d_min_disp, (size_t)c->cap_b));
    A(dalloc(&c->d_points, npix)); A(dalloc(&c->d_blk_count, npix / 1024 + 2)); A(dalloc(&c->d_counters, 16));
    A(dalloc(&c->d_table, slots));
    c->spill_cap = std::min<size_t>(std::max<size_t>(npix, (size_t)1 << 16), (size_t)1 << 22);
    A(dalloc(&c->d_spill[0], c->spill_cap)); A(dalloc(&c->d_spill[1], c->spill_cap));
    A(cudaHostAlloc(reinterpret_cast<void**>(&c->h_mirror), 8 * sizeof(uint32_t), cudaHostAllocDefault));
    if (c->h_mirror) memset(c->h_mirror, 0, 8 * sizeof(uint32_t));
    if (e != cudaSuccess) {
        free_all(c);
        delete c;
        return cuda_fail(e, "ssm_create allocation");
    }
    fill_dev_params(c, c->cap_w, c->cap_h);
    {   // uniqueness: S[d]*(100-u) < minS*100  <=>  S[d] < thr[minS]   (exact integer threshold, capped at 32768)
        std::vector<uint32_t> thr(32768);
        const int u = c->dp.uniq;
        for (int ms = 0; ms < 32768; ++ms) {
            long long t = u >= 100 ? (ms > 0 ? 32768 : 0) : ((long long)ms * 100 + (100 - u) - 1) / (100 - u);
            thr[ms] = (uint32_t)std::min<long long>(t, 32768);
        }
        if (cudaMemcpy(c->d_uniq_thr, thr.data(), thr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            free_all(c);
            delete c;
            return cuda_fail(cudaGetLastError(), "uniqueness table upload");
        }
    }
    {   // colour -> class table: 255 everywhere, then the palette entries (highest id first, so the lowest id wins a duplicate colour)
        cudaError_t le = cudaMemset(c->d_label_lut, SSM_LABEL_UNKNOWN, (size_t)1 << 24);
        for (int i = p->num_labels - 1; i >= 0 && le == cudaSuccess; --i) {
            const uint8_t id = (uint8_t)i;
            const size_t key = (size_t)p->palette_bgr[i][0] | ((size_t)p->palette_bgr[i][1] << 8) | ((size_t)p->palette_bgr[i][2] << 16);
            le = cudaMemcpy(c->d_label_lut + key, &id, 1, cudaMemcpyHostToDevice);
        }
        if (le != cudaSuccess) {
            free_all(c);
            delete c;
            return cuda_fail(le, "label table upload");
        }
    }
    rc = launch_map_clear(c, c->stream);
    if (rc == SSM_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "ssm_create");
    if (rc) {
        free_all(c);
        delete c;
        return rc;
    }
    *out = c;
    return SSM_OK;
}

void ssm_destroy(ssm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    ssm_comm_destroy(c);
    cues_free(c);
    labels_free(c);
    ingest_free(c);
    keyframes_free(c);
    free_all(c);
    delete c;
}

int ssm_set_stage_timing(ssm_ctx* c, int enabled)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    c->timing = enabled != 0;
    c->ev_used = 0;
    c->ev_set = 0;
    return SSM_OK;
}
int ssm_stage_time_ms(ssm_ctx* c, int stage, float* ms)
{
    if (!c || !ms || stage < 0 || stage >= SSM_STAGE_COUNT) return fail(SSM_ERR_INVALID_ARGUMENT, "bad stage");
    SSM_ENTER(c);
    if (!c->timing || c->ev_used == 0) return fail(SSM_ERR_INVALID_ARGUMENT, "stage timing is off or no pipeline call was timed");
    int rc = finish_timing(c);   // waits for the last timed pipeline call
    if (rc) return rc;
    *ms = c->stage_ms[stage];
    return SSM_OK;
}
int ssm_synchronize(ssm_ctx* c)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return sync_route(c);
}

int ssm_set_route_overlap(ssm_ctx* c, int enabled)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    int rc = sync_route(c);
    if (rc) return rc;
    c->route_overlap = enabled != 0;
    return SSM_OK;
}

// ---- stereo.h -------------------------------------------------------------------------------------
int ssm_sgbm_batch_device(ssm_ctx* c, int batch, const uint8_t* dL, const uint8_t* dR, int w, int h, int16_t* d_disp, void* stream)
{
    if (!c || !dL || !dR || !d_disp) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, batch);
    if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    return run_sgbm(c, batch, dL, dR, d_disp, s);
}

int ssm_sgbm(ssm_ctx* c, const uint8_t* left, const uint8_t* right, int w, int h, size_t stride, int16_t* disp, size_t disp_stride)
{
    if (!c || !left || !right || !disp) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    if (stride < (size_t)w || disp_stride < (size_t)w * 2) return fail(SSM_ERR_INVALID_ARGUMENT, "stride smaller than a row");
    int rc = set_shape(c, w, h, 1);
    if (rc) return rc;
    SSM_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpy2DAsync(c->d_left, w, left, stride, w, h, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpy2DAsync(c->d_right, w, right, stride, w, h, cudaMemcpyHostToDevice, s));
    if ((rc = run_sgbm(c, 1, c->d_left, c->d_right, c->d_disp, s))) return rc;
    SSM_CUDA(cudaMemcpy2DAsync(disp, disp_stride, c->d_disp, (size_t)w * 2, (size_t)w * 2, h, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

int ssm_debug_copy_volume(ssm_ctx* c, int which, int bi, void* dst, size_t bytes)
{
    if (!c || !dst || bi < 0 || bi >= c->cap_b) return fail(SSM_ERR_INVALID_ARGUMENT, "bad argument");
    SSM_ENTER(c);
    const DevParams& p = c->dp;
    const size_t cells = (size_t)p.H * p.W1 * p.D, lcells = (size_t)p.H * p.W1 * p.Dl, npix = (size_t)p.H * p.W;
    const void* src = nullptr;
    size_t need = 0;
    switch (which) {
        case 0: src = c->d_C + lcells * bi; need = cells * 2; break;
        case 1: src = c->d_S + lcells * bi; need = cells * 2; break;
        case 2: src = c->d_disp_lr + npix * bi; need = npix * 2; break;
        case 3: src = c->d_disp_med + npix * bi; need = npix * 2; break;
        default: return fail(SSM_ERR_INVALID_ARGUMENT, "unknown volume id");
    }
    if (bytes < need) return fail(SSM_ERR_INVALID_ARGUMENT, "destination too small");
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    if (which <= 1 && p.Dl != p.D)   // padded layout: the caller gets the D valid disparities of every column, densely packed
        SSM_CUDA(cudaMemcpy2D(dst, (size_t)p.D * 2, src, (size_t)p.Dl * 2, (size_t)p.D * 2, (size_t)p.H * p.W1, cudaMemcpyDeviceToHost));
    else
        SSM_CUDA(cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    return SSM_OK;
}

// ---- glue: disparity -> depth ------------------------------------------------------------------------
int ssm_disparity_to_depth(ssm_ctx* c, const int16_t* disp, int w, int h, size_t disp_stride, uint16_t* depth, size_t depth_stride)
{
    if (!c || !disp || !depth) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, 1);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpy2DAsync(c->d_disp, (size_t)w * 2, disp, disp_stride, (size_t)w * 2, h, cudaMemcpyHostToDevice, s));
    if ((rc = launch_depth(c, 1, c->d_disp, c->d_depth, s))) return rc;
    SSM_CUDA(cudaMemcpy2DAsync(depth, depth_stride, c->d_depth, (size_t)w * 2, (size_t)w * 2, h, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

// ---- mapper.h ----------------------------------------------------------------------------------------
int ssm_semantic_motion_fuse(ssm_ctx* c, const uint8_t* sem, int w, int h, size_t stride, uint8_t* mask, size_t mask_stride)
{
    if (!c || !sem || !mask) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, 1);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    SSM_CUDA(cudaMemcpy2DAsync(c->d_sem, (size_t)w * 3, sem, stride, (size_t)w * 3, h, cudaMemcpyHostToDevice, s));
    if ((rc = launch_labels_mask(c, 1, c->d_sem, s))) return rc;
    SSM_CUDA(cudaMemcpy2DAsync(mask, mask_stride, c->d_mask, w, w, h, cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    return SSM_OK;
}

static int upload_frame(ssm_ctx* c, const uint16_t* depth, const uint8_t* sem, const uint8_t* rgb, int w, int h, const double* T,
                        cudaStream_t s)
{
    const size_t npix = (size_t)w * h;
    SSM_CUDA(cudaMemcpyAsync(c->d_depth, depth, npix * 2, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_sem, sem, npix * 3, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_rgb, rgb, npix * 3, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_pose, T, 16 * sizeof(double), cudaMemcpyHostToDevice, s));
    return SSM_OK;
}

int ssm_generate_point_cloud(ssm_ctx* c, const uint16_t* depth, const uint8_t* sem, const uint8_t* rgb, int w, int h,
                             const double* T, float* xyz, uint32_t* rgba, uint8_t* label, int max_points, int* n_points)
{
    if (!c || !depth || !sem || !rgb || !T || !n_points) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, 1);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    if ((rc = upload_frame(c, depth, sem, rgb, w, h, T, s))) return rc;
    if ((rc = launch_labels_mask(c, 1, c->d_sem, s))) return rc;
    if ((rc = launch_points(c, 1, c->d_depth, c->d_sem, c->d_rgb, c->d_pose, false, s))) return rc;
    uint32_t n = 0;
    SSM_CUDA(cudaMemcpyAsync(&n, c->d_counters, sizeof(n), cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    *n_points = (int)n;
    const uint32_t m = std::min<uint32_t>(n, (uint32_t)std::max(max_points, 0));
    if (m && (xyz || rgba || label)) {
        std::vector<Point> pts(m);
        SSM_CUDA(cudaMemcpy(pts.data(), c->d_points, sizeof(Point) * m, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < m; ++i) {
            if (xyz) { xyz[3 * i] = pts[i].x; xyz[3 * i + 1] = pts[i].y; xyz[3 * i + 2] = pts[i].z; }
            if (rgba) rgba[i] = pts[i].rgba;
            if (label) label[i] = (uint8_t)pts[i].label;
        }
    }
    return SSM_OK;
}

int ssm_map_integrate_frame(ssm_ctx* c, const uint16_t* depth, const uint8_t* sem, const uint8_t* rgb, int w, int h, const double* T)
{
    if (!c || !depth || !sem || !rgb || !T) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, 1);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    if ((rc = upload_frame(c, depth, sem, rgb, w, h, T, s))) return rc;
    if ((rc = launch_labels_mask(c, 1, c->d_sem, s))) return rc;
    if (c->nranks > 1 && c->p2p) {
        if ((rc = points_route_p2p(c, 1, c->d_depth, c->d_sem, c->d_rgb, c->d_pose, s))) return rc;
    } else if (c->nranks > 1) {
        if ((rc = launch_points(c, 1, c->d_depth, c->d_sem, c->d_rgb, c->d_pose, false, s))) return rc;
        if ((rc = route_and_fuse(c, s))) return rc;
    } else if ((rc = launch_points(c, 1, c->d_depth, c->d_sem, c->d_rgb, c->d_pose, true, s))) {
        return rc;
    }
    return check_overflow(c, s, nullptr);
}

// ---- keyframe cache + redraw (mapper.cpp:17-20, :121-149; poses are rewritten by the pose graph, pose_graph.cpp:253-260) ----
namespace {
struct Keyframe {
    Point* d_pts = nullptr;    // camera-frame cloud, row-major order
    uint32_t n = 0;
    double T[16];
    bool live = false;
};
struct KeyframeStore { std::vector<Keyframe> kf; };
KeyframeStore* kf_store(ssm_ctx* c)
{
    if (!c->keyframes) c->keyframes = new KeyframeStore();
    return static_cast<KeyframeStore*>(c->keyframes);
}
}  // namespace

static void keyframes_free(ssm_ctx* c)
{
    KeyframeStore* st = static_cast<KeyframeStore*>(c->keyframes);
    if (!st) return;
    for (auto& k : st->kf)
        if (k.d_pts) cudaFree(k.d_pts);
    delete st;
    c->keyframes = nullptr;
}

int ssm_keyframe_add(ssm_ctx* c, const uint16_t* depth, const uint8_t* sem, const uint8_t* rgb, int w, int h, const double* T, int* id_out)
{
    if (!c || !depth || !sem || !rgb || !T || !id_out) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    if (c->nranks > 1) return fail(SSM_ERR_UNSUPPORTED, "the keyframe cache is per context; with a communicator use ssm_map_integrate_frame");
    int rc = set_shape(c, w, h, 1);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if ((rc = upload_frame(c, depth, sem, rgb, w, h, I, s))) return rc;   // identity pose: the compact cloud stays in camera coordinates
    if ((rc = launch_labels_mask(c, 1, c->d_sem, s))) return rc;
    if ((rc = launch_points(c, 1, c->d_depth, c->d_sem, c->d_rgb, c->d_pose, false, s))) return rc;
    uint32_t n = 0;
    SSM_CUDA(cudaMemcpyAsync(&n, c->d_counters, sizeof(n), cudaMemcpyDeviceToHost, s));
    SSM_CUDA(cudaStreamSynchronize(s));
    Keyframe k;
    k.n = n;
    k.live = true;
    std::memcpy(k.T, T, sizeof(k.T));
    if (n) {
        SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&k.d_pts), sizeof(Point) * n));
        SSM_CUDA(cudaMemcpyAsync(k.d_pts, c->d_points, sizeof(Point) * n, cudaMemcpyDeviceToDevice, s));
        SSM_CUDA(cudaStreamSynchronize(s));
    }
    KeyframeStore* st = kf_store(c);
    st->kf.push_back(k);
    *id_out = (int)st->kf.size() - 1;
    return SSM_OK;
}

static Keyframe* kf_get(ssm_ctx* c, int id)
{
    KeyframeStore* st = static_cast<KeyframeStore*>(c->keyframes);
    if (!st || id < 0 || id >= (int)st->kf.size() || !st->kf[id].live) return nullptr;
    return &st->kf[id];
}

int ssm_keyframe_set_pose(ssm_ctx* c, int id, const double* T)
{
    if (!c || !T) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    Keyframe* k = kf_get(c, id);
    if (!k) return fail(SSM_ERR_INVALID_ARGUMENT, "unknown keyframe id");
    std::memcpy(k->T, T, sizeof(k->T));
    return SSM_OK;
}

int ssm_keyframe_release(ssm_ctx* c, int id)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    Keyframe* k = kf_get(c, id);
    if (!k) return fail(SSM_ERR_INVALID_ARGUMENT, "unknown keyframe id");
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    if (k->d_pts) cudaFree(k->d_pts);
    k->d_pts = nullptr; k->n = 0; k->live = false;
    return SSM_OK;
}

int ssm_keyframe_count(ssm_ctx* c, int* n_live, uint64_t* n_points)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    KeyframeStore* st = static_cast<KeyframeStore*>(c->keyframes);
    int live = 0;
    uint64_t pts = 0;
    if (st)
        for (auto& k : st->kf)
            if (k.live) { ++live; pts += k.n; }
    if (n_live) *n_live = live;
    if (n_points) *n_points = pts;
    return SSM_OK;
}

// fuse the cached clouds of the given keyframes (ids == NULL: every live keyframe) under their current poses
static int integrate_keyframes(ssm_ctx* c, const int* ids, int n, bool clear_first)
{
    if (!c || n < 0) return fail(SSM_ERR_INVALID_ARGUMENT, "bad argument");
    SSM_ENTER(c);
    KeyframeStore* st = static_cast<KeyframeStore*>(c->keyframes);
    cudaStream_t s = c->stream;
    int rc;
    if (ids)
        for (int i = 0; i < n; ++i)
            if (!kf_get(c, ids[i])) return fail(SSM_ERR_INVALID_ARGUMENT, "unknown keyframe id");
    if (clear_first && (rc = launch_map_clear(c, s))) return rc;
    const int count = ids ? n : (st ? (int)st->kf.size() : 0);
    for (int i = 0; i < count; ++i) {
        const Keyframe* k = ids ? kf_get(c, ids[i]) : (st->kf[i].live ? &st->kf[i] : nullptr);
        if (!k) continue;
        if ((rc = launch_transform_fuse(c, k->d_pts, k->n, k->T, s))) return rc;
    }
    return check_overflow(c, s, nullptr);
}

int ssm_map_redraw(ssm_ctx* c, const int* ids, int n) { return integrate_keyframes(c, ids, n, true); }
int ssm_map_integrate_keyframes(ssm_ctx* c, const int* ids, int n) { return integrate_keyframes(c, ids, n, false); }

int ssm_map_integrate_points(ssm_ctx* c, const float* xyz, const uint32_t* rgba, const uint8_t* label, int n)
{
    if (!c || (n > 0 && (!xyz || !rgba || !label)) || n < 0) return fail(SSM_ERR_INVALID_ARGUMENT, "bad argument");
    SSM_ENTER(c);
    const size_t cap = (size_t)c->cap_w * c->cap_h * c->cap_b;
    cudaStream_t s = c->stream;
    std::vector<Point> pts;
    // With a communicator every round is collective (bucket, all-gather of the counts, grouped send / recv), so the ranks
    // agree on the number of rounds first: the largest ceil(n / cap) of any rank, at least one; a rank without points left
    // takes part with empty rounds.
    size_t rounds = ((size_t)n + cap - 1) / cap;
    if (c->nranks > 1) {
        uint32_t r32 = (uint32_t)std::max<size_t>(rounds, 1);
        int rc0 = comm_allreduce_max(c, &r32, s);
        if (rc0) return rc0;
        rounds = r32;
    }
    for (size_t round = 0; round < rounds; ++round) {
        const size_t base = std::min(round * cap, (size_t)n);
        const size_t m = std::min(cap, (size_t)n - base);
        pts.resize(m);
        for (size_t i = 0; i < m; ++i) {
            const size_t q = base + i;
            pts[i] = Point{xyz[3 * q], xyz[3 * q + 1], xyz[3 * q + 2], rgba[q] & 0xffffffu, (uint32_t)label[q]};
        }
        if (m) SSM_CUDA(cudaMemcpyAsync(c->d_points, pts.data(), sizeof(Point) * m, cudaMemcpyHostToDevice, s));
        int rc;
        if (c->nranks > 1) {
            const uint32_t mm = (uint32_t)m;
            SSM_CUDA(cudaMemcpyAsync(c->d_counters, &mm, sizeof(mm), cudaMemcpyHostToDevice, s));
            if ((rc = route_and_fuse(c, s))) return rc;
        } else if ((rc = launch_fuse_points(c, c->d_points, nullptr, (uint32_t)m, s))) {
            return rc;
        }
        SSM_CUDA(cudaStreamSynchronize(s));
    }
    return check_overflow(c, s, nullptr);
}

int ssm_map_clear(ssm_ctx* c)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    int rc = sync_route(c);
    if (rc) return rc;
    if (c->user_stream) SSM_CUDA(cudaStreamSynchronize(c->user_stream));
    rc = launch_map_clear(c, c->stream);
    if (rc) return rc;
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_mirror) memset(c->h_mirror, 0, 8 * sizeof(uint32_t));
    c->mirror_prev = 0;
    c->pipeline_calls = 0;
    return SSM_OK;
}

int ssm_map_size(ssm_ctx* c, uint64_t* n)
{
    if (!c || !n) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    return check_overflow(c, c->stream, n);
}

int ssm_map_export(ssm_ctx* c, const ssm_voxel_export* out, uint64_t max_voxels, int sorted, uint64_t* n_out)
{
    if (!c || !out || !n_out) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    uint64_t n = 0;
    int rc = check_overflow(c, c->stream, &n);
    if (rc) return rc;
    // K9 on the device: index the occupied records, radix-sort them into pcl::VoxelGrid's output order, finalize, copy out
    return export_records_device(c, c->d_table, c->table_slots, n, sorted != 0, out, max_voxels, n_out, c->stream, &c->last_export_ms);
}

int ssm_map_export_device_ms(ssm_ctx* c, float* ms)
{
    if (!c || !ms) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    *ms = c->last_export_ms;
    return SSM_OK;
}

// Every rank calls this; rank 0 receives the union of the ranks' tables (disjoint by ownership) in pcl::VoxelGrid order.
int ssm_map_export_gathered(ssm_ctx* c, const ssm_voxel_export* out, uint64_t max_voxels, int sorted, uint64_t* n_out)
{
    if (!c || !n_out) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    if (c->nranks <= 1) return ssm_map_export(c, out, max_voxels, sorted, n_out);
    uint64_t n = 0;
    int rc = check_overflow(c, c->stream, &n);   // (an error here leaves the peers waiting in the collective: the caller tears down)
    if (rc) return rc;
    Voxel* d_all = nullptr;
    uint64_t n_all = 0;
    if ((rc = comm_gather_records(c, &d_all, &n_all, c->stream))) return rc;
    *n_out = 0;
    if (c->rank == 0) {
        if (!out) rc = fail(SSM_ERR_INVALID_ARGUMENT, "rank 0 needs the output arrays");
        else rc = export_records_device(c, d_all, n_all, n_all, sorted != 0, out, max_voxels, n_out, c->stream, &c->last_export_ms);
    }
    if (d_all) cudaFree(d_all);
    return rc;
}

int ssm_map_save_pcd(ssm_ctx* c, const char* path)
{
    if (!c || !path) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    uint64_t n = 0;
    int rc = check_overflow(c, c->stream, &n);
    if (rc) return rc;
    std::vector<float> xyz(3 * n);
    std::vector<uint32_t> rgba(n);
    ssm_voxel_export ex = {};
    ex.xyz = xyz.data();
    ex.rgba = rgba.data();
    uint64_t got = 0;
    if ((rc = export_records_device(c, c->d_table, c->table_slots, n, true, &ex, n, &got, c->stream, nullptr))) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SSM_ERR_INVALID_ARGUMENT, std::string("cannot open ") + path);
    // pcl::PCDWriter::write (ASCII header, binary body) for PointXYZRGBA: x y z rgba
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\n"
               "COUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", (size_t)got, (size_t)got);
    for (uint64_t q = 0; q < got; ++q) {
        fwrite(&xyz[3 * q], sizeof(float), 3, f);
        fwrite(&rgba[q], sizeof(uint32_t), 1, f);
    }
    fclose(f);
    return SSM_OK;
}

int ssm_map_reserve(ssm_ctx* c, uint64_t slots)
{
    if (!c) return fail(SSM_ERR_INVALID_ARGUMENT, "null ctx");
    SSM_ENTER(c);
    int rc = check_overflow(c, c->stream, nullptr);
    if (rc) return rc;
    if (slots <= c->table_slots) return SSM_OK;
    if ((rc = table_grow(c, slots, c->stream))) return rc;
    SSM_CUDA(cudaStreamSynchronize(c->stream));
    return SSM_OK;
}

int ssm_map_stats(ssm_ctx* c, ssm_map_statistics* st)
{
    if (!c || !st) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    uint64_t n = 0;
    int rc = check_overflow(c, c->stream, &n);
    if (rc) return rc;
    unsigned long long t[3] = {0, 0, 0};
    if ((rc = table_stats(c, c->stream, t))) return rc;
    st->slots = c->table_slots;
    st->voxels = t[0];
    st->load_factor = (double)t[0] / (double)c->table_slots;
    st->mean_probe = t[0] ? (double)t[1] / (double)t[0] : 0.0;
    st->max_probe = t[2];
    st->grow_steps = c->grows;
    st->table_bytes = c->table_slots * sizeof(Voxel);
    return SSM_OK;
}

// ---- the whole path, batched --------------------------------------------------------------------------
int ssm_pipeline_batch_device(ssm_ctx* c, int batch, const uint8_t* dL, const uint8_t* dR, const uint8_t* d_sem,
                              const uint8_t* d_rgb, const double* d_poses, int w, int h, int16_t* d_disp_out, void* stream)
{
    if (!c || !dL || !dR || !d_sem || !d_rgb || !d_poses) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, batch);
    if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    c->user_stream = stream ? (cudaStream_t)stream : nullptr;
    int16_t* d_disp = d_disp_out ? d_disp_out : c->d_disp;
    return run_pipeline(c, batch, dL, dR, d_sem, d_rgb, d_poses, d_disp, s);
}

int ssm_pipeline_batch_host_async(ssm_ctx* c, int batch, const uint8_t* left, const uint8_t* right, const uint8_t* sem,
                                  const uint8_t* rgb, const double* poses, int w, int h, uint32_t* n_voxels_pinned)
{
    if (!c || !left || !right || !sem || !rgb || !poses) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, batch);
    if (rc) return rc;
    SSM_CUDA(cudaSetDevice(c->device));
    const size_t cap = (size_t)c->cap_w * c->cap_h * c->cap_b;
    if (!c->copy_stream) {
        SSM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            SSM_CUDA(cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
            SSM_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[k], cudaEventDisableTiming));
            SSM_CUDA(cudaMalloc(&c->stage_left[k], cap));
            SSM_CUDA(cudaMalloc(&c->stage_right[k], cap));
            SSM_CUDA(cudaMalloc(&c->stage_sem[k], cap * 3));
            SSM_CUDA(cudaMalloc(&c->stage_rgb[k], cap * 3));
            SSM_CUDA(cudaMalloc(&c->stage_pose[k], sizeof(double) * 16 * c->cap_b));
        }
    }
    const int k = (int)(c->async_calls & 1u);
    const bool reused = c->async_calls >= 2;
    c->async_calls++;
    cudaStream_t cs = c->copy_stream, s = c->stream;
    const size_t npix = (size_t)w * h * batch;
    if (reused) SSM_CUDA(cudaStreamWaitEvent(cs, c->ev_consumed[k], 0));   // the kernels that read this staging set are done
    SSM_CUDA(cudaMemcpyAsync(c->stage_left[k], left, npix, cudaMemcpyHostToDevice, cs));
    SSM_CUDA(cudaMemcpyAsync(c->stage_right[k], right, npix, cudaMemcpyHostToDevice, cs));
    SSM_CUDA(cudaMemcpyAsync(c->stage_sem[k], sem, npix * 3, cudaMemcpyHostToDevice, cs));
    SSM_CUDA(cudaMemcpyAsync(c->stage_rgb[k], rgb, npix * 3, cudaMemcpyHostToDevice, cs));
    SSM_CUDA(cudaMemcpyAsync(c->stage_pose[k], poses, sizeof(double) * 16 * batch, cudaMemcpyHostToDevice, cs));
    SSM_CUDA(cudaEventRecord(c->ev_copied[k], cs));
    SSM_CUDA(cudaStreamWaitEvent(s, c->ev_copied[k], 0));
    if ((rc = run_pipeline(c, batch, c->stage_left[k], c->stage_right[k], c->stage_sem[k], c->stage_rgb[k], c->stage_pose[k], c->d_disp, s)))
        return rc;
    // with route overlap the exchange of this batch runs on the route stream and reads the staged semantic / rgb / poses
    cudaStream_t done = (c->route_pending && c->route_stream) ? c->route_stream : s;
    SSM_CUDA(cudaEventRecord(c->ev_consumed[k], done));
    if (n_voxels_pinned) SSM_CUDA(cudaMemcpyAsync(n_voxels_pinned, c->d_counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, done));
    return SSM_OK;
}

int ssm_pipeline_batch_host(ssm_ctx* c, int batch, const uint8_t* left, const uint8_t* right, const uint8_t* sem,
                            const uint8_t* rgb, const double* poses, int w, int h, int16_t* disp_out, uint64_t* n_voxels_out)
{
    if (!c || !left || !right || !sem || !rgb || !poses) return fail(SSM_ERR_INVALID_ARGUMENT, "null argument");
    SSM_ENTER(c);
    int rc = set_shape(c, w, h, batch);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    const size_t npix = (size_t)w * h * batch;
    SSM_CUDA(cudaMemcpyAsync(c->d_left, left, npix, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_right, right, npix, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_sem, sem, npix * 3, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_rgb, rgb, npix * 3, cudaMemcpyHostToDevice, s));
    SSM_CUDA(cudaMemcpyAsync(c->d_pose, poses, sizeof(double) * 16 * batch, cudaMemcpyHostToDevice, s));
    if ((rc = run_pipeline(c, batch, c->d_left, c->d_right, c->d_sem, c->d_rgb, c->d_pose, c->d_disp, s))) return rc;
    if (disp_out) SSM_CUDA(cudaMemcpyAsync(disp_out, c->d_disp, npix * 2, cudaMemcpyDeviceToHost, s));
    return check_overflow(c, s, n_voxels_out);
}

}  // extern "C"
