// mapper.cu -- disparity -> depth -> labelled cloud -> voxel-hash fusion, sm_100a.
//
// Replaces, for the drop-in path:
//   FrameReader::next depth loop              /root/reference src/rgbdframe.cpp:85-116   (k_min_disp, k_depth)
//   Mapper::semantic_motion_fuse              src/mapper.cpp:189-216                     (k_labels, k_moving_mask)
//   Mapper::generatePointCloud                src/mapper.cpp:12-94                       (k_points_*)
//   RGBDFrame::project2dTo3d                  include/rgbdframe.h:63-75
//   pcl::transformPointCloud(Matrix4d)        src/mapper.cpp:90-91
//   pcl::VoxelGrid<PointXYZRGBA>::filter      src/mapper.cpp:106-107,154-155 (k_fuse, k_export)
// fp64 expressions use explicit __dmul_rn/__dadd_rn/__ddiv_rn so that no FMA contraction can change a bit
// relative to the reference's (unfused) double arithmetic.
#include <algorithm>

#include "ssm_internal.cuh"

namespace ssm {

// DevParams travels as a __grid_constant__ kernel argument (constant bank, no global state between contexts).
#define SSM_DP const __grid_constant__ DevParams p

// ------------------------------------------------------------------------------------------------
// rgbdframe.cpp:85  cv::minMaxIdx(disp_sgbm, &minDisparity)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_min_disp(const int16_t* __restrict__ disp, int* __restrict__ min_disp,
                                                  size_t per_frame, int B)
{
    const int b = blockIdx.y;
    if (b >= B) return;
    const int16_t* d = disp + (size_t)b * per_frame;
    int m = 32767;
    // 8 values per 16-byte load where the frame is aligned; scalar otherwise / for the tail
    const size_t nvec = (reinterpret_cast<uintptr_t>(d) & 15) == 0 ? per_frame / 8 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4*>(d)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) m = min(m, min((int)(int16_t)(w[k] & 0xffffu), (int)(int16_t)(w[k] >> 16)));
    }
    for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_frame; i += (size_t)gridDim.x * blockDim.x)
        m = min(m, (int)d[i]);
    m = __reduce_min_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMin(&min_disp[b], m);
}

// rgbdframe.cpp:97-116.  One thread makes 4 consecutive pixels of the flat [B][H][W] array (8-byte load / store when the
// buffers are aligned).
__device__ __forceinline__ uint16_t depth_of(const DevParams& p, int d, int dmin, int u, int v)
{
    if (d == 0 || d == dmin) return 0;
    const double pw = __ddiv_rn(p.baseline, (double)d);
    const double px = __dmul_rn(__dmul_rn(__dsub_rn((double)u, p.cx), pw), 16.0);
    const double py = __dmul_rn(__dmul_rn(__dsub_rn((double)v, p.cy), pw), 16.0);
    const double pz = __dmul_rn(__dmul_rn(p.fx, pw), 16.0);
    if (fabs(px) < p.roix && fabs(py) < p.roiy && fabs(pz) < p.roiz && pz > 0)
        return (uint16_t)(int)__dmul_rn(pz, p.scale);   // truncation, pz*scale < 65536 because roiz*scale is checked at create
    return 0;
}
__global__ void __launch_bounds__(256) k_depth(const int16_t* __restrict__ disp, const int* __restrict__ min_disp,
                                               uint16_t* __restrict__ depth, size_t total, int aligned8, SSM_DP)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= total) return;
    int u = (int)(i0 % p.W);
    const size_t r = i0 / p.W;
    int v = (int)(r % p.H);
    int b = (int)(r / p.H);
    int d[4];
    const bool full = i0 + 4 <= total;
    if (full && aligned8) {
        const uint2 q = *reinterpret_cast<const uint2*>(disp + i0);
        d[0] = (int16_t)(q.x & 0xffffu); d[1] = (int16_t)(q.x >> 16); d[2] = (int16_t)(q.y & 0xffffu); d[3] = (int16_t)(q.y >> 16);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) d[k] = i0 + k < total ? disp[i0 + k] : 0;
    }
    uint32_t o[4];
    int dmin = min_disp[b];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o[k] = depth_of(p, d[k], dmin, u, v);
        if (++u == p.W) {
            u = 0;
            if (++v == p.H) { v = 0; ++b; if (i0 + k + 1 < total) dmin = min_disp[b]; }
        }
    }
    if (full && aligned8) {
        *reinterpret_cast<uint2*>(depth + i0) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < total) depth[i0 + k] = (uint16_t)o[k];
    }
}

// ------------------------------------------------------------------------------------------------
// semantic BGR -> class id (palette lookup), then the dilated dynamic-object mask
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int label_of(const DevParams& p, uint32_t bgr)
{
    int l = SSM_LABEL_UNKNOWN;
#pragma unroll 4
    for (int i = p.num_labels - 1; i >= 0; --i)
        if (p.palette[i] == bgr) l = i;   // lowest matching id wins, like the oracle's first-match scan
    return l;
}

// One thread labels 4 consecutive pixels of the flat [B][H][W] array (three 4-byte loads, one 4-byte store when the
// buffers are aligned) and records the "dynamic class" predicate as one bit per pixel in `dynbits`
// ([B][H][ceil(W/32)] words, zeroed beforehand; bit i of word wx <-> pixel x = 32 * wx + i).
__global__ void __launch_bounds__(256) k_labels(const uint8_t* __restrict__ sem, uint8_t* __restrict__ label, uint32_t* __restrict__ dynbits,
                                                const uint8_t* __restrict__ lut, size_t total, int aligned4, SSM_DP)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= total) return;
    const bool full = i0 + 4 <= total;
    uint32_t bgr[4];
    if (full && aligned4) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(sem + i0 * 3);
        const uint32_t a = q[0], b = q[1], c = q[2];
        bgr[0] = a & 0xffffffu; bgr[1] = (a >> 24) | ((b & 0xffffu) << 8); bgr[2] = (b >> 16) | ((c & 0xffu) << 16); bgr[3] = c >> 8;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* t = sem + (i0 + k) * 3;
            bgr[k] = i0 + k < total ? ((uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16)) : 0xffffffffu;
        }
    }
    uint32_t l[4], dyn = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // 2^24-entry colour -> class table (only num_labels of its lines are ever touched: cache resident)
        l[k] = lut ? (bgr[k] <= 0xffffffu ? (uint32_t)lut[bgr[k]] : (uint32_t)SSM_LABEL_UNKNOWN) : (uint32_t)label_of(p, bgr[k]);
        if (l[k] != SSM_LABEL_UNKNOWN && ((p.dynamic_mask >> l[k]) & 1u) && i0 + k < total) dyn |= 1u << k;
    }
    if (full && aligned4) {
        *reinterpret_cast<uint32_t*>(label + i0) = l[0] | (l[1] << 8) | (l[2] << 16) | (l[3] << 24);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < total) label[i0 + k] = (uint8_t)l[k];
    }
    if (dyn) {
        // the 4 pixels usually share one word: one atomic for the run, a second only where the run crosses a word / row end
        const int wpr = (p.W + 31) >> 5;
        size_t cur = (size_t)-1;
        uint32_t bits = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((dyn >> k) & 1u) {
                const size_t idx = i0 + k;
                const int x = (int)(idx % p.W);
                const size_t row = idx / p.W;              // b * H + y
                const size_t word = row * wpr + (x >> 5);
                if (word != cur) {
                    if (bits) atomicOr(&dynbits[cur], bits);
                    cur = word; bits = 0u;
                }
                bits |= 1u << (x & 31);
            }
        if (bits) atomicOr(&dynbits[cur], bits);
    }
}

// cv::dilate(img, img, ones(3,3), anchor centre, iterations) == (2*it+1)^2 max filter, out-of-image ignored, on the
// bit image, separably: k_dilate_rows ORs the 2R+1 rows in reach (one thread per word), then every output word is 2R
// shifts of that row-dilated word with carry from its two neighbour words.
__global__ void __launch_bounds__(256) k_dilate_rows(const uint32_t* __restrict__ dynbits, uint32_t* __restrict__ vbits, int wpr, int H, int R,
                                                     size_t nwords)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const int wx = (int)(i % wpr);
    const size_t row = i / wpr;
    const int y = (int)(row % H);
    uint32_t c = 0u;
    for (int dy = -R; dy <= R; ++dy) {
        const int yy = y + dy;
        if (yy >= 0 && yy < H) c |= dynbits[(row - y + yy) * wpr + wx];
    }
    vbits[i] = c;
}
__device__ __forceinline__ uint32_t dilated_word(const uint32_t* __restrict__ vbits, int wpr, size_t row, int wx, int R)
{
    const uint32_t* w = vbits + row * wpr + wx;
    const uint32_t c = w[0], l = wx > 0 ? w[-1] : 0u, r = wx + 1 < wpr ? w[1] : 0u;
    uint32_t dil = c;
    for (int k = 1; k <= R; ++k) dil |= (c << k) | (l >> (32 - k)) | (c >> k) | (r << (32 - k));
    return dil;
}
// one thread writes 4 consecutive mask bytes of the flat array
__global__ void __launch_bounds__(256) k_moving_mask(const uint32_t* __restrict__ dynbits, uint8_t* __restrict__ mask, size_t total,
                                                     int aligned4, SSM_DP)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= total) return;
    const int wpr = (p.W + 31) >> 5;
    const int R = p.dilate_radius;
    uint32_t out = 0u;
    int cur_wx = -1;
    size_t cur_row = (size_t)-1;
    uint32_t dil = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t idx = i0 + k;
        if (idx >= total) break;
        const int x = (int)(idx % p.W);
        const size_t row = idx / p.W;
        if (row != cur_row || (x >> 5) != cur_wx) {
            cur_row = row; cur_wx = x >> 5;
            dil = dilated_word(dynbits, wpr, row, cur_wx, R);
        }
        if ((dil >> (x & 31)) & 1u) out |= 0xffu << (8 * k);
    }
    if (i0 + 4 <= total && aligned4) {
        *reinterpret_cast<uint32_t*>(mask + i0) = out;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k < total) mask[i0 + k] = (uint8_t)(out >> (8 * k));
    }
}

// ------------------------------------------------------------------------------------------------
// per-pixel point generation (mapper.cpp:22-86 + rgbdframe.h:63-75 + transformPointCloud)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool make_point_from(const DevParams& p, size_t idx, int d, int mask_v, int l, const uint8_t* __restrict__ sem,
                                                const uint8_t* __restrict__ rgb, const double* __restrict__ pose, Point& out);
__device__ __forceinline__ bool make_point(const DevParams& p, size_t idx, const uint16_t* __restrict__ depth, const uint8_t* __restrict__ label,
                                           const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sem,
                                           const uint8_t* __restrict__ rgb, const double* __restrict__ pose, Point& out)
{
    return make_point_from(p, idx, depth[idx], mask[idx], label[idx], sem, rgb, pose, out);
}
// the same with the pixel's depth / mask / label already in registers (callers that keep several pixels' loads in flight)
__device__ __forceinline__ bool make_point_from(const DevParams& p, size_t idx, int d, int mask_v, int l, const uint8_t* __restrict__ sem,
                                                const uint8_t* __restrict__ rgb, const double* __restrict__ pose, Point& out)
{
    if (d == 0) return false;                                  // mapper.cpp:28
    if ((double)d > p.max_depth_units) return false;           // :30  d > max_distance * camera.scale
    if (mask_v == 255) return false;                           // :32
    if (l != SSM_LABEL_UNKNOWN && ((p.drop_mask >> l) & 1u)) return false;   // :41-55
    const int u = (int)(idx % p.W);
    const size_t r = idx / p.W;
    const int v = (int)(r % p.H);
    const double* T = pose + (r / p.H) * 16;
    // rgbdframe.h:71-73 -- each component is rounded to float on assignment
    const float z = (float)__ddiv_rn((double)d, p.scale);
    const float x = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)u, p.cx), (double)z), p.fx);
    const float y = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)v, p.cy), (double)z), p.fy);
    // Eigen Affine3d * Vector3d, evaluated left to right without contraction
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double acc = __dmul_rn(T[4 * k], (double)x);
        acc = __dadd_rn(acc, __dmul_rn(T[4 * k + 1], (double)y));
        acc = __dadd_rn(acc, __dmul_rn(T[4 * k + 2], (double)z));
        acc = __dadd_rn(acc, T[4 * k + 3]);
        w[k] = (float)acc;
    }
    const uint8_t* c = (p.colour_source == 0 ? rgb : sem) + idx * 3;   // mapper.cpp:72-84 vs mapper.cpp~:60
    out.x = w[0]; out.y = w[1]; out.z = w[2];
    out.rgba = ((uint32_t)c[2] << 16) | ((uint32_t)c[1] << 8) | (uint32_t)c[0];
    out.label = (uint32_t)l;
    return true;
}

// ---- voxel hash insert -------------------------------------------------------------------------
// Warp-collective insert (every lane of a converged warp calls it; `ok` says whether the lane has a point).  Adjacent lanes
// are adjacent pixels, and at map resolution several of them fall into the same voxel: lanes are grouped into runs of equal
// keys, the run's sums are formed with segmented shuffles, and only the run's first lane probes the table and issues the
// atomics -- the kernel is bound by the L2 atomic units, not by issue slots.  All accumulators are integers, so the table is
// bit-identical to the point-by-point insert.
// An insert that finds neither its key nor a free slot within kMaxProbe probes parks the run's points in the spill list
// (TableRef::spill); the host grows the table and re-inserts them (voxel_table.cu).
__device__ __forceinline__ void fuse_point_warp(const DevParams& p, const TableRef& tr, const Point& pt, bool ok)
{
    Voxel* __restrict__ table = tr.table;
    const uint64_t mask = tr.mask;
    uint32_t* counters = tr.counters;
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned long long key = kEmptyKey;
    unsigned long long fx = 0ull, fy = 0ull, fz = 0ull, col = 0ull;
    if (ok && !(isfinite(pt.x) && isfinite(pt.y) && isfinite(pt.z))) ok = false;   // cloud.is_dense == false (mapper.cpp:92)
    if (ok) {
        // pcl::VoxelGrid: ijk = floor(coord * inverse_leaf_size) in fp32
        const int i = (int)floorf(__fmul_rn(pt.x, p.inv_leaf));
        const int j = (int)floorf(__fmul_rn(pt.y, p.inv_leaf));
        const int k = (int)floorf(__fmul_rn(pt.z, p.inv_leaf));
        if (i < -(1 << 20) || i >= (1 << 20) || j < -(1 << 20) || j >= (1 << 20) || k < -(1 << 20) || k >= (1 << 20)) {
            atomicOr(&counters[2], 2u);   // coordinate outside the 21-bit key range
            ok = false;
        } else {
            key = pack_key(i, j, k);
            fx = (unsigned long long)__double2ll_rn(__dmul_rn((double)pt.x, kFixScale));
            fy = (unsigned long long)__double2ll_rn(__dmul_rn((double)pt.y, kFixScale));
            fz = (unsigned long long)__double2ll_rn(__dmul_rn((double)pt.z, kFixScale));
            col = (unsigned long long)((pt.rgba >> 16) & 255u) | ((unsigned long long)((pt.rgba >> 8) & 255u) << 20) |
                  ((unsigned long long)(pt.rgba & 255u) << 40);   // r, g, b in 20-bit fields: 32 lanes x 255 fits
        }
    }
    const uint32_t label = ok && pt.label < (uint32_t)p.num_labels ? pt.label : 0xffffffffu;
    // segment starts: a lane without a point is a segment of its own, so runs never reach across it
    const unsigned long long kprev = __shfl_up_sync(full, key, 1);
    const uint32_t lprev = __shfl_up_sync(full, label, 1);
    const bool head = ok && (lane == 0 || kprev != key);
    const bool vhead = ok && (head || lprev != label);               // start of a run of equal (voxel, label)
    const uint32_t bounds = __ballot_sync(full, head || !ok);
    const uint32_t vbounds = __ballot_sync(full, vhead || !ok);
    const uint32_t after = lane == 31 ? 0u : (bounds >> (lane + 1));
    const int run_to = after ? lane + __ffs(after) - 1 : 31;          // last lane of my run
    const uint32_t vafter = lane == 31 ? 0u : (vbounds >> (lane + 1));
    const uint32_t vcount = (uint32_t)(vafter ? __ffs(vafter) : 32 - lane);
    const int head_lane = 31 - __clz(bounds & (0xffffffffu >> (31 - lane)));   // first lane of my run (bit `lane` or below is set)
    // segmented suffix sums by doubling; as many rounds as the longest run needs
    const int longest = (int)__reduce_max_sync(full, head ? (uint32_t)(run_to - lane + 1) : 0u);
    for (int s = 1; s < longest; s <<= 1) {
        const unsigned long long ox = __shfl_down_sync(full, fx, s), oy = __shfl_down_sync(full, fy, s),
                                 oz = __shfl_down_sync(full, fz, s), oc = __shfl_down_sync(full, col, s);
        if (lane + s <= run_to) { fx += ox; fy += oy; fz += oz; col += oc; }
    }
    // the run's first lane finds (or claims) the voxel's slot
    uint32_t slot32 = 0xffffffffu;
    if (head) {
        uint64_t slot = mix64(key) & mask;
        const uint64_t probe_end = mask < (uint64_t)kMaxProbe ? mask : (uint64_t)kMaxProbe;
        for (uint64_t probe = 0; probe <= probe_end; ++probe, slot = (slot + 1) & mask) {
            Voxel* v = table + slot;
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&v->key);
            if (cur == kEmptyKey) {
                cur = atomicCAS(&v->key, kEmptyKey, key);
                if (cur == kEmptyKey) {
                    atomicAdd(&counters[1], 1u);
                    cur = key;
                }
            }
            if (cur == key) {
                const unsigned long long n = (unsigned long long)(run_to - lane + 1);
                atomicAdd(&v->sx, fx);
                atomicAdd(&v->sy, fy);
                atomicAdd(&v->sz, fz);
                // (n, sr) and (sg, sb) are adjacent 32-bit fields on 8-byte boundaries: one 64-bit add each (no field can carry
                // into its neighbour before 2^32 points / 2^24 points of full intensity)
                atomicAdd(reinterpret_cast<unsigned long long*>(&v->n), n | ((col & 0xfffffull) << 32));
                atomicAdd(reinterpret_cast<unsigned long long*>(&v->sg), ((col >> 20) & 0xfffffull) | (((col >> 40) & 0xfffffull) << 32));
                slot32 = (uint32_t)slot;
                break;
            }
        }
    }
    __syncwarp();
    slot32 = __shfl_sync(full, slot32, head_lane);
    if (vhead && label != 0xffffffffu && slot32 != 0xffffffffu) atomicAdd(&table[slot32].votes[label], vcount);
    if (ok && slot32 == 0xffffffffu) {   // no slot within reach: park the point (rare; every lane of the run parks its own)
        const uint32_t o = tr.spill ? atomicAdd(&counters[4], 1u) : 0xffffffffu;
        if (o < tr.spill_cap) tr.spill[o] = pt;
        else atomicOr(&counters[2], 1u);   // spill list full as well: the table is full for good
    }
}

// L2 prefetch of the record a point's first probe will touch
__device__ __forceinline__ void prefetch_voxel_line(const Voxel* __restrict__ table, uint64_t slot_mask, const Point& pt, float inv_leaf)
{
    const int i = (int)floorf(__fmul_rn(pt.x, inv_leaf)), j = (int)floorf(__fmul_rn(pt.y, inv_leaf)), k = (int)floorf(__fmul_rn(pt.z, inv_leaf));
    const Voxel* v = table + (mix64(pack_key(i, j, k)) & slot_mask);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(v));
}
// FUSE mode: straight from pixels into the hash (single-GPU pipeline).  The kernel waits on memory (the voxel records live
// in a multi-GB table: every probe is a DRAM access), so a thread takes kPixPerThread pixels -- lane l of a warp takes pixels
// l, l + 32, ... of the warp's span, so that adjacent lanes stay adjacent pixels -- with all their input loads, and then all
// their table-line prefetches, in flight together before the first insert.
constexpr int kPixPerThread = 4;
__global__ void __launch_bounds__(256) k_points_fuse(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ label,
                                                     const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sem,
                                                     const uint8_t* __restrict__ rgb, const double* __restrict__ pose,
                                                     TableRef tr, size_t total, SSM_DP)
{
    const int lane = threadIdx.x & 31;
    const size_t span = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane)) * kPixPerThread;   // first pixel of my warp
    int d[kPixPerThread], m[kPixPerThread], l[kPixPerThread];
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        const size_t idx = span + k * 32 + lane;
        const bool in = idx < total;
        d[k] = in ? (int)depth[idx] : 0;
        m[k] = in ? (int)mask[idx] : 0;
        l[k] = in ? (int)label[idx] : 0;
    }
    Point pt[kPixPerThread];
    bool ok[kPixPerThread];
    uint32_t made = 0;
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        const size_t idx = span + k * 32 + lane;
        ok[k] = idx < total && make_point_from(p, idx, d[k], m[k], l[k], sem, rgb, pose, pt[k]);
        if (ok[k]) {
            prefetch_voxel_line(tr.table, tr.mask, pt[k], p.inv_leaf);   // same slot as the insert's first probe; a stray prefetch is harmless
            ++made;
        }
    }
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) fuse_point_warp(p, tr, pt[k], ok[k]);
    // points of this call: one atomic per warp instead of one per point
    made = __reduce_add_sync(0xffffffffu, made);
    if (lane == 0 && made) atomicAdd(&tr.counters[0], made);
}

// P2P mode (multi-GPU): straight from pixels; locally owned points go into this rank's hash, the others are appended
// to their owner's inbox through the peer mapping (NVLink stores).  The compute step (back-projection, transform) and the
// dispatch all-to-all are one kernel.  Inbox slots are reserved in two levels: the lanes of a warp that share an owner
// take consecutive places in the CTA's block for that owner (shared-memory atomic, one per (warp, owner) group), and one
// system-scope atomic per (CTA, owner) -- issued by nranks threads in parallel -- places the block in the peer's inbox:
// one NVLink round trip per CTA instead of one per group in every warp.
__global__ void __launch_bounds__(256) k_points_p2p(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ label,
                                                    const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sem,
                                                    const uint8_t* __restrict__ rgb, const double* __restrict__ pose,
                                                    TableRef tr,
                                                    void* const* __restrict__ peer_base, int rank, int nranks, int parity,
                                                    uint32_t inbox_cap, size_t total, SSM_DP)
{
    __shared__ uint32_t s_cnt[ssm_ctx::kMaxPeers], s_base[ssm_ctx::kMaxPeers];
    const int lane = threadIdx.x & 31;
    if ((int)threadIdx.x < nranks) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const size_t span = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane)) * kPixPerThread;   // first pixel of my warp
    int d[kPixPerThread], m[kPixPerThread], l[kPixPerThread];
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        const size_t idx = span + k * 32 + lane;
        const bool in = idx < total;
        d[k] = in ? (int)depth[idx] : 0;
        m[k] = in ? (int)mask[idx] : 0;
        l[k] = in ? (int)label[idx] : 0;
    }
    Point pt[kPixPerThread];
    int owner[kPixPerThread];
    uint32_t pos[kPixPerThread];
    uint32_t made = 0;
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        const size_t idx = span + k * 32 + lane;
        owner[k] = -1;
        pos[k] = 0u;
        if (idx < total && make_point_from(p, idx, d[k], m[k], l[k], sem, rgb, pose, pt[k]) && isfinite(pt[k].x) && isfinite(pt[k].y) &&
            isfinite(pt[k].z)) {
            const int i = (int)floorf(__fmul_rn(pt[k].x, p.inv_leaf));
            const int j = (int)floorf(__fmul_rn(pt[k].y, p.inv_leaf));
            const int kk = (int)floorf(__fmul_rn(pt[k].z, p.inv_leaf));
            owner[k] = voxel_owner(i, j, kk, nranks);
            ++made;
            if (owner[k] == rank) prefetch_voxel_line(tr.table, tr.mask, pt[k], p.inv_leaf);
        }
        // remote points: the warp's lanes with the same owner take consecutive places in the CTA's block for that owner
        uint32_t pending = __ballot_sync(0xffffffffu, owner[k] >= 0 && owner[k] != rank);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int o = __shfl_sync(0xffffffffu, owner[k], leader);
            const uint32_t group = __ballot_sync(0xffffffffu, owner[k] == o) & pending;
            uint32_t p0 = 0u;
            if (lane == leader) p0 = atomicAdd(&s_cnt[o], (uint32_t)__popc(group));
            p0 = __shfl_sync(0xffffffffu, p0, leader);
            if ((group >> lane) & 1u) pos[k] = p0 + __popc(group & ((1u << lane) - 1u));
            pending &= ~group;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nranks && (int)threadIdx.x != rank && s_cnt[threadIdx.x] != 0u)
        s_base[threadIdx.x] = atomicAdd_system(reinterpret_cast<uint32_t*>(peer_base[threadIdx.x]) + parity, s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
        if (owner[k] >= 0 && owner[k] != rank) {
            char* base = static_cast<char*>(peer_base[owner[k]]);
            const uint32_t slot = s_base[owner[k]] + pos[k];
            if (slot < inbox_cap) {
                // 16-byte inbox records (the label rides in the free top byte of 0x00RRGGBB): one aligned 128-bit store over NVLink per point
                uint4* dst = reinterpret_cast<uint4*>(base + kInboxHeader) + (size_t)parity * inbox_cap + slot;
                *dst = make_uint4(__float_as_uint(pt[k].x), __float_as_uint(pt[k].y), __float_as_uint(pt[k].z), pt[k].rgba | (pt[k].label << 24));
            } else {
                atomicOr_system(reinterpret_cast<uint32_t*>(base) + 2, 1u);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) fuse_point_warp(p, tr, pt[k], owner[k] == rank);
    made = __reduce_add_sync(0xffffffffu, made);
    if (lane == 0 && made) atomicAdd(&tr.counters[0], made);
}

// Step barrier of the peer-memory exchange without a collective: thread t tells rank t "my points of step `step` are in your
// inbox" (a store into t's header over NVLink, behind a system-scope fence; the point stores themselves were issued by the
// previous kernel on this stream) and then waits until rank t has said the same here.  One small CTA, no NCCL launch; a peer
// that never arrives (a rank that died or lost step) trips the time-out and raises a sticky flag instead of hanging the GPU.
__global__ void __launch_bounds__(64) k_flag_barrier(void* const* __restrict__ peer_base, int rank, int nranks, uint32_t step,
                                                     uint32_t* __restrict__ counters, long long timeout_cycles)
{
    const int t = threadIdx.x;
    if (t >= nranks) return;
    __threadfence_system();
    volatile uint32_t* theirs = reinterpret_cast<volatile uint32_t*>(static_cast<char*>(peer_base[t]) + kInboxFlags) + rank;
    *theirs = step;
    volatile uint32_t* mine = reinterpret_cast<volatile uint32_t*>(static_cast<char*>(peer_base[rank]) + kInboxFlags) + t;
    const long long t0 = clock64();
    while ((int32_t)(*mine - step) < 0) {
        if (clock64() - t0 > timeout_cycles) {
            atomicOr(&counters[2], 4u);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// COMPACT mode (ordered, row-major like the reference's push_back loop): count per block, scan, scatter
constexpr int kCompactBlock = 1024;
__global__ void __launch_bounds__(kCompactBlock) k_points_count(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ label,
                                                               const uint8_t* __restrict__ mask, uint32_t* __restrict__ blk_count,
                                                               size_t total, SSM_DP)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (idx < total) {
        const int d = depth[idx];
        const int l = label[idx];
        ok = d != 0 && !((double)d > p.max_depth_units) && mask[idx] != 255 &&
             !(l != SSM_LABEL_UNKNOWN && ((p.drop_mask >> l) & 1u));
    }
    const int c = __syncthreads_count(ok);
    if (threadIdx.x == 0) blk_count[blockIdx.x] = (uint32_t)c;
}
// single-CTA exclusive scan over the per-block counts; total goes to counters[0]
__global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t* __restrict__ blk, int n, uint32_t* __restrict__ counters)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n ? blk[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) >= o) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u;
        if (i < n) blk[i] = carry + before + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) counters[0] = carry;
}
__global__ void __launch_bounds__(kCompactBlock) k_points_scatter(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ label,
                                                                 const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sem,
                                                                 const uint8_t* __restrict__ rgb, const double* __restrict__ pose,
                                                                 const uint32_t* __restrict__ blk_off, Point* __restrict__ out,
                                                                 size_t total, SSM_DP)
{
    __shared__ uint32_t warp_off[32];
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Point pt;
    const bool ok = idx < total && make_point(p, idx, depth, label, mask, sem, rgb, pose, pt);
    const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_off[lane], s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        warp_off[lane] = s - w;
    }
    __syncthreads();
    if (ok) out[blk_off[blockIdx.x] + warp_off[warp] + __popc(ballot & ((1u << lane) - 1u))] = pt;
}

// fuse an explicit list of points (host-provided clouds, or points received from other ranks)
template <bool PACKED /* 16-byte inbox records instead of 20-byte points */>
__global__ void __launch_bounds__(256) k_fuse_list(const void* __restrict__ pts_, const uint32_t* __restrict__ count,
                                                   uint32_t max_count, TableRef tr, SSM_DP)
{
    const uint32_t n = count ? min(*count, max_count) : max_count;
    const uint32_t lane = threadIdx.x & 31;
    // a warp takes 32 * kPixPerThread consecutive points per trip (warp-uniform trip count): loads, then table-line prefetches, then inserts
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane)) * kPixPerThread; base < n;
         base += (uint64_t)gridDim.x * blockDim.x * kPixPerThread) {
        Point pt[kPixPerThread];
        bool ok[kPixPerThread];
#pragma unroll
        for (int k = 0; k < kPixPerThread; ++k) {
            const uint64_t i = base + k * 32 + lane;
            ok[k] = i < n;
            pt[k] = Point{};
            if (ok[k]) {
                if constexpr (PACKED) {
                    const uint4 q = static_cast<const uint4*>(pts_)[i];
                    pt[k].x = __uint_as_float(q.x); pt[k].y = __uint_as_float(q.y); pt[k].z = __uint_as_float(q.z);
                    pt[k].rgba = q.w & 0x00ffffffu; pt[k].label = q.w >> 24;
                } else {
                    pt[k] = static_cast<const Point*>(pts_)[i];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kPixPerThread; ++k)
            if (ok[k]) prefetch_voxel_line(tr.table, tr.mask, pt[k], p.inv_leaf);
#pragma unroll
        for (int k = 0; k < kPixPerThread; ++k) fuse_point_warp(p, tr, pt[k], ok[k]);
    }
}

// Cached keyframe clouds (mapper.cpp:17-20 keeps frame->pointcloud in camera coordinates; :90-91 re-transforms it by
// the frame's current pose on every redraw): pcl::transformPointCloud in double, rounded to float, then the hash insert.
struct Pose12 { double m[12]; };
__global__ void __launch_bounds__(256) k_transform_fuse(const Point* __restrict__ pts, uint32_t n, Pose12 T, TableRef tr, SSM_DP)
{
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x - lane); base < n; base += gridDim.x * blockDim.x) {   // warp-uniform trips
        const uint32_t i = base + lane;
        Point pt = {};
        if (i < n) pt = pts[i];
        float w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double acc = __dmul_rn(T.m[4 * k], (double)pt.x);
            acc = __dadd_rn(acc, __dmul_rn(T.m[4 * k + 1], (double)pt.y));
            acc = __dadd_rn(acc, __dmul_rn(T.m[4 * k + 2], (double)pt.z));
            acc = __dadd_rn(acc, T.m[4 * k + 3]);
            w[k] = (float)acc;
        }
        pt.x = w[0]; pt.y = w[1]; pt.z = w[2];
        fuse_point_warp(p, tr, pt, i < n);
    }
}

__global__ void __launch_bounds__(256) k_table_clear(Voxel* __restrict__ table, size_t words16)
{
    uint4* t = reinterpret_cast<uint4*>(table);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words16; i += (size_t)gridDim.x * blockDim.x) {
        // key = all ones, everything else zero; the key is the first 8 bytes of each 128-byte record
        t[i] = (i & 7) == 0 ? make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// export: compact occupied records (unordered; the host sorts when PCL order is requested)
__global__ void __launch_bounds__(256) k_export(const Voxel* __restrict__ table, uint64_t slots, Voxel* __restrict__ out,
                                                uint32_t* __restrict__ counters, uint32_t max_out)
{
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (uint64_t)gridDim.x * blockDim.x) {
        if (table[s].key == kEmptyKey) continue;
        const uint32_t o = atomicAdd(&counters[3], 1u);
        if (o < max_out) out[o] = table[s];
    }
}

// ------------------------------------------------------------------------------------------------
int launch_depth(ssm_ctx* c, int B, const int16_t* d_disp, uint16_t* d_depth, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t per_frame = (size_t)p.W * p.H, total = per_frame * B;
    SSM_CUDA(cudaMemsetAsync(c->d_min_disp, 0x7f, sizeof(int) * B, s));
    k_min_disp<<<dim3(64, B), 256, 0, s>>>(d_disp, c->d_min_disp, per_frame, B);
    SSM_LAUNCH_CHECK(c);
    const int aligned8 = ((reinterpret_cast<uintptr_t>(d_disp) | reinterpret_cast<uintptr_t>(d_depth)) & 7) == 0;
    k_depth<<<(unsigned)((total / 4 + 256) / 256), 256, 0, s>>>(d_disp, c->d_min_disp, d_depth, total, aligned8, p);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_labels_mask(ssm_ctx* c, int B, const uint8_t* d_sem, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t total = (size_t)p.W * p.H * B;
    const unsigned grid = (unsigned)((total / 4 + 256) / 256);
    // the speckle filter's size array is free once the post stage is done: it holds the one-bit-per-pixel predicate and
    // its row-dilated copy
    const int wpr = (p.W + 31) >> 5;
    const size_t nwords = (size_t)wpr * p.H * B;
    uint32_t* dynbits = reinterpret_cast<uint32_t*>(c->d_cc_size);
    uint32_t* vbits = dynbits + nwords;
    SSM_CUDA(cudaMemsetAsync(dynbits, 0, nwords * sizeof(uint32_t), s));
    const int a_sem = ((reinterpret_cast<uintptr_t>(d_sem) | reinterpret_cast<uintptr_t>(c->d_label)) & 3) == 0;
    k_labels<<<grid, 256, 0, s>>>(d_sem, c->d_label, dynbits, c->d_label_lut, total, a_sem, p);
    SSM_LAUNCH_CHECK(c);
    k_dilate_rows<<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(dynbits, vbits, wpr, p.H, p.dilate_radius, nwords);
    SSM_LAUNCH_CHECK(c);
    const int a_mask = (reinterpret_cast<uintptr_t>(c->d_mask) & 3) == 0;
    k_moving_mask<<<grid, 256, 0, s>>>(vbits, c->d_mask, total, a_mask, p);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_points(ssm_ctx* c, int B, const uint16_t* d_depth, const uint8_t* d_sem, const uint8_t* d_rgb,
                  const double* d_pose, bool fuse_into_map, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t total = (size_t)p.W * p.H * B;
    SSM_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(uint32_t), s));   // counters[0] = points of this call
    if (fuse_into_map) {
        k_points_fuse<<<(unsigned)((total + 256 * kPixPerThread - 1) / (256 * kPixPerThread)), 256, 0, s>>>(d_depth, c->d_label, c->d_mask, d_sem, d_rgb, d_pose,
                                                                      table_ref(c), total, p);
        SSM_LAUNCH_CHECK(c);
        return SSM_OK;
    }
    const int nblk = (int)((total + kCompactBlock - 1) / kCompactBlock);
    k_points_count<<<nblk, kCompactBlock, 0, s>>>(d_depth, c->d_label, c->d_mask, c->d_blk_count, total, p);
    SSM_LAUNCH_CHECK(c);
    k_scan_blocks<<<1, 1024, 0, s>>>(c->d_blk_count, nblk, c->d_counters);
    SSM_LAUNCH_CHECK(c);
    k_points_scatter<<<nblk, kCompactBlock, 0, s>>>(d_depth, c->d_label, c->d_mask, d_sem, d_rgb, d_pose, c->d_blk_count,
                                                    c->d_points, total, p);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_points_p2p(ssm_ctx* c, int B, const uint16_t* d_depth, const uint8_t* d_sem, const uint8_t* d_rgb, const double* d_pose,
                      void* const* d_peer_base, int parity, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t total = (size_t)p.W * p.H * B;
    SSM_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(uint32_t), s));
    k_points_p2p<<<(unsigned)((total + 256 * kPixPerThread - 1) / (256 * kPixPerThread)), 256, 0, s>>>(d_depth, c->d_label, c->d_mask, d_sem, d_rgb, d_pose, table_ref(c),
                                                                 d_peer_base, c->rank, c->nranks, parity,
                                                                 (uint32_t)c->inbox_cap, total, p);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_flag_barrier(ssm_ctx* c, uint32_t step, cudaStream_t s)
{
    k_flag_barrier<<<1, 64, 0, s>>>(c->d_peer_base, c->rank, c->nranks, step, c->d_counters, 20000000000ll /* ~10 s */);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_fuse_inbox(ssm_ctx* c, int parity, cudaStream_t s)
{
    char* base = static_cast<char*>(c->ipc_base);
    const uint4* pts = reinterpret_cast<const uint4*>(base + kInboxHeader) + (size_t)parity * c->inbox_cap;
    uint32_t* count = reinterpret_cast<uint32_t*>(base) + parity;
    k_fuse_list<true><<<c->sm_count * 16, 256, 0, s>>>(pts, count, (uint32_t)c->inbox_cap, table_ref(c), c->dp);
    SSM_LAUNCH_CHECK(c);
    SSM_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t), s));   // ready for the step after next
    return SSM_OK;
}

int launch_fuse_points(ssm_ctx* c, const Point* d_pts, const uint32_t* d_count, uint32_t max_count, cudaStream_t s)
{
    if (max_count == 0) return SSM_OK;
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)max_count + 255) / 256, (size_t)c->sm_count * 16);
    k_fuse_list<false><<<grid, 256, 0, s>>>(d_pts, d_count, max_count, table_ref(c), c->dp);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_transform_fuse(ssm_ctx* c, const Point* d_pts, uint32_t n, const double* T16, cudaStream_t s)
{
    if (n == 0) return SSM_OK;
    Pose12 T;
    for (int i = 0; i < 12; ++i) T.m[i] = T16[i];
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)n + 255) / 256, (size_t)c->sm_count * 16);
    k_transform_fuse<<<grid, 256, 0, s>>>(d_pts, n, T, table_ref(c), c->dp);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_map_clear(ssm_ctx* c, cudaStream_t s)
{
    const size_t words16 = c->table_slots * (sizeof(Voxel) / 16);
    k_table_clear<<<c->sm_count * 8, 256, 0, s>>>(c->d_table, words16);
    SSM_LAUNCH_CHECK(c);
    SSM_CUDA(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(uint32_t), s));   // points, voxels, flags, export count, parked points
    return SSM_OK;
}

int launch_export(ssm_ctx* c, Voxel* d_out, uint32_t max_out, cudaStream_t s)
{
    SSM_CUDA(cudaMemsetAsync(c->d_counters + 3, 0, sizeof(uint32_t), s));
    k_export<<<c->sm_count * 8, 256, 0, s>>>(c->d_table, c->table_slots, d_out, c->d_counters, max_out);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

}  // namespace ssm
