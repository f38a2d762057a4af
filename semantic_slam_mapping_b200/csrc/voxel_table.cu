// voxel_table.cu -- the voxel hash as a container: K9 voxel_finalize + PCL-ordered export, growth, statistics.  sm_100a.
//
// Replaces, for the drop-in path, the tail of pcl::VoxelGrid<PointXYZRGBA>::applyFilter as Mapper::viewer uses it
// (/root/reference src/mapper.cpp:154-159: voxel.filter(*tmp); globalMap->swap(*tmp); viewer.showCloud(globalMap)):
//   * per voxel: centroid = sums / n, colour = truncated mean (int(r) << 16 | int(g) << 8 | int(b), alpha 0), SURVEY App. B-2;
//   * the build's addition: majority label of the vote histogram (ties -> lowest id; no votes -> 255), SURVEY App. C-8;
//   * output order = ascending idx = i + j * dx + k * dx * dy  <=>  lexicographic (k, j, i): a device LSD radix sort of
//     (min-subtracted key, slot) pairs over exactly as many 8-bit digits as the map's extent needs.
// The pipeline runs over any array of Voxel records with empty slots marked by kEmptyKey: the hash table itself, or the
// dense record list a rank has gathered from its peers (ssm_map_export_gathered).
//
// Growth: the reference's map is a std::vector that grows without bound, and pcl::VoxelGrid refuses grids of more than
// 2^31 cells (App. B-2) -- the 0.02 m stress config of BASELINE.json is the one PCL cannot run.  Here an insert that does not
// find a free slot within kMaxProbe probes parks its point in a spill list instead of failing; the next pipeline call (or
// any blocking map call) doubles the table -- record-level re-insertion ordered on the stream, no host stall -- and drains
// the list.  Nothing is lost unless the spill list itself overflows (reported as SSM_ERR_CAPACITY).
#include <algorithm>
#include <vector>

#include "ssm_internal.cuh"

namespace ssm {

// ------------------------------------------------------------------------------------------------
// export step 1: index the occupied records (slot list + packed keys) and the map's extent
// ------------------------------------------------------------------------------------------------
// ext[0..2] = min i, j, k (as biased 21-bit fields), ext[3..5] = max.  Warp-aggregated: one atomic per warp and word.
__global__ void __launch_bounds__(256) k_export_index(const Voxel* __restrict__ recs, uint64_t slots, uint32_t* __restrict__ slot_out,
                                                      unsigned long long* __restrict__ key_out, uint32_t* __restrict__ counter,
                                                      uint32_t* __restrict__ ext, uint32_t max_out)
{
    const int lane = threadIdx.x & 31;
    uint32_t lo[3] = {0x1fffffu, 0x1fffffu, 0x1fffffu}, hi[3] = {0u, 0u, 0u};
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < slots; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = base + lane;
        const unsigned long long key = s < slots ? recs[s].key : kEmptyKey;
        const bool occ = key != kEmptyKey;
        const uint32_t m = __ballot_sync(0xffffffffu, occ);
        if (!m) continue;
        uint32_t o = 0;
        if (lane == 0) o = atomicAdd(counter, (uint32_t)__popc(m));
        o = __shfl_sync(0xffffffffu, o, 0) + __popc(m & ((1u << lane) - 1u));
        if (occ) {
            if (o < max_out) { slot_out[o] = (uint32_t)s; key_out[o] = key; }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const uint32_t f = (uint32_t)(key >> (21 * a)) & 0x1fffffu;
                lo[a] = min(lo[a], f); hi[a] = max(hi[a], f);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const uint32_t l = __reduce_min_sync(0xffffffffu, lo[a]), h = __reduce_max_sync(0xffffffffu, hi[a]);
        if (lane == 0 && l <= h) { atomicMin(&ext[a], l); atomicMax(&ext[3 + a], h); }
    }
}

// sort key = (i - i_min) | (j - j_min) << bi | (k - k_min) << (bi + bj): ascending == pcl::VoxelGrid's idx order
__global__ void __launch_bounds__(256) k_export_sortkeys(unsigned long long* __restrict__ keys, uint32_t n, uint32_t i0, uint32_t j0, uint32_t k0,
                                                         int bi, int bj)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned long long key = keys[t];
    const unsigned long long i = (key & 0x1fffffu) - i0, j = ((key >> 21) & 0x1fffffu) - j0, k = ((key >> 42) & 0x1fffffu) - k0;
    keys[t] = i | (j << bi) | (k << (bi + bj));
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass, stable.  Per pass: tile histograms -> exclusive scan over the
// bin-major histogram array -> scatter with in-tile ranks from warp match groups.
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256, kSortItems = 16, kSortTile = kSortThreads * kSortItems, kSortWarps = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift,
                                                            uint32_t* __restrict__ hist, uint32_t ntiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll 4
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan in three steps (chunks of 4096 words per CTA; the chunk totals are scanned by one CTA)
constexpr int kScanChunk = 4096;
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* warp_sums, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], ws = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        warp_sums[lane] = ws - w;
        if (lane == 31) *total = ws;
    }
    __syncthreads();
    const uint32_t r = warp_sums[warp] + s - v;
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(1024) k_scan_chunks(uint32_t* __restrict__ a, size_t n, uint32_t* __restrict__ chunk_total)
{
    __shared__ uint32_t ws[32], tot;
    const size_t i0 = (size_t)blockIdx.x * kScanChunk + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k < n ? a[i0 + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t ex = block_exclusive_scan_1024(mine, ws, &tot);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < n) a[i0 + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) chunk_total[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_scan_totals(uint32_t* __restrict__ t, uint32_t n)
{
    __shared__ uint32_t ws[32], tot, carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? t[i] : 0u;
        const uint32_t ex = block_exclusive_scan_1024(v, ws, &tot);
        if (i < n) t[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024) k_scan_add(uint32_t* __restrict__ a, size_t n, const uint32_t* __restrict__ chunk_off)
{
    const uint32_t off = chunk_off[blockIdx.x];
    const size_t i0 = (size_t)blockIdx.x * kScanChunk + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < n) a[i0 + k] += off;
}

// warp w of a tile owns keys [w * 512, (w + 1) * 512) of it, 16 trips of 32 consecutive keys: ranks within the tile are
// (earlier warps) + (earlier trips of this warp) + (lower lanes with the same digit in this trip) -- stable
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const unsigned long long* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                               unsigned long long* __restrict__ kout, uint32_t* __restrict__ vout, uint32_t n,
                                                               int shift, const uint32_t* __restrict__ hist, uint32_t ntiles)
{
    __shared__ uint32_t cnt[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile + warp * (kSortTile / kSortWarps);
    unsigned long long key[kSortItems];
    uint32_t val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        key[r] = i < n ? kin[i] : 0ull;
        val[r] = i < n ? vin[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        const bool in = i < n;
        const uint32_t d = in ? (uint32_t)(key[r] >> shift) & 255u : 256u + lane;   // out-of-range lanes form groups of their own
        const uint32_t grp = __match_any_sync(0xffffffffu, d);
        const uint32_t before = __popc(grp & ((1u << lane) - 1u));
        uint32_t c0 = 0u;
        if (in) c0 = cnt[warp][d];
        __syncwarp();
        if (in && before == 0u) cnt[warp][d] = c0 + __popc(grp);
        __syncwarp();
        rank[r] = c0 + before;
    }
    __syncthreads();
    {   // per digit: global offset of this tile, then the warps' exclusive offsets
        const uint32_t b = threadIdx.x;
        uint32_t acc = hist[(size_t)b * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t t = cnt[w][b];
            cnt[w][b] = acc;
            acc += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t pos = cnt[warp][(uint32_t)(key[r] >> shift) & 255u] + rank[r];
            kout[pos] = key[r];
            vout[pos] = val[r];
        }
    }
}

// sorts (keys, vals) over the low `bits` bits; the result is in (keys, vals) again.  tmp buffers: same sizes; hist: 256 * ntiles
// + chunk totals.
static int radix_sort_pairs(ssm_ctx* c, unsigned long long* keys, uint32_t* vals, unsigned long long* keys_tmp, uint32_t* vals_tmp,
                            uint32_t* hist, uint32_t n, int bits, cudaStream_t s)
{
    if (n < 2 || bits <= 0) return SSM_OK;
    const uint32_t ntiles = (n + kSortTile - 1) / kSortTile;
    const size_t hn = (size_t)256 * ntiles;
    const uint32_t nchunks = (uint32_t)((hn + kScanChunk - 1) / kScanChunk);
    uint32_t* totals = hist + hn;
    unsigned long long *ka = keys, *kb = keys_tmp;
    uint32_t *va = vals, *vb = vals_tmp;
    int passes = 0;
    for (int shift = 0; shift < bits; shift += 8, ++passes) {
        k_sort_hist<<<ntiles, kSortThreads, 0, s>>>(ka, n, shift, hist, ntiles);
        SSM_LAUNCH_CHECK(c);
        k_scan_chunks<<<nchunks, 1024, 0, s>>>(hist, hn, totals);
        SSM_LAUNCH_CHECK(c);
        if (nchunks > 1) {
            k_scan_totals<<<1, 1024, 0, s>>>(totals, nchunks);
            SSM_LAUNCH_CHECK(c);
            k_scan_add<<<nchunks, 1024, 0, s>>>(hist, hn, totals);
            SSM_LAUNCH_CHECK(c);
        }
        k_sort_scatter<<<ntiles, kSortThreads, 0, s>>>(ka, va, kb, vb, n, shift, hist, ntiles);
        SSM_LAUNCH_CHECK(c);
        std::swap(ka, kb);
        std::swap(va, vb);
    }
    if (passes & 1) {   // odd number of passes: the result sits in the tmp pair
        SSM_CUDA(cudaMemcpyAsync(keys, keys_tmp, sizeof(unsigned long long) * n, cudaMemcpyDeviceToDevice, s));
        SSM_CUDA(cudaMemcpyAsync(vals, vals_tmp, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s));
    }
    return SSM_OK;
}

// ------------------------------------------------------------------------------------------------
// K9 voxel_finalize: eight lanes read one 128-byte record (one 16-byte word each, coalesced), exchange what the outputs need
// with shuffles and write the compact arrays.  Arithmetic mirrors pcl::VoxelGrid: centroid = sum / n with
// the sums held as 2^-24 m fixed point in 64-bit integers (exact, order independent), colour = int(float(sum) / float(n)).
// ------------------------------------------------------------------------------------------------
struct ExportPtrs {
    int32_t* ijk; float* xyz; uint32_t* rgba; uint8_t* label; uint32_t* count; uint32_t* votes;
};
__global__ void __launch_bounds__(256) k_voxel_finalize(const Voxel* __restrict__ recs, const uint32_t* __restrict__ slots, uint32_t n,
                                                        ExportPtrs out, int num_labels)
{
    const int q = threadIdx.x & 7;
    const uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool in = v < n;                    // uniform over the eight lanes of a record; whole groups may be idle in the last warp
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (in) w = reinterpret_cast<const uint4*>(recs + slots[v])[q];
    const uint32_t full = 0xffffffffu;
    const int g0 = (threadIdx.x & 31) & ~7;   // first lane of my group
    // q0 = {key, sx}, q1 = {sy, sz}, q2 = {n, sr, sg, sb}, q3..q7 = votes[4 (q - 3) ..]
    const uint32_t key_lo = __shfl_sync(full, w.x, g0), key_hi = __shfl_sync(full, w.y, g0);
    const uint32_t sx_lo = __shfl_sync(full, w.z, g0), sx_hi = __shfl_sync(full, w.w, g0);
    const uint32_t sy_lo = __shfl_sync(full, w.x, g0 + 1), sy_hi = __shfl_sync(full, w.y, g0 + 1);
    const uint32_t sz_lo = __shfl_sync(full, w.z, g0 + 1), sz_hi = __shfl_sync(full, w.w, g0 + 1);
    const uint32_t cn = __shfl_sync(full, w.x, g0 + 2), sr = __shfl_sync(full, w.y, g0 + 2), sg = __shfl_sync(full, w.z, g0 + 2),
                   sb = __shfl_sync(full, w.w, g0 + 2);
    // majority label: maximum of (votes << 8 | 255 - id) over the record's 20 bins
    unsigned long long best = 0ull;
    if (q >= 3) {
        const uint32_t vv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int id = 4 * (q - 3) + t;
            if (id < num_labels && vv[t] > 0u) best = max(best, ((unsigned long long)vv[t] << 8) | (unsigned long long)(255 - id));
        }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) best = max(best, __shfl_xor_sync(full, best, o));
    if (!in) return;
    if (q >= 3 && out.votes) {
        const uint32_t vv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int id = 4 * (q - 3) + t;
            if (id < num_labels) out.votes[(size_t)v * num_labels + id] = vv[t];
        }
    }
    if (q < 3) {
        const unsigned long long key = ((unsigned long long)key_hi << 32) | key_lo;
        const uint32_t field = (uint32_t)(key >> (21 * q)) & 0x1fffffu;
        if (out.ijk) out.ijk[(size_t)v * 3 + q] = (int)field - (1 << 20);
        if (out.xyz) {
            const uint32_t lo = q == 0 ? sx_lo : (q == 1 ? sy_lo : sz_lo), hi = q == 0 ? sx_hi : (q == 1 ? sy_hi : sz_hi);
            const long long sum = (long long)(((unsigned long long)hi << 32) | lo);
            out.xyz[(size_t)v * 3 + q] = (float)__ddiv_rn(__ddiv_rn(__ll2double_rn(sum), kFixScale), (double)cn);
        }
    } else if (q == 3) {
        if (out.rgba) {
            const float fn = (float)cn;
            const float r = __fdiv_rn((float)sr, fn), g = __fdiv_rn((float)sg, fn), b = __fdiv_rn((float)sb, fn);
            out.rgba[v] = ((uint32_t)(int)r << 16) | ((uint32_t)(int)g << 8) | (uint32_t)(int)b;
        }
    } else if (q == 4) {
        if (out.label) out.label[v] = best ? (uint8_t)(255u - (uint32_t)(best & 255ull)) : (uint8_t)SSM_LABEL_UNKNOWN;
    } else if (q == 5) {
        if (out.count) out.count[v] = cn;
    }
}

// ------------------------------------------------------------------------------------------------
// growth: every record of the old table is re-inserted into the (cleared) new one.  Eight lanes move one record; the
// group's first lane claims the slot.  Keys are unique, so a claimed slot is written by its claimer alone.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rehash(const Voxel* __restrict__ old_t, uint64_t old_slots, Voxel* __restrict__ new_t, uint64_t new_mask,
                                                uint32_t* __restrict__ counters)
{
    const int q = threadIdx.x & 7;
    const int g0 = (threadIdx.x & 31) & ~7;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) >> 3; base < old_slots;
         base += ((uint64_t)gridDim.x * blockDim.x) >> 3) {
        const uint64_t s = base + ((threadIdx.x & 31) >> 3);
        uint4 w = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
        if (s < old_slots) w = reinterpret_cast<const uint4*>(old_t + s)[q];
        const unsigned long long key = ((unsigned long long)__shfl_sync(0xffffffffu, w.y, g0) << 32) | __shfl_sync(0xffffffffu, w.x, g0);
        uint64_t slot = ~0ull;
        if (q == 0 && key != kEmptyKey) {
            slot = mix64(key) & new_mask;
            for (uint64_t probe = 0; probe <= new_mask; ++probe, slot = (slot + 1) & new_mask) {
                if (atomicCAS(&new_t[slot].key, kEmptyKey, key) == kEmptyKey) break;
            }
        }
        slot = __shfl_sync(0xffffffffu, slot, g0);
        if (slot != ~0ull && q != 0) reinterpret_cast<uint4*>(new_t + slot)[q] = w;
        if (slot != ~0ull && q == 0) new_t[slot].sx = ((unsigned long long)w.w << 32) | w.z;   // the key half of word 0 is in place
    }
}

// probe statistics: stats[0] = occupied, stats[1] = sum of displacements, stats[2] = max displacement
__global__ void __launch_bounds__(256) k_table_stats(const Voxel* __restrict__ table, uint64_t slots, unsigned long long* __restrict__ stats)
{
    const uint64_t mask = slots - 1;
    unsigned long long occ = 0, sum = 0, mx = 0;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = table[s].key;
        if (key == kEmptyKey) continue;
        const unsigned long long d = (s - (mix64(key) & mask)) & mask;
        ++occ; sum += d; mx = max(mx, d);
    }
    for (int o = 16; o; o >>= 1) {
        occ += __shfl_xor_sync(0xffffffffu, occ, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && occ) {
        atomicAdd(&stats[0], occ);
        atomicAdd(&stats[1], sum);
        atomicMax(&stats[2], mx);
    }
}

__global__ void __launch_bounds__(256) k_table_clear2(Voxel* __restrict__ table, size_t words16)
{
    uint4* t = reinterpret_cast<uint4*>(table);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words16; i += (size_t)gridDim.x * blockDim.x)
        t[i] = (i & 7) == 0 ? make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
}

// Old tables wait here until the kernels that moved their records are done (cudaFree synchronises the whole device, so it
// is only called for tables whose move has completed, or from blocking calls).
int table_reap(ssm_ctx* c, bool wait)
{
    for (int i = 0; i < ssm_ctx::kRetired; ++i) {
        if (!c->retired[i]) continue;
        if (!wait && cudaEventQuery(c->retired_ev[i]) != cudaSuccess) { cudaGetLastError(); continue; }
        if (wait) SSM_CUDA(cudaEventSynchronize(c->retired_ev[i]));
        SSM_CUDA(cudaFree(c->retired[i]));
        c->retired[i] = nullptr;
    }
    return SSM_OK;
}

// Doubles the table until it has at least `min_slots` slots, stream-ordered on `s`: allocate (plain cudaMalloc: 6 ms for
// 68 GB on a B200, against 1.8 s from the stream-ordered allocator -- scripts/micro/alloc_time.cu), clear, move the records,
// retire the old allocation, then re-insert the parked points.  Every stream that touches the table must already be ordered
// before `s` (the pipeline joins its sub-batch streams into `s`; callers wait for the route stream).
int table_grow(ssm_ctx* c, uint64_t min_slots, cudaStream_t s)
{
    uint64_t slots = c->table_slots;
    while (slots < min_slots) slots <<= 1;
    if (slots == c->table_slots) slots <<= 1;
    if (slots > (1ull << 32)) {
        set_error("voxel hash cannot grow beyond 2^32 slots");
        return SSM_ERR_CAPACITY;
    }
    int rc = table_reap(c, false);
    if (rc) return rc;
    int free_slot = -1;
    for (int i = 0; i < ssm_ctx::kRetired; ++i)
        if (!c->retired[i]) free_slot = i;
    if (free_slot < 0) {   // every slot still busy (three growth steps within two batches): wait for them
        if ((rc = table_reap(c, true))) return rc;
        free_slot = 0;
    }
    Voxel* nt = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&nt), sizeof(Voxel) * slots);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("voxel hash is full and the device has no room for a table of twice the size");
        return SSM_ERR_CAPACITY;
    }
    k_table_clear2<<<c->sm_count * 8, 256, 0, s>>>(nt, slots * (sizeof(Voxel) / 16));
    SSM_LAUNCH_CHECK(c);
    k_rehash<<<c->sm_count * 8, 256, 0, s>>>(c->d_table, c->table_slots, nt, slots - 1, c->d_counters);
    SSM_LAUNCH_CHECK(c);
    if (!c->retired_ev[free_slot]) SSM_CUDA(cudaEventCreateWithFlags(&c->retired_ev[free_slot], cudaEventDisableTiming));
    SSM_CUDA(cudaEventRecord(c->retired_ev[free_slot], s));
    c->retired[free_slot] = c->d_table;
    c->d_table = nt;
    c->table_slots = slots;
    c->grows++;
    return spill_drain(c, s);
}

// re-insert the parked points (after a growth step) and flip to the other spill buffer
int spill_drain(ssm_ctx* c, cudaStream_t s)
{
    if (!c->d_spill[0]) return SSM_OK;
    const int cur = c->spill_cur;
    // counters[4] = points parked in the current buffer: move the count aside, zero it, fuse from the old buffer while new
    // spills (none are expected right after a growth step) go to the other one
    SSM_CUDA(cudaMemcpyAsync(c->d_counters + 5, c->d_counters + 4, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    SSM_CUDA(cudaMemsetAsync(c->d_counters + 4, 0, sizeof(uint32_t), s));
    c->spill_cur = cur ^ 1;
    return launch_fuse_points(c, c->d_spill[cur], c->d_counters + 5, (uint32_t)c->spill_cap, s);
}

int table_stats(ssm_ctx* c, cudaStream_t s, unsigned long long out[3])
{
    unsigned long long* d = nullptr;
    SSM_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), 3 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), s);
    if (e == cudaSuccess) {
        k_table_stats<<<c->sm_count * 8, 256, 0, s>>>(c->d_table, c->table_slots, d);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "table statistics");
    return SSM_OK;
}

// ------------------------------------------------------------------------------------------------
// the export pipeline over a record array: index -> (sort) -> finalize -> D2H into the caller's arrays
// ------------------------------------------------------------------------------------------------
static int bits_for(uint32_t range) { int b = 0; while (b < 32 && (range >> b)) ++b; return b; }

int export_records_device(ssm_ctx* c, const Voxel* d_recs, uint64_t slots, uint64_t n_expected, bool sorted, const ssm_voxel_export* out,
                          uint64_t max_voxels, uint64_t* n_out, cudaStream_t s, float* ms_device)
{
    *n_out = 0;
    if (n_expected == 0) return SSM_OK;
    if (n_expected > 0xfffffff0ull) {
        set_error("more than 2^32 voxels in one export");
        return SSM_ERR_CAPACITY;
    }
    const uint32_t n = (uint32_t)n_expected;
    const int L = c->p.num_labels;
    const uint32_t ntiles = (n + kSortTile - 1) / kSortTile;
    const size_t hist_words = (size_t)256 * ntiles + ((size_t)256 * ntiles + kScanChunk - 1) / kScanChunk + 16;
    // one workspace allocation: slot list, keys, their sort doubles, histograms, 8 header words, then the output arrays
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_hdr = take(64), o_keys = take(8ull * n), o_keys2 = take(sorted ? 8ull * n : 0), o_slots = take(4ull * n),
                 o_slots2 = take(sorted ? 4ull * n : 0), o_hist = take(sorted ? 4 * hist_words : 0);
    const uint64_t m = std::min<uint64_t>(n, max_voxels);
    const size_t o_ijk = take(out->ijk ? 12ull * n : 0), o_xyz = take(out->xyz ? 12ull * n : 0), o_rgba = take(out->rgba ? 4ull * n : 0),
                 o_label = take(out->label ? n : 0), o_count = take(out->count ? 4ull * n : 0), o_votes = take(out->votes ? 4ull * L * n : 0);
    if (off > c->export_ws_bytes) {   // grow-only workspace, kept by the context (an export per map update must not pay cudaMalloc)
        if (c->export_ws) SSM_CUDA(cudaFree(c->export_ws));
        c->export_ws = nullptr;
        c->export_ws_bytes = 0;
        const size_t want = off + off / 4;
        SSM_CUDA(cudaMalloc(&c->export_ws, want));
        c->export_ws_bytes = want;
    }
    char* ws = static_cast<char*>(c->export_ws);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = SSM_OK;
    auto body = [&]() -> int {
        uint32_t* hdr = reinterpret_cast<uint32_t*>(ws + o_hdr);   // [0] count, [1..3] min fields, [4..6] max fields
        auto* keys = reinterpret_cast<unsigned long long*>(ws + o_keys);
        auto* slot_list = reinterpret_cast<uint32_t*>(ws + o_slots);
        if (ms_device) {
            SSM_CUDA(cudaEventCreate(&e0));
            SSM_CUDA(cudaEventCreate(&e1));
            SSM_CUDA(cudaEventRecord(e0, s));
        }
        static const uint32_t init[8] = {0u, 0x1fffffu, 0x1fffffu, 0x1fffffu, 0u, 0u, 0u, 0u};
        SSM_CUDA(cudaMemcpyAsync(hdr, init, sizeof(init), cudaMemcpyHostToDevice, s));
        k_export_index<<<c->sm_count * 8, 256, 0, s>>>(d_recs, slots, slot_list, keys, hdr, hdr + 1, n);
        SSM_LAUNCH_CHECK(c);
        uint32_t h[8];
        SSM_CUDA(cudaMemcpyAsync(h, hdr, sizeof(h), cudaMemcpyDeviceToHost, s));
        SSM_CUDA(cudaStreamSynchronize(s));
        if (h[0] != n) {
            set_error("voxel table changed during the export (calls on a context must be serialised)");
            return SSM_ERR_INVALID_ARGUMENT;
        }
        if (sorted) {
            const int bi = bits_for(h[4] - h[1]), bj = bits_for(h[5] - h[2]), bk = bits_for(h[6] - h[3]);
            k_export_sortkeys<<<(n + 255) / 256, 256, 0, s>>>(keys, n, h[1], h[2], h[3], bi, bj);
            SSM_LAUNCH_CHECK(c);
            int r = radix_sort_pairs(c, keys, slot_list, reinterpret_cast<unsigned long long*>(ws + o_keys2),
                                     reinterpret_cast<uint32_t*>(ws + o_slots2), reinterpret_cast<uint32_t*>(ws + o_hist), n, bi + bj + bk, s);
            if (r) return r;
        }
        ExportPtrs ep;
        ep.ijk = out->ijk ? reinterpret_cast<int32_t*>(ws + o_ijk) : nullptr;
        ep.xyz = out->xyz ? reinterpret_cast<float*>(ws + o_xyz) : nullptr;
        ep.rgba = out->rgba ? reinterpret_cast<uint32_t*>(ws + o_rgba) : nullptr;
        ep.label = out->label ? reinterpret_cast<uint8_t*>(ws + o_label) : nullptr;
        ep.count = out->count ? reinterpret_cast<uint32_t*>(ws + o_count) : nullptr;
        ep.votes = out->votes ? reinterpret_cast<uint32_t*>(ws + o_votes) : nullptr;
        k_voxel_finalize<<<(unsigned)(((size_t)n * 8 + 255) / 256), 256, 0, s>>>(d_recs, slot_list, n, ep, L);
        SSM_LAUNCH_CHECK(c);
        if (ms_device) SSM_CUDA(cudaEventRecord(e1, s));
        if (out->ijk) SSM_CUDA(cudaMemcpyAsync(out->ijk, ep.ijk, 12ull * m, cudaMemcpyDeviceToHost, s));
        if (out->xyz) SSM_CUDA(cudaMemcpyAsync(out->xyz, ep.xyz, 12ull * m, cudaMemcpyDeviceToHost, s));
        if (out->rgba) SSM_CUDA(cudaMemcpyAsync(out->rgba, ep.rgba, 4ull * m, cudaMemcpyDeviceToHost, s));
        if (out->label) SSM_CUDA(cudaMemcpyAsync(out->label, ep.label, m, cudaMemcpyDeviceToHost, s));
        if (out->count) SSM_CUDA(cudaMemcpyAsync(out->count, ep.count, 4ull * m, cudaMemcpyDeviceToHost, s));
        if (out->votes) SSM_CUDA(cudaMemcpyAsync(out->votes, ep.votes, 4ull * L * m, cudaMemcpyDeviceToHost, s));
        SSM_CUDA(cudaStreamSynchronize(s));
        if (ms_device) SSM_CUDA(cudaEventElapsedTime(ms_device, e0, e1));
        return SSM_OK;
    };
    rc = body();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc == SSM_OK) *n_out = n;
    return rc;
}

}  // namespace ssm
