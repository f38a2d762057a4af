// ssm_internal.cuh -- shared declarations of libssm.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ssm.h"

namespace ssm {

// ---------------------------------------------------------------------------------------------
// constants of the SGBM chain
// ---------------------------------------------------------------------------------------------
constexpr int kDispScale = 16;
constexpr int kInvalidDisp = -16;           // (minDisparity - 1) * 16 with minDisparity == 0
constexpr int kMaxCost = 32767;
constexpr uint32_t kBig = 0xC000u;          // "+inf" for 16-bit path costs held in unsigned lanes
constexpr uint32_t kBigW = 0xC000C000u;
constexpr uint32_t kSatW = 0x7fff7fffu;     // saturation bound of S, both lanes

// ---------------------------------------------------------------------------------------------
// voxel record: one 128-byte line per voxel (key + all accumulators), SURVEY section 8a-K8
// ---------------------------------------------------------------------------------------------
constexpr int kMaxVotes = 20;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr double kFixScale = 16777216.0;    // 2^24 fixed-point metres for centroid sums
struct __align__(128) Voxel {
    unsigned long long key;                 // 21-bit biased i | j<<21 | k<<42
    unsigned long long sx, sy, sz;          // sum of llrint(coord * 2^24), two's complement
    uint32_t n;                             // points fused
    uint32_t sr, sg, sb;                    // colour sums
    uint32_t votes[kMaxVotes];              // label histogram
};
static_assert(sizeof(Voxel) == 128, "voxel record must be one 128-byte line");

__host__ __device__ inline uint64_t pack_key(int i, int j, int k)
{
    return (uint64_t)(uint32_t)(i + (1 << 20)) | ((uint64_t)(uint32_t)(j + (1 << 20)) << 21) |
           ((uint64_t)(uint32_t)(k + (1 << 20)) << 42);
}
__host__ __device__ inline void unpack_key(uint64_t key, int& i, int& j, int& k)
{
    i = (int)(key & 0x1FFFFF) - (1 << 20);
    j = (int)((key >> 21) & 0x1FFFFF) - (1 << 20);
    k = (int)((key >> 42) & 0x1FFFFF) - (1 << 20);
}
__host__ __device__ inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
// spatial ownership: 8^3-voxel bricks hashed over the ranks
constexpr int kBrickShift = 3;
__host__ __device__ inline int voxel_owner(int i, int j, int k, int nranks)
{
    if (nranks <= 1) return 0;
    const uint64_t b = pack_key(i >> kBrickShift, j >> kBrickShift, k >> kBrickShift);
    return (int)(mix64(b ^ 0x9E3779B97F4A7C15ull) % (uint64_t)nranks);
}

// device-side constants of one context
struct DevParams {
    int W, H, D, W1;
    int Dl;                                 // disparities per column in the C / S_v layout: D, or D padded to 64 / 128 so that the
                                            // full-width kernels apply (lanes at d >= D are inactive; their cells are never read)
    int bs, P1, P2, uniq, d12, ftzero, speckle_win, speckle_diff;
    double cx, cy, fx, fy, baseline, scale, roix, roiy, roiz, max_depth_units;
    float inv_leaf;
    int num_labels;
    uint32_t palette[SSM_MAX_LABELS];       // b | g<<8 | r<<16
    uint32_t drop_mask, dynamic_mask;
    int dilate_radius, colour_source;
};

// inbox layout: [512-byte header: u32 count[2] (per parity), u32 overflow; at byte 256: u32 arrived[64], one word per source rank
// = the last step whose points that rank has stored here][parity 0 points][parity 1 points]
constexpr size_t kInboxHeader = 512;
constexpr size_t kInboxFlags = 256;

struct Point {                              // routed between ranks, 20 bytes
    float x, y, z;
    uint32_t rgba;                          // 0x00RRGGBB
    uint32_t label;
};

// what an insert kernel needs of the voxel hash: the table, the device counters ([0] points of the call, [1] voxels, [2] flags:
// bit 0 = table and spill list full, bit 1 = coordinate outside the key range, [3] export count, [4] parked points, [5] drain
// count) and the spill list that takes the points of inserts that found no slot within kMaxProbe probes
constexpr int kMaxProbe = 1024;
struct TableRef {
    Voxel* table;
    uint64_t mask;
    uint32_t* counters;
    Point* spill;
    uint32_t spill_cap;
};

}  // namespace ssm

// ---------------------------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------------------------
struct ssm_ctx {
    ssm_params p;
    ssm::DevParams dp;
    int device = 0;
    int sm_count = 148;
    int max_cluster = 16, min_cluster = 1;   // thread-block cluster sizes the vertical kernel may use (SSM_MAX_CLUSTER / SSM_MIN_CLUSTER)
    int tune[4] = {0, 0, 0, 0};           // SSM_TUNE0..3: experiment knobs (see the kernels that read them)
    bool force_legacy_hsweep = false;     // SSM_LEGACY_HSWEEP=1: one-kernel horizontal sweep (S_f through HBM) instead of checkpointed recomputation
    bool force_legacy_cost = false;       // SSM_LEGACY_COST=1: k_pix_hsum + k_vsum instead of the fused cost kernel
    bool force_legacy_vertical = false;   // SSM_LEGACY_VERTICAL=1: per-direction kernels instead of the cluster kernel
    bool no_pad = false;                  // SSM_NO_PAD=1: keep the cost-volume layout at exactly D disparities per column
    int select_rows = 0;                  // SSM_SELECT_ROWS: rows per band of the fused selection kernel (0 = 8)
    bool force_legacy_select = false;     // SSM_LEGACY_SELECT=1: finalize / L-R check / median / speckle as separate kernels instead of the fused band kernel
    cudaStream_t stream = nullptr;
    // sub-batch streams of the split pipeline (SSM_TUNE3 = number of concurrent sub-batches): kernels bound by different
    // units (shared-memory LSU, issue, HBM) overlap across sub-batches
    static constexpr int kMaxSplit = 4;
    cudaStream_t sub_stream[kMaxSplit] = {};
    cudaEvent_t sub_fork = nullptr, sub_join[kMaxSplit] = {};
    // staggered sub-batches (SSM_STAGGER=0 turns it off): sub-batch i + 1 starts its cost stage when sub-batch i has finished its
    // own, so that it runs next to i's vertical cluster kernel (which cannot use 16 of the 148 SMs) instead of in front of it:
    // 4830 -> 4900 frames/s at 66 KITTI frames per step
    int stagger = 1;
    cudaEvent_t sub_cost[kMaxSplit] = {};
    cudaEvent_t ev_after_cost = nullptr;   // recorded by run_sgbm behind the cost stage when set
    uint64_t launches = 0;
    bool timing = false;
    // stage timing: a ring of event sets, one per timed pipeline call; read back (averaged) by ssm_stage_time_ms
    static constexpr int kEvSets = 64;
    cudaEvent_t ev[kEvSets][SSM_STAGE_COUNT + 1] = {};
    int ev_set = 0;          // set used by the call being enqueued
    int ev_used = 0;         // sets recorded since the last ssm_set_stage_timing(1)
    float stage_ms[SSM_STAGE_COUNT] = {};

    // asynchronous host pipeline (ssm_pipeline_batch_host_async): a copy stream and two input staging sets, so that the
    // H2D copy of batch k+1 overlaps the kernels of batch k
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {}, ev_consumed[2] = {};
    uint8_t *stage_left[2] = {}, *stage_right[2] = {}, *stage_sem[2] = {}, *stage_rgb[2] = {};
    double* stage_pose[2] = {};
    uint64_t async_calls = 0;

    int vert_wave = 0, vert_wave_w = 0, vert_wave_h = 0;   // vertical_wave_frames() cache (per frame shape)
    bool vert_query = false;                              // launch_vertical_t: report the occupancy, do not launch
    int cap_w = 0, cap_h = 0, cap_b = 0;
    // stereo
    uint8_t *d_left = nullptr, *d_right = nullptr;   // [B][H][W] staging for host calls
    uint4 *d_recL = nullptr, *d_recR = nullptr;      // [B][H][W] prefilter + BT records
    uint32_t* d_ptab = nullptr;                      // [B][H][6][ptab_pitch] right image in table format (k_prefilter_tab -> k_cost_tma), 128-disparity layouts
    int ptab_pitch = 0, ptab_margin = 0;             // words per (row, quantity); zero words left of pixel 0 (padded layouts)
    bool no_cost_tma = false;                        // SSM_NO_COST_TMA=1: k_cost_fused (tables built per row inside the kernel) instead of k_cost_tma
    uint16_t* d_hs = nullptr;                        // [B][H][W1][D] horizontal window sums (aliases d_S)
    int16_t* d_C = nullptr;                          // [B][H][W1][D] matching cost
    uint16_t* d_S = nullptr;                         // [B][H][W1][D] aggregated cost
    int16_t *d_disp_raw = nullptr, *d_disp_lr = nullptr, *d_disp_med = nullptr, *d_disp = nullptr;  // [B][H][W]
    uint32_t* d_disp2key = nullptr;                  // [B][H][W]
    uint64_t* d_wta_rec = nullptr;                   // [B][H][W1] winner-take-all records, 8 bytes (sgbm_hsweep.cu) or 16 bytes (sgbm_hsweep2.cu) each
    uint32_t* d_ck = nullptr;                        // [B][H][blocks][state] left-to-right path checkpoints (sgbm_hsweep2.cu)
    uint32_t* d_uniq_thr = nullptr;                  // [32768] uniqueness threshold per minS
    int32_t *d_cc_label = nullptr, *d_cc_size = nullptr;  // speckle filter
    // mapper
    uint16_t* d_depth = nullptr;                     // [B][H][W]
    uint8_t *d_label = nullptr, *d_mask = nullptr;   // [B][H][W]
    uint8_t* d_label_lut = nullptr;                  // [2^24] semantic colour (B | G << 8 | R << 16) -> class id, 255 = not in the palette
    uint8_t *d_sem = nullptr, *d_rgb = nullptr;      // [B][H][W][3] staging
    double* d_pose = nullptr;                        // [B][16]
    int32_t* d_min_disp = nullptr;                   // [B]
    ssm::Point* d_points = nullptr;                  // [B*H*W] compacted cloud
    uint32_t* d_blk_count = nullptr;                 // per-block counts / offsets for ordered compaction
    uint32_t* d_counters = nullptr;                  // [8] misc device counters (0: n_points, 1: n_voxels, 2: overflow flag, 3: export count)
    // voxel hash (voxel_table.cu: growth, export)
    ssm::Voxel* d_table = nullptr;
    uint64_t table_slots = 0;
    ssm::Point* d_spill[2] = {};                     // points whose insert found no slot within kMaxProbe probes (current / draining)
    size_t spill_cap = 0;
    int spill_cur = 0;
    uint32_t* h_mirror = nullptr;                    // pinned copy of d_counters, refreshed behind every pipeline call (stale by <= 2 batches)
    uint64_t mirror_prev = 0, mirror_dmax = 0;       // voxel count seen at the previous pipeline call; largest increase per call
    static constexpr int kBatchRing = 4;
    cudaEvent_t ev_batch[kBatchRing] = {};           // completion of the latest pipeline calls (the host runs at most two batches ahead)
    uint64_t pipeline_calls = 0;                     // pipeline calls since the map was last cleared
    static constexpr int kRetired = 3;
    ssm::Voxel* retired[kRetired] = {};              // tables replaced by a growth step, freed once their records have moved
    cudaEvent_t retired_ev[kRetired] = {};
    uint64_t grows = 0;                              // growth steps so far
    bool auto_grow = true;                           // SSM_NO_GROW=1: a full table is SSM_ERR_CAPACITY, as in round 1
    void* export_ws = nullptr;                       // export workspace (grow-only)
    size_t export_ws_bytes = 0;
    float last_export_ms = 0.f;                      // device time of the latest export's kernels (index, sort, finalize)
    cudaStream_t user_stream = nullptr;              // stream of the latest ssm_pipeline_batch_device call, if the caller passed one
    // multi-GPU
    void* comm = nullptr;                            // ncclComm_t
    int rank = 0, nranks = 1;
    ssm::Point *d_send = nullptr, *d_recv = nullptr; // routing buffers
    uint32_t *d_send_counts = nullptr;               // [nranks] + offsets
    size_t route_cap = 0;
    // peer-memory routing (ssm_comm_ipc_*): every rank owns an inbox = {header, two point buffers (step parity)} that
    // its peers append to directly over NVLink
    static constexpr int kMaxPeers = 64;
    void* ipc_base = nullptr;                        // this rank's inbox allocation (exported through CUDA IPC)
    void* peer_base[kMaxPeers] = {};                 // peers' inbox allocations as mapped into this process (own entry = ipc_base)
    void** d_peer_base = nullptr;                    // the same table on the device
    size_t inbox_cap = 0;                            // points per parity buffer
    bool p2p = false;
    uint64_t p2p_step = 0;                           // parity source
    // route overlap (ssm_set_route_overlap): the per-batch exchange runs on its own stream behind the next batch's SGBM
    cudaStream_t route_stream = nullptr;
    cudaEvent_t ev_route_done = nullptr;
    bool route_overlap = false, route_pending = false;
    void* keyframes = nullptr;                       // cached camera-space keyframe clouds (api.cu: ssm_keyframe_*)
    void* labels_ws = nullptr;                       // label production workspace (labels.cu)
    void* cues_ws = nullptr;                         // dense motion cues workspace (cues.cu), allocated on first use
    void* ingest_ws = nullptr;                       // PNG ingest staging (ingest.cu), allocated on first use
};

namespace ssm {

void set_error(const std::string& s);
void cues_free(ssm_ctx* c);   // cues.cu
void labels_free(ssm_ctx* c); // labels.cu
void ingest_free(ssm_ctx* c); // ingest.cu
int cuda_fail(cudaError_t e, const char* what);

#define SSM_CUDA(expr)                                                      \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return ::ssm::cuda_fail(_e, #expr);          \
    } while (0)

// first statement of every extern "C" entry point that takes a context: the caller's thread may have another device current
#define SSM_ENTER(ctx) SSM_CUDA(cudaSetDevice((ctx)->device))

#define SSM_LAUNCH_CHECK(ctx)                                               \
    do {                                                                    \
        (ctx)->launches++;                                                  \
        cudaError_t _e = cudaGetLastError();                                \
        if (_e != cudaSuccess) return ::ssm::cuda_fail(_e, "kernel launch"); \
    } while (0)

// stage launchers (each returns an ssm_status); all buffers densely packed
int launch_prefilter(ssm_ctx* c, int B, const uint8_t* dL, const uint8_t* dR, cudaStream_t s);
int launch_cost_volume(ssm_ctx* c, int B, cudaStream_t s);
int launch_aggregate_vertical(ssm_ctx* c, int B, cudaStream_t s);     // down, down-right, down-left -> S_v
int launch_aggregate_horizontal(ssm_ctx* c, int B, cudaStream_t s);   // right, left, winner-take-all records
int launch_vertical(ssm_ctx* c, int B, cudaStream_t s, bool* done);   // cluster kernel; *done = false -> caller falls back
int vertical_wave_frames(ssm_ctx* c);   // frames whose vertical clusters are co-resident (one full wave of the cluster kernel); 0 = unknown
int launch_select(ssm_ctx* c, int B, cudaStream_t s);
int launch_hsweep(ssm_ctx* c, int B, cudaStream_t s);
int launch_hsweep2(ssm_ctx* c, int B, cudaStream_t s);       // checkpointed recomputation (D <= 128)
int launch_wta_finalize2(ssm_ctx* c, int B, cudaStream_t s);
bool hsweep2_supported(const ssm_ctx* c);
size_t hsweep2_ck_words(int W1, int D, int H, int B);
int launch_wta_finalize(ssm_ctx* c, int B, cudaStream_t s);
int launch_post(ssm_ctx* c, int B, int16_t* d_out, cudaStream_t s);
int launch_depth(ssm_ctx* c, int B, const int16_t* d_disp, uint16_t* d_depth, cudaStream_t s);
int launch_labels_mask(ssm_ctx* c, int B, const uint8_t* d_sem, cudaStream_t s);
int launch_points(ssm_ctx* c, int B, const uint16_t* d_depth, const uint8_t* d_sem, const uint8_t* d_rgb,
                  const double* d_pose, bool fuse_into_map, cudaStream_t s);
int launch_fuse_points(ssm_ctx* c, const Point* d_pts, const uint32_t* d_count, uint32_t max_count, cudaStream_t s);
int launch_map_clear(ssm_ctx* c, cudaStream_t s);
int launch_transform_fuse(ssm_ctx* c, const Point* d_pts, uint32_t n, const double* T16, cudaStream_t s);
int launch_export(ssm_ctx* c, Voxel* d_out, uint32_t max_out, cudaStream_t s);   // dense, unordered copy of the occupied records
// voxel_table.cu
inline TableRef table_ref(const ssm_ctx* c)
{
    return TableRef{c->d_table, c->table_slots - 1, c->d_counters, c->d_spill[c->spill_cur], (uint32_t)c->spill_cap};
}
int table_grow(ssm_ctx* c, uint64_t min_slots, cudaStream_t s);   // stream-ordered: allocate, move the records, free, drain the spill list
int spill_drain(ssm_ctx* c, cudaStream_t s);
int table_reap(ssm_ctx* c, bool wait);   // free the tables earlier growth steps replaced (wait: block until their moves are done)
int table_stats(ssm_ctx* c, cudaStream_t s, unsigned long long out[3]);   // occupied, sum of probe displacements, longest
// index -> (sort by (k, j, i)) -> K9 finalize -> D2H; d_recs = the hash table or a dense record list; ms_device (optional) = the
// device time of the kernels
int export_records_device(ssm_ctx* c, const Voxel* d_recs, uint64_t slots, uint64_t n_expected, bool sorted, const ssm_voxel_export* out,
                          uint64_t max_voxels, uint64_t* n_out, cudaStream_t s, float* ms_device);
int comm_gather_records(ssm_ctx* c, Voxel** d_all, uint64_t* n_all, cudaStream_t s);   // comm.cu: every rank's records on rank 0
int launch_route_bucket(ssm_ctx* c, uint32_t max_points, cudaStream_t s);   // d_points -> d_send grouped by owner rank
int route_and_fuse(ssm_ctx* c, cudaStream_t s);   // multi-GPU: bucket by owner, NCCL all-to-all, fuse received
// multi-GPU, peer-memory path: one kernel makes the points, fuses the locally owned ones and appends the others to
// their owner's inbox over NVLink; then a stream-ordered NCCL barrier and the fusion of this rank's inbox
int points_route_p2p(ssm_ctx* c, int B, const uint16_t* d_depth, const uint8_t* d_sem, const uint8_t* d_rgb, const double* d_pose,
                     cudaStream_t s);
int launch_points_p2p(ssm_ctx* c, int B, const uint16_t* d_depth, const uint8_t* d_sem, const uint8_t* d_rgb, const double* d_pose,
                      void* const* d_peer_base, int parity, cudaStream_t s);
int launch_fuse_inbox(ssm_ctx* c, int parity, cudaStream_t s);
int comm_barrier(ssm_ctx* c, cudaStream_t s);
int launch_flag_barrier(ssm_ctx* c, uint32_t step, cudaStream_t s);   // mapper.cu: peer-memory arrival flags instead of a collective
int comm_allreduce_max(ssm_ctx* c, uint32_t* value, cudaStream_t s);   // blocking: *value = max over the ranks

// packed 16x2 helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dup16(int v) { return (uint32_t)(v & 0xffff) * 0x10001u; }
__device__ __forceinline__ uint32_t fsl16(uint32_t lo_word, uint32_t hi_word)
{   // elements (lo_word.hi, hi_word.lo): shifts the pair stream by one 16-bit element towards higher d
    return __funnelshift_l(lo_word, hi_word, 16);
}
__device__ __forceinline__ uint32_t fsr16(uint32_t lo_word, uint32_t hi_word)
{   // elements (lo_word.hi, hi_word.lo) as well, expressed with the right funnel
    return __funnelshift_r(lo_word, hi_word, 16);
}

}  // namespace ssm
