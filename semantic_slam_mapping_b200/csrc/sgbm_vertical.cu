// sgbm_vertical.cu -- the three top-down SGM paths (down, down-right, down-left) in ONE launch.
//
// Replaces three of the five path passes inside cv::StereoSGBM (called from /root/reference src/stereo.cpp:30).
// The per-direction kernel (sgbm_aggregate.cu) streams C once per direction and read-modify-writes S twice:
// 16N bytes of HBM traffic for 4N algorithmic bytes.  Here a thread-block CLUSTER owns one frame and walks it
// row by row; each CTA owns a strip of T columns and keeps the D-wide state of all three directions for its
// strip in shared memory (u16x2 words, one 64*NR-disparity line per column and direction), so every cost row is
// read once and S_v = sat(L_down + L_dr + L_dl) is written once.
//
//   * a warp takes a column: one coalesced 2*D-byte load of C(y, x, .), three path_step recurrences (packed
//     VIADDMNMX.U16x2, d+-1 by SHFL + funnel shift, min over d by one REDUX each), one coalesced store of S_v;
//   * the diagonal directions are stored in a ring indexed by (x - y) mod T resp. (x + y) mod T, so that the
//     predecessor of column x in the previous row sits in the very slot column x is about to overwrite:
//     the update is in place, no second copy of the state;
//   * the only cross-CTA dependency is one column of state per side and row: the owner writes it straight into
//     the neighbour's halo buffer through distributed shared memory (st.shared::cluster), double-buffered by row
//     parity; one barrier.cluster per row orders both the intra-CTA and the inter-CTA hand-over;
//   * C rows are prefetched G columns ahead in registers (the next row's first group is in flight across the
//     barrier).
// Out-of-image predecessors are the all-zero state (m = 0), as in the per-direction kernel.
#include <cooperative_groups.h>

#include <algorithm>

#include "sgbm_path.cuh"

namespace cg = cooperative_groups;

namespace ssm {

// ---- mbarrier / st.async helpers (shared::cluster addresses are 32-bit) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// asynchronous store into a peer CTA's shared memory; the peer's mbarrier is credited with the bytes on arrival
template <int NR>
__device__ __forceinline__ void st_async_words(uint32_t remote_addr, const uint32_t (&w)[NR], uint32_t remote_bar)
{
    if constexpr (NR == 1) {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(w[0]), "r"(remote_bar) : "memory");
    } else if constexpr (NR == 2) {
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "r"(w[0]), "r"(w[1]), "r"(remote_bar) : "memory");
    } else {
#pragma unroll
        for (int q = 0; q < NR / 4; ++q)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr + 16 * q),
                         "r"(w[4 * q]), "r"(w[4 * q + 1]), "r"(w[4 * q + 2]), "r"(w[4 * q + 3]), "r"(remote_bar) : "memory");
    }
}
__device__ __forceinline__ void st_async_word(uint32_t remote_addr, uint32_t v, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v), "r"(remote_bar) : "memory");
}

template <int NR, int NWARPS, bool FULL /* D == 2 * NR * LANES: every lane owns disparities */, int LANES = 32 /* lanes per column */>
__global__ void __launch_bounds__(NWARPS * 32, NWARPS <= 8 ? 2 : 1) k_vertical3(const int16_t* __restrict__ C, uint16_t* __restrict__ S, int W1, int H,
                                                             int D /* disparities per column in the layout */, int P1, int P2, int T,
                                                             uint32_t one, int pf_rows, int Dv /* valid disparities <= D */)
{
    extern __shared__ __align__(16) uint32_t smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int frame = blockIdx.x / CS;
    // a "virtual warp" of LANES lanes takes a column; with LANES == 16 a physical warp walks two columns at once
    constexpr int VPW = 32 / LANES;              // virtual warps per warp
    constexpr int NVW = NWARPS * VPW;            // virtual warps per CTA
    const int lane = threadIdx.x & (LANES - 1);  // lane within the virtual warp
    const int warp = (int)(threadIdx.x / LANES); // virtual warp
    const int pwarp = threadIdx.x >> 5;          // physical warp
    constexpr int LW = LANES * NR;               // words per (column, direction) line
    constexpr int LWB = LW * 4;                  // bytes per line
    constexpr uint32_t kHandBytes = LWB + 4;     // one hand-over: a state line + its packed minimum

    const int x0 = rank * T;
    const int Tc = max(0, min(T, W1 - x0));      // columns of this CTA (the last strips may be shorter or empty)
    const int R = Tc + 2;                        // ring length of the diagonal directions: two spare slots, because the
                                                 // neighbour that fills the halo slot may run one row ahead of this CTA
    const bool has_left = rank > 0 && Tc > 0;
    const bool has_right = Tc > 0 && (x0 + T < W1);
    const int Tn = has_right ? min(T, W1 - (x0 + T)) : 0;   // columns of the right neighbour (the left one always has T)

    // shared memory: st0 [T] lines (down), st1 / st2 [T+2] lines (rings), the three arrays of packed minima, two mbarriers
    const int n_state = (3 * T + 4) * LW;
    const int n_min = (3 * (T + 2) + 1) & ~1;     // keeps the mbarriers 8-byte aligned
    const int n_all = n_state + n_min + 4;
    for (int i = threadIdx.x; i < n_all; i += NWARPS * 32) smem[i] = 0u;
    char* const sm = reinterpret_cast<char*>(smem);
    const int b0 = lane * NR * 4, b1 = b0 + T * LWB, b2 = b1 + (T + 2) * LWB;   // byte offsets incl. this lane's words
    const int mb0 = n_state * 4, mb1 = mb0 + (T + 2) * 4, mb2 = mb1 + (T + 2) * 4;
    const uint32_t sm_a = smem_u32(smem);
    uint32_t a0 = sm_a + b0, a1 = sm_a + b1, a2 = sm_a + b2;                   // shared-window addresses of this lane's words
    uint32_t am0 = sm_a + mb0, am1 = sm_a + mb1, am2 = sm_a + mb2;
    keep(a0); keep(a1); keep(a2); keep(am0); keep(am1); keep(am2);
    const uint32_t bar_l = sm_a + (uint32_t)(n_state + n_min) * 4;      // credited by the left neighbour (ring 1 halo)
    const uint32_t bar_r = bar_l + 8;                                          // credited by the right neighbour (ring 2 halo)
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_init(bar_l, 1);
        mbar_init(bar_r, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // phase 0 collects what the neighbours produce during their row 0
        if (has_left) mbar_arrive_expect_tx(bar_l, kHandBytes);
        if (has_right) mbar_arrive_expect_tx(bar_r, kHandBytes);
    }
    // peers' windows onto this layout
    const uint32_t right_a = has_right ? mapa_u32(sm_a, (uint32_t)(rank + 1)) : 0u;
    const uint32_t left_a = has_left ? mapa_u32(sm_a, (uint32_t)(rank - 1)) : 0u;
    const uint32_t right_bar_l = has_right ? mapa_u32(bar_l, (uint32_t)(rank + 1)) : 0u;   // my ring-1 data lands at the right peer's bar_l
    const uint32_t left_bar_r = has_left ? mapa_u32(bar_r, (uint32_t)(rank - 1)) : 0u;
    cluster.sync();

    const int d0 = lane * 2 * NR;
    const bool active = FULL || d0 < Dv;
    const uint32_t P1w = (uint32_t)P1 * 0x10001u, P2w = (uint32_t)P2 * 0x10001u;
    const uint32_t padC = (kBig - (uint32_t)P2) * 0x10001u;
    const PathLane pl = make_path_lane<LANES>((int)(threadIdx.x & 31), one, (uint32_t)P1);
    const size_t rowbytes = (size_t)W1 * D * 2;
    const char* Crow = reinterpret_cast<const char*>(C) + ((size_t)frame * H * W1 + x0) * D * 2 + d0 * 2;
    char* Srow = reinterpret_cast<char*>(S) + ((size_t)frame * H * W1 + x0) * D * 2 + d0 * 2;
    const int nk = Tc > warp ? (Tc - warp + NVW - 1) / NVW : 0;         // my columns: warp, warp + NVW, ...
    const int nk_w = Tc > pwarp * VPW ? (Tc - pwarp * VPW + NVW - 1) / NVW : 0;   // trip count of the physical warp (its first virtual warp's)
    const uint32_t cstep = (uint32_t)NVW * D * 2;                       // bytes between my consecutive columns
    const int wl = Tc > 0 ? (Tc - 1) % NVW : -1;                        // the virtual warp that owns the strip's last column

    uint32_t Cw[NR], Cn[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { Cw[r] = padC; Cn[r] = padC; }
    if (active && nk > 0) load_words_cs<NR>(Crow + (size_t)warp * D * 2, Cw);

    // ring phases: my column c lives in slot (c + o1) mod R of ring 1 and (c + o2) mod R of ring 2
    int o1 = 0, o2 = 0;
    int h1 = Tc + 1;                             // ring-1 slot of my halo column -1 for this row's hand-over: (-1 - y) mod R
    int h2 = Tc;                                 // ring-2 slot of my halo column Tc: (Tc + y) mod R
    int n1 = Tn + 1;                             // the same slot in the right neighbour's ring 1: (-1 - y) mod (Tn + 2)
    int l2 = T;                                  // and in the left neighbour's ring 2: (T + y) mod (T + 2)
    // L2 prefetch (one TMA bulk instruction per row and CTA) pf_rows rows ahead of the register prefetch below
    const uint32_t seg_bytes = (uint32_t)Tc * D * 2;            // my strip of one cost row is contiguous in memory
    const char* const Cseg = Crow - d0 * 2;
    if (threadIdx.x == 0 && Tc > 0 && pf_rows > 0)
        for (int q = 1; q < pf_rows && q < H; ++q) l2_prefetch_bulk(Cseg + (size_t)q * rowbytes, seg_bytes);
    for (int y = 0; y < H; ++y) {
        if (threadIdx.x == 0 && Tc > 0 && pf_rows > 0 && y + pf_rows < H) l2_prefetch_bulk(Cseg + (size_t)(y + pf_rows) * rowbytes, seg_bytes);
        // the hand-overs of the neighbours' previous row must have landed before the border columns are touched
        if (y > 0) {
            const uint32_t parity = (uint32_t)(y - 1) & 1u;
            if (pwarp == 0 && has_left) {
                mbar_wait_cluster(bar_l, parity);
                if (threadIdx.x == 0) mbar_arrive_expect_tx(bar_l, kHandBytes);      // arm the next phase (this row's hand-over)
            }
            if (pwarp == wl / VPW && has_right) {
                mbar_wait_cluster(bar_r, parity);
                if ((threadIdx.x & 31) == 0) mbar_arrive_expect_tx(bar_r, kHandBytes);
            }
        }
        uint32_t off = (uint32_t)warp * D * 2;
        int c = warp;
        int s1 = c + o1; s1 = (int)min((unsigned)s1, (unsigned)(s1 - R));
        int s2 = c + o2; s2 = (int)min((unsigned)s2, (unsigned)(s2 - R));
        for (int k = 0; k < nk_w; ++k) {
            const bool mine = VPW == 1 || k < nk;    // the last trip of a warp may cover only its first column(s)
            // prefetch the cost of my next column (next row's first one at the end of the row)
            if (active) {
                if (k + 1 < nk) load_words_cs<NR>(Crow + off + cstep, Cn);
                else if (k + 1 == nk && y + 1 < H) load_words_cs<NR>(Crow + rowbytes + (size_t)warp * D * 2, Cn);
            }
            const uint32_t p0 = a0 + c * LWB, p1 = a1 + s1 * LWB, p2 = a2 + s2 * LWB;
            const uint32_t q0 = am0 + c * 4, q1 = am1 + s1 * 4, q2 = am2 + s2 * 4;
            uint32_t L0[NR], L1[NR], L2[NR];
            uint32_t m0 = 0u, m1 = 0u, m2 = 0u;
            if (mine) {
                lds_words<NR>(p0, L0);
                lds_words<NR>(p1, L1);
                lds_words<NR>(p2, L2);
                m0 = lds_u32(q0); m1 = lds_u32(q1); m2 = lds_u32(q2);
            } else {
#pragma unroll
                for (int r = 0; r < NR; ++r) { L0[r] = 0u; L1[r] = 0u; L2[r] = 0u; }
            }
            m0 = path_step<NR, LANES>(L0, Cw, m0, P1w, P2w, pl);
            m1 = path_step<NR, LANES>(L1, Cw, m1, P1w, P2w, pl);
            m2 = path_step<NR, LANES>(L2, Cw, m2, P1w, P2w, pl);
            if (mine) {
                sts_words<NR>(p0, L0);
                sts_words<NR>(p1, L1);
                sts_words<NR>(p2, L2);
                if (lane == 0) { sts_u32(q0, m0); sts_u32(q1, m1); sts_u32(q2, m2); }
            }
            if (active && mine) {
                uint32_t o[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) o[r] = __viaddmin_u16x2(__viaddmin_u16x2(L0[r], L1[r], kSatW), L2[r], kSatW);
                store_words_cs<NR>(Srow + off, o);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) Cw[r] = Cn[r];
            off += cstep;
            c += NVW;
            s1 += NVW; s1 = (int)min((unsigned)s1, (unsigned)(s1 - R));
            s2 += NVW; s2 = (int)min((unsigned)s2, (unsigned)(s2 - R));
        }
        // hand the strip's border columns to the neighbours (asynchronous stores that credit the peer's mbarrier; no
        // fence, no cluster barrier), or feed zeros at the image border.  Nothing is sent after the last row: the
        // peer may already have left.
        if (warp == wl) {
            // my last column's down-right state is the right neighbour's halo column -1 for the next row
            int sl = Tc - 1 + o1; sl = (int)min((unsigned)sl, (unsigned)(sl - R));
            if (has_right) {
                if (y + 1 < H) {
                    uint32_t Lx[NR];
                    load_words<NR>(sm + sl * LWB + b1, Lx);
                    st_async_words<NR>(right_a + (uint32_t)(n1 * LWB + b1), Lx, right_bar_l);
                    if (lane == 0) st_async_word(right_a + (uint32_t)(mb1 + n1 * 4), *(reinterpret_cast<uint32_t*>(sm + mb1) + sl), right_bar_l);
                }
            } else if (Tc > 0) {
                // image border on the right: the down-left predecessor of my last column is the zero state
                uint32_t Z[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) Z[r] = 0u;
                store_words<NR>(sm + h2 * LWB + b2, Z);
                if (lane == 0) *(reinterpret_cast<uint32_t*>(sm + mb2) + h2) = 0u;
            }
        }
        if (warp == 0 && Tc > 0) {
            if (has_left) {
                if (y + 1 < H) {
                    uint32_t Lx[NR];
                    load_words<NR>(sm + o2 * LWB + b2, Lx);         // my column 0 lives in ring-2 slot (0 + o2)
                    st_async_words<NR>(left_a + (uint32_t)(l2 * LWB + b2), Lx, left_bar_r);
                    if (lane == 0) st_async_word(left_a + (uint32_t)(mb2 + l2 * 4), *(reinterpret_cast<uint32_t*>(sm + mb2) + o2), left_bar_r);
                }
            } else {
                uint32_t Z[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) Z[r] = 0u;
                store_words<NR>(sm + h1 * LWB + b1, Z);
                if (lane == 0) *(reinterpret_cast<uint32_t*>(sm + mb1) + h1) = 0u;
            }
        }
        Crow += rowbytes;
        Srow += rowbytes;
        // next row: ring 1 rotates one slot back, ring 2 one slot forward
        o1 = o1 == 0 ? R - 1 : o1 - 1;
        o2 = o2 + 1 == R ? 0 : o2 + 1;
        h1 = h1 == 0 ? R - 1 : h1 - 1;
        h2 = h2 + 1 == R ? 0 : h2 + 1;
        n1 = n1 == 0 ? Tn + 1 : n1 - 1;
        l2 = l2 == T + 1 ? 0 : l2 + 1;
        __syncthreads();                         // the strip's own columns: row y is complete before row y + 1 reads it
    }
    cluster.sync();                              // nobody leaves while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------
struct VerticalPlan {
    int cluster = 0, T = 0;
    size_t smem = 0;
};

template <int NR, int NWARPS, bool FULL, int LANES = 32>
static int launch_vertical_t(ssm_ctx* c, int B, const VerticalPlan& plan, cudaStream_t s, bool* done)
{
    const DevParams& p = c->dp;
    auto kern = k_vertical3<NR, NWARPS, FULL, LANES>;
    SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    if (plan.cluster > 8) SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * plan.cluster));
    cfg.blockDim = dim3(NWARPS * 32);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)plan.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    const cudaError_t occ = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    static const bool dbg = getenv("SSM_DEBUG_CLUSTERS") != nullptr;
    if (dbg) fprintf(stderr, "k_vertical3<%d,%d>: cluster %d, T %d, smem %zu, max active clusters %d (%s)\n", NR, NWARPS, plan.cluster, plan.T, plan.smem, nclusters, cudaGetErrorString(occ));
    if (c->vert_query) {             // vertical_wave_frames(): the answer is the occupancy, nothing is launched
        c->vert_wave = occ == cudaSuccess ? nclusters : 0;
        if (occ != cudaSuccess) cudaGetLastError();
        *done = true;
        return SSM_OK;
    }
    if (occ != cudaSuccess || nclusters < 1) {
        cudaGetLastError();      // this cluster shape cannot be co-scheduled on this device: fall back
        return SSM_OK;
    }
    // L2 prefetch distance of the cost rows (SSM_TUNE0): two rows ahead, one for the two-columns-per-warp kernels of 128-disparity
    // layouts (measured at KITTI size, 18 warps: 0 / 1 / 2 / 3 / 4 / 6 rows -> 2.39 / 2.12 / 2.15 / 2.28 / 2.41 / 2.47 ms per 33 frames)
    const int pf_rows = c->tune[0] >= 0 ? c->tune[0] : (NR == 4 && LANES == 16 ? 1 : 2);
    SSM_CUDA(cudaLaunchKernelEx(&cfg, kern, (const int16_t*)c->d_C, c->d_S, p.W1, p.H, p.Dl, p.P1, p.P2, plan.T, 1u, pf_rows, p.D));
    SSM_LAUNCH_CHECK(c);
    *done = true;
    return SSM_OK;
}

static size_t vertical_smem(int NR, int T)
{
    const int LW = 32 * NR;
    return sizeof(uint32_t) * ((size_t)(3 * T + 4) * LW + ((3 * (T + 2) + 1) & ~1) + 4);
}

// Smallest cluster whose per-CTA strip fits shared memory -- or, for the few-frame calls of the drop-in API (one frame per
// calDisparity_SGBM call), a larger one: a frame is walked row by row, so its latency is rows x trips per row, and a cluster of
// 16 narrow strips makes three trips per row where a cluster of 4 makes nine.  Larger clusters are only taken while all of
// them are co-resident (at most 64 CTAs: four 16-CTA or eight 8-CTA clusters fit the GPCs).  Returns false when no cluster
// shape fits (caller falls back to the per-direction kernels).
static bool plan_vertical(const ssm_ctx* c, int B, VerticalPlan& plan)
{
    const DevParams& p = c->dp;
    const int NR = p.Dl <= 64 ? 1 : (p.Dl <= 128 ? 2 : (p.Dl <= 256 ? 4 : 8));
    const size_t limit = 225 * 1024;
    bool found = false;
    // powers of two, plus an explicitly requested size (SSM_MIN_CLUSTER = 3, 5, 6, 7 ...: any strip count works)
    for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16}) {
        if (cs > c->max_cluster) break;
        if (cs < c->min_cluster) continue;
        if ((cs & (cs - 1)) != 0 && cs != c->min_cluster) continue;
        const int T = (p.W1 + cs - 1) / cs;
        const size_t need = vertical_smem(NR, T);
        if (need > limit) continue;
        if (found && (B * cs > 64 || T < 32)) break;     // a larger cluster than needed: only for small batches and strips of a full trip
        plan.cluster = cs; plan.T = T; plan.smem = need;
        found = true;
    }
    return found;
}

static int launch_vertical_plan(ssm_ctx* c, int B, const VerticalPlan& plan, cudaStream_t s, bool* done)
{
    const int D = c->dp.Dl;                       // the layout picks the kernel; lanes at d >= dp.D are inactive (FULL = false)
    // padded layouts: the cost kernel wrote the "+inf" cost into the cells at d >= dp.D, so every lane runs the same code
    const bool full = c->dp.D != D || (D == 64 || D == 128 || D == 256 || D == 512);
    if (D <= 64) {
        // D == 64 / 32: two columns per warp as well (16 lanes x 2 words / x 1 word); SSM_TUNE2=3 keeps one column per warp
        if (D == 64 && c->tune[2] != 3)
            return full ? launch_vertical_t<2, 16, true, 16>(c, B, plan, s, done) : launch_vertical_t<2, 16, false, 16>(c, B, plan, s, done);
        if (D == 32 && c->tune[2] != 3) return launch_vertical_t<1, 16, true, 16>(c, B, plan, s, done);
        return full ? launch_vertical_t<1, 32, true>(c, B, plan, s, done) : launch_vertical_t<1, 32, false>(c, B, plan, s, done);
    }
    if (D <= 128) {
        // D == 128: two columns per warp (16 lanes x 4 words each) halve the per-column overhead of the recurrence
        if (D == 128 && c->tune[2] == 0) {
            // 17 warps walk 34 columns per trip: taken when that saves a trip per row (strips of 289..306 columns, e.g. the
            // 291 of a 1241-pixel frame at the reference's 80 disparities: nine trips instead of ten)
            if ((plan.T + 33) / 34 < (plan.T + 31) / 32)
                return full ? launch_vertical_t<4, 17, true, 16>(c, B, plan, s, done) : launch_vertical_t<4, 17, false, 16>(c, B, plan, s, done);
            // ... and 18 warps (36 columns per trip) when only they do: the 279-column strips of a 1241-pixel frame at 128
            // disparities take eight trips instead of nine (2.175 -> 2.152 ms per 33 frames; 20 and 24 warps measure slower)
            if (full && (plan.T + 35) / 36 < (plan.T + 31) / 32) return launch_vertical_t<4, 18, true, 16>(c, B, plan, s, done);
            return full ? launch_vertical_t<4, 16, true, 16>(c, B, plan, s, done) : launch_vertical_t<4, 16, false, 16>(c, B, plan, s, done);
        }
        if (D == 128 && full && c->tune[2] == 2) return launch_vertical_t<4, 24, true, 16>(c, B, plan, s, done);
        if (D == 128 && full && c->tune[2] == 4) return launch_vertical_t<4, 18, true, 16>(c, B, plan, s, done);
        if (D == 128 && full && c->tune[2] == 6) return launch_vertical_t<4, 8, true, 16>(c, B, plan, s, done);   // two CTAs per SM (8-CTA clusters)
        if (D == 128 && full && c->tune[2] == 5) return launch_vertical_t<4, 20, true, 16>(c, B, plan, s, done);
        if (full && c->tune[2] == 1) return launch_vertical_t<2, 24, true>(c, B, plan, s, done);
        return full ? launch_vertical_t<2, 32, true>(c, B, plan, s, done) : launch_vertical_t<2, 32, false>(c, B, plan, s, done);
    }
    if (D <= 256) return full ? launch_vertical_t<4, 16, true>(c, B, plan, s, done) : launch_vertical_t<4, 16, false>(c, B, plan, s, done);
    return full ? launch_vertical_t<8, 16, true>(c, B, plan, s, done) : launch_vertical_t<8, 16, false>(c, B, plan, s, done);
}

int launch_vertical(ssm_ctx* c, int B, cudaStream_t s, bool* done)
{
    *done = false;
    if (c->force_legacy_vertical) return SSM_OK;
    VerticalPlan plan, smallest;
    if (!plan_vertical(c, B, plan)) return SSM_OK;
    int rc = launch_vertical_plan(c, B, plan, s, done);
    if (rc || *done) return rc;
    // the latency-oriented cluster shape cannot be co-scheduled here: the smallest shape that fits shared memory
    if (plan_vertical(c, 1 << 20, smallest) && smallest.cluster != plan.cluster) return launch_vertical_plan(c, B, smallest, s, done);
    return SSM_OK;
}

// One wave of the cluster kernel = as many frames as clusters are co-resident (33 four-CTA clusters at KITTI size with 128
// disparities, 7 sixteen-CTA clusters at 2048 x 1024 with 256): the kernel's duration is rows x trips per row whatever the number of
// clusters in flight, so a batch is best cut into sub-batches of whole waves (api.cu: run_pipeline).
int vertical_wave_frames(ssm_ctx* c)
{
    if (c->vert_wave_w == c->dp.W && c->vert_wave_h == c->dp.H) return c->vert_wave;
    c->vert_wave_w = c->dp.W; c->vert_wave_h = c->dp.H; c->vert_wave = 0;
    VerticalPlan plan;
    if (c->force_legacy_vertical || !plan_vertical(c, 1 << 20, plan)) return 0;
    bool done = false;
    c->vert_query = true;
    const uint64_t launches = c->launches;
    launch_vertical_plan(c, 1 << 20, plan, nullptr, &done);
    c->vert_query = false;
    c->launches = launches;
    return c->vert_wave;
}

}  // namespace ssm
