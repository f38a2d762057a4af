// sgbm_hsweep2.cu -- the two horizontal SGM paths + winner-take-all with checkpointed recomputation (SURVEY.md
// App. A-4, A-5), sm_100a.  Used for layouts of up to 256 disparities (NR = 1, 2, 4 words per lane); sgbm_hsweep.cu remains the generic path.
//
// The one-kernel sweep (sgbm_hsweep.cu) writes S_f = S_v + L(->) for the whole row and reads it back on the way
// home: 10N bytes of HBM traffic for 4N algorithmic bytes, at 5.2 TB/s -- it is HBM-bound.  Here the left-to-right
// path is computed twice instead of being stored:
//   k_hfwd   one warp per row walks left to right through C only and drops a checkpoint (its D-wide state, 256 B)
//            every kBlk columns: reads 2N, writes ~1 %.
//   k_hrev   one warp per row walks the blocks right to left.  Per block: one elected lane pulls the block's C and
//            S_v rows into shared memory with two TMA bulk copies (the rows are contiguous: kBlk * D * 2 bytes)
//            tracked by an mbarrier; the warp re-runs the left-to-right recurrence over the block from its
//            checkpoint, folding it into the staged rows in place (T = sat(S_v + L->)); then walks the block right
//            to left with the right-to-left recurrence, S = sat(T + L<-) in registers, winner-take-all.
//            Reads 4N (+ checkpoints), writes 16 B per pixel.
// Winner-take-all on the ALU-pipe diet: first minimum by one REDUX on (S << 16 | d) keys; the uniqueness test
// "exists d outside [best-1, best+1] with S[d] * (100 - u) < minS * 100" as a packed compare against the exact
// integer threshold (computed with a multiply-high, no table load) -> one flag byte per disparity -> window bytes
// masked by a 64-bit shift -> one vote.  The sub-pixel neighbours are NOT extracted in the serial loop: the lane that
// holds the winner stores its packed costs and the two values across its lane borders (one 16-byte record), and the
// per-pixel finalize kernel picks S[best-1], S[best+1] from it.
#include <type_traits>

#include "sgbm_path.cuh"
#include "sgbm_wta.cuh"
#include "ssm_tma.cuh"

namespace ssm {

constexpr int kBlk = 16;    // columns per checkpoint block

// ---- pass A: left-to-right checkpoints -----------------------------------------------------------------
// ck[row][j][LW + 32] words, j = 1 .. nb-1: the state entering block j (after column j*kBlk - 1), packed minimum at [LW].
// One warp per row.  The row's cost blocks (kBlk columns = kBlk * D * 2 contiguous bytes) stream through a two-stage
// shared-memory ring filled by TMA bulk copies (one elected lane, one mbarrier per stage); the kBlk steps of a block
// are fully unrolled, so every shared-memory load has an immediate offset and the loop carries no address arithmetic.
// Only full blocks are walked: the last block of a row is never needed as a checkpoint source.
template <int NR, bool FULL /* D == 64 * NR: every lane owns disparities */>
__global__ void __launch_bounds__(128) k_hfwd(const int16_t* __restrict__ C, uint32_t* __restrict__ ck, int W1, int D /* layout */, int P1, int P2,
                                              int nrows, int nb, uint32_t one, int Dv /* valid disparities <= D */)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    const uint32_t colbytes = (uint32_t)D * 2, blkbytes = colbytes * kBlk;
    uint8_t* mine = smem_raw + (size_t)warp * 2 * blkbytes;            // two stages
    const uint32_t bar0 = smem_addr(smem_raw + (size_t)(blockDim.x >> 5) * 2 * blkbytes) + warp * 16;
    if (lane == 0) { mbar_init1(bar0); mbar_init1(bar0 + 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (row >= nrows) return;
    constexpr int LW = 32 * NR;
    const int d0 = lane * 2 * NR;
    const bool active = FULL || d0 < Dv;
    const PathLane pl = make_path_lane(lane, one, (uint32_t)P1);
    const uint32_t P1w = (uint32_t)P1 * 0x10001u, P2w = (uint32_t)P2 * 0x10001u;
    const uint32_t padC = (kBig - (uint32_t)P2) * 0x10001u;
    const uint8_t* Cg = reinterpret_cast<const uint8_t*>(C) + (size_t)row * W1 * colbytes;
    uint32_t* ckrow = ck + (size_t)row * nb * (LW + 32);
    const int nwalk = nb - 1;                    // blocks 0 .. nb-2, all full
    uint32_t sA = smem_addr(mine) + (uint32_t)lane * NR * 4;
    keep(sA);

    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < 2; ++st)
            if (st < nwalk) {
                mbar_expect(bar0 + 8 * st, blkbytes);
                tma_load_1d(smem_addr(mine) + st * blkbytes, Cg + (size_t)st * blkbytes, blkbytes, bar0 + 8 * st);
            }
    }
    uint32_t L[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) L[r] = 0u;
    uint32_t m = 0u;
    for (int j = 0; j < nwalk; ++j) {
        const int st = j & 1;
        mbar_wait(bar0 + 8 * st, (uint32_t)(j >> 1) & 1u);
        const uint32_t base = sA + st * blkbytes;
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            uint32_t cw[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) cw[r] = padC;
            if (active) lds_words<NR>(base + i * colbytes, cw);
            m = path_step<NR>(L, cw, m, P1w, P2w, pl);
        }
        uint32_t* dst = ckrow + (size_t)(j + 1) * (LW + 32);           // state entering block j + 1
        store_words<NR>(dst + lane * NR, L);
        if (lane == 0) dst[LW] = m;
        __syncwarp();                                                   // everyone has read this stage
        if (lane == 0 && j + 2 < nwalk) {
            mbar_expect(bar0 + 8 * st, blkbytes);
            tma_load_1d(smem_addr(mine) + st * blkbytes, Cg + (size_t)(j + 2) * blkbytes, blkbytes, bar0 + 8 * st);
        }
    }
}

// ---- pass B: blocks right to left, recompute + reverse path + winner-take-all -----------------------------------
struct HrevArgs {
    const int16_t* C;
    const uint16_t* Sv;
    const uint32_t* ck;
    uint4* rec;
    int W1, D, Dv, P1, P2, nrows, nb;   // D: disparities per column in the layout, Dv <= D: valid ones
    uint32_t one;
    // uniqueness threshold thr = min(umulhi(minS * mul + add, magic), 32768): exact ceil(minS * 100 / (100 - ratio)) with
    // magic = ceil(2^32 / den) (floor(n / den) == umulhi(n, magic) for n < 2^32 / den); see the launcher for ratio >= 100
    uint32_t uniq_mul, uniq_add, uniq_magic;
    int pf;                 // L2 prefetch of the next block while the current one is processed
};

template <int NR, bool FULL, int MINB = 5 /* CTAs per SM the register budget is cut for: 5 -> 96 registers, 6 -> 80 (no spills) */,
          bool PADDED = false /* padded layout (with FULL): all lanes load, the lanes at d >= Dv are kept out of the winner-take-all */>
__global__ void __launch_bounds__(128, MINB) k_hrev(const HrevArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    constexpr int LOG_DPL = NR == 1 ? 1 : (NR == 2 ? 2 : 3);     // log2(disparities per lane)
    const int D = a.D, W1 = a.W1;
    const uint32_t colbytes = (uint32_t)D * 2;   // one column of C / S_v
    const uint32_t blkbytes = colbytes * kBlk;
    // per warp: [C block][S block] + one mbarrier (at the end of the CTA's dynamic shared memory)
    uint8_t* myC = smem_raw + (size_t)warp * 2 * blkbytes;
    uint8_t* myS = myC + blkbytes;
    const uint32_t bar = smem_addr(smem_raw + (size_t)(blockDim.x >> 5) * 2 * blkbytes) + warp * 8;
    if (lane == 0) mbar_init1(bar);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (row >= a.nrows) return;

    constexpr int LW = 32 * NR;
    const int d0 = lane * 2 * NR;
    const bool active = FULL || d0 < a.Dv;
    const PathLane pl = make_path_lane(lane, a.one, (uint32_t)a.P1);
    const uint32_t P1w = (uint32_t)a.P1 * 0x10001u, P2w = (uint32_t)a.P2 * 0x10001u;
    const uint32_t padC = (kBig - (uint32_t)a.P2) * 0x10001u;
    const bool valid = PADDED ? d0 < a.Dv : active;   // lanes that hold disparities of the image
    const uint32_t lanemask = valid ? (NR >= 2 ? 0x80808080u : 0x00008080u) : 0u;   // per flag word (4 disparities)
    const uint8_t* Cg = reinterpret_cast<const uint8_t*>(a.C) + (size_t)row * W1 * colbytes;
    const uint8_t* Sg = reinterpret_cast<const uint8_t*>(a.Sv) + (size_t)row * W1 * colbytes;
    const uint32_t* ckrow = a.ck + (size_t)row * a.nb * (LW + 32);
    uint4* rrow = a.rec + (size_t)row * W1 * WtaRec<NR>::kQuads;
    const uint32_t loff = (uint32_t)lane * NR * 4;   // this lane's words inside a column
    uint32_t cA = smem_addr(myC) + loff, sA = smem_addr(myS) + loff;   // shared-window addresses of this lane's words
    keep(cA); keep(sA);

    uint32_t Lr[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) Lr[r] = 0u;
    uint32_t mr = 0u;
    uint32_t phase = 0u;

    for (int j = a.nb - 1; j >= 0; --j) {
        const int xs = j * kBlk, n = min(kBlk, W1 - xs);
        // stage the block: two bulk copies, one mbarrier phase
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my earlier generic writes to this buffer precede the async writes
            const uint32_t bytes = colbytes * (uint32_t)n;
            mbar_expect(bar, 2 * bytes);
            tma_load_1d(smem_addr(myC), Cg + (size_t)xs * colbytes, bytes, bar);
            tma_load_1d(smem_addr(myS), Sg + (size_t)xs * colbytes, bytes, bar);
            // the warp's buffer is single (six CTAs per SM leave 8 KB per warp), so the copy of block j - 1 cannot start before
            // block j is done.  Pulling its lines into L2 at this point (SSM_TUNE1=16) was measured and does not pay: the 24 warps of
            // an SM cover the copy latency between them
            if (a.pf && j > 0) {
                l2_prefetch_bulk(Cg + (size_t)(xs - kBlk) * colbytes, blkbytes);
                l2_prefetch_bulk(Sg + (size_t)(xs - kBlk) * colbytes, blkbytes);
            }
        }
        // the state entering the block, while the copies fly
        uint32_t Lf[NR];
        uint32_t mf = 0u;
#pragma unroll
        for (int r = 0; r < NR; ++r) Lf[r] = 0u;
        if (j > 0) {
            const uint32_t* src = ckrow + (size_t)j * (LW + 32);
            load_words<NR>(src + lane * NR, Lf);
            mf = src[LW];
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        // left to right again: T = sat(S_v + L->) replaces S_v in place
        auto walk = [&](auto nn) {
        const int n_ = decltype(nn)::value > 0 ? decltype(nn)::value : n;   // compile-time kBlk for full blocks
#pragma unroll
        for (int i = 0; i < n_; ++i) {
            uint32_t cw[NR], sw[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) { cw[r] = padC; sw[r] = 0u; }
            if (active) {
                lds_words<NR>(cA + i * colbytes, cw);
                lds_words<NR>(sA + i * colbytes, sw);
            }
            mf = path_step<NR>(Lf, cw, mf, P1w, P2w, pl);
            if (active) {
#pragma unroll
                for (int r = 0; r < NR; ++r) sw[r] = __viaddmin_u16x2(sw[r], Lf[r], kSatW);
                sts_words<NR>(sA + i * colbytes, sw);
            }
        }
        // right to left: the fifth path, the full sum, winner-take-all
#pragma unroll
        for (int i = n_ - 1; i >= 0; --i) {
            uint32_t cw[NR], tw[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) { cw[r] = padC; tw[r] = 0u; }
            if (active) {
                lds_words<NR>(cA + i * colbytes, cw);
                lds_words<NR>(sA + i * colbytes, tw);
            }
            mr = path_step<NR>(Lr, cw, mr, P1w, P2w, pl);
            uint32_t Sw[NR];
            uint32_t kmin = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                Sw[r] = valid ? __viaddmin_u16x2(tw[r], Lr[r], kSatW) : 0xffffffffu;
                const uint32_t klo = Sw[r] * pl.sh16 + (uint32_t)(d0 + 2 * r);
                const uint32_t khi = (Sw[r] & 0xffff0000u) | (uint32_t)(d0 + 2 * r + 1);
                kmin = __vimin3_u32(kmin, klo, khi);                       // smaller d wins ties: first minimum
            }
            kmin = __reduce_min_sync(0xffffffffu, kmin);
            const uint32_t minS = kmin >> 16, best = kmin & 0xffffu;
            // S[d] * (100 - u) < minS * 100  <=>  S[d] < thr, thr = ceil(minS * 100 / (100 - u)) (exact; capped at 32768)
            uint32_t thr;
            thr = min(__umulhi(minS * a.uniq_mul + a.uniq_add, a.uniq_magic), 32768u);
            // flag byte per disparity: bit 7 set <=> S < thr.  (S + 0x8000 - thr never leaves its 16-bit half.)
            const uint32_t kt = 0x80008000u - thr * 0x10001u;
            uint32_t flags;
            if constexpr (NR >= 2) flags = __byte_perm(Sw[0] * pl.one + kt, Sw[1] * pl.one + kt, 0x7531);
            else flags = __byte_perm(Sw[0] * pl.one + kt, 0u, 0x4431);
            const uint32_t below = ~flags & lanemask;
            // bytes of [best - 1, best + 1] inside this lane: 0x808080 shifted by whole bytes (64-bit shift clamps to zero)
            const uint32_t k8 = (uint32_t)((int)d0 + 4 - (int)best) * 8u;  // 8 * (4 - rel), rel = best - d0; wraps huge when negative
            unsigned long long win;
            asm("shr.u64 %0, %1, %2;" : "=l"(win) : "l"(0x0000808080000000ull), "r"(k8));
            uint32_t outside = below & ~(uint32_t)win;
            if constexpr (NR == 4) {   // the lane's second group of four disparities, d0 + 4 .. d0 + 7
                const uint32_t flags1 = __byte_perm(Sw[2] * pl.one + kt, Sw[3] * pl.one + kt, 0x7531);
                const uint32_t k81 = (uint32_t)((int)d0 + 8 - (int)best) * 8u;
                unsigned long long win1;
                asm("shr.u64 %0, %1, %2;" : "=l"(win1) : "l"(0x0000808080000000ull), "r"(k81));
                outside |= ~flags1 & lanemask & ~(uint32_t)win1;
            }
            const uint32_t reject = __any_sync(0xffffffffu, outside != 0u) ? 1u : 0u;
            // values across the lane borders, for the sub-pixel fit when the winner sits at a lane edge
            const uint32_t upv = __shfl_up_sync(0xffffffffu, Sw[NR - 1] >> 16, 1);
            const uint32_t dnv = __shfl_down_sync(0xffffffffu, Sw[0] & 0xffffu, 1);
            if (lane == (int)(best >> LOG_DPL)) {
                if constexpr (NR == 4) {
                    rrow[2 * (xs + i)] = make_uint4(minS | (best << 16) | (reject << 31), Sw[0], Sw[1], Sw[2]);
                    rrow[2 * (xs + i) + 1] = make_uint4(Sw[3], (upv & 0xffffu) | (dnv << 16), 0u, 0u);
                } else {
                    rrow[xs + i] = make_uint4(minS | (best << 16) | (reject << 31), Sw[0], NR == 2 ? Sw[NR - 1] : 0u, (upv & 0xffffu) | (dnv << 16));
                }
            }
        }
        };
        if (n == kBlk) walk(std::integral_constant<int, kBlk>{});
        else walk(std::integral_constant<int, 0>{});
        __syncwarp();
    }
}

// sub-pixel fit, disp2 candidates, raw disparity: one thread per pixel of the valid region (16-byte records)
template <int NR>
__global__ void __launch_bounds__(256) k_wta_finalize2(const uint4* __restrict__ rec, int16_t* __restrict__ disp_raw,
                                                       uint32_t* __restrict__ disp2key, int W, int D, size_t total)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int W1 = W - D;
    const int xp = (int)(idx % W1);
    const size_t row = idx / W1;
    const uint4 r = rec[idx * WtaRec<NR>::kQuads];
    const uint4 r2 = NR == 4 ? rec[idx * 2 + 1] : r;
    const int x = xp + D;
    int minS, best;
    bool valid;
    const int out = wta2_decode<NR>(r, r2, D, minS, best, valid);
    if (valid) atomicMin(&disp2key[row * W + (x - best)], ((uint32_t)minS << 16) | (uint32_t)(0xffff - x));
    disp_raw[row * W + x] = (int16_t)out;
}

// ------------------------------------------------------------------------------------------------
size_t hsweep2_ck_words(int W1, int D, int H, int B)
{
    const int NR = D <= 64 ? 1 : (D <= 128 ? 2 : 4);
    const int nb = (W1 + kBlk - 1) / kBlk;
    return (size_t)B * H * nb * (32 * NR + 32);
}

bool hsweep2_supported(const ssm_ctx* c) { return c->dp.Dl <= 256 && !c->force_legacy_hsweep && c->d_ck != nullptr; }

template <int NR>
static int launch_hsweep2_t(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const int nrows = B * p.H;
    const int nb = (p.W1 + kBlk - 1) / kBlk;
    const int wpb = 4;
    const unsigned grid = (unsigned)((nrows + wpb - 1) / wpb);
    if (nb > 1) {
        const size_t smem_f = (size_t)wpb * 2 * kBlk * p.Dl * 2 + wpb * 16;
        if (p.D == 64 * NR || p.Dl != p.D) {   // full-width layout, or a padded one (its cells at d >= D hold the "+inf" cost)
            SSM_CUDA(cudaFuncSetAttribute(k_hfwd<NR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
            k_hfwd<NR, true><<<grid, wpb * 32, smem_f, s>>>(c->d_C, c->d_ck, p.W1, p.Dl, p.P1, p.P2, nrows, nb, 1u, p.D);
        } else {
            SSM_CUDA(cudaFuncSetAttribute(k_hfwd<NR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
            k_hfwd<NR, false><<<grid, wpb * 32, smem_f, s>>>(c->d_C, c->d_ck, p.W1, p.Dl, p.P1, p.P2, nrows, nb, 1u, p.D);
        }
        SSM_LAUNCH_CHECK(c);
    }
    HrevArgs a;
    a.C = c->d_C; a.Sv = c->d_S; a.ck = c->d_ck; a.rec = reinterpret_cast<uint4*>(c->d_wta_rec);
    a.W1 = p.W1; a.D = p.Dl; a.Dv = p.D; a.P1 = p.P1; a.P2 = p.P2; a.nrows = nrows; a.nb = nb; a.one = 1u;
    a.pf = c->tune[1] == 16 ? 1 : 0;   // measured: 2.10 ms with the prefetch, 2.07 ms without (per 33 frames) -- off
    if (p.uniq < 100) {   // thr = ceil(minS * 100 / den) = floor((minS * 100 + den - 1) / den)
        const uint32_t den = (uint32_t)(100 - p.uniq);
        a.uniq_mul = 100u; a.uniq_add = den - 1u;
        a.uniq_magic = (uint32_t)(((1ull << 32) + den - 1) / den);
        if (den == 1) { a.uniq_mul = 200u; a.uniq_add = 0u; a.uniq_magic = 0x80000000u; }   // 2^32 / 1 does not fit: n / 1 = 2n / 2
    } else {              // thr = minS ? 32768 : 0  ==  min(floor(minS * 32768 / 1), 32768) with the division by 1 as umulhi(n << ..)
        a.uniq_mul = 65536u; a.uniq_add = 0u; a.uniq_magic = 0x80000000u;   // umulhi(minS << 16, 2^31) = minS << 15 >= 32768 for minS >= 1
    }
    const size_t smem = (size_t)wpb * 2 * kBlk * p.Dl * 2 + wpb * 8;
    auto go = [&](auto kern) -> int {
        SSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, wpb * 32, smem, s>>>(a);
        return SSM_OK;
    };
    int rc;
    if (p.Dl != p.D) rc = NR == 4 ? go(k_hrev<NR, true, 3, true>) : go(k_hrev<NR, true, 6, true>);   // padded layout
    else if constexpr (NR == 4) rc = p.D == 64 * NR ? go(k_hrev<NR, true, 3>) : go(k_hrev<NR, false, 3>);   // 64 KB of staging per CTA: three per SM
    else if (p.D == 64 * NR) rc = c->tune[1] == 5 ? go(k_hrev<NR, true, 5>) : go(k_hrev<NR, true, 6>);   // 80 registers: six CTAs per SM (SSM_TUNE1=5: 96 registers, five)
    else rc = go(k_hrev<NR, false, 6>);
    if (rc) return rc;
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

int launch_hsweep2(ssm_ctx* c, int B, cudaStream_t s)
{
    return c->dp.Dl <= 64 ? launch_hsweep2_t<1>(c, B, s) : (c->dp.Dl <= 128 ? launch_hsweep2_t<2>(c, B, s) : launch_hsweep2_t<4>(c, B, s));
}

int launch_wta_finalize2(ssm_ctx* c, int B, cudaStream_t s)
{
    const DevParams& p = c->dp;
    const size_t total = (size_t)B * p.H * p.W1;
    const unsigned grid = (unsigned)((total + 255) / 256);
    const uint4* rec = reinterpret_cast<const uint4*>(c->d_wta_rec);
    if (p.Dl <= 64) k_wta_finalize2<1><<<grid, 256, 0, s>>>(rec, c->d_disp_raw, c->d_disp2key, p.W, p.D, total);
    else if (p.Dl <= 128) k_wta_finalize2<2><<<grid, 256, 0, s>>>(rec, c->d_disp_raw, c->d_disp2key, p.W, p.D, total);
    else k_wta_finalize2<4><<<grid, 256, 0, s>>>(rec, c->d_disp_raw, c->d_disp2key, p.W, p.D, total);
    SSM_LAUNCH_CHECK(c);
    return SSM_OK;
}

}  // namespace ssm
