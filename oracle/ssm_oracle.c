/*
 * ssm_oracle.c -- CPU ORACLE (test infrastructure only; see ssm_oracle.h).
 *
 * Plain-C restatement of the reference hot path.  Each function cites the reference
 * file:line it follows; the SGBM chain follows SURVEY.md Appendix A (the semantics of
 * the un-vendored cv::StereoSGBM that src/stereo.cpp:13-30 calls), and is pinned
 * bit-exactly against cv2 4.13 (tests/test_oracle_golden.py: committed vectors from
 * tests/golden/make_golden.py, plus live comparisons where cv2 is importable).
 * The glue (rgbdframe.cpp:85-116, rgbdframe.h:63-75, mapper.cpp:12-94,189-216) and the
 * motion cues (stereo.cpp:41-192, uvdisparity.cpp:195-366) are pinned to the reference's
 * own source files compiled untouched into oracle/_ref (tests/test_oracle_mapper_ref.py,
 * tests/test_oracle_cues.py; vectors in tests/golden/mapper_ref.npz, cues_ref.npz).
 * PARITY UNPINNED for the PCL 1.7 part only (pcl::VoxelGrid / pcl::transformPointCloud,
 * SURVEY.md Appendix B): PCL cannot be built or run here, the voxel fusion below restates
 * its published algorithm; this affects centroid rounding (tolerance 1e-5) and the output
 * order, not voxel membership, counts or votes.
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off: the fp64 glue must not be fused).
 */
#include "ssm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAX_COST 32767
#define DISP_SHIFT 4
#define DISP_SCALE 16

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------
 * A-1 prefilter: x-Sobel clipped to [0, 2*ftzero] and the raw intensity, both forced to
 * ftzero in the first and last column.
 * ------------------------------------------------------------------------------------------ */
static void prefilter_row(const uint8_t* img, int W, int H, size_t stride, int y, int ftzero,
                          uint8_t* grad, uint8_t* raw)
{
    const uint8_t* cur = img + (size_t)y * stride;
    const uint8_t* up = img + (size_t)(y > 0 ? y - 1 : y) * stride;
    const uint8_t* dn = img + (size_t)(y < H - 1 ? y + 1 : y) * stride;
    for (int x = 1; x < W - 1; ++x) {
        int s = (cur[x + 1] - cur[x - 1]) * 2 + (up[x + 1] - up[x - 1]) + (dn[x + 1] - dn[x - 1]);
        grad[x] = (uint8_t)(iclamp(s, -ftzero, ftzero) + ftzero);
        raw[x] = cur[x];
    }
    grad[0] = grad[W - 1] = (uint8_t)ftzero;
    raw[0] = raw[W - 1] = (uint8_t)ftzero;
}

/* A-2 helper: half-sample interval [lo,hi] of a 1-D signal (Birchfield-Tomasi). */
static void bt_interval(const uint8_t* p, int W, uint8_t* lo, uint8_t* hi)
{
    for (int x = 0; x < W; ++x) {
        int c = p[x];
        int a = x > 0 ? (c + p[x - 1]) / 2 : c;
        int b = x < W - 1 ? (c + p[x + 1]) / 2 : c;
        lo[x] = (uint8_t)imin(imin(a, b), c);
        hi[x] = (uint8_t)imax(imax(a, b), c);
    }
}

/* A-2: pixel cost for one row, pix[x'][d], x' = x - D. */
static void pixel_cost_row(const uint8_t* left, const uint8_t* right, int W, int H, size_t stride, int y,
                           int D, int ftzero, uint8_t* scratch /* 12*W */, uint16_t* pix /* W1*D */)
{
    uint8_t* gl = scratch;
    uint8_t* rl = gl + W;
    uint8_t* gr = rl + W;
    uint8_t* rr = gr + W;
    uint8_t* lo[4];
    uint8_t* hi[4];
    for (int i = 0; i < 4; ++i) {
        lo[i] = rr + W + (size_t)(2 * i) * W;
        hi[i] = lo[i] + W;
    }
    prefilter_row(left, W, H, stride, y, ftzero, gl, rl);
    prefilter_row(right, W, H, stride, y, ftzero, gr, rr);
    const uint8_t* sigL[2] = {gl, rl};
    const uint8_t* sigR[2] = {gr, rr};
    bt_interval(gl, W, lo[0], hi[0]);
    bt_interval(rl, W, lo[1], hi[1]);
    bt_interval(gr, W, lo[2], hi[2]);
    bt_interval(rr, W, lo[3], hi[3]);
    const int W1 = W - D;
    memset(pix, 0, sizeof(uint16_t) * (size_t)W1 * D);
    for (int c = 0; c < 2; ++c) {
        const int shift = c == 0 ? 0 : 2;
        const uint8_t *pl = sigL[c], *pr = sigR[c];
        const uint8_t *loL = lo[c], *hiL = hi[c], *loR = lo[2 + c], *hiR = hi[2 + c];
        for (int x = D; x < W; ++x) {
            const int u = pl[x], u0 = loL[x], u1 = hiL[x];
            uint16_t* out = pix + (size_t)(x - D) * D;
            for (int d = 0; d < D; ++d) {
                const int xr = x - d;
                const int v = pr[xr], v0 = loR[xr], v1 = hiR[xr];
                const int c0 = imax(0, imax(u - v1, v0 - u));
                const int c1 = imax(0, imax(v - u1, u0 - v));
                out[d] = (uint16_t)(out[d] + (imin(c0, c1) >> shift));
            }
        }
    }
}

/* A-1..A-3: block-summed matching cost C[y][x'][d]. */
static void cost_volume(const uint8_t* left, const uint8_t* right, int W, int H, size_t stride, int D, int bs,
                        int ftzero, int16_t* C)
{
    const int W1 = W - D, r = bs / 2;
    const size_t row = (size_t)W1 * D;
    uint16_t* pix = (uint16_t*)malloc(sizeof(uint16_t) * row);
    uint16_t* hs = (uint16_t*)malloc(sizeof(uint16_t) * row * (size_t)H); /* horizontal sums, all rows */
    uint8_t* scratch = (uint8_t*)malloc((size_t)12 * W);
    for (int y = 0; y < H; ++y) {
        pixel_cost_row(left, right, W, H, stride, y, D, ftzero, scratch, pix);
        uint16_t* h = hs + row * y;
        /* window sum over x' with replicate clamping, maintained incrementally along x' */
        for (int d = 0; d < D; ++d) {
            int s = 0;
            for (int k = -r; k <= r; ++k) s += pix[(size_t)iclamp(k, 0, W1 - 1) * D + d];
            h[d] = (uint16_t)s;
        }
        for (int xp = 1; xp < W1; ++xp) {
            const uint16_t* in = pix + (size_t)iclamp(xp + r, 0, W1 - 1) * D;
            const uint16_t* out = pix + (size_t)iclamp(xp - r - 1, 0, W1 - 1) * D;
            const uint16_t* hp = h + (size_t)(xp - 1) * D;
            uint16_t* hc = h + (size_t)xp * D;
            for (int d = 0; d < D; ++d) hc[d] = (uint16_t)(hp[d] + in[d] - out[d]);
        }
    }
    /* window sum over y with replicate clamping, maintained incrementally down the rows */
    {
        int32_t* acc = (int32_t*)calloc(row, sizeof(int32_t));
        for (int k = -r; k <= r; ++k) {
            const uint16_t* h = hs + row * (size_t)iclamp(k, 0, H - 1);
            for (size_t i = 0; i < row; ++i) acc[i] += h[i];
        }
        for (int y = 0; y < H; ++y) {
            int16_t* c = C + row * y;
            if (y > 0) {
                const uint16_t* in = hs + row * (size_t)iclamp(y + r, 0, H - 1);
                const uint16_t* out = hs + row * (size_t)iclamp(y - r - 1, 0, H - 1);
                for (size_t i = 0; i < row; ++i) acc[i] += in[i] - out[i];
            }
            for (size_t i = 0; i < row; ++i) c[i] = (int16_t)acc[i];
        }
        free(acc);
    }
    free(scratch);
    free(hs);
    free(pix);
}

/* A-4 one step of the path recurrence.  prev == NULL means "predecessor outside the image". */
static inline int path_step(const int16_t* Cp, const int16_t* prev, int prev_min, int D, int P1, int P2,
                            int legacy, int16_t* out)
{
    int m = MAX_COST;
    if (!prev) {
        /* zero state: min(0, P1, P1, P2) - 0 (paper form) or - P2 (legacy form) */
        for (int d = 0; d < D; ++d) {
            int L = Cp[d] + (legacy ? -P2 : 0);
            out[d] = (int16_t)L;
            m = imin(m, L);
        }
        return m;
    }
    const int delta = prev_min + P2;
    for (int d = 0; d < D; ++d) {
        const int a = prev[d];
        const int b = (d > 0 ? prev[d - 1] : MAX_COST) + P1;
        const int c = (d < D - 1 ? prev[d + 1] : MAX_COST) + P1;
        const int L = Cp[d] + imin(imin(a, b), imin(c, delta)) - (legacy ? delta : prev_min);
        out[d] = (int16_t)L;
        m = imin(m, L);
    }
    return m;
}

static inline int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

/* A-4..A-6: 5-direction single-pass aggregation, WTA, uniqueness, sub-pixel, disp2, L-R check. */
static void aggregate_and_select(const int16_t* C, int W, int H, int D, const osgbm_params* p, int16_t* disp,
                                 size_t dstride, int16_t* S_out, int16_t* Sf_out, int16_t* Sv_out)
{
    const int W1 = W - D, minX1 = D;
    const int P1 = p->p1 > 0 ? p->p1 : 2;
    const int P2 = imax(p->p2 > 0 ? p->p2 : 5, P1 + 1);
    const int uniq = p->uniqueness_ratio >= 0 ? p->uniqueness_ratio : 10;
    const int d12 = p->disp12_max_diff > 0 ? p->disp12_max_diff : 1;
    const int legacy = p->legacy_p2_form;
    const int INVALID = -DISP_SCALE; /* (minD - 1) * 16, minD = 0 */
    const size_t row = (size_t)W1 * D;

    /* previous/current row of L for the three top-down directions, plus min per pixel */
    int16_t* Lprev[3];
    int16_t* Lcur[3];
    int* mprev[3];
    int* mcur[3];
    for (int i = 0; i < 3; ++i) {
        Lprev[i] = (int16_t*)malloc(sizeof(int16_t) * row);
        Lcur[i] = (int16_t*)malloc(sizeof(int16_t) * row);
        mprev[i] = (int*)malloc(sizeof(int) * W1);
        mcur[i] = (int*)malloc(sizeof(int) * W1);
    }
    int16_t* L0 = (int16_t*)malloc(sizeof(int16_t) * D * 2);
    int16_t* S = (int16_t*)malloc(sizeof(int16_t) * row);
    int* disp2 = (int*)malloc(sizeof(int) * W);
    int* disp2cost = (int*)malloc(sizeof(int) * W);

    for (int y = 0; y < H; ++y) {
        const int16_t* Cy = C + row * y;
        int16_t* drow = disp + dstride * y;
        /* forward sweep: r0 = (-1,0), r1 = (-1,-1), r2 = (0,-1), r3 = (+1,-1) */
        int m0 = 0;
        for (int xp = 0; xp < W1; ++xp) {
            const int16_t* Cp = Cy + (size_t)xp * D;
            int16_t* cur0 = L0 + (size_t)(xp & 1) * D;
            const int16_t* prev0 = xp > 0 ? L0 + (size_t)((xp - 1) & 1) * D : NULL;
            m0 = path_step(Cp, prev0, m0, D, P1, P2, legacy, cur0);
            const int has_up = y > 0;
            const int16_t* p1 = (has_up && xp > 0) ? Lprev[0] + (size_t)(xp - 1) * D : NULL;
            const int16_t* p2 = has_up ? Lprev[1] + (size_t)xp * D : NULL;
            const int16_t* p3 = (has_up && xp < W1 - 1) ? Lprev[2] + (size_t)(xp + 1) * D : NULL;
            mcur[0][xp] = path_step(Cp, p1, p1 ? mprev[0][xp - 1] : 0, D, P1, P2, legacy, Lcur[0] + (size_t)xp * D);
            mcur[1][xp] = path_step(Cp, p2, p2 ? mprev[1][xp] : 0, D, P1, P2, legacy, Lcur[1] + (size_t)xp * D);
            mcur[2][xp] = path_step(Cp, p3, p3 ? mprev[2][xp + 1] : 0, D, P1, P2, legacy, Lcur[2] + (size_t)xp * D);
            int16_t* Sp = S + (size_t)xp * D;
            for (int d = 0; d < D; ++d) {
                int s = sat16(cur0[d] + Lcur[0][(size_t)xp * D + d]);
                s = sat16(s + Lcur[1][(size_t)xp * D + d]);
                s = sat16(s + Lcur[2][(size_t)xp * D + d]);
                Sp[d] = (int16_t)s;
                if (Sv_out) /* the three top-down directions alone (a device-side intermediate, for stage parity) */
                    Sv_out[row * y + (size_t)xp * D + d] = (int16_t)sat16(
                        sat16(Lcur[0][(size_t)xp * D + d] + Lcur[1][(size_t)xp * D + d]) + Lcur[2][(size_t)xp * D + d]);
            }
        }
        if (Sf_out) memcpy(Sf_out + row * y, S, sizeof(int16_t) * row); /* S_f = sat(L0+L1+L2+L3) */
        /* reverse sweep r = (+1,0) fused with the selection (A-5) */
        for (int x = 0; x < W; ++x) {
            drow[x] = (int16_t)INVALID;
            disp2[x] = INVALID;
            disp2cost[x] = MAX_COST;
        }
        int mr = 0;
        for (int xp = W1 - 1; xp >= 0; --xp) {
            const int16_t* Cp = Cy + (size_t)xp * D;
            int16_t* cur = L0 + (size_t)(xp & 1) * D;
            const int16_t* prev = xp < W1 - 1 ? L0 + (size_t)((xp + 1) & 1) * D : NULL;
            mr = path_step(Cp, prev, mr, D, P1, P2, legacy, cur);
            int16_t* Sp = S + (size_t)xp * D;
            int minS = MAX_COST, best = -1;
            for (int d = 0; d < D; ++d) {
                const int s = sat16(Sp[d] + cur[d]);
                Sp[d] = (int16_t)s;
                if (s < minS) {
                    minS = s;
                    best = d;
                }
            }
            if (best < 0) best = 0; /* every S saturated at MAX_COST: cv2 keeps bestDisp=-1 -> see note */
            int d;
            for (d = 0; d < D; ++d)
                if (Sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
            if (d < D) continue;
            const int x = xp + minX1;
            const int x2 = x - best;
            if (disp2cost[x2] > minS) {
                disp2cost[x2] = minS;
                disp2[x2] = best;
            }
            int d16;
            if (0 < best && best < D - 1) {
                const int denom2 = imax(Sp[best - 1] + Sp[best + 1] - 2 * Sp[best], 1);
                d16 = best * DISP_SCALE + ((Sp[best - 1] - Sp[best + 1]) * DISP_SCALE + denom2) / (denom2 * 2);
            } else {
                d16 = best * DISP_SCALE;
            }
            drow[x] = (int16_t)d16;
        }
        if (S_out) memcpy(S_out + row * y, S, sizeof(int16_t) * row);
        /* A-6 left-right consistency */
        for (int x = minX1; x < W; ++x) {
            const int d1 = drow[x];
            if (d1 == INVALID) continue;
            const int dlo = d1 >> DISP_SHIFT, dhi = (d1 + DISP_SCALE - 1) >> DISP_SHIFT;
            const int xlo = x - dlo, xhi = x - dhi;
            if (0 <= xlo && xlo < W && disp2[xlo] >= 0 && abs(disp2[xlo] - dlo) > d12 && 0 <= xhi && xhi < W &&
                disp2[xhi] >= 0 && abs(disp2[xhi] - dhi) > d12)
                drow[x] = (int16_t)INVALID;
        }
        for (int i = 0; i < 3; ++i) {
            int16_t* t = Lprev[i];
            Lprev[i] = Lcur[i];
            Lcur[i] = t;
            int* tm = mprev[i];
            mprev[i] = mcur[i];
            mcur[i] = tm;
        }
    }
    for (int i = 0; i < 3; ++i) {
        free(Lprev[i]);
        free(Lcur[i]);
        free(mprev[i]);
        free(mcur[i]);
    }
    free(L0);
    free(S);
    free(disp2);
    free(disp2cost);
}

/* A-6 tail: cv::medianBlur(disp, disp, 3) on int16, replicate border. */
void oracle_median3x3_s16(const int16_t* src, int16_t* dst, int W, int H)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int16_t v[9];
            int n = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    v[n++] = src[(size_t)iclamp(y + dy, 0, H - 1) * W + iclamp(x + dx, 0, W - 1)];
            for (int i = 1; i < 9; ++i) { /* insertion sort */
                int16_t t = v[i];
                int j = i - 1;
                while (j >= 0 && v[j] > t) {
                    v[j + 1] = v[j];
                    --j;
                }
                v[j + 1] = t;
            }
            dst[(size_t)y * W + x] = v[4];
        }
}

/* A-7: cv::filterSpeckles == 4-connected components with edge predicate |a-b| <= max_diff. */
void oracle_filter_speckles(int16_t* img, int W, int H, int new_val, int max_speckle_size, int max_diff)
{
    const size_t n = (size_t)W * H;
    uint8_t* seen = (uint8_t*)calloc(n, 1);
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * n);
    int32_t* comp = (int32_t*)malloc(sizeof(int32_t) * n);
    for (size_t s = 0; s < n; ++s) {
        if (seen[s] || img[s] == new_val) continue;
        size_t top = 0, cnt = 0;
        stack[top++] = (int32_t)s;
        seen[s] = 1;
        while (top) {
            const int32_t q = stack[--top];
            comp[cnt++] = q;
            const int x = q % W, y = q / W;
            const int v = img[q];
            const int nx[4] = {x - 1, x + 1, x, x};
            const int ny[4] = {y, y, y - 1, y + 1};
            for (int k = 0; k < 4; ++k) {
                if (nx[k] < 0 || nx[k] >= W || ny[k] < 0 || ny[k] >= H) continue;
                const size_t t = (size_t)ny[k] * W + nx[k];
                if (seen[t] || img[t] == new_val || abs(img[t] - v) > max_diff) continue;
                seen[t] = 1;
                stack[top++] = (int32_t)t;
            }
        }
        if ((int)cnt <= max_speckle_size)
            for (size_t i = 0; i < cnt; ++i) img[comp[i]] = (int16_t)new_val;
    }
    /* NB: pixels reset to new_val keep seen=1, matching a labelling done on the input image. */
    free(comp);
    free(stack);
    free(seen);
}

int oracle_sgbm(const uint8_t* left, const uint8_t* right, int W, int H, size_t stride, const osgbm_params* p,
                int16_t* disp, size_t dstride, int16_t* C_out, int16_t* S_out, int16_t* disp_raw_out,
                int16_t* disp_median_out, int16_t* Sf_out, int16_t* Sv_out)
{
    const int D = p->num_disparities;
    if (D <= 0 || D % 16 != 0 || W <= D || H <= 0 || p->block_size < 1 || p->block_size % 2 == 0) return -1;
    const int W1 = W - D;
    const int ftzero = imax(p->pre_filter_cap, 15) | 1;
    int16_t* C = C_out ? C_out : (int16_t*)malloc(sizeof(int16_t) * (size_t)W1 * D * H);
    cost_volume(left, right, W, H, stride, D, p->block_size, ftzero, C);

    int16_t* raw = (int16_t*)malloc(sizeof(int16_t) * (size_t)W * H);
    aggregate_and_select(C, W, H, D, p, raw, (size_t)W, S_out, Sf_out, Sv_out);
    if (!C_out) free(C);
    if (disp_raw_out) memcpy(disp_raw_out, raw, sizeof(int16_t) * (size_t)W * H);

    int16_t* med = (int16_t*)malloc(sizeof(int16_t) * (size_t)W * H);
    oracle_median3x3_s16(raw, med, W, H);
    if (disp_median_out) memcpy(disp_median_out, med, sizeof(int16_t) * (size_t)W * H);
    if (p->speckle_window_size > 0)
        oracle_filter_speckles(med, W, H, -DISP_SCALE, p->speckle_window_size, DISP_SCALE * p->speckle_range);
    for (int y = 0; y < H; ++y) memcpy(disp + dstride * y, med + (size_t)W * y, sizeof(int16_t) * W);
    free(med);
    free(raw);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Glue: FrameReader::next disparity -> depth  (src/rgbdframe.cpp:85-116)
 * ------------------------------------------------------------------------------------------ */
void oracle_disparity_to_depth(const int16_t* disp, int W, int H, const omap_params* p, uint16_t* depth)
{
    int min_disp = 32767; /* rgbdframe.cpp:85 cv::minMaxIdx */
    for (size_t i = 0; i < (size_t)W * H; ++i)
        if (disp[i] < min_disp) min_disp = disp[i];
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            const short d = disp[(size_t)v * W + u];
            uint16_t out = 0; /* rgbdframe.cpp:86 */
            if (d != 0 && d != min_disp) { /* :103, :110 (FLT_EPSILON tests on integers) */
                const double pw = p->baseline / (1.0 * (double)d);
                const double px = (((double)u - p->cx) * pw) * 16.0;
                const double py = (((double)v - p->cy) * pw) * 16.0;
                const double pz = (p->fx * pw) * 16.0;
                if (fabs(px) < p->roix && fabs(py) < p->roiy && fabs(pz) < p->roiz && pz > 0)
                    out = (uint16_t)(pz * p->scale); /* :113 truncation */
            }
            depth[(size_t)v * W + u] = out;
        }
}

uint8_t oracle_label_of(const omap_params* p, uint8_t b, uint8_t g, uint8_t r)
{
    for (int i = 0; i < p->num_labels; ++i)
        if (p->palette_bgr[i][0] == b && p->palette_bgr[i][1] == g && p->palette_bgr[i][2] == r) return (uint8_t)i;
    return 255;
}

/* Mapper::semantic_motion_fuse (src/mapper.cpp:189-216); dilate with all-ones 3x3, out-of-image ignored. */
void oracle_moving_mask(const uint8_t* sem, int W, int H, const omap_params* p, uint8_t* mask)
{
    const size_t n = (size_t)W * H;
    uint8_t* a = (uint8_t*)malloc(n);
    uint8_t* b = (uint8_t*)malloc(n);
    for (size_t i = 0; i < n; ++i) {
        const uint8_t l = oracle_label_of(p, sem[3 * i], sem[3 * i + 1], sem[3 * i + 2]);
        a[i] = (l != 255 && ((p->dynamic_mask >> l) & 1u)) ? 255 : 0;
    }
    for (int it = 0; it < p->dilate_iterations; ++it) {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                uint8_t m = 0;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        if (a[(size_t)yy * W + xx] > m) m = a[(size_t)yy * W + xx];
                    }
                b[(size_t)y * W + x] = m;
            }
        uint8_t* t = a;
        a = b;
        b = t;
    }
    memcpy(mask, a, n);
    free(a);
    free(b);
}

/* Mapper::generatePointCloud (src/mapper.cpp:12-94) + RGBDFrame::project2dTo3d (include/rgbdframe.h:63-75)
 * + pcl::transformPointCloud with a Matrix4d. */
int oracle_generate_point_cloud(const uint16_t* depth, const uint8_t* sem, const uint8_t* rgb, int W, int H,
                                const omap_params* p, const double* T, float* xyz, float* xyz_cam, uint32_t* rgba,
                                uint8_t* label, int32_t* pix)
{
    uint8_t* mask = (uint8_t*)malloc((size_t)W * H);
    oracle_moving_mask(sem, W, H, p, mask);
    int n = 0;
    for (int m = 0; m < H; ++m)
        for (int c = 0; c < W; ++c) {
            const size_t i = (size_t)m * W + c;
            const uint16_t d = depth[i];
            if (d == 0) continue;                                   /* mapper.cpp:28 */
            if ((double)d > p->max_distance * p->scale) continue;   /* :30 */
            if (mask[i] == 255) continue;                           /* :32 */
            const uint8_t sb = sem[3 * i], sg = sem[3 * i + 1], sr = sem[3 * i + 2];
            const uint8_t l = oracle_label_of(p, sb, sg, sr);
            if (l != 255 && ((p->drop_mask >> l) & 1u)) continue;   /* :41-55 */
            /* rgbdframe.h:71-73: each assignment rounds to float */
            const float z = (float)((double)d / p->scale);
            const float x = (float)(((double)c - p->cx) * (double)z / p->fx);
            const float y = (float)(((double)m - p->cy) * (double)z / p->fy);
            if (xyz_cam) {
                xyz_cam[3 * n] = x;
                xyz_cam[3 * n + 1] = y;
                xyz_cam[3 * n + 2] = z;
            }
            /* pcl::transformPointCloud(Matrix4d): double, left to right, no FMA */
            for (int r = 0; r < 3; ++r) {
                double acc = T[4 * r] * (double)x;
                acc = acc + T[4 * r + 1] * (double)y;
                acc = acc + T[4 * r + 2] * (double)z;
                acc = acc + T[4 * r + 3];
                xyz[3 * n + r] = (float)acc;
            }
            uint8_t cb, cg, cr;
            if (p->colour_source == 0) { /* mapper.cpp:72-84 */
                cb = rgb[3 * i];
                cg = rgb[3 * i + 1];
                cr = rgb[3 * i + 2];
            } else { /* mapper.cpp~:60 */
                cb = sb;
                cg = sg;
                cr = sr;
            }
            rgba[n] = ((uint32_t)cr << 16) | ((uint32_t)cg << 8) | (uint32_t)cb;
            label[n] = l;
            if (pix) pix[n] = (int32_t)i;
            ++n;
        }
    free(mask);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Voxel map: pcl::VoxelGrid<PointXYZRGBA> semantics + label histogram.
 * Chained hash keyed by (i,j,k); fp32 sums in insertion order (PCL sums in sorted order, which
 * is unspecified for equal keys -> centroids carry the 1e-5 tolerance), plus fp64 sums.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t i, j, k;
    uint32_t n;
    float sx, sy, sz, sr, sg, sb;
    double dx, dy, dz;
    uint32_t votes[32];
    int64_t next;
} ovoxel;

struct ovoxel_map {
    float inv_leaf;
    int num_labels;
    ovoxel* v;
    int64_t size, cap;
    int64_t* buckets;
    int64_t nbuckets;
};

static uint64_t okey_hash(int32_t i, int32_t j, int32_t k)
{
    uint64_t h = (uint64_t)(uint32_t)i * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)(uint32_t)j * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (uint64_t)(uint32_t)k * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return h;
}

ovoxel_map* oracle_map_create(double leaf, int num_labels)
{
    ovoxel_map* m = (ovoxel_map*)calloc(1, sizeof(ovoxel_map));
    m->inv_leaf = 1.0f / (float)leaf; /* VoxelGrid::setLeafSize: inverse_leaf_size_ in fp32 */
    m->num_labels = num_labels;
    m->cap = 1 << 16;
    m->v = (ovoxel*)malloc(sizeof(ovoxel) * m->cap);
    m->nbuckets = 1 << 17;
    m->buckets = (int64_t*)malloc(sizeof(int64_t) * m->nbuckets);
    for (int64_t i = 0; i < m->nbuckets; ++i) m->buckets[i] = -1;
    return m;
}

void oracle_map_destroy(ovoxel_map* m)
{
    if (!m) return;
    free(m->v);
    free(m->buckets);
    free(m);
}

void oracle_map_clear(ovoxel_map* m)
{
    m->size = 0;
    for (int64_t i = 0; i < m->nbuckets; ++i) m->buckets[i] = -1;
}

static void omap_rehash(ovoxel_map* m)
{
    m->nbuckets *= 4;
    m->buckets = (int64_t*)realloc(m->buckets, sizeof(int64_t) * m->nbuckets);
    for (int64_t i = 0; i < m->nbuckets; ++i) m->buckets[i] = -1;
    for (int64_t i = 0; i < m->size; ++i) {
        const uint64_t b = okey_hash(m->v[i].i, m->v[i].j, m->v[i].k) & (uint64_t)(m->nbuckets - 1);
        m->v[i].next = m->buckets[b];
        m->buckets[b] = i;
    }
}

void oracle_map_insert(ovoxel_map* m, const float* xyz, const uint32_t* rgba, const uint8_t* label, int n)
{
    for (int q = 0; q < n; ++q) {
        const float x = xyz[3 * q], y = xyz[3 * q + 1], z = xyz[3 * q + 2];
        if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue; /* is_dense == false, mapper.cpp:92 */
        const int32_t i = (int32_t)floorf(x * m->inv_leaf);
        const int32_t j = (int32_t)floorf(y * m->inv_leaf);
        const int32_t k = (int32_t)floorf(z * m->inv_leaf);
        const uint64_t b = okey_hash(i, j, k) & (uint64_t)(m->nbuckets - 1);
        int64_t e = m->buckets[b];
        while (e >= 0 && !(m->v[e].i == i && m->v[e].j == j && m->v[e].k == k)) e = m->v[e].next;
        if (e < 0) {
            if (m->size == m->cap) {
                m->cap *= 2;
                m->v = (ovoxel*)realloc(m->v, sizeof(ovoxel) * m->cap);
            }
            e = m->size++;
            memset(&m->v[e], 0, sizeof(ovoxel));
            m->v[e].i = i;
            m->v[e].j = j;
            m->v[e].k = k;
            m->v[e].next = m->buckets[b];
            m->buckets[b] = e;
            if (m->size > m->nbuckets) omap_rehash(m);
        }
        ovoxel* v = &m->v[e];
        v->n += 1;
        v->sx += x;
        v->sy += y;
        v->sz += z;
        v->dx += (double)x;
        v->dy += (double)y;
        v->dz += (double)z;
        v->sr += (float)((rgba[q] >> 16) & 255u);
        v->sg += (float)((rgba[q] >> 8) & 255u);
        v->sb += (float)(rgba[q] & 255u);
        if (label[q] < m->num_labels) v->votes[label[q]] += 1;
    }
}

int64_t oracle_map_size(const ovoxel_map* m) { return m->size; }

static int ovoxel_cmp(const void* a, const void* b)
{
    const ovoxel *p = (const ovoxel*)a, *q = (const ovoxel*)b;
    if (p->k != q->k) return p->k < q->k ? -1 : 1;
    if (p->j != q->j) return p->j < q->j ? -1 : 1;
    if (p->i != q->i) return p->i < q->i ? -1 : 1;
    return 0;
}

int64_t oracle_map_export(const ovoxel_map* m, int32_t* ijk, float* centroid, double* centroid_d, uint32_t* rgba,
                          uint32_t* count, uint32_t* votes, uint8_t* label)
{
    ovoxel* s = (ovoxel*)malloc(sizeof(ovoxel) * (size_t)(m->size > 0 ? m->size : 1));
    memcpy(s, m->v, sizeof(ovoxel) * (size_t)m->size);
    qsort(s, (size_t)m->size, sizeof(ovoxel), ovoxel_cmp);
    const int L = m->num_labels;
    for (int64_t q = 0; q < m->size; ++q) {
        const ovoxel* v = &s[q];
        const float fn = (float)v->n;
        if (ijk) {
            ijk[3 * q] = v->i;
            ijk[3 * q + 1] = v->j;
            ijk[3 * q + 2] = v->k;
        }
        if (centroid) {
            centroid[3 * q] = v->sx / fn;
            centroid[3 * q + 1] = v->sy / fn;
            centroid[3 * q + 2] = v->sz / fn;
        }
        if (centroid_d) {
            centroid_d[3 * q] = v->dx / (double)v->n;
            centroid_d[3 * q + 1] = v->dy / (double)v->n;
            centroid_d[3 * q + 2] = v->dz / (double)v->n;
        }
        if (rgba) {
            const float r = v->sr / fn, g = v->sg / fn, b = v->sb / fn;
            rgba[q] = ((uint32_t)(int)r << 16) | ((uint32_t)(int)g << 8) | (uint32_t)(int)b;
        }
        if (count) count[q] = v->n;
        uint32_t bestv = 0;
        int bestl = 255;
        for (int l = 0; l < L; ++l) {
            if (votes) votes[(size_t)q * L + l] = v->votes[l];
            if (v->votes[l] > bestv) {
                bestv = v->votes[l];
                bestl = l;
            }
        }
        if (label) label[q] = (uint8_t)bestl;
    }
    free(s);
    return m->size;
}

/* ================================================================================================
 * Dense motion cues: the other dense consumer of the disparity map, on the tracker thread
 * (src/track.cpp:67-79 -> UVDisparity::Process, src/uvdisparity.cpp:842-903).
 * ================================================================================================ */

/* cvRound: round half to even (lrint under the default rounding mode) */
static inline int ocv_round(double v) { return (int)nearbyint(v); }

/* src/stereo.cpp:41-118.  xyz: [H][W][10] fp32 = X, Y, Z, u, v, disparity, intensity, I_u, I_v, motion mark.
 * Both branches of the ROI test (:88-113) store the same values, so roi does not influence the result. */
void oracle_triangulate10d(const uint8_t* img, const int16_t* disp, int W, int H, double f, double cx, double cy, double b,
                           float* xyz)
{
    double min_disp = DBL_MAX; /* :58-59 cv::minMaxIdx */
    for (size_t i = 0; i < (size_t)W * H; ++i)
        if ((double)disp[i] < min_disp) min_disp = (double)disp[i];
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const short d = disp[(size_t)i * W + j];
            const double pw = b / (1.0 * (double)d);               /* :78  (d == 0 -> +-inf) */
            double px = (((double)j - cx) * pw) * 16.0;            /* :79  16.0f promotes to double */
            double py = (((double)i - cy) * pw) * 16.0;            /* :80 */
            double pz = (f * pw) * 16.0;                           /* :81 */
            if (fabs((double)d - min_disp) <= (double)FLT_EPSILON) /* :83-88 missing values -> +inf */
                px = py = pz = (double)INFINITY;
            float* o = xyz + ((size_t)i * W + j) * 10;
            o[0] = (float)px; o[1] = (float)py; o[2] = (float)pz;
            o[3] = (float)j; o[4] = (float)i;
            o[5] = (float)d / 16.0f;                               /* :95 */
            o[6] = (float)(int)img[(size_t)i * W + j];             /* :96 */
            o[7] = 0.f; o[8] = 0.f; o[9] = 0.f;
        }
}

/* src/stereo.cpp:127-181: rotate Y/Z by the first pitch angle for 0 < round(disparity) < 100, clear the intensity
 * channel outside the ROI (and for every other disparity).  pitch2 is accepted and unused, as in the reference. */
void oracle_correct_3d_points(float* xyz, int W, int H, double roi_x, double roi_y, double roi_z, double pitch1, double pitch2)
{
    (void)pitch2;
    const double cos_p1 = cos(pitch1), sin_p1 = sin(pitch1);
    for (size_t k = 0; k < (size_t)W * H; ++k) {
        float* o = xyz + k * 10;
        const float yp = o[1], zp = o[2];
        const int d = ocv_round((double)o[5]);                     /* :146 cvRound(float) */
        if (d > 0 && d < 100) {                                    /* :148 and :161 are the same body */
            o[1] = (float)(cos_p1 * (double)yp + sin_p1 * (double)zp);
            o[2] = (float)(cos_p1 * (double)zp - sin_p1 * (double)yp);
            if ((double)o[0] > roi_x || (double)o[1] > roi_y || (double)o[2] > roi_z) o[6] = 0.f;
        } else {
            o[6] = 0.f;                                            /* :174 */
        }
    }
}

/* src/stereo.cpp:183-192: roi_mask = convertScaleAbs(channel 6) = saturate_cast<uchar>(round(|intensity|)) */
void oracle_set_image_roi(const float* xyz, int W, int H, uint8_t* roi_mask)
{
    for (size_t k = 0; k < (size_t)W * H; ++k) {
        const int r = ocv_round((double)fabsf(xyz[k * 10 + 6]));
        roi_mask[k] = (uint8_t)(r > 255 ? 255 : r);
    }
}

static int disp_max_ceil(const int16_t* disp, int W, int H, double* max_dis_out)
{
    int mx = -32768;
    for (size_t i = 0; i < (size_t)W * H; ++i)
        if (disp[i] > mx) mx = disp[i];
    const double max_dis = (double)mx / 16;                        /* uvdisparity.cpp:197-199 / 279-282 */
    if (max_dis_out) *max_dis_out = max_dis;
    const int c = (int)ceil(max_dis);
    return c < 0 ? 0 : c;                                          /* canonical: a negative size is an empty map */
}

/* UVDisparity::calVDisparity, src/uvdisparity.cpp:277-366.  Outputs: v_dis_int [H][v_cols] int32, v_dis [H][v_cols] u8
 * (both sized by the caller for v_cols <= cap_cols), xyz channel 8.  Returns v_cols = cvCeil(max(disp)/16).
 * Canonical choices for the reference's out-of-bounds accesses (both buffers are continuous cv::Mat allocations):
 *   - bin id = min(v_cols, round(d/16)) can equal v_cols (:309-311): the increment lands on the flat element
 *     i * v_cols + v_cols, i.e. bin 0 of the next row; past the end of the matrix it is dropped;
 *   - v_dis_.at<int>(v, d) on the 8-bit map (:352) reads the 4 bytes at flat byte offset v * v_cols + 4 * d as a
 *     little-endian int; bytes past the end of the map read as 0;
 *   - uchar = int * float (:331) is truncation toward zero, then modulo 256 (x86 cvttss2si + byte store). */
int oracle_v_disparity(const int16_t* disp, int W, int H, float* xyz, int32_t* v_dis_int, uint8_t* v_dis, int cap_cols)
{
    const int v_cols = disp_max_ceil(disp, W, H, NULL);
    if (v_cols > cap_cols) return -1;
    const size_t n = (size_t)H * v_cols;
    memset(v_dis_int, 0, n * sizeof(int32_t));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const short d = disp[(size_t)i * W + j];
            if (d > 0) {                                           /* :304 (Inf / NaN tests on a short are vacuous) */
                const int dis = ocv_round((double)((float)d / 16.0f)); /* :306 */
                const int id = imax(0, imin(v_cols, dis));         /* :307 */
                const size_t flat = (size_t)i * v_cols + id;
                if (flat < n) v_dis_int[flat]++;
            }
        }
    const float scale = 255 * 1.0f / (float)W;                     /* :319 xyz.cols */
    for (size_t k = 0; k < n; ++k) v_dis[k] = (uint8_t)(int)((float)v_dis_int[k] * scale);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float* o = xyz + ((size_t)i * W + j) * 10;
            const int v = ocv_round((double)o[4]), d = ocv_round((double)o[5]);   /* :347-348 */
            float out = 0.f;
            if (d > 0) {
                uint32_t word = 0;
                for (int q = 0; q < 4; ++q) {
                    const size_t off = (size_t)v * v_cols + 4 * (size_t)d + q;
                    if (off < n) word |= (uint32_t)v_dis[off] << (8 * q);
                }
                out = (float)(int32_t)word;                        /* :352-353 */
            }
            o[8] = out;
        }
    return v_cols;
}

/* UVDisparity::calUDisparity, src/uvdisparity.cpp:195-274.  Outputs: u_dis_int [u_rows][W] int32, u_dis [u_rows][W] u8
 * (u_rows <= cap_rows), xyz channel 7.  Returns u_rows = cvCeil(max(disp)/16) + 1.
 * Canonical choice: u_dis_.at<uchar>(d, u) with d < 0 (invalid pixels carry disparity -1, :267-269) reads 0. */
int oracle_u_disparity(const int16_t* disp, int W, int H, float* xyz, const uint8_t* roi_mask, const uint8_t* ground_mask,
                       int32_t* u_dis_int, uint8_t* u_dis, int cap_rows)
{
    const int u_rows = disp_max_ceil(disp, W, H, NULL) + 1;
    if (u_rows > cap_rows) return -1;
    const size_t n = (size_t)u_rows * W;
    memset(u_dis_int, 0, n * sizeof(int32_t));
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const short d = disp[(size_t)i * W + j];
            if (d > 0) {
                const int dis = d / 16;                            /* :219 cvRound(d/16): integer division first */
                if (roi_mask[(size_t)i * W + j] > 0 && ground_mask[(size_t)i * W + j] > 0 && dis > 0)
                    u_dis_int[(size_t)dis * W + j]++;
            }
        }
    const float scale = 255 * 1.0f / (float)H;                     /* :235 xyz.rows */
    for (size_t k = 0; k < n; ++k) u_dis[k] = (uint8_t)(int)((float)u_dis_int[k] * scale);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float* o = xyz + ((size_t)i * W + j) * 10;
            const int u = ocv_round((double)o[3]), d = ocv_round((double)o[5]);   /* :265-266 */
            o[7] = (d >= 0 && d < u_rows && u >= 0 && u < W) ? (float)u_dis[(size_t)d * W + u] : 0.f;
        }
    return u_rows;
}
