// Iterative pyramidal Lucas-Kanade, ONE WARP PER FEATURE (sm_100a) -- the throughput / low-latency form of vf_lk.cu.
//
// Same arithmetic as vf_lk.cu, bit for bit (OpenCV's LKTrackerInvoker, restated in oracle/vf_oracle.c): W_BITS = 14 bilinear
// patches, Scharr derivatives from the staged uint8 tile, float sums of A and b in OpenCV's SIMD accumulation order, the
// exact-integer shortcut for b (a chain whose terms satisfy sum|t| < 2^24 equals its integer sum, whatever the order).
// What changes is the work decomposition (north_star: "one warp per feature, the 21x21 bilinear patch staged in shared
// memory and warp-shuffle reductions of the 2x2 normal equations"):
//   * a warp owns a feature from the first level to the last: no block barrier anywhere, only __syncwarp between the
//     phases of a level, and nothing of the scalar solve / exit logic is replicated across four warps;
//   * the 441 window pixels are 63 row segments of 7 pixels, two per lane: a lane fetches its 8 + 8 tap bytes of J as
//     three aligned 32-bit words per row, realigns them with funnel shifts and interpolates with IDP.2A (two taps per
//     instruction, signed 16-bit weights x unsigned 8-bit pixels) -- 6 shared-memory loads and 14 dot products per
//     segment instead of 28 byte loads and 28 multiplies;
//   * the source patch (I, Ix, Iy of its 14 pixels) stays in the lane's registers for the whole level; b and sum|b| are
//     reduced with REDUX, every lane solves the 2x2 step itself;
//   * levels are processed one at a time in ~10 KB of shared memory per warp (the 4-warp kernel keeps all levels of a
//     pass resident: 54 KB per feature), so 16 features are resident per SM instead of 4; the next level's tile of I and a
//     speculative tile of J are fetched into registers while the current level iterates.
// Used for every launch unless vf_debug_set_lk_kernel says otherwise; vf_lk.cu remains as the A/B reference.
#include "vf_camera.cuh"
#include "vf_common.cuh"
#include "vf_kernels.h"
#include "vf_lk_common.cuh"

namespace vf {

namespace {

constexpr int kSegPx = 7;                    // pixels per row segment
constexpr int kSegs = 63;                    // 21 rows x 3 segments; lane l owns segments l and l + 32
constexpr int kPx = 2 * kSegPx;              // pixels per lane
constexpr int kTJ = kTileJ * kTileJ;

struct alignas(16) WarpLk {                  // shared memory of one warp
    uint8_t tileJ[2][kTJ + 16];              // staged target tile, double-buffered (+16: the last row's word fetch reads one word on)
    uint8_t tileI[kTileI * kTileI + 16];     // 24x24 source pixels around the window
    short2 dtile[kTileD * kTileD];           // Scharr (Ix, Iy) at the 22x22 tap positions
    short2 dI[kWinPx + 3];                   // interpolated derivative patch (the ordered chains of A read it)
    int termX[kWinPx + 3], termY[kWinPx + 3];   // per-pixel terms of b, spilled only when the exactness bound fails
};

__device__ __forceinline__ int dp2a_lo(int a, uint32_t b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi(int a, uint32_t b, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// Eight + one tap bytes starting at byte offset `a` of a row in shared memory, as words of bytes 0..3, 4..7 (r) and
// 1..4, 5..8 (s): pixel k of the segment reads taps k, k + 1 = the low / high half of r[k / 4] (k even) or s[(k - 1) / 4] (k odd).
struct TapWords { uint32_t r0, r1, s0, s1; };
__device__ __forceinline__ TapWords tap_words(const uint8_t* base, int a) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(base) + (a >> 2);
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const int m = 8 * (a & 3);
    TapWords t;
    t.r0 = __funnelshift_r(w0, w1, m); t.r1 = __funnelshift_r(w1, w2, m);
    t.s0 = __funnelshift_rc(w0, w1, m + 8); t.s1 = __funnelshift_rc(w1, w2, m + 8);
    return t;
}
// bilinear value of segment pixel K from the tap words of its two rows: sum + rnd, not yet shifted
template <int K>
__device__ __forceinline__ int interp_px(const TapWords& t0, const TapWords& t1, int w_top, int w_bot, int rnd) {
    constexpr bool odd = (K & 1) != 0;
    constexpr int word = (odd ? K - 1 : K) / 4, hi = ((odd ? K - 1 : K) & 2) != 0;
    const uint32_t a = odd ? (word ? t0.s1 : t0.s0) : (word ? t0.r1 : t0.r0);
    const uint32_t b = odd ? (word ? t1.s1 : t1.s0) : (word ? t1.r1 : t1.r0);
    return hi ? dp2a_hi(w_top, a, dp2a_hi(w_bot, b, rnd)) : dp2a_lo(w_top, a, dp2a_lo(w_bot, b, rnd));
}

struct WarpPass {   // result of one calcOpticalFlowPyrLK for one point + its statistics
    float x, y;
    int status, iters, fast;
};

// One calcOpticalFlowPyrLK for one point, executed by one warp (all lanes pass identical arguments and return identical
// results).  pI / pJ: source and target pyramids.
__device__ __noinline__ WarpPass lk_track_point_warp(WarpLk& sm, const LevelRef* __restrict__ pI, const LevelRef* __restrict__ pJ,
                                                     float2 prev_pt, float2 init_pt, bool use_init, int max_level) {
    const int lane = threadIdx.x & 31;
    const float half = 10.f;
    const float FLT_SCALE = 1.f / (1 << 20);
    int status = 1, iters = 0, fast = 0;
    float outx = init_pt.x, outy = init_pt.y;

    // the lane's two row segments: window row, first column
    int seg_e[2], seg_oj[2], seg_oi[2], seg_od[2];
    bool seg_on[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int s = lane + 32 * k;
        seg_on[k] = s < kSegs;
        const int r = seg_on[k] ? s / 3 : 0, c0 = seg_on[k] ? kSegPx * (s - 3 * (s / 3)) : 0;
        seg_e[k] = r * kWin + c0;                       // first pixel of the segment in the 21x21 patch
        seg_oj[k] = r * kTileJ + c0;                    // ... in the J tile, relative to the window origin
        seg_oi[k] = (r + 1) * kTileI + c0 + 1;          // ... in the I tile
        seg_od[k] = r * kTileD + c0;                    // ... in the derivative tile
    }

    // ---- top level: its tile of I (lanes 0..23: one row each) and the J tile around the start point go out first ----
    uint32_t pfI[kTileI / 4 + 1];
    bool pfI_valid;
    {
        const LevelRef LI = pI[max_level];
        const float scale = pow2_neg(max_level);
        const int ipx = __float2int_rd(__fsub_rn(__fmul_rn(prev_pt.x, scale), half)), ipy = __float2int_rd(__fsub_rn(__fmul_rn(prev_pt.y, scale), half));
        pfI_valid = !(ipx < -kWin || ipx >= LI.w || ipy < -kWin || ipy >= LI.h);
        if (pfI_valid && lane < kTileI) issue_row_words<kTileI / 4>(LI, ipx - 1, ipy - 1 + lane, pfI);
    }
    int jx0, jy0, jbuf = 0;
    __syncwarp();                                        // a previous pass of this warp is done with the shared buffers
    {
        const LevelRef LT = pJ[max_level];
        const float scale = pow2_neg(max_level);
        const float sx = use_init ? init_pt.x : prev_pt.x, sy = use_init ? init_pt.y : prev_pt.y;
        jx0 = clamp_tile_origin(__float2int_rd(__fsub_rn(__fmul_rn(sx, scale), half)) - kJMargin, LT.w);
        jy0 = clamp_tile_origin(__float2int_rd(__fsub_rn(__fmul_rn(sy, scale), half)) - kJMargin, LT.h);
        uint32_t v[kTileJ / 4];
        load_row_words<kTileJ / 4>(LT, jx0, jy0 + lane, v);
        uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[0] + lane * kTileJ);
        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
        dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
    bool tile_ready = true;

#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        const LevelRef LI = pI[level], LJ = pJ[level];
        const int jcols = LJ.w, jrows = LJ.h;
        const float scale = pow2_neg(level);
        const float ppx = __fsub_rn(__fmul_rn(prev_pt.x, scale), half), ppy = __fsub_rn(__fmul_rn(prev_pt.y, scale), half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        const bool valid = pfI_valid;                    // the window intersects the padded level (lkpyramid.cpp bounds test)
        float nx, ny;
        if (level == max_level) {
            if (use_init) { nx = __fmul_rn(outx, scale); ny = __fmul_rn(outy, scale); }
            else { nx = __fmul_rn(prev_pt.x, scale); ny = __fmul_rn(prev_pt.y, scale); }
        } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;

        // ---- P1: publish this level's tile of I; issue the next level's ----
        __syncwarp();                                    // the previous level / pass is done with the shared buffers
        if (valid && lane < kTileI) {
            const int sh = 8 * ((ipx - 1) & 3);
            uint32_t* dst = reinterpret_cast<uint32_t*>(sm.tileI + lane * kTileI);
#pragma unroll
            for (int k = 0; k < kTileI / 4; ++k) dst[k] = __funnelshift_r(pfI[k], pfI[k + 1], sh);
        }
        if (level > 0) {
            const LevelRef LN = pI[level - 1];
            const float sc = pow2_neg(level - 1);
            const int qx = __float2int_rd(__fsub_rn(__fmul_rn(prev_pt.x, sc), half)), qy = __float2int_rd(__fsub_rn(__fmul_rn(prev_pt.y, sc), half));
            pfI_valid = !(qx < -kWin || qx >= LN.w || qy < -kWin || qy >= LN.h);
            if (pfI_valid && lane < kTileI) issue_row_words<kTileI / 4>(LN, qx - 1, qy - 1 + lane, pfI);
        }
        if (!valid) {
            if (level == 0) status = 0;
            tile_ready = false;
            continue;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);
        __syncwarp();

        // ---- P2: Scharr (Ix, Iy) at the 22x22 tap positions, zero outside the image: lane = tile column, rolling down the rows ----
        if (lane < kTileD) {
            const uint8_t* col = sm.tileI + lane;                       // tile columns lane, lane + 1, lane + 2
            const int x = ipx + lane;
            const bool in_x = (unsigned)x < (unsigned)LI.w;
            int a0 = col[0], a1 = col[1], a2 = col[2];
            int b0 = col[kTileI], b1 = col[kTileI + 1], b2 = col[kTileI + 2];
#pragma unroll 2
            for (int ty = 0; ty < kTileD; ++ty) {
                const uint8_t* p = col + (ty + 2) * kTileI;
                const int c0 = p[0], c1 = p[1], c2 = p[2];
                const bool inside = in_x && (unsigned)(ipy + ty) < (unsigned)LI.h;
                const int t0m = (a0 + c0) * 3 + b0 * 10, t0p = (a2 + c2) * 3 + b2 * 10;
                const int t1m = c0 - a0, t1c = c1 - a1, t1p = c2 - a2;
                sm.dtile[ty * kTileD + lane] = make_short2((short)(inside ? t0p - t0m : 0), (short)(inside ? (t1m + t1p) * 3 + t1c * 10 : 0));
                a0 = b0; a1 = b1; a2 = b2; b0 = c0; b1 = c1; b2 = c2;
            }
        }
        __syncwarp();

        // ---- P3: the lane's 14 pixels of the interpolated patches: I (2^5 fixed point) and (Ix, Iy) stay in registers for
        //      the level, (Ix, Iy) also goes to shared memory for the chains of A ----
        int Iw[kPx], gx[kPx], gy[kPx];
        {
            const int w_top = (w00 & 0xffff) | (w01 << 16), w_bot = (w10 & 0xffff) | (w11 << 16);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const TapWords t0 = tap_words(sm.tileI, seg_oi[k]), t1 = tap_words(sm.tileI, seg_oi[k] + kTileI);
                int iv[kSegPx];
                iv[0] = interp_px<0>(t0, t1, w_top, w_bot, 256); iv[1] = interp_px<1>(t0, t1, w_top, w_bot, 256);
                iv[2] = interp_px<2>(t0, t1, w_top, w_bot, 256); iv[3] = interp_px<3>(t0, t1, w_top, w_bot, 256);
                iv[4] = interp_px<4>(t0, t1, w_top, w_bot, 256); iv[5] = interp_px<5>(t0, t1, w_top, w_bot, 256);
                iv[6] = interp_px<6>(t0, t1, w_top, w_bot, 256);
                const short2* d0 = sm.dtile + seg_od[k];
                short2 top[kSegPx + 1], bot[kSegPx + 1];
#pragma unroll
                for (int q = 0; q <= kSegPx; ++q) { top[q] = d0[q]; bot[q] = d0[kTileD + q]; }
#pragma unroll
                for (int q = 0; q < kSegPx; ++q) {
                    const int ixval = VF_DESCALE(top[q].x * w00 + top[q + 1].x * w01 + bot[q].x * w10 + bot[q + 1].x * w11, 14);
                    const int iyval = VF_DESCALE(top[q].y * w00 + top[q + 1].y * w01 + bot[q].y * w10 + bot[q + 1].y * w11, 14);
                    const bool on = seg_on[k];
                    Iw[kSegPx * k + q] = on ? iv[q] >> 9 : 0;
                    gx[kSegPx * k + q] = on ? ixval : 0;
                    gy[kSegPx * k + q] = on ? iyval : 0;
                    if (on) sm.dI[seg_e[k] + q] = make_short2((short)ixval, (short)iyval);
                }
            }
        }
        __syncwarp();

        // ---- P4: A = sums of (Ix Ix, Ix Iy, Iy Iy): 15 ordered float chains, one LANE per chain, in OpenCV's accumulation
        //      order -- SIMD lane c = x & 3 sees columns c, 4 + c, 8 + c, 12 + c of every row in turn (the fifth step of such a
        //      chain adds +0, an exact identity), the scalar tail sees columns 16..20 of every row ----
        float A11, A12, A22, D;
        int skip;
        {
            const int ch = min(lane, 14), m = ch / 5, c = ch - 5 * m;
            const int xstep = c < 4 ? 4 : 1;
            const short2* dIp = sm.dI + (c < 4 ? c : 16);
            float acc = 0.f;
#pragma unroll 3
            for (int y = 0; y < kWin; ++y) {
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const short2 d = dIp[y * kWin + t * xstep];               // columns c + 16 <= 19 and 16 + 4 = 20: in range
                    const int fa = m < 2 ? d.x : d.y, fb = m == 0 ? d.x : d.y;
                    acc = __fadd_rn(acc, (c < 4 && t == 4) ? 0.f : (float)(fa * fb));
                }
            }
            float r[3];
#pragma unroll
            for (int mm = 0; mm < 3; ++mm) {                                   // iA += v_reduce_sum(qA): tail + ((q0 + q2) + (q1 + q3))
                const float q0 = __shfl_sync(0xffffffffu, acc, 5 * mm + 0), q1 = __shfl_sync(0xffffffffu, acc, 5 * mm + 1);
                const float q2 = __shfl_sync(0xffffffffu, acc, 5 * mm + 2), q3 = __shfl_sync(0xffffffffu, acc, 5 * mm + 3);
                const float tl = __shfl_sync(0xffffffffu, acc, 5 * mm + 4);
                r[mm] = __fadd_rn(tl, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));
            }
            A11 = __fmul_rn(r[0], FLT_SCALE); A12 = __fmul_rn(r[1], FLT_SCALE); A22 = __fmul_rn(r[2], FLT_SCALE);
            const float det = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
            const float dA = __fsub_rn(A11, A22);
            const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
            const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * kWin * kWin));
            skip = (minEig < 1e-4f || det < FLT_EPSILON) ? 1 : 0;
            D = __fdiv_rn(1.f, det);
        }
        if (skip) {   // minEig / det below threshold (lkpyramid.cpp)
            if (level == 0) status = 0;
            tile_ready = false;
            continue;
        }
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);

        // ---- J tile of this level: the prefetched one if it covers the window, else staged now ----
        {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (!tile_ready || inx < jx0 || iny < jy0 || inx + kTileD > jx0 + kTileJ || iny + kTileD > jy0 + kTileJ) {
                jx0 = clamp_tile_origin(inx - kJMargin, jcols); jy0 = clamp_tile_origin(iny - kJMargin, jrows);
                jbuf ^= 1;
                uint32_t v[kTileJ / 4];
                load_row_words<kTileJ / 4>(LJ, jx0, jy0 + lane, v);
                uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[jbuf] + lane * kTileJ);
                dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                __syncwarp();
            }
        }
        // speculative tile of the next level around the current estimate: the words stay in registers while this level iterates
        uint32_t pfJ[kTileJ / 4 + 1];
        int px0 = 0, py0 = 0;
        if (level > 0) {
            const LevelRef LN = pJ[level - 1];
            px0 = clamp_tile_origin(__float2int_rd(__fadd_rn(__fmul_rn(nx, 2.f), half)) - kJMargin, LN.w);   // 2 (nx + half) - half
            py0 = clamp_tile_origin(__float2int_rd(__fadd_rn(__fmul_rn(ny, 2.f), half)) - kJMargin, LN.h);
            issue_row_words<kTileJ / 4>(LN, px0, py0 + lane, pfJ);
        }

        const uint8_t* tj = sm.tileJ[jbuf];
        float pdx = 0.f, pdy = 0.f;
        // window origins for which an iteration needs no special handling: inside the level's search range (lkpyramid.cpp
        // bounds test) AND covered by the staged tile
        int lox = max(-kWin, jx0), hix = min(jcols - 1, jx0 + kTileJ - kTileD);
        int loy = max(-kWin, jy0), hiy = min(jrows - 1, jy0 + kTileJ - kTileD);
#pragma unroll 1
        for (int j = 0;; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < lox || inx > hix || iny < loy || iny > hiy) {
                if (inx < -kWin || inx >= jcols || iny < -kWin || iny >= jrows) {
                    if (level == 0) status = 0;
                    break;
                }
                jx0 = inx - kJMargin; jy0 = iny - kJMargin;      // the window left the staged tile: re-centre it (uniform branch)
                __syncwarp();
                {
                    uint32_t v[kTileJ / 4];
                    load_row_words<kTileJ / 4>(LJ, jx0, jy0 + lane, v);
                    uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[jbuf] + lane * kTileJ);
                    dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                }
                __syncwarp();
                lox = max(-kWin, jx0); hix = min(jcols - 1, jx0 + kTileJ - kTileD);
                loy = max(-kWin, jy0); hiy = min(jrows - 1, jy0 + kTileJ - kTileD);
            }
            ++iters;
            int v00, v01, v10, v11;
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), v00, v01, v10, v11);
            const int w_top = (v00 & 0xffff) | (v01 << 16), w_bot = (v10 & 0xffff) | (v11 << 16);
            const int base = (iny - jy0) * kTileJ + (inx - jx0);
            // the lane's 14 differences J - I and its share of b = sum (J - I) (Ix, Iy), exact in int32
            int diff[kPx];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const TapWords t0 = tap_words(tj, base + seg_oj[k]), t1 = tap_words(tj, base + seg_oj[k] + kTileJ);
                diff[kSegPx * k + 0] = (interp_px<0>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 0];
                diff[kSegPx * k + 1] = (interp_px<1>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 1];
                diff[kSegPx * k + 2] = (interp_px<2>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 2];
                diff[kSegPx * k + 3] = (interp_px<3>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 3];
                diff[kSegPx * k + 4] = (interp_px<4>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 4];
                diff[kSegPx * k + 5] = (interp_px<5>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 5];
                diff[kSegPx * k + 6] = (interp_px<6>(t0, t1, w_top, w_bot, 256) >> 9) - Iw[kSegPx * k + 6];
            }
            int tx = 0, ty = 0;
            unsigned int abx = 0, aby = 0;                      // <= 14 * 8160 * 4080 < 2^29 per lane
#pragma unroll
            for (int q = 0; q < kPx; ++q) {
                const int ex = diff[q] * gx[q], ey = diff[q] * gy[q];
                tx += ex; ty += ey;
                abx += (unsigned int)abs(ex); aby += (unsigned int)abs(ey);
            }
            const int bx = __reduce_add_sync(0xffffffffu, tx), by = __reduce_add_sync(0xffffffffu, ty);
            // b1 only adds x-terms, b2 only y-terms: each has its own exactness bound (lane sums clamped: the total stays below 2^32)
            const unsigned int bnd = __reduce_add_sync(0xffffffffu, min(max(abx, aby), (unsigned int)kExact));
            float ib1, ib2;
            if (bnd < (unsigned int)kExact) {
                // every partial sum of every chain, in any order, is an exactly representable integer
                ib1 = (float)bx; ib2 = (float)by;
                ++fast;
            } else {
                // spill the terms and run OpenCV's 10 ordered float chains, one lane per chain: chains 2c + comp (c = 0..3) take the
                // pmaddwd pair sums of columns (8g + c, 8g + c + 4), g = 0, 1, row by row; chains 8 / 9 the scalar tail (columns 16..20)
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (seg_on[k]) {
#pragma unroll
                        for (int q = 0; q < kSegPx; ++q) {
                            sm.termX[seg_e[k] + q] = diff[kSegPx * k + q] * gx[kSegPx * k + q];
                            sm.termY[seg_e[k] + q] = diff[kSegPx * k + q] * gy[kSegPx * k + q];
                        }
                    }
                __syncwarp();
                float acc = 0.f;
                {
                    const int q = min(lane, 9);
                    const int* T = (q & 1) ? sm.termY : sm.termX;
                    if (q < 8) {
                        const int c = q >> 1;
#pragma unroll 3
                        for (int y = 0; y < kWin; ++y) {
                            acc = __fadd_rn(acc, (float)(T[y * kWin + c] + T[y * kWin + c + 4]));            // v_cvt_f32 of the pair sum
                            acc = __fadd_rn(acc, (float)(T[y * kWin + 8 + c] + T[y * kWin + 12 + c]));
                        }
                    } else {
#pragma unroll 3
                        for (int y = 0; y < kWin; ++y)
#pragma unroll
                            for (int x = 16; x < kWin; ++x) acc = __fadd_rn(acc, (float)T[y * kWin + x]);
                    }
                }
                // qb0 + qb1 -> (s0, s1, s2, s3); ib1 += (s0 + 0) + (s2 + 0); ib2 += (s1 + 0) + (s3 + 0)
                const float k0x = __shfl_sync(0xffffffffu, acc, 0), k0y = __shfl_sync(0xffffffffu, acc, 1);
                const float k0z = __shfl_sync(0xffffffffu, acc, 2), k0w = __shfl_sync(0xffffffffu, acc, 3);
                const float k1x = __shfl_sync(0xffffffffu, acc, 4), k1y = __shfl_sync(0xffffffffu, acc, 5);
                const float k1z = __shfl_sync(0xffffffffu, acc, 6), k1w = __shfl_sync(0xffffffffu, acc, 7);
                const float k2x = __shfl_sync(0xffffffffu, acc, 8), k2y = __shfl_sync(0xffffffffu, acc, 9);
                const float s0 = __fadd_rn(k0x, k1x), s1 = __fadd_rn(k0y, k1y), s2 = __fadd_rn(k0z, k1z), s3 = __fadd_rn(k0w, k1w);
                ib1 = __fadd_rn(k2x, __fadd_rn(s0, s2));
                ib2 = __fadd_rn(k2y, __fadd_rn(s1, s3));
            }
            const float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            // The three ways out of the loop (lkpyramid.cpp, in this order), folded into ONE branch on the hot path:
            //   delta.ddot(delta) <= eps^2 (in double; a float estimate settles all but borderline cases);
            //   j > 0 and |delta + prevDelta| < 0.01 on both axes (std::abs on a float widened to double:
            //       < 0.01  <=>  <= 0.01f, the largest float below 0.01) -> step back by delta / 2;
            //   maxCount = 30 iterations.
            const float ff = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const bool osc = j > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f;
            if (ff < 1.0001e-4f || osc || j == 29) {
                bool conv = ff < 0.9999e-4f;
                if (!conv && ff < 1.0001e-4f)
                    conv = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= 0.01 * 0.01;
                if (conv) break;
                if (osc) {
                    outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                    break;
                }
                if (j == 29) break;
            }
            pdx = dx; pdy = dy;
        }
        tile_ready = false;
        if (level > 0) {   // publish the prefetched tile of the next level
            jbuf ^= 1;
            const int sh = 8 * (px0 & 3);
            uint32_t v[kTileJ / 4];
#pragma unroll
            for (int k = 0; k < kTileJ / 4; ++k) v[k] = __funnelshift_r(pfJ[k], pfJ[k + 1], sh);
            uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[jbuf] + lane * kTileJ);
            dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
            dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
            jx0 = px0; jy0 = py0;
            tile_ready = true;
        }
        if (status && level == 0) {   // the caller passes an err vector: final window test (lkpyramid.cpp)
            const int ix = __float2int_rd(__fsub_rn(outx, half)), iy = __float2int_rd(__fsub_rn(outy, half));
            if (ix < -kWin || ix >= jcols || iy < -kWin || iy >= jrows) status = 0;
        }
    }
    WarpPass res;
    res.x = outx; res.y = outy; res.status = status; res.iters = iters; res.fast = fast;
    return res;
}

extern __shared__ __align__(16) unsigned char lkw_smem_raw[];

// pyramids of stream s into shared memory (threads 0 .. 2 kMaxLevels - 1 of the CTA), then one block barrier
__device__ __forceinline__ void load_pyrs(LevelRef (*pyr)[kMaxLevels], const PyrView& a, const PyrView& c, const size_t* stride, int s) {
    const int t = threadIdx.x;
    if (t < 2 * kMaxLevels) {
        const PyrView& pv = t < kMaxLevels ? a : c;
        const int l = t & (kMaxLevels - 1);
        LevelRef r;
        const bool ok = l < pv.n_levels;
        r.p = ok ? pv.lvl[l] + (stride ? (size_t)s * stride[l] : 0) : nullptr;
        r.w = ok ? pv.w[l] : 0; r.h = ok ? pv.h[l] : 0; r.pitch = ok ? pv.pitch[l] : 0;
        pyr[t / kMaxLevels][l] = r;
    }
    __syncthreads();
}

// ---- temporal: prev -> cur (maxLevel L) then cur -> prev (maxLevel 1, initial flow), feature_tracker.cpp:37-67 ----
__global__ void __launch_bounds__(128) lk_temporal_warp_kernel(const __grid_constant__ DevBatch b, int parity_arg, int cur, int prev) {
    pdl_wait(); pdl_release();
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;
    __shared__ LevelRef pyr[2][kMaxLevels];
    TraceScope trace(b.trace, TR_LK_TEMPORAL + 16 * tphase);
    const int s = blockIdx.y, wid = threadIdx.x >> 5, i = blockIdx.x * (blockDim.x >> 5) + wid, cap = b.cap;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    if (blockIdx.x * (blockDim.x >> 5) >= n_prev) return;       // whole CTA beyond the feature count
    load_pyrs(pyr, b.left[prev], b.left[cur], b.lvl_stream_stride, s);   // 0 = previous left, 1 = current left
    if (i >= n_prev || i >= cap) return;
    WarpLk& sm = reinterpret_cast<WarpLk*>(lkw_smem_raw)[wid];
    const float2 p0 = b.kps[parity ^ 1][(size_t)s * cap + i];
    const int L = min(b.lk_max_level, b.left[cur].n_levels - 1);
    const WarpPass f = lk_track_point_warp(sm, pyr[0], pyr[1], p0, p0, false, L);
    const float2 fwd = make_float2(f.x, f.y);
    float2 rev = p0;
    int st = f.status, rst = 0, iters = f.iters, fast = f.fast;
    int ok = st;
    if (b.flow_back) {
        // the reference also runs the reverse pass on points whose forward status is already 0 and then discards the
        // result (:55-59); the device skips that work, the outcome is the same
        if (st) {
            const WarpPass r = lk_track_point_warp(sm, pyr[1], pyr[0], fwd, p0, true, min(1, L));
            rev = make_float2(r.x, r.y); rst = r.status; iters += r.iters; fast += r.fast;
        }
        ok = st && rst && pt_distance(p0, rev) <= 0.5;
    }
    if (ok && !in_border(fwd, b.H, b.W)) ok = 0;
    if ((threadIdx.x & 31) == 0) {
        const size_t o = (size_t)s * cap + i;
        b.fwd_pts[o] = fwd; b.fwd_st[o] = (uint8_t)st;
        b.rev_pts[o] = rev; b.rev_st[o] = (uint8_t)rst;
        b.temporal_st[o] = (uint8_t)ok;
        b.lk_iters[o] = iters | (fast << 16);   // low 16 bits: iterations, high: of which on the exact-integer path
    }
}

// ---- stereo: left -> right then right -> left (both maxLevel L), status (:112-135), then RemoveDistortion +
//      ComputeKpsVelocity for the right camera and the right half of the featureFrame record ----
__global__ void __launch_bounds__(128) lk_stereo_warp_kernel(const __grid_constant__ DevBatch b, int parity_arg, int cur, const double* __restrict__ dt_host) {
    pdl_wait(); pdl_release();
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;
    __shared__ LevelRef pyr[2][kMaxLevels];
    TraceScope trace(b.trace, TR_LK_STEREO + 16 * tphase);
    const int s = blockIdx.y, wid = threadIdx.x >> 5, i = blockIdx.x * (blockDim.x >> 5) + wid, cap = b.cap;
    const int n = b.counters[parity][s * C_COUNT + C_N_TOTAL];
    if (blockIdx.x * (blockDim.x >> 5) >= n) return;
    load_pyrs(pyr, b.left[cur], b.right[parity], b.lvl_stream_stride, s);   // 0 = current left, 1 = current right
    if (i >= n || i >= cap) return;
    WarpLk& sm = reinterpret_cast<WarpLk*>(lkw_smem_raw)[wid];
    const size_t o = (size_t)s * cap + i;
    const float2 p0 = b.kps[parity][o];
    const int pidx = b.prev_idx[parity][o];
    // previous_right_un_distortion_ids_kps_map_.find(id): select_tracked recorded where the feature sat in the previous
    // frame's arrays (-1: new corner); it is in the right map iff it had a right observation there
    const int found = (pidx >= 0 && b.right_ok[parity ^ 1][(size_t)s * cap + pidx]) ? pidx : -1;
    const int L = min(b.lk_max_level, b.right[parity].n_levels - 1);
    const WarpPass f = lk_track_point_warp(sm, pyr[0], pyr[1], p0, p0, false, L);
    const float2 rp = make_float2(f.x, f.y);
    float2 back = p0;
    int st = f.status, rst = 0, iters = f.iters, fast = f.fast;
    int ok = st;
    if (b.flow_back) {
        if (st) {
            const WarpPass r = lk_track_point_warp(sm, pyr[1], pyr[0], rp, rp, false, L);
            back = make_float2(r.x, r.y); rst = r.status; iters += r.iters; fast += r.fast;
        }
        ok = st && rst && in_border(rp, b.H, b.W) && pt_distance(p0, back) <= 0.5;
    }
    if ((threadIdx.x & 31) == 0) {
        vf_feature* rec = &b.out[tphase % 3][o];     // mapped pinned host memory (ring of three frames)
        rec->has_right = ok;
        float2 unr = make_float2(0.f, 0.f), vr = make_float2(0.f, 0.f);
        if (ok) {
            unr = undistort_point(b.cam[1], rp);
            if (found >= 0) vr = velocity(unr, b.un_right[parity ^ 1][(size_t)s * cap + found], dt_host[s]);
        }
        rec->xr = unr.x; rec->yr = unr.y; rec->ur = ok ? rp.x : 0.f; rec->vr = ok ? rp.y : 0.f; rec->vxr = vr.x; rec->vyr = vr.y;
        b.un_right[parity][o] = unr;
        b.right_ok[parity][o] = (uint8_t)ok;
        b.lr_pts[o] = rp; b.lr_st[o] = (uint8_t)st; b.rl_pts[o] = back; b.rl_st[o] = (uint8_t)rst;
        b.lk_iters_stereo[o] = iters | (fast << 16);
        if (ok) atomicAdd(&b.counters[parity][s * C_COUNT + C_N_RIGHT], 1);
    }
}

// ---- stand-alone calcOpticalFlowPyrLK (vf_op_lk) ----
__global__ void __launch_bounds__(128) lk_op_warp_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts,
                                                         float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                         int* __restrict__ iters_out, int n, int max_level, int use_init) {
    __shared__ LevelRef pyr[2][kMaxLevels];
    const int wid = threadIdx.x >> 5, i = blockIdx.x * (blockDim.x >> 5) + wid;
    load_pyrs(pyr, prev, next, nullptr, 0);
    if (i >= n) return;
    WarpLk& sm = reinterpret_cast<WarpLk*>(lkw_smem_raw)[wid];
    const int L = min(max_level, min(prev.n_levels, next.n_levels) - 1);
    const WarpPass r = lk_track_point_warp(sm, pyr[0], pyr[1], prev_pts[i], next_pts[i], use_init != 0, L);
    if ((threadIdx.x & 31) == 0) {
        next_pts[i] = make_float2(r.x, r.y); status[i] = (uint8_t)r.status;
        if (iters_out) iters_out[i] = r.iters;
    }
}

// Warps (= features) per CTA: one for a launch that does not fill the GPU (every feature on its own SM: 150 features on 148
// SMs), four once there are more features than that (fewer CTAs to schedule, the pyramid descriptors shared by four warps).
int warps_per_cta(long long n_features) { return n_features <= 2 * kNumSms ? 1 : 4; }

}  // namespace

static int g_lk_mode = 0;   // 0 = one warp per feature (default), 1 = the 4-warp-per-feature kernels of vf_lk.cu
void set_lk_kernel_mode(int mode) { g_lk_mode = mode; }
bool lk_use_warp_kernel() { return g_lk_mode != 1; }

cudaError_t lk_warp_prepare() {
    cudaError_t e;
    const int bytes = 4 * (int)sizeof(WarpLk);
    if ((e = cudaFuncSetAttribute(lk_temporal_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lk_stereo_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lk_op_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lk_temporal_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lk_stereo_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(lk_op_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout);
}

void launch_lk_temporal_warp(const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream) {
    const int wpc = warps_per_cta((long long)h.cap * h.n_streams);
    dim3 grid((h.cap + wpc - 1) / wpc, h.n_streams);
    launch_kernel(lk_temporal_warp_kernel, grid, dim3(32 * wpc), sizeof(WarpLk) * wpc, stream, h, parity, cur, prev);
}
void launch_lk_stereo_warp(const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream) {
    const int wpc = warps_per_cta((long long)h.cap * h.n_streams);
    dim3 grid((h.cap + wpc - 1) / wpc, h.n_streams);
    launch_kernel(lk_stereo_warp_kernel, grid, dim3(32 * wpc), sizeof(WarpLk) * wpc, stream, h, parity, cur, dt_host);
}
void launch_lk_op_warp(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status, int* iters, int n,
                       int max_level, int use_initial_flow, cudaStream_t stream) {
    if (n <= 0) return;
    const int wpc = warps_per_cta(n);
    lk_op_warp_kernel<<<(n + wpc - 1) / wpc, 32 * wpc, sizeof(WarpLk) * wpc, stream>>>(prev, next, prev_pts, next_pts, status, iters, n, max_level,
                                                                                   use_initial_flow ? 1 : 0);
}

}  // namespace vf
