// Iterative pyramidal Lucas-Kanade with WARP-LEVEL passes (sm_100a) -- the low-latency and the throughput form of vf_lk.cu.
//
// Same arithmetic as vf_lk.cu, bit for bit (OpenCV's LKTrackerInvoker, restated in oracle/vf_oracle.c): W_BITS = 14 bilinear
// patches, Scharr derivatives from the staged uint8 tile, float sums of A and b in OpenCV's SIMD accumulation order, the
// exact-integer shortcut for b (a chain whose terms satisfy sum|t| < 2^24 equals its integer sum, whatever the order).
// What changes is the work decomposition (north_star: "one warp per feature, the 21x21 bilinear patch staged in shared
// memory and warp-shuffle reductions of the 2x2 normal equations"):
//   * the Gauss-Newton iterations of a feature are run by ONE warp: no block barrier in the loop, nothing of the scalar
//     solve / exit logic replicated across warps.  The 441 window pixels are 63 row segments of 7 pixels, two per lane: a
//     lane fetches its 8 + 8 tap bytes of J as three aligned 32-bit words per row, realigns them with funnel shifts and
//     interpolates with IDP.2A (two taps per instruction, signed 16-bit weights x unsigned 8-bit pixels) -- 6 shared-memory
//     loads and 14 dot products per segment instead of 28 byte loads and 28 multiplies; the lane's 14 source pixels
//     (I, Ix, Iy) stay in registers for the level; b and sum|b| are reduced with REDUX and every lane solves the 2x2 step;
//   * the per-level prologue (tile of I, Scharr, interpolated patches, the 15 ordered float chains of A over terms staged in
//     accumulation order) is warp-level code as well, and there are two ways to schedule it:
//       - BATCH (many features per SM: 64-stream batches, 4K frames): one warp per feature does everything, level after
//         level, in 14 KB of shared memory (the 4-warp kernel of vf_lk.cu keeps all levels resident: 54 KB per feature), so
//         three to four times as many features are resident per SM; the next level's tile of I and a speculative tile of J
//         are fetched into registers while the current level iterates;
//       - COOP (a launch that does not fill the GPU: one stream, 150 features on 148 SMs): a CTA of four warps per feature;
//         the prologues of ALL levels of a pass run side by side, one warp per level, then ONE block barrier, then warp 0
//         walks the levels and iterates.  The critical path of a pass is one prologue + the iterations.
// vf_lk.cu (one 4-warp CTA per feature, block barrier per iteration) remains as the A/B reference (vf_debug_set_lk_kernel).
#include "vf_camera.cuh"
#include "vf_common.cuh"
#include "vf_kernels.h"
#include "vf_lk_common.cuh"

// Optional phase timing (build with -DVF_LKW_TIMING, python -m vins_fast_b200.build --timing): lane 0 accumulates clock64 deltas
// into vf_lkw_timing[k] (sum) and [16 + k] (count); read and cleared by vf_debug_lkw_timing.
#ifdef VF_LKW_TIMING
__device__ unsigned long long vf_lkw_timing[32];
#define LKW_T0(v) const long long v = clock64()
#define LKW_ADD(k, v) do { if ((threadIdx.x & 31) == 0) { atomicAdd(&vf_lkw_timing[k], (unsigned long long)(clock64() - (v))); atomicAdd(&vf_lkw_timing[16 + (k)], 1ull); } } while (0)
#define LKW_CNT(k, n) do { if ((threadIdx.x & 31) == 0) atomicAdd(&vf_lkw_timing[k], (unsigned long long)(n)); } while (0)
#else
#define LKW_T0(v) do { } while (0)
#define LKW_ADD(k, v) do { } while (0)
#define LKW_CNT(k, n) do { } while (0)
#endif

namespace vf {

namespace {

constexpr int kSegPx = 7;                    // pixels per row segment
constexpr int kSegs = 63;                    // 21 rows x 3 segments; lane l owns segments l and l + 32
constexpr int kPx = 2 * kSegPx;              // pixels per lane
constexpr int kTJ = kTileJ * kTileJ;
constexpr int kChains = 15;                  // 3 matrix entries x (4 SIMD-lane accumulators + 1 scalar tail)
constexpr int kProd = kWinPx + 3;            // floats per product plane (row-major 21x21)
constexpr int kPP = 24;                      // pitch of the weighted-pixel plane P (23 x 23 values)
constexpr int kCoopWarps = 4;
#ifndef VF_LKW_BATCH_WARPS
#define VF_LKW_BATCH_WARPS 4                // BATCH: warps (= features) per CTA
#endif
#ifdef VF_LKW_MAXNREG                        // A/B aid: cap the registers directly (e.g. 144 with 2-warp CTAs: 14 warps per SM)
#define VF_LKW_BOUNDS(COOP) __maxnreg__(VF_LKW_MAXNREG)
#else
#define VF_LKW_BOUNDS(COOP) __launch_bounds__(COOP ? 128 : 32 * VF_LKW_BATCH_WARPS, COOP ? 1 : VF_LKW_MIN_BLOCKS)
#endif
#ifndef VF_LKW_MIN_BLOCKS
#define VF_LKW_MIN_BLOCKS 3                  // BATCH: resident 4-warp CTAs per SM the register budget is sized for
#endif

struct alignas(16) LkLevel {                 // what the prologue of one level leaves for the iterations
    uint8_t tileI[kTileI * kTileI + 16];     // 24x24 source pixels around the window (+16: word fetches of the last row)
    union {
        short2 dtile[kTileD * kTileD];       // generic path: Scharr (Ix, Iy) at the 22x22 tap positions
        int P[23 * kPP];                     // fast path: bilinear-weighted pixels P(u, v), u, v = 0..22
    };
    short2 dI[kWinPx + 3];                   // interpolated derivative patch
    short Iw[kWinPx + 7];                    // interpolated I, 2^5 fixed point
    float A11, A12, A22, Dinv;               // spatial gradient matrix (x 2^-20) and 1 / det
    int valid, skip, pad0, pad1;             // window intersects the padded level; minEig / det below threshold
};
struct alignas(16) LkWork {                  // scratch of one warp
    uint8_t tileJ[2][kTJ + 16];              // staged target tile, double-buffered (+16: the last row's word fetch reads one word on)
    float prod[3 * kProd];                   // Ix Ix, Ix Iy, Iy Iy of the derivative patch, row-major (prologue); the iterations
                                             // reuse the space for the per-pixel terms of b when the exactness bound fails
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
__device__ __forceinline__ void interp_segment(const uint8_t* tile, int off, int pitch, int w_top, int w_bot, int (&v)[kSegPx]) {
    const TapWords t0 = tap_words(tile, off), t1 = tap_words(tile, off + pitch);
    v[0] = interp_px<0>(t0, t1, w_top, w_bot, 256) >> 9; v[1] = interp_px<1>(t0, t1, w_top, w_bot, 256) >> 9;
    v[2] = interp_px<2>(t0, t1, w_top, w_bot, 256) >> 9; v[3] = interp_px<3>(t0, t1, w_top, w_bot, 256) >> 9;
    v[4] = interp_px<4>(t0, t1, w_top, w_bot, 256) >> 9; v[5] = interp_px<5>(t0, t1, w_top, w_bot, 256) >> 9;
    v[6] = interp_px<6>(t0, t1, w_top, w_bot, 256) >> 9;
}

// the lane's two row segments of the 21x21 window
struct Segs {
    int e[2];      // first pixel of the segment in the patch (row * 21 + column)
    int oj[2];     // ... in the J tile, relative to the window origin
    bool on[2];    // lane 31 owns one segment only
};
__device__ __forceinline__ Segs make_segs(int lane) {
    Segs g;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int s = lane + 32 * k;
        g.on[k] = s < kSegs;
        const int r = g.on[k] ? s / 3 : 0, c0 = g.on[k] ? kSegPx * (s - 3 * (s / 3)) : 0;
        g.e[k] = r * kWin + c0;
        g.oj[k] = r * kTileJ + c0;
    }
    return g;
}

// window origin of a point on a level: floor(pt * 2^-level - half) and the bounds test of lkpyramid.cpp
__device__ __forceinline__ bool level_origin(const LevelRef& L, float2 pt, int level, float& ppx, float& ppy, int& ipx, int& ipy) {
    const float scale = pow2_neg(level);
    ppx = __fsub_rn(__fmul_rn(pt.x, scale), 10.f); ppy = __fsub_rn(__fmul_rn(pt.y, scale), 10.f);
    ipx = __float2int_rd(ppx); ipy = __float2int_rd(ppy);
    return !(ipx < -kWin || ipx >= L.w || ipy < -kWin || ipy >= L.h);
}

// One row of the I tile as issue_row_words left it (lane = row), handed to the prologue BY VALUE: an array passed by address
// would live in local memory, and the store that puts it there waits for the loads -- which defeats the prefetch.
struct RawRow {
    uint32_t w[kTileI / 4 + 1];
    int have;                                  // 0: the prologue fetches the tile itself
};

// ---- prologue of one level, one warp: everything that depends on the source point only ----
__device__ __noinline__ void lk_prologue(LkLevel& B, LkWork& wk, const LevelRef& LI, float2 prev_pt, int level, const RawRow raw) {
    const int lane = threadIdx.x & 31;
    const float FLT_SCALE = 1.f / (1 << 20);
    LKW_T0(t_pro);
    float ppx, ppy;
    int ipx, ipy;
    const bool valid = level_origin(LI, prev_pt, level, ppx, ppy, ipx, ipy);
    if (!valid) {
        if (lane == 0) { B.valid = 0; B.skip = 1; }
        __syncwarp();
        return;
    }
    // ---- P1: 24x24 tile of I at (ipx - 1, ipy - 1), lane = row ----
    if (lane < kTileI) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(B.tileI + lane * kTileI);
        if (raw.have) {
            const int sh = 8 * ((ipx - 1) & 3);
#pragma unroll
            for (int k = 0; k < kTileI / 4; ++k) dst[k] = __funnelshift_r(raw.w[k], raw.w[k + 1], sh);
        } else {
            uint32_t v[kTileI / 4];
            load_row_words<kTileI / 4>(LI, ipx - 1, ipy - 1 + lane, v);
#pragma unroll
            for (int k = 0; k < kTileI / 4; ++k) dst[k] = v[k];
        }
    }
    int w00, w01, w10, w11;
    bilinear_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);
    const int w_top = (w00 & 0xffff) | (w01 << 16), w_bot = (w10 & 0xffff) | (w11 << 16);
    __syncwarp();
    LKW_ADD(1, t_pro);                                   // tile of I in shared memory (includes the wait for the global loads)
    // the three product planes of a pixel (exact in float: |Ix|, |Iy| <= 4080)
    auto put = [&](int e, int ival, int ixval, int iyval) {
        B.Iw[e] = (short)ival;
        B.dI[e] = make_short2((short)ixval, (short)iyval);
        wk.prod[e] = (float)(ixval * ixval);
        wk.prod[kProd + e] = (float)(ixval * iyval);
        wk.prod[2 * kProd + e] = (float)(iyval * iyval);
    };
    if (ipx >= 0 && ipx + kTileD <= LI.w && ipy >= 0 && ipy + kTileD <= LI.h) {
        // ---- fast path: every derivative tap lies inside the image.  Interpolation and the Scharr stencil are both integer
        //      linear maps of the pixels, so they commute exactly: first the bilinear-weighted pixels
        //          P(u, v) = w00 I(u, v) + w01 I(u+1, v) + w10 I(u, v+1) + w11 I(u+1, v+1)        (tile coordinates, 23 x 23)
        //      with IDP.2A from aligned words (69 row segments of 8 values), then per window pixel
        //          sum w Ix(tap) = [3 (P(x+2, y) + P(x+2, y+2)) + 10 P(x+2, y+1)] - [the same at x]
        //          sum w Iy(tap) = 3 [(P(x, y+2) - P(x, y)) + (P(x+2, y+2) - P(x+2, y))] + 10 (P(x+1, y+2) - P(x+1, y))
        //          sum w I(tap)  = P(x+1, y+1)
        //      -- the same integers the tap-by-tap form produces, without ever forming the 484 derivative taps ----
#pragma unroll 1
        for (int task = lane; task < 69; task += 32) {
            const int v = task / 3, g = task - 3 * v;
            const uint32_t* r0 = reinterpret_cast<const uint32_t*>(B.tileI + v * kTileI + 8 * g);   // 8-byte aligned
            const uint32_t* r1 = r0 + kTileI / 4;
            const uint32_t a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2];
            const TapWords t0 = {a0, a1, __funnelshift_r(a0, a1, 8), __funnelshift_r(a1, a2, 8)};
            const TapWords t1 = {b0, b1, __funnelshift_r(b0, b1, 8), __funnelshift_r(b1, b2, 8)};
            int4* dst = reinterpret_cast<int4*>(B.P + v * kPP + 8 * g);
            dst[0] = make_int4(interp_px<0>(t0, t1, w_top, w_bot, 0), interp_px<1>(t0, t1, w_top, w_bot, 0),
                               interp_px<2>(t0, t1, w_top, w_bot, 0), interp_px<3>(t0, t1, w_top, w_bot, 0));
            dst[1] = make_int4(interp_px<4>(t0, t1, w_top, w_bot, 0), interp_px<5>(t0, t1, w_top, w_bot, 0),
                               interp_px<6>(t0, t1, w_top, w_bot, 0), interp_px<7>(t0, t1, w_top, w_bot, 0));
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int s = lane + 32 * k;
            if (s >= kSegs) break;
            const int row = s / 3, c0 = kSegPx * (s - 3 * row);
            const int* p0 = B.P + row * kPP + c0;
            int V[kSegPx + 2], Dv[kSegPx + 2], mid[kSegPx + 2];
#pragma unroll
            for (int q = 0; q < kSegPx + 2; ++q) {
                const int a = p0[q], b = p0[kPP + q], c = p0[2 * kPP + q];
                V[q] = (a + c) * 3 + b * 10; Dv[q] = c - a; mid[q] = b;
            }
#pragma unroll
            for (int q = 0; q < kSegPx; ++q)
                put(row * kWin + c0 + q, VF_DESCALE(mid[q + 1], 9), VF_DESCALE(V[q + 2] - V[q], 14),
                    VF_DESCALE((Dv[q] + Dv[q + 2]) * 3 + Dv[q + 1] * 10, 14));
        }
    } else {
        // ---- generic path (the window reaches over the image border, where OpenCV's derivative image is zero):
        //      Scharr (Ix, Iy) at the 22x22 tap positions, 484 positions over the 32 lanes, then tap-by-tap interpolation ----
#pragma unroll 4
        for (int r = 0; r < (kTileD * kTileD + 31) / 32; ++r) {
            const int e = min(lane + 32 * r, kTileD * kTileD - 1);
            const int ty = (e * 2979) >> 16, tx = e - kTileD * ty;            // e / 22 for e < 484
            const uint8_t* p = B.tileI + (ty + 1) * kTileI + (tx + 1);
            const int a0 = p[-kTileI - 1], a1 = p[-kTileI], a2 = p[-kTileI + 1], b0 = p[-1], b2 = p[1];
            const int c0 = p[kTileI - 1], c1 = p[kTileI], c2 = p[kTileI + 1];
            const bool inside = (unsigned)(ipx + tx) < (unsigned)LI.w && (unsigned)(ipy + ty) < (unsigned)LI.h;
            const int t0m = (a0 + c0) * 3 + b0 * 10, t0p = (a2 + c2) * 3 + b2 * 10;
            const int t1m = c0 - a0, t1c = c1 - a1, t1p = c2 - a2;
            B.dtile[e] = make_short2((short)(inside ? t0p - t0m : 0), (short)(inside ? (t1m + t1p) * 3 + t1c * 10 : 0));
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int s = lane + 32 * k;
            if (s >= kSegs) break;
            const int row = s / 3, c0 = kSegPx * (s - 3 * row);
            int iv[kSegPx];
            interp_segment(B.tileI, (row + 1) * kTileI + c0 + 1, kTileI, w_top, w_bot, iv);
            const short2* d0 = B.dtile + row * kTileD + c0;
            short2 top[kSegPx + 1], bot[kSegPx + 1];
#pragma unroll
            for (int q = 0; q <= kSegPx; ++q) { top[q] = d0[q]; bot[q] = d0[kTileD + q]; }
#pragma unroll
            for (int q = 0; q < kSegPx; ++q)
                put(row * kWin + c0 + q, iv[q], VF_DESCALE(top[q].x * w00 + top[q + 1].x * w01 + bot[q].x * w10 + bot[q + 1].x * w11, 14),
                    VF_DESCALE(top[q].y * w00 + top[q + 1].y * w01 + bot[q].y * w10 + bot[q + 1].y * w11, 14));
        }
    }
    __syncwarp();
    LKW_ADD(2, t_pro);                                   // ... + patches
    // ---- A = sums of (Ix Ix, Ix Iy, Iy Iy): 15 ordered float chains, one LANE per chain (acc = ((0 + t0) + t1) + ...), in
    //      OpenCV's accumulation order -- SIMD lane c = x & 3 sees columns c, 4 + c, 8 + c, 12 + c of every row in turn (the fifth
    //      step of such a chain adds +0, an exact identity), the scalar tail sees columns 16..20 of every row ----
    {
        const int ch = min(lane, kChains - 1), m = ch / 5, c = ch - 5 * m;
        const int xstep = c < 4 ? 4 : 1;
        const float* src = wk.prod + m * kProd + (c < 4 ? c : 16);
        float acc = 0.f;
#pragma unroll 7
        for (int y = 0; y < kWin; ++y) {
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const float v = src[y * kWin + t * xstep];                     // columns c + 16 <= 19 and 16 + 4 = 20: in range
                acc = __fadd_rn(acc, (c < 4 && t == 4) ? 0.f : v);
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
        if (lane == 0) {
            const float A11 = __fmul_rn(r[0], FLT_SCALE), A12 = __fmul_rn(r[1], FLT_SCALE), A22 = __fmul_rn(r[2], FLT_SCALE);
            const float det = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
            const float dA = __fsub_rn(A11, A22);
            const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
            const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * kWin * kWin));
            B.A11 = A11; B.A12 = A12; B.A22 = A22; B.Dinv = __fdiv_rn(1.f, det);
            B.valid = 1;
            B.skip = (minEig < 1e-4f || det < FLT_EPSILON) ? 1 : 0;
        }
    }
    __syncwarp();
    LKW_ADD(0, t_pro);                                   // whole prologue
}

struct WarpPass {   // result of one calcOpticalFlowPyrLK for one point + its statistics
    float x, y;
    int status, iters, fast;
};

// state of the iterating warp across the levels of a pass
struct IterState {
    float outx, outy;
    int status, iters, fast;
    int jx0, jy0, jbuf;
    bool tile_ready;
};

// one row of the 32x32 J tile per lane, around (x0, y0)
__device__ __forceinline__ void stage_tile_j(uint8_t* tile, const LevelRef& L, int x0, int y0) {
    const int lane = threadIdx.x & 31;
    uint32_t v[kTileJ / 4];
    load_row_words<kTileJ / 4>(L, x0, y0 + lane, v);
    uint4* dst = reinterpret_cast<uint4*>(tile + lane * kTileJ);
    dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
    dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// Cold path of an iteration (the exactness bound failed): OpenCV's 10 ordered float chains over the spilled terms, one lane per
// chain.  Chains 2c + comp (c = 0..3) take the pmaddwd pair sums of columns (8g + c, 8g + c + 4), g = 0, 1, row by row; chains
// 8 / 9 the scalar tail (columns 16..20).  Returns (ib1, ib2) in every lane.  Out of line on purpose.
__device__ __noinline__ float2 lkw_ordered_chains(const int* termX, const int* termY) {
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    {
        const int q = min(lane, 9);
        const int* T = (q & 1) ? termY : termX;
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
    return make_float2(__fadd_rn(k2x, __fadd_rn(s0, s2)), __fadd_rn(k2y, __fadd_rn(s1, s3)));
}
// Cold path of an iteration: the window left the staged tile, re-stage it around the new origin.  Out of line.
__device__ __noinline__ void lkw_restage_tile(uint8_t* tile, const LevelRef& L, int x0, int y0) {
    __syncwarp();
    stage_tile_j(tile, L, x0, y0);
    __syncwarp();
}

// ---- the iterations of one level, one warp (all lanes hold identical scalars) ----
// LN: the next (finer) level of the target pyramid, for the speculative tile fetch; unused at level 0.
__device__ __forceinline__ void lk_iterate(IterState& st, const LkLevel& B, LkWork& wk, const Segs& sg, const LevelRef& LJ, const LevelRef& LN,
                                           float2 prev_pt, bool use_init, int level, int max_level) {
    const int lane = threadIdx.x & 31;
    const float half = 10.f;
    const float FLT_SCALE = 1.f / (1 << 20);
    LKW_T0(t_it);
    const int jcols = LJ.w, jrows = LJ.h;
    const float scale = pow2_neg(level);
    float nx, ny;
    if (level == max_level) {
        if (use_init) { nx = __fmul_rn(st.outx, scale); ny = __fmul_rn(st.outy, scale); }
        else { nx = __fmul_rn(prev_pt.x, scale); ny = __fmul_rn(prev_pt.y, scale); }
    } else { nx = __fmul_rn(st.outx, 2.f); ny = __fmul_rn(st.outy, 2.f); }
    st.outx = nx; st.outy = ny;
    const int4 flags = *reinterpret_cast<const int4*>(&B.valid);
    if (!flags.x || flags.y) {   // window outside the level, or minEig / det below threshold (lkpyramid.cpp)
        if (level == 0) st.status = 0;
        st.tile_ready = false;
        return;
    }
    const float4 gA = *reinterpret_cast<const float4*>(&B.A11);
    const float A11 = gA.x, A12 = gA.y, A22 = gA.z, D = gA.w;
    nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
    // the lane's 14 source pixels stay in registers for the whole level
    int Iw[kPx], gx[kPx], gy[kPx];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int q = 0; q < kSegPx; ++q) {
            const short2 d = B.dI[sg.e[k] + q];
            Iw[kSegPx * k + q] = sg.on[k] ? (int)B.Iw[sg.e[k] + q] : 0;
            gx[kSegPx * k + q] = sg.on[k] ? (int)d.x : 0;
            gy[kSegPx * k + q] = sg.on[k] ? (int)d.y : 0;
        }
    // ---- J tile of this level: the prefetched one if it covers the window, else staged now ----
    {
        const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
        if (!st.tile_ready || inx < st.jx0 || iny < st.jy0 || inx + kTileD > st.jx0 + kTileJ || iny + kTileD > st.jy0 + kTileJ) {
            st.jx0 = clamp_tile_origin(inx - kJMargin, jcols); st.jy0 = clamp_tile_origin(iny - kJMargin, jrows);
            st.jbuf ^= 1;
            LKW_T0(t_stage);
            stage_tile_j(wk.tileJ[st.jbuf], LJ, st.jx0, st.jy0);
            __syncwarp();
            LKW_ADD(6, t_stage);                         // J tile staged on demand (prefetch missed / none)
        }
    }
    // speculative tile of the next level around the current estimate: the words stay in registers while this level iterates
    uint32_t pfJ[kTileJ / 4 + 1];
    int px0 = 0, py0 = 0;
    if (level > 0) {
        px0 = clamp_tile_origin(__float2int_rd(__fadd_rn(__fmul_rn(nx, 2.f), half)) - kJMargin, LN.w);   // 2 (nx + half) - half
        py0 = clamp_tile_origin(__float2int_rd(__fadd_rn(__fmul_rn(ny, 2.f), half)) - kJMargin, LN.h);
        issue_row_words<kTileJ / 4>(LN, px0, py0 + lane, pfJ);
    }
    int* termX = reinterpret_cast<int*>(wk.prod);
    int* termY = termX + kWinPx + 3;
    const uint8_t* tj = wk.tileJ[st.jbuf];
    int jx0 = st.jx0, jy0 = st.jy0, status = st.status, iters = 0, fast = 0;
    float outx = st.outx, outy = st.outy;
    float pdx = 0.f, pdy = 0.f;
    // window origins for which an iteration needs no special handling: inside the level's search range (lkpyramid.cpp
    // bounds test) AND covered by the staged tile
    int lox = max(-kWin, jx0), hix = min(jcols - 1, jx0 + kTileJ - kTileD);
    int loy = max(-kWin, jy0), hiy = min(jrows - 1, jy0 + kTileJ - kTileD);
    LKW_ADD(4, t_it);                                    // level set-up: pixels into registers, J tile check, prefetch issue
    LKW_T0(t_loop);
#pragma unroll 1
    for (int j = 0;; ++j) {
        const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
        if (inx < lox || inx > hix || iny < loy || iny > hiy) {
            if (inx < -kWin || inx >= jcols || iny < -kWin || iny >= jrows) {
                if (level == 0) status = 0;
                break;
            }
            jx0 = inx - kJMargin; jy0 = iny - kJMargin;          // the window left the staged tile: re-centre it (uniform branch)
            lkw_restage_tile(wk.tileJ[st.jbuf], LJ, jx0, jy0);
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
            int jv[kSegPx];
            interp_segment(tj, base + sg.oj[k], kTileJ, w_top, w_bot, jv);
#pragma unroll
            for (int q = 0; q < kSegPx; ++q) diff[kSegPx * k + q] = jv[q] - Iw[kSegPx * k + q];
        }
        int tx = 0, ty = 0;
        unsigned int abx = 0, aby = 0;                          // <= 14 * 8160 * 4080 < 2^29 per lane
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
                if (sg.on[k]) {
#pragma unroll
                    for (int q = 0; q < kSegPx; ++q) {
                        termX[sg.e[k] + q] = diff[kSegPx * k + q] * gx[kSegPx * k + q];
                        termY[sg.e[k] + q] = diff[kSegPx * k + q] * gy[kSegPx * k + q];
                    }
                }
            __syncwarp();
            const float2 ib = lkw_ordered_chains(termX, termY);      // cold, out of line (keeps the loop compact in the L0 I-cache)
            ib1 = ib.x; ib2 = ib.y;
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
    LKW_ADD(5, t_loop);                                  // the iteration loop of the level
    LKW_CNT(7, iters); LKW_CNT(8, fast);
    st.tile_ready = false;
    if (level > 0) {   // publish the prefetched tile of the next level
        st.jbuf ^= 1;
        const int sh = 8 * (px0 & 3);
        uint32_t v[kTileJ / 4];
#pragma unroll
        for (int k = 0; k < kTileJ / 4; ++k) v[k] = __funnelshift_r(pfJ[k], pfJ[k + 1], sh);
        uint4* dst = reinterpret_cast<uint4*>(wk.tileJ[st.jbuf] + lane * kTileJ);
        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
        dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
        jx0 = px0; jy0 = py0;
        st.tile_ready = true;
        __syncwarp();
    }
    if (status && level == 0) {   // the caller passes an err vector: final window test (lkpyramid.cpp)
        const int ix = __float2int_rd(__fsub_rn(outx, half)), iy = __float2int_rd(__fsub_rn(outy, half));
        if (ix < -kWin || ix >= jcols || iy < -kWin || iy >= jrows) status = 0;
    }
    LKW_ADD(3, t_it);                                    // whole level
    st.outx = outx; st.outy = outy; st.status = status; st.iters += iters; st.fast += fast; st.jx0 = jx0; st.jy0 = jy0;
}

__device__ __forceinline__ void iter_begin(IterState& st, LkWork& wk, const LevelRef& LT, float2 prev_pt, float2 init_pt, bool use_init, int max_level) {
    st.outx = init_pt.x; st.outy = init_pt.y; st.status = 1; st.iters = 0; st.fast = 0; st.jbuf = 0;
    const float scale = pow2_neg(max_level);
    const float sx = use_init ? init_pt.x : prev_pt.x, sy = use_init ? init_pt.y : prev_pt.y;
    st.jx0 = clamp_tile_origin(__float2int_rd(__fsub_rn(__fmul_rn(sx, scale), 10.f)) - kJMargin, LT.w);
    st.jy0 = clamp_tile_origin(__float2int_rd(__fsub_rn(__fmul_rn(sy, scale), 10.f)) - kJMargin, LT.h);
    stage_tile_j(wk.tileJ[0], LT, st.jx0, st.jy0);        // the top level's J tile only depends on the start point
    st.tile_ready = true;
}

// One calcOpticalFlowPyrLK for one point by ONE warp, level after level (BATCH scheduling).  All lanes pass identical
// arguments and return identical results.
__device__ __noinline__ WarpPass lk_pass_warp(LkLevel& B, LkWork& wk, const LevelRef* __restrict__ pI, const LevelRef* __restrict__ pJ,
                                              float2 prev_pt, float2 init_pt, bool use_init, int max_level) {
    const int lane = threadIdx.x & 31;
    const Segs sg = make_segs(lane);
    __syncwarp();                                          // a previous pass of this warp is done with the shared buffers
    // the top level's tile of I goes out first, then the J tile around the start point
    RawRow raw;
    raw.have = 1;
    {
        float ppx, ppy; int ipx, ipy;
        if (level_origin(pI[max_level], prev_pt, max_level, ppx, ppy, ipx, ipy) && lane < kTileI)
            issue_row_words<kTileI / 4>(pI[max_level], ipx - 1, ipy - 1 + lane, raw.w);
    }
    IterState st;
    iter_begin(st, wk, pJ[max_level], prev_pt, init_pt, use_init, max_level);
#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        __syncwarp();
        lk_prologue(B, wk, pI[level], prev_pt, level, raw);          // consumes `raw`
        // the next level's tile of I goes out now: its words travel while this level iterates (issued behind the prologue
        // call, not in front of it -- values that are live across a call would have to be complete to be saved)
        if (level > 0) {
            float ppx, ppy; int ipx, ipy;
            if (level_origin(pI[level - 1], prev_pt, level - 1, ppx, ppy, ipx, ipy) && lane < kTileI)
                issue_row_words<kTileI / 4>(pI[level - 1], ipx - 1, ipy - 1 + lane, raw.w);
        }
        lk_iterate(st, B, wk, sg, pJ[level], pJ[level > 0 ? level - 1 : 0], prev_pt, use_init, level, max_level);
    }
    WarpPass res;
    res.x = st.outx; res.y = st.outy; res.status = st.status; res.iters = st.iters; res.fast = st.fast;
    return res;
}

// One calcOpticalFlowPyrLK for one point by a CTA of kCoopWarps warps (COOP scheduling): the prologues of all levels side by
// side, one block barrier, then warp 0 iterates through the levels.  blk: [max_level + 1] level blocks, wk: [kCoopWarps].
// Every thread returns the result (through shared memory).
__device__ __noinline__ WarpPass lk_pass_coop(LkLevel* blk, LkWork* wk, WarpPass* result, const LevelRef* __restrict__ pI, const LevelRef* __restrict__ pJ,
                                              float2 prev_pt, float2 init_pt, bool use_init, int max_level) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();                                       // the previous pass is done with the shared buffers (and with *result)
    LKW_T0(t_pass);
    IterState st;
    if (wid == 0) iter_begin(st, wk[0], pJ[max_level], prev_pt, init_pt, use_init, max_level);
    // level l -> warp (max_level - l) % kCoopWarps: the top level, which the iterations need first, is warp 0's own
    for (int level = max_level - wid; level >= 0; level -= kCoopWarps) { RawRow none; none.have = 0; lk_prologue(blk[level], wk[wid], pI[level], prev_pt, level, none); }
    __syncthreads();
    if (wid == 0) LKW_ADD(9, t_pass);                      // COOP: start of the pass -> all prologues done
    if (wid == 0) {
        const Segs sg = make_segs(lane);
#pragma unroll 1
        for (int level = max_level; level >= 0; --level)
            lk_iterate(st, blk[level], wk[0], sg, pJ[level], pJ[level > 0 ? level - 1 : 0], prev_pt, use_init, level, max_level);
        if (lane == 0) { result->x = st.outx; result->y = st.outy; result->status = st.status; result->iters = st.iters; result->fast = st.fast; }
    }
    __syncthreads();
    if (wid == 0) LKW_ADD(10, t_pass);                     // COOP: whole pass
    return *result;
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

// The two schedulings behind one call: COOP = the CTA works on feature blockIdx.x, BATCH = warp w of the CTA on its own.
template <bool COOP>
struct Pass {
    LkLevel* blk; LkWork* wk; WarpPass* result;
    __device__ __forceinline__ Pass(int n_levels, WarpPass* res) : result(res) {
        if (COOP) {
            blk = reinterpret_cast<LkLevel*>(lkw_smem_raw);
            wk = reinterpret_cast<LkWork*>(lkw_smem_raw + sizeof(LkLevel) * n_levels);
        } else {
            unsigned char* mine = lkw_smem_raw + (sizeof(LkLevel) + sizeof(LkWork)) * (threadIdx.x >> 5);
            blk = reinterpret_cast<LkLevel*>(mine);
            wk = reinterpret_cast<LkWork*>(mine + sizeof(LkLevel));
        }
    }
    __device__ __forceinline__ WarpPass run(const LevelRef* pI, const LevelRef* pJ, float2 prev_pt, float2 init_pt, bool use_init, int L) const {
        return COOP ? lk_pass_coop(blk, wk, result, pI, pJ, prev_pt, init_pt, use_init, L) : lk_pass_warp(*blk, *wk, pI, pJ, prev_pt, init_pt, use_init, L);
    }
    // index of the feature this thread works on
    __device__ __forceinline__ static int feature() { return COOP ? blockIdx.x : blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); }
    __device__ __forceinline__ static int first_feature() { return COOP ? blockIdx.x : blockIdx.x * (blockDim.x >> 5); }
    __device__ __forceinline__ static bool writer() { return COOP ? threadIdx.x == 0 : (threadIdx.x & 31) == 0; }
};

// ---- temporal: prev -> cur (maxLevel L) then cur -> prev (maxLevel 1, initial flow), feature_tracker.cpp:37-67 ----
template <bool COOP>
__device__ __forceinline__ void lk_temporal_warp_body(const DevBatch& b, int parity_arg, int cur, int prev, LevelRef (*pyr)[kMaxLevels], WarpPass& result) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;
    TraceScope trace(b.trace, TR_LK_TEMPORAL + 16 * tphase);
    const int s = blockIdx.y, i = Pass<COOP>::feature(), cap = b.cap;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    if (Pass<COOP>::first_feature() >= n_prev) return;           // whole CTA beyond the feature count
    load_pyrs(pyr, b.left[prev], b.left[cur], b.lvl_stream_stride, s);   // 0 = previous left, 1 = current left
    if (i >= n_prev || i >= cap) return;
    const Pass<COOP> pass(b.n_levels, &result);
    const float2 p0 = b.kps[parity ^ 1][(size_t)s * cap + i];
    const int L = min(b.lk_max_level, b.left[cur].n_levels - 1);
    const WarpPass f = pass.run(pyr[0], pyr[1], p0, p0, false, L);
    const float2 fwd = make_float2(f.x, f.y);
    float2 rev = p0;
    int st = f.status, rst = 0, iters = f.iters, fast = f.fast;
    int ok = st;
    if (b.flow_back) {
        // the reference also runs the reverse pass on points whose forward status is already 0 and then discards the
        // result (:55-59); the device skips that work, the outcome is the same
        if (st) {
            const WarpPass r = pass.run(pyr[1], pyr[0], fwd, p0, true, min(1, L));
            rev = make_float2(r.x, r.y); rst = r.status; iters += r.iters; fast += r.fast;
        }
        ok = st && rst && pt_distance(p0, rev) <= 0.5;
    }
    if (ok && !in_border(fwd, b.H, b.W)) ok = 0;
    if (Pass<COOP>::writer()) {
        const size_t o = (size_t)s * cap + i;
        b.fwd_pts[o] = fwd; b.fwd_st[o] = (uint8_t)st;
        b.rev_pts[o] = rev; b.rev_st[o] = (uint8_t)rst;
        b.temporal_st[o] = (uint8_t)ok;
        b.lk_iters[o] = iters | (fast << 16);   // low 16 bits: iterations, high: of which on the exact-integer path
    }
}

// ---- stereo: left -> right then right -> left (both maxLevel L), status (:112-135), then RemoveDistortion +
//      ComputeKpsVelocity for the right camera and the right half of the featureFrame record ----
template <bool COOP>
__device__ __forceinline__ void lk_stereo_warp_body(const DevBatch& b, int parity_arg, int cur, const double* __restrict__ dt_host, LevelRef (*pyr)[kMaxLevels],
                                                    WarpPass& result) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;
    TraceScope trace(b.trace, TR_LK_STEREO + 16 * tphase);
    const int s = blockIdx.y, i = Pass<COOP>::feature(), cap = b.cap;
    const int n = b.counters[parity][s * C_COUNT + C_N_TOTAL];
    if (Pass<COOP>::first_feature() >= n) return;
    load_pyrs(pyr, b.left[cur], b.right[parity], b.lvl_stream_stride, s);   // 0 = current left, 1 = current right
    if (i >= n || i >= cap) return;
    const Pass<COOP> pass(b.n_levels, &result);
    const size_t o = (size_t)s * cap + i;
    const float2 p0 = b.kps[parity][o];
    const int pidx = b.prev_idx[parity][o];
    // previous_right_un_distortion_ids_kps_map_.find(id): select_tracked recorded where the feature sat in the previous
    // frame's arrays (-1: new corner); it is in the right map iff it had a right observation there
    const int found = (pidx >= 0 && b.right_ok[parity ^ 1][(size_t)s * cap + pidx]) ? pidx : -1;
    const int L = min(b.lk_max_level, b.right[parity].n_levels - 1);
    const WarpPass f = pass.run(pyr[0], pyr[1], p0, p0, false, L);
    const float2 rp = make_float2(f.x, f.y);
    float2 back = p0;
    int st = f.status, rst = 0, iters = f.iters, fast = f.fast;
    int ok = st;
    if (b.flow_back) {
        if (st) {
            const WarpPass r = pass.run(pyr[1], pyr[0], rp, rp, false, L);
            back = make_float2(r.x, r.y); rst = r.status; iters += r.iters; fast += r.fast;
        }
        ok = st && rst && in_border(rp, b.H, b.W) && pt_distance(p0, back) <= 0.5;
    }
    if (Pass<COOP>::writer()) {
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

// ONE entry point per scheduling for both pairs of LK calls (bit 2 of parity_arg: 0 = temporal, 1 = stereo): the passes are
// __noinline__ functions, linked into every kernel that calls them, and in overlapped frames the temporal launch of frame k+1
// shares the SMs with the stereo launch of frame k -- two kernels meant two 65 KB copies of the same code fighting for the
// instruction caches (vf_lk.cu: lk_pair_kernel).
template <bool COOP>
__global__ void VF_LKW_BOUNDS(COOP) lk_pair_warp_kernel(const __grid_constant__ DevBatch b, int parity_arg, int cur, int prev, const double* __restrict__ dt_host) {
    pdl_wait(); pdl_release();
    __shared__ LevelRef pyr[2][kMaxLevels];
    __shared__ WarpPass result;
    if (parity_arg & 4) lk_stereo_warp_body<COOP>(b, parity_arg & ~4, cur, dt_host, pyr, result);
    else lk_temporal_warp_body<COOP>(b, parity_arg, cur, prev, pyr, result);
}

// ---- stand-alone calcOpticalFlowPyrLK (vf_op_lk) ----
template <bool COOP>
__global__ void __launch_bounds__(128) lk_op_warp_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts,
                                                         float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                         int* __restrict__ iters_out, int n, int max_level, int use_init) {
    __shared__ LevelRef pyr[2][kMaxLevels];
    __shared__ WarpPass result;
    const int i = Pass<COOP>::feature();
    load_pyrs(pyr, prev, next, nullptr, 0);
    if (i >= n) return;
    const int L = min(max_level, min(prev.n_levels, next.n_levels) - 1);
    const Pass<COOP> pass(L + 1, &result);
    const WarpPass r = pass.run(pyr[0], pyr[1], prev_pts[i], next_pts[i], use_init != 0, L);
    if (Pass<COOP>::writer()) {
        next_pts[i] = make_float2(r.x, r.y); status[i] = (uint8_t)r.status;
        if (iters_out) iters_out[i] = r.iters;
    }
}

constexpr int kBatchWarps = VF_LKW_BATCH_WARPS;     // features per CTA in BATCH scheduling
size_t smem_batch() { return (sizeof(LkLevel) + sizeof(LkWork)) * kBatchWarps; }
size_t smem_coop(int n_levels) { return sizeof(LkLevel) * (size_t)(n_levels > 1 ? n_levels : 1) + sizeof(LkWork) * kCoopWarps; }

}  // namespace

// 0 = by launch size: the 4-warp-per-feature kernels of vf_lk.cu (block barrier per iteration) while a feature has an SM to itself
// -- measured on one 752x480 stream: temporal / stereo 27 / 42 us there, 30 / 46 us with COOP, 40 / 60 us with one warp per
// feature -- and the warp-level passes with BATCH scheduling once features queue up per SM (64 streams: 1.41 -> 1.08 ms per
// step); 1 = vf_lk.cu whatever the size; 2 = BATCH whatever the size; 3 = COOP whatever the size.
static int g_lk_mode = 0;
void set_lk_kernel_mode(int mode) { g_lk_mode = mode; }
bool lk_use_warp_kernel(long long n_features) { return g_lk_mode >= 2 || (g_lk_mode == 0 && n_features > 4LL * kNumSms); }
static bool pick_coop(long long n_features) { (void)n_features; return g_lk_mode == 3; }

cudaError_t lk_warp_prepare() {
    cudaError_t e;
    const int coop = (int)smem_coop(kMaxLevels), batch = (int)smem_batch();
#define VF_PREP(kernel, bytes)                                                                                            \
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;  \
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    VF_PREP(lk_pair_warp_kernel<true>, coop) VF_PREP(lk_pair_warp_kernel<false>, batch)
    VF_PREP(lk_op_warp_kernel<true>, coop) VF_PREP(lk_op_warp_kernel<false>, batch)
#undef VF_PREP
    return cudaSuccess;
}

void launch_lk_temporal_warp(const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream) {
    if (pick_coop((long long)h.cap * h.n_streams)) {
        launch_kernel(lk_pair_warp_kernel<true>, dim3(h.cap, h.n_streams), dim3(32 * kCoopWarps), smem_coop(h.n_levels), stream, h, parity, cur, prev, (const double*)nullptr);
    } else {
        dim3 grid((h.cap + kBatchWarps - 1) / kBatchWarps, h.n_streams);
        launch_kernel(lk_pair_warp_kernel<false>, grid, dim3(32 * kBatchWarps), smem_batch(), stream, h, parity, cur, prev, (const double*)nullptr);
    }
}
void launch_lk_stereo_warp(const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream) {
    if (pick_coop((long long)h.cap * h.n_streams)) {
        launch_kernel(lk_pair_warp_kernel<true>, dim3(h.cap, h.n_streams), dim3(32 * kCoopWarps), smem_coop(h.n_levels), stream, h, parity | 4, cur, 0, dt_host);
    } else {
        dim3 grid((h.cap + kBatchWarps - 1) / kBatchWarps, h.n_streams);
        launch_kernel(lk_pair_warp_kernel<false>, grid, dim3(32 * kBatchWarps), smem_batch(), stream, h, parity | 4, cur, 0, dt_host);
    }
}
void launch_lk_op_warp(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status, int* iters, int n,
                       int max_level, int use_initial_flow, cudaStream_t stream) {
    if (n <= 0) return;
    const int nlv = (max_level < prev.n_levels - 1 ? max_level : prev.n_levels - 1) + 1;
    if (pick_coop(n))
        lk_op_warp_kernel<true><<<n, 32 * kCoopWarps, smem_coop(nlv), stream>>>(prev, next, prev_pts, next_pts, status, iters, n, max_level, use_initial_flow ? 1 : 0);
    else
        lk_op_warp_kernel<false><<<(n + kBatchWarps - 1) / kBatchWarps, 32 * kBatchWarps, smem_batch(), stream>>>(prev, next, prev_pts, next_pts, status, iters, n,
                                                                                                                 max_level, use_initial_flow ? 1 : 0);
}

}  // namespace vf

#ifdef VF_LKW_TIMING
// measurement aid (timing build only): read and clear the phase counters
extern "C" int vf_debug_lkw_timing(unsigned long long* out32) {
    unsigned long long zero[32] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(out32, vf_lkw_timing, sizeof(zero)) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(vf_lkw_timing, zero, sizeof(zero)) != cudaSuccess;
}
#endif
