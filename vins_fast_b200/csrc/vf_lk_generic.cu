// Pyramidal Lucas-Kanade for window sizes other than 21x21 (`lk_win`: any odd width 5..31).
//
// The reference passes the literal cv::Size(21, 21) at its four calcOpticalFlowPyrLK call sites
// (vins/estimator/feature_tracker.cpp:42,51,119,127); the north star keeps the window size as a parameter, so this file carries
// the same four calls (forward + flow-back, left -> right -> left, status logic :55-67,128-133) for a window given at run time.
// The arithmetic is OpenCV's LKTrackerInvoker (video/lkpyramid.cpp) for that width, bit for bit:
//   * 14-bit fixed-point bilinear patch of I (2^5) and of its Scharr derivatives, REFLECT_101 outside the level for I, zero
//     outside the level for the derivatives;
//   * the float32 sums of A and b in the order of OpenCV's 128-bit SIMD loops for a row of `win` columns: the first
//     (win / 8) * 8 columns in 8-wide groups (A: four lane accumulators per matrix entry, lane = column & 3; b: eight lane
//     accumulators of pmaddwd pair sums (column c, c + 4)), the remaining columns in one scalar accumulator per sum, then
//     v_reduce_sum's tree -- so the chain LAYOUT is a function of the width and is built here from `win`;
//   * the exact-integer shortcut of the 21x21 kernels (vf_lk.cu): b1 only adds x-terms and b2 only y-terms, so if the sum of
//     the |terms| of each stays below 2^24 every partial sum of every chain is an exactly representable integer and the
//     ordered chains equal the integer sums.
// One 256-thread CTA per feature, level after level, all tiles in 25 KB of static shared memory; tiles that reach beyond the
// filled part of a level's border are fetched with index reflection (the 21x21 kernels' aligned word fetches need 26 of the 28
// filled border pixels; a 31x31 window can reach 36), the others with aligned words like theirs.  The tuned kernels of vf_lk.cu /
// vf_lk_warp.cu stay the path for the reference's own width; this one trades their per-level pipelining for generality.
// Checked against the C oracle for every odd width 5..31 (tests/test_gpu_stages.py, tests/test_gpu_tracker.py), and the oracle
// against cv2 for the same widths (tests/golden/lk_windows_cv2_4_13.npz).
#include "vf_camera.cuh"
#include "vf_kernels.h"
#include "vf_lk_common.cuh"

namespace vf {

namespace {

constexpr int kGenWinMax = 31;
constexpr int kGenThreads = 256;                                // two warps per SM sub-partition: a 31-wide window is 2.2 x the pixels of a 21-wide one
constexpr int kGenWarps = kGenThreads / 32;
constexpr int kGenMargin = 4;                                   // staged J tile = (win + 1) taps + this margin on every side
constexpr int kGenTI = kGenWinMax + 3;                          // I tile: window + bilinear tap + Scharr ring (pitch: next multiple of 4)
constexpr int kGenTD = kGenWinMax + 1;                          // derivative tile: window + bilinear tap
constexpr int kGenTJ = kGenWinMax + 1 + 2 * kGenMargin;         // J tile (pitch: next multiple of 4)
constexpr int kGenPx = kGenWinMax * kGenWinMax;
constexpr int kGenIpt = 3;                                      // work items per thread: pair items + tail items <= 589 (win 31)
// staged terms of A: per matrix entry four SIMD-lane chains of win * (simd_w / 4) terms and one tail chain of win * tail
// terms, each padded to a multiple of four -- largest at win 31: 3 * (4 * 188 + 220)
constexpr int kGenScratch = 3 * (4 * 188 + 220);
static_assert(kGenScratch >= 2 * (4 * 96 + 220), "the staged terms of b (2 x (4 x win * (simd_w / 8) + win * tail), padded) must fit the scratch area");

struct GenSmem {
    LevelRef pyr[2][kMaxLevels];
    alignas(16) uint8_t tileI[((kGenTI + 3) & ~3) * kGenTI];
    alignas(16) short2 dtile[kGenTD * kGenTD];
    alignas(16) short Iw[kGenPx + 3];
    alignas(16) short2 dI[kGenPx + 1];
    alignas(16) uint8_t tileJ[((kGenTJ + 3) & ~3) * kGenTJ];
    // the terms of A staged in accumulation order (prologue of a level) and, later in the same level, the spilled terms of b
    // (iterations that leave the exact-integer path): never live at the same time
    alignas(16) int scratch[kGenScratch];
    alignas(16) int4 part[2][kGenWarps];
    alignas(16) float A[4];                                     // A11, A12, A22, 1 / det
    float2 ib;
    int skip;
    int found;
};

// REFLECT_101 for the coordinates a window can reach (one reflection; a level is always wider than the window), clamped for the
// tile positions nothing ever reads (the Scharr ring around a tap that lies outside the level, the margin of a J tile).
__device__ __forceinline__ int refl_clamp(int i, int n) {
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    return min(max(i, 0), n - 1);
}

// side x side pixels of a level from (x0, y0) into a shared-memory tile whose pitch is a multiple of 4.  Where the rows lie inside
// the level's filled REFLECT_101 border (vf_common.cuh: kPyrPadFill; always, for windows up to 21) they are fetched as aligned
// 32-bit words and realigned with a funnel shift; otherwise pixel by pixel with index reflection.
__device__ __forceinline__ void gen_stage_tile(uint8_t* dst, int pitch, int side, const LevelRef& L, int x0, int y0) {
    const int tid = threadIdx.x, m = x0 & 3, xa = x0 - m, ow = pitch >> 2;
    if (xa >= -kPyrPadFill && xa + 4 * (ow + 1) <= L.w + kPyrPadFill && y0 >= -kPyrPadFill && y0 + side <= L.h + kPyrPadFill) {
        for (int y = tid >> 3; y < side; y += kGenThreads / 8)
            for (int k = tid & 7; k < ow; k += 8) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(L.p + (ptrdiff_t)(y0 + y) * L.pitch + xa) + k;   // origin, pitch: 16-byte aligned
                reinterpret_cast<uint32_t*>(dst + y * pitch)[k] = __funnelshift_r(__ldg(wp), __ldg(wp + 1), 8 * m);
            }
    } else {
        for (int y = tid >> 4; y < side; y += kGenThreads / 16) {
            const uint8_t* row = L.p + (ptrdiff_t)refl_clamp(y0 + y, L.h) * L.pitch;
            for (int x = tid & 15; x < side; x += 16) dst[y * pitch + x] = __ldg(row + refl_clamp(x0 + x, L.w));
        }
    }
}

// One calcOpticalFlowPyrLK call for one point; all threads of the CTA take part and return the same result.
__device__ __noinline__ LkResult lk_track_generic(GenSmem& sm, int ia, int ja, float2 prev_pt, float2 init_pt, bool use_init,
                                                  int max_level, int win, int* iters_io) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float half = __fmul_rn((float)(win - 1), 0.5f);
    const float FLT_SCALE = 1.f / (1 << 20);
    const int TIs = win + 3, TD = win + 1, TJs = win + 1 + 2 * kGenMargin;   // tile sides
    const int TI = (TIs + 3) & ~3, TJ = (TJs + 3) & ~3;                     // tile pitches
    const int simd_w = (win >> 3) << 3, ngrp = simd_w >> 3, tailw = win - simd_w;
    const int n_pair = win * ngrp * 4, n_item = n_pair + win * tailw;
    const int n_q = (n_item + kGenThreads - 1) / kGenThreads;   // item slots in use (uniform): a 15-wide window needs 1 of the 3
    int status = 1, iters = 0, fast = 0;
    float outx = init_pt.x, outy = init_pt.y;
    const int Ls = win * (simd_w >> 2), Lsp = (Ls + 3) & ~3, Lt = win * tailw, Ltp = (Lt + 3) & ~3, Lm = 4 * Lsp + Ltp;
    float* termsA = reinterpret_cast<float*>(sm.scratch);
    // staged terms of b (iterations off the exact-integer path): per component four SIMD-lane chains of win * (simd_w / 8) pair
    // sums and one tail chain, in accumulation order
    const int Lb = win * ngrp, Lbp = (Lb + 3) & ~3, Lbm = 4 * Lbp + Ltp;
    float* termsB = reinterpret_cast<float*>(sm.scratch);

    // Work items of an iteration (level-invariant): pair items = the column pairs (c, c + 4) of every 8-wide group of every
    // row (OpenCV's pmaddwd pair sums), tail items = single pixels of the scalar tail columns.
    int e0[kGenIpt], e1[kGenIpt], o0[kGenIpt], o1[kGenIpt], bs[kGenIpt];   // bs: slot of the item's term in its chain of b
#pragma unroll
    for (int q = 0; q < kGenIpt; ++q) {
        const int it = tid + q * kGenThreads;
        e0[q] = -1; e1[q] = -1; bs[q] = 0;
        if (it < n_pair) {
            const int y = it / (4 * ngrp), r = it - y * (4 * ngrp);
            e0[q] = y * win + 8 * (r >> 2) + (r & 3); e1[q] = e0[q] + 4;
            bs[q] = (r & 3) * Lbp + y * ngrp + (r >> 2);
        } else if (it < n_item) {
            const int t = it - n_pair, y = t / tailw;
            e0[q] = y * win + simd_w + (t - y * tailw);
            bs[q] = 4 * Lbp + t;
        }
        o0[q] = e0[q] < 0 ? 0 : (e0[q] / win) * TJ + (e0[q] - (e0[q] / win) * win);
        o1[q] = e1[q] < 0 ? 0 : (e1[q] / win) * TJ + (e1[q] - (e1[q] / win) * win);
    }

#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        const LevelRef LI = sm.pyr[ia][level], LJ = sm.pyr[ja][level];
        const int jcols = LJ.w, jrows = LJ.h;
        const float scale = pow2_neg(level);
        float nx, ny;
        if (level == max_level) {
            if (use_init) { nx = __fmul_rn(outx, scale); ny = __fmul_rn(outy, scale); }
            else { nx = __fmul_rn(prev_pt.x, scale); ny = __fmul_rn(prev_pt.y, scale); }
        } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;
        const float ppx = __fsub_rn(__fmul_rn(prev_pt.x, scale), half), ppy = __fsub_rn(__fmul_rn(prev_pt.y, scale), half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -win || ipx >= LI.w || ipy < -win || ipy >= LI.h) {   // window outside the level (lkpyramid.cpp)
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);

        // ---- source side of the level: I tile, Scharr taps, interpolated patches.  The J tile around the start estimate goes
        //      out in the same batch of loads (one global round trip per level instead of two) ----
        int jx0 = __float2int_rd(__fsub_rn(nx, half)) - kGenMargin, jy0 = __float2int_rd(__fsub_rn(ny, half)) - kGenMargin;
        __syncthreads();                                        // the previous level / pass may still read the tiles
        gen_stage_tile(sm.tileI, TI, TIs, LI, ipx - 1, ipy - 1);
        gen_stage_tile(sm.tileJ, TJ, TJs, LJ, jx0, jy0);
        if (tid < 15) {                                         // +0 in the padding slots of the 15 chains of A (an exact identity)
            const int m = tid / 5, c = tid - 5 * m;
            float* ch = termsA + m * Lm + (c < 4 ? c * Lsp : 4 * Lsp);
            for (int i = (c < 4 ? Ls : Lt); i < (c < 4 ? Lsp : Ltp); ++i) ch[i] = 0.f;
        }
        __syncthreads();
        for (int ty = tid >> 4; ty < TD; ty += kGenThreads / 16)
        for (int tx = tid & 15; tx < TD; tx += 16) {
            const int e = ty * TD + tx, x = ipx + tx, y = ipy + ty;
            const bool inside = (unsigned)x < (unsigned)LI.w && (unsigned)y < (unsigned)LI.h;
            const uint8_t* p = sm.tileI + (ty + 1) * TI + (tx + 1);
            const int a0 = p[-TI - 1], a1 = p[-TI], a2 = p[-TI + 1], a3 = p[-1], a4 = p[1], a5 = p[TI - 1], a6 = p[TI], a7 = p[TI + 1];
            const int t0m = (a0 + a5) * 3 + a3 * 10, t0p = (a2 + a7) * 3 + a4 * 10;
            const int t1m = a5 - a0, t1c = a6 - a1, t1p = a7 - a2;
            sm.dtile[e] = make_short2((short)(inside ? t0p - t0m : 0), (short)(inside ? (t1m + t1p) * 3 + t1c * 10 : 0));
        }
        __syncthreads();
        for (int y = tid >> 4; y < win; y += kGenThreads / 16)
        for (int x = tid & 15; x < win; x += 16) {
            const int e = y * win + x;
            const uint8_t* s = sm.tileI + (y + 1) * TI + (x + 1);
            const short2* dt = sm.dtile + y * TD + x;
            const short2 d00 = dt[0], d01 = dt[1], d10 = dt[TD], d11 = dt[TD + 1];
            sm.Iw[e] = (short)VF_DESCALE(s[0] * w00 + s[1] * w01 + s[TI] * w10 + s[TI + 1] * w11, 9);
            const int ixv = VF_DESCALE(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, 14);
            const int iyv = VF_DESCALE(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, 14);
            sm.dI[e] = make_short2((short)ixv, (short)iyv);
            // the three products in the slot of their chain: |Ix|, |Iy| <= 4080, so each is an exact float
            float* ta = termsA + (x < simd_w ? (x & 3) * Lsp + y * (simd_w >> 2) + (x >> 2) : 4 * Lsp + y * tailw + (x - simd_w));
            ta[0] = (float)(ixv * ixv); ta[Lm] = (float)(ixv * iyv); ta[2 * Lm] = (float)(iyv * iyv);
        }
        __syncthreads();
        // ---- the 15 ordered float chains of A, one lane per chain: per matrix entry four SIMD lanes (lane l sees columns
        //      l, l + 4, ... < simd_w of every row in turn) and the scalar tail (columns simd_w .. win - 1 of every row), read
        //      back in that order with 16-byte loads ----
        if (wid == 0) {
            const int ch = min(lane, 14), m = ch / 5, c = ch - 5 * m;
            const float4* src = reinterpret_cast<const float4*>(termsA + m * Lm + (c < 4 ? c * Lsp : 4 * Lsp));
            const int n4 = (c < 4 ? Lsp : Ltp) >> 2;
            float acc = 0.f;
#pragma unroll 4
            for (int i = 0; i < n4; ++i) {
                const float4 v = src[i];
                acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
            }
            float r[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float q0 = __shfl_sync(0xffffffffu, acc, 5 * k + 0), q1 = __shfl_sync(0xffffffffu, acc, 5 * k + 1);
                const float q2 = __shfl_sync(0xffffffffu, acc, 5 * k + 2), q3 = __shfl_sync(0xffffffffu, acc, 5 * k + 3);
                const float tl = __shfl_sync(0xffffffffu, acc, 5 * k + 4);
                r[k] = __fadd_rn(tl, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));   // iA += v_reduce_sum(qA)
            }
            if (lane == 0) {
                const float A11 = __fmul_rn(r[0], FLT_SCALE), A12 = __fmul_rn(r[1], FLT_SCALE), A22 = __fmul_rn(r[2], FLT_SCALE);
                const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
                const float dA = __fsub_rn(A11, A22);
                const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
                const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * win * win));
                sm.A[0] = A11; sm.A[1] = A12; sm.A[2] = A22; sm.A[3] = __fdiv_rn(1.f, D);
                sm.skip = (minEig < 1e-4f || D < FLT_EPSILON) ? 1 : 0;
            }
        }
        __syncthreads();
        if (sm.skip) {                                          // minEig / det below threshold (lkpyramid.cpp)
            if (level == 0) status = 0;
            continue;
        }
        const float A11 = sm.A[0], A12 = sm.A[1], A22 = sm.A[2], D = sm.A[3];
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        // this thread's source pixels stay in registers for the whole level
        int I0[kGenIpt], I1[kGenIpt];
        short2 g0[kGenIpt], g1[kGenIpt];
#pragma unroll
        for (int q = 0; q < kGenIpt; ++q) {
            if (q >= n_q) break;
            I0[q] = sm.Iw[max(e0[q], 0)]; I1[q] = sm.Iw[max(e1[q], 0)];
            g0[q] = e0[q] >= 0 ? sm.dI[e0[q]] : make_short2(0, 0); g1[q] = e1[q] >= 0 ? sm.dI[e1[q]] : make_short2(0, 0);
        }
        int pbuf = 0;
        float pdx = 0.f, pdy = 0.f;
#pragma unroll 1
        for (int j = 0; ; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -win || inx >= jcols || iny < -win || iny >= jrows) {
                if (level == 0) status = 0;
                break;
            }
            if (inx < jx0 || iny < jy0 || inx + TD > jx0 + TJs || iny + TD > jy0 + TJs) {
                jx0 = inx - kGenMargin; jy0 = iny - kGenMargin;     // (re-)centre the staged tile on the window (uniform branch)
                __syncthreads();
                gen_stage_tile(sm.tileJ, TJ, TJs, LJ, jx0, jy0);
                __syncthreads();
            }
            ++iters;
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const uint8_t* jt = sm.tileJ + (iny - jy0) * TJ + (inx - jx0);
            // exact int32 products (idle items own zero gradients); abx / aby: this thread's sum of |term| per component over the
            // terms the chains would add (pair sums for pair items), saturated so that the CTA's total stays below 2^32
            int ex0[kGenIpt], ey0[kGenIpt], ex1[kGenIpt], ey1[kGenIpt];
            int tx = 0, ty = 0;
            unsigned int abx = 0, aby = 0;
#pragma unroll
            for (int q = 0; q < kGenIpt; ++q) {
                if (q >= n_q) break;
                const uint8_t* p0 = jt + o0[q];
                const uint8_t* p1 = jt + o1[q];
                const int d0 = VF_DESCALE(p0[0] * w00 + p0[1] * w01 + p0[TJ] * w10 + p0[TJ + 1] * w11, 9) - I0[q];
                const int d1 = VF_DESCALE(p1[0] * w00 + p1[1] * w01 + p1[TJ] * w10 + p1[TJ + 1] * w11, 9) - I1[q];
                ex0[q] = d0 * g0[q].x; ey0[q] = d0 * g0[q].y; ex1[q] = d1 * g1[q].x; ey1[q] = d1 * g1[q].y;
                const int sx_ = ex0[q] + ex1[q], sy_ = ey0[q] + ey1[q];
                tx += sx_; ty += sy_;
                abx += (unsigned int)min(abs(sx_), kExact); aby += (unsigned int)min(abs(sy_), kExact);
            }
            abx = min(abx, (unsigned int)kExact); aby = min(aby, (unsigned int)kExact);
            const int sx = __reduce_add_sync(0xffffffffu, tx), sy = __reduce_add_sync(0xffffffffu, ty);
            const unsigned int sax = __reduce_add_sync(0xffffffffu, abx), say = __reduce_add_sync(0xffffffffu, aby);
            if (lane == 0) sm.part[pbuf][wid] = make_int4(sx, sy, (int)sax, (int)say);
            __syncthreads();
            int bx = 0, by = 0;
            unsigned int bz = 0, bw = 0;
#pragma unroll
            for (int w = 0; w < kGenWarps; ++w) {
                const int4 q4 = sm.part[pbuf][w];
                bx += q4.x; by += q4.y; bz += (unsigned int)q4.z; bw += (unsigned int)q4.w;
            }
            pbuf ^= 1;
            float ib1, ib2;
            if (max(bz, bw) < (unsigned int)kExact) {
                ib1 = (float)bx; ib2 = (float)by;               // every partial sum of every chain is an exact integer
                ++fast;
            } else {
                // ordered chains, one lane per chain: per component the four SIMD lane accumulators (qb0 / qb1 of lkpyramid.cpp:
                // column offset c of an 8-wide group with its partner c + 4, one pmaddwd pair sum per group and row) and the scalar
                // tail.  Every thread converts its items' terms (v_cvt_f32 of the exact int32 pair sum; a tail item's partner is
                // zero) and parks them in accumulation order; the lanes read them back with 16-byte loads.
#pragma unroll
                for (int q = 0; q < kGenIpt; ++q) {
                    if (q >= n_q) break;
                    if (e0[q] >= 0) { termsB[bs[q]] = (float)(ex0[q] + ex1[q]); termsB[Lbm + bs[q]] = (float)(ey0[q] + ey1[q]); }
                }
                if (tid < 10) {                                 // +0 in the padding slots (the area also holds the terms of A)
                    const int comp = tid / 5, c = tid - 5 * comp;
                    float* ch = termsB + comp * Lbm + (c < 4 ? c * Lbp : 4 * Lbp);
                    for (int i = (c < 4 ? Lb : Lt); i < (c < 4 ? Lbp : Ltp); ++i) ch[i] = 0.f;
                }
                __syncthreads();
                if (wid == 0) {
                    const int ch = min(lane, 9), comp = ch / 5, c = ch - 5 * comp;
                    const float4* src = reinterpret_cast<const float4*>(termsB + comp * Lbm + (c < 4 ? c * Lbp : 4 * Lbp));
                    const int n4 = (c < 4 ? Lbp : Ltp) >> 2;
                    float acc = 0.f;
#pragma unroll 4
                    for (int i = 0; i < n4; ++i) {
                        const float4 v = src[i];
                        acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
                    }
                    // s_k = qb0[k] + qb1[k]; ib += v_reduce_sum({s0, s2, 0, 0}) resp. {s1, s3, 0, 0}
                    float r[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float m0 = __shfl_sync(0xffffffffu, acc, 5 * k + 0), m1 = __shfl_sync(0xffffffffu, acc, 5 * k + 1);
                        const float m2 = __shfl_sync(0xffffffffu, acc, 5 * k + 2), m3 = __shfl_sync(0xffffffffu, acc, 5 * k + 3);
                        const float tl = __shfl_sync(0xffffffffu, acc, 5 * k + 4);
                        const float s0 = __fadd_rn(m0, m2), s2 = __fadd_rn(m1, m3);
                        r[k] = __fadd_rn(tl, __fadd_rn(__fadd_rn(s0, 0.f), __fadd_rn(s2, 0.f)));
                    }
                    if (lane == 0) sm.ib = make_float2(r[0], r[1]);
                }
                __syncthreads();
                ib1 = sm.ib.x; ib2 = sm.ib.y;
            }
            const float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            // the three ways out (lkpyramid.cpp, in this order): |delta|^2 <= eps^2 in double; j > 0 and |delta + prevDelta| < 0.01
            // on both axes (-> step back by delta / 2); maxCount = 30 iterations
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= 0.01 * 0.01) break;
            if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            if (j == 29) break;
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {   // the caller passes an err vector: final window test (lkpyramid.cpp)
            const int ix = __float2int_rd(__fsub_rn(outx, half)), iy = __float2int_rd(__fsub_rn(outy, half));
            if (ix < -win || ix >= jcols || iy < -win || iy >= jrows) status = 0;
        }
    }
    if (tid == 0) { iters_io[0] += iters; iters_io[1] += fast; }
    LkResult res;
    res.x = outx; res.y = outy; res.status = status;
    return res;
}

__device__ __forceinline__ void gen_load_pyr(GenSmem& sm, int slot, const PyrView& pv, const size_t* stride, int s) {
    const int l = threadIdx.x - slot * kMaxLevels;
    if (l >= 0 && l < kMaxLevels) {
        LevelRef r;
        const bool ok = l < pv.n_levels;
        r.p = ok ? pv.lvl[l] + (stride ? (size_t)s * stride[l] : 0) : nullptr;
        r.w = ok ? pv.w[l] : 0; r.h = ok ? pv.h[l] : 0; r.pitch = ok ? pv.pitch[l] : 0;
        sm.pyr[slot][l] = r;
    }
}

// ---- temporal: prev -> cur (maxLevel L) then cur -> prev (maxLevel 1, initial flow), feature_tracker.cpp:37-67 ----
__device__ __forceinline__ void gen_temporal_body(GenSmem& sm, int* stat, const DevBatch& b, int parity_arg, int cur, int prev) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;
    TraceScope trace(b.trace, TR_LK_TEMPORAL + 16 * tphase);
    const int s = blockIdx.y, i = blockIdx.x, cap = b.cap;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    if (i >= n_prev) return;
    const float2 p0 = b.kps[parity ^ 1][(size_t)s * cap + i];
    if (threadIdx.x == 0) { stat[0] = 0; stat[1] = 0; }
    gen_load_pyr(sm, 0, b.left[prev], b.lvl_stream_stride, s);
    gen_load_pyr(sm, 1, b.left[cur], b.lvl_stream_stride, s);
    __syncthreads();
    const int L = min(b.lk_max_level, b.left[cur].n_levels - 1);
    const LkResult f = lk_track_generic(sm, 0, 1, p0, p0, false, L, b.lk_win, stat);
    const float2 fwd = make_float2(f.x, f.y);
    float2 rev = p0;
    int st = f.status, rst = 0;
    int ok = st;
    if (b.flow_back) {      // the reference also runs the reverse pass on lost points and discards the result (:55-59)
        if (st) {
            const LkResult r = lk_track_generic(sm, 1, 0, fwd, p0, true, min(1, L), b.lk_win, stat);
            rev = make_float2(r.x, r.y); rst = r.status;
        }
        ok = st && rst && pt_distance(p0, rev) <= 0.5;
    }
    if (ok && !in_border(fwd, b.H, b.W)) ok = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t o = (size_t)s * cap + i;
        b.fwd_pts[o] = fwd; b.fwd_st[o] = (uint8_t)st;
        b.rev_pts[o] = rev; b.rev_st[o] = (uint8_t)rst;
        b.temporal_st[o] = (uint8_t)ok;
        b.lk_iters[o] = stat[0] | (stat[1] << 16);
    }
}

// ---- stereo: left -> right then right -> left (both maxLevel L), status (:112-135), right RemoveDistortion + velocity ----
__device__ __forceinline__ void gen_stereo_body(GenSmem& sm, int* stat, const DevBatch& b, int parity_arg, int cur,
                                                const double* __restrict__ dt_host) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;
    TraceScope trace(b.trace, TR_LK_STEREO + 16 * tphase);
    const int s = blockIdx.y, i = blockIdx.x, cap = b.cap, tid = threadIdx.x;
    const int n = b.counters[parity][s * C_COUNT + C_N_TOTAL];
    if (i >= n) return;
    const size_t o = (size_t)s * cap + i;
    const float2 p0 = b.kps[parity][o];
    const int pidx = b.prev_idx[parity][o];
    if (tid == 0) {
        stat[0] = 0; stat[1] = 0;
        // previous_right_un_distortion_ids_kps_map_.find(id): in the map iff the feature had a right observation one frame ago
        sm.found = (pidx >= 0 && b.right_ok[parity ^ 1][(size_t)s * cap + pidx]) ? pidx : -1;
    }
    gen_load_pyr(sm, 0, b.left[cur], b.lvl_stream_stride, s);
    gen_load_pyr(sm, 1, b.right[parity], b.lvl_stream_stride, s);
    __syncthreads();
    const int L = min(b.lk_max_level, b.right[parity].n_levels - 1);
    const LkResult f = lk_track_generic(sm, 0, 1, p0, p0, false, L, b.lk_win, stat);
    const float2 rp = make_float2(f.x, f.y);
    float2 back = p0;
    int st = f.status, rst = 0;
    int ok = st;
    if (b.flow_back) {
        if (st) {
            const LkResult r = lk_track_generic(sm, 1, 0, rp, rp, false, L, b.lk_win, stat);
            back = make_float2(r.x, r.y); rst = r.status;
        }
        ok = st && rst && in_border(rp, b.H, b.W) && pt_distance(p0, back) <= 0.5;
    }
    __syncthreads();
    if (tid == 0) {
        vf_feature* rec = &b.out[tphase % 3][o];
        rec->has_right = ok;
        float2 unr = make_float2(0.f, 0.f), vr = make_float2(0.f, 0.f);
        if (ok) {
            unr = undistort_point(b.cam[1], rp);
            if (sm.found >= 0) vr = velocity(unr, b.un_right[parity ^ 1][(size_t)s * cap + sm.found], dt_host[s]);
        }
        rec->xr = unr.x; rec->yr = unr.y; rec->ur = ok ? rp.x : 0.f; rec->vr = ok ? rp.y : 0.f; rec->vxr = vr.x; rec->vyr = vr.y;
        b.un_right[parity][o] = unr;
        b.right_ok[parity][o] = (uint8_t)ok;
        b.lr_pts[o] = rp; b.lr_st[o] = (uint8_t)st; b.rl_pts[o] = back; b.rl_st[o] = (uint8_t)rst;
        b.lk_iters_stereo[o] = stat[0] | (stat[1] << 16);
        if (ok) atomicAdd(&b.counters[parity][s * C_COUNT + C_N_RIGHT], 1);
    }
}

// one entry point for both pairs of calls (bit 2 of parity_arg: 0 = temporal, 1 = stereo), as lk_pair_kernel
__global__ void __launch_bounds__(kGenThreads) lk_pair_generic_kernel(const __grid_constant__ DevBatch b, int parity_arg, int cur, int prev,
                                                                      const double* __restrict__ dt_host) {
    __shared__ GenSmem sm;
    __shared__ int stat[2];
    pdl_wait(); pdl_release();
    if (parity_arg & 4) gen_stereo_body(sm, stat, b, parity_arg & ~4, cur, dt_host);
    else gen_temporal_body(sm, stat, b, parity_arg, cur, prev);
}

__global__ void __launch_bounds__(kGenThreads) lk_op_generic_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts,
                                                                    float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                                    int* __restrict__ iters_out, int n, int max_level, int use_init, int win) {
    __shared__ GenSmem sm;
    __shared__ int stat[2];
    const int i = blockIdx.x;
    if (i >= n) return;
    const int L = min(max_level, min(prev.n_levels, next.n_levels) - 1);
    if (threadIdx.x == 0) { stat[0] = 0; stat[1] = 0; }
    gen_load_pyr(sm, 0, prev, nullptr, 0);
    gen_load_pyr(sm, 1, next, nullptr, 0);
    __syncthreads();
    const LkResult r = lk_track_generic(sm, 0, 1, prev_pts[i], next_pts[i], use_init != 0, L, win, stat);
    __syncthreads();
    if (threadIdx.x == 0) {
        next_pts[i] = make_float2(r.x, r.y); status[i] = (uint8_t)r.status;
        if (iters_out) iters_out[i] = stat[0];
    }
}

}  // namespace

bool lk_window_supported(int win) { return win >= 5 && win <= kGenWinMax && (win & 1) == 1; }

cudaError_t lk_generic_prepare() {   // same carveout as every other kernel of the library (see corners_prepare)
    cudaError_t e = cudaFuncSetAttribute(lk_pair_generic_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(lk_op_generic_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout);
}

void launch_lk_temporal_generic(const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream) {
    launch_kernel(lk_pair_generic_kernel, dim3(h.cap, h.n_streams), dim3(kGenThreads), 0, stream, h, parity, cur, prev, (const double*)nullptr);
}
void launch_lk_stereo_generic(const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream) {
    launch_kernel(lk_pair_generic_kernel, dim3(h.cap, h.n_streams), dim3(kGenThreads), 0, stream, h, parity | 4, cur, 0, dt_host);
}
void launch_lk_op_generic(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status, int* iters,
                          int n, int max_level, int use_initial_flow, int win, cudaStream_t stream) {
    if (n <= 0) return;
    lk_op_generic_kernel<<<n, kGenThreads, 0, stream>>>(prev, next, prev_pts, next_pts, status, iters, n, max_level, use_initial_flow ? 1 : 0, win);
}

}  // namespace vf
