// Iterative pyramidal Lucas-Kanade for sm_100a: forward + reverse (flow_back) temporal tracking and
// left->right->left stereo tracking, one CTA per feature, one launch per pair of LK calls.
//
// Replaces the four cv::calcOpticalFlowPyrLK calls of the reference
// (vins/estimator/feature_tracker.cpp:41-42, 50-53, 117-119, 125-127) and the status logic around them
// (:55-67, :128-133), plus RemoveDistortion / ComputeKpsVelocity (:310-363) in the stereo epilogue.
//
// Arithmetic = OpenCV's LKTrackerInvoker (modules/video/src/lkpyramid.cpp), restated in oracle/vf_oracle.c:
//   * 21x21 bilinear patch of I and of its Scharr derivative in W_BITS=14 fixed point (integers);
//   * A11,A12,A22 and b1,b2 are FLOAT sums accumulated in the exact order OpenCV's 128-bit SIMD loop uses
//     (4 / 8 interleaved lane accumulators over columns 0..15, a scalar tail over columns 16..20, then the
//     v_reduce_sum tree), so positions and status flags are bit-identical to the OpenCV CPU path instead of
//     "within 0.01 px".  The terms of those sums are produced by all 128 threads; one warp then runs the
//     order-dependent chains (<= 105 dependent FADDs), which is the latency floor of an iteration.
//   * the Scharr derivative is never materialised in HBM: each level stages a 24x24 uint8 tile of I in shared
//     memory and derives the 22x22 (Ix,Iy) tile from it (zero outside the image, reflect-101 inside), so the
//     kernel reads only the uint8 pyramids.
// All float expressions use explicit round-to-nearest intrinsics (no FMA contraction).
#include <float.h>

#include "vf_camera.cuh"
#include "vf_common.cuh"
#include "vf_kernels.h"

namespace vf {

constexpr int kLkThreads = 256;
constexpr int kTileI = 24;                   // staged I tile (window + bilinear tap + Scharr ring)
constexpr int kTileD = 22;                   // derivative tile (window + bilinear tap)
constexpr int kJMargin = 5;                  // staged J tile = 22x22 taps + this margin on every side
constexpr int kTileJ = 22 + 2 * kJMargin;    // 32
constexpr int kTail = 105;                   // 21 rows x 5 tail columns
constexpr int kChain = 108;                  // every summation chain is padded with zeros to 27 float4
constexpr int kChainsA = 15;                 // 3 x 4 SIMD-lane accumulators (84 terms) + 3 scalar tails (105 terms)
constexpr int kChainsB = 10;                 // 8 SIMD-lane accumulators (42 terms) + 2 scalar tails (105 terms)

struct LevelRef {
    const uint8_t* p;
    int w, h, pitch;
};

struct LkSmem {
    alignas(16) float termsA[kChainsA * kChain];
    alignas(16) float termsB[kChainsB * kChain];
    uint8_t tileIs[kMaxLevels][kTileI * kTileI];   // I tiles of every level of the current pass (prefetched together)
    uint8_t tileJ[kTileJ * kTileJ];
    short2 dtile[kTileD * kTileD];
    short Iw[kWinPx + 1];
    short2 dI[kWinPx];
    LevelRef pyr[2][kMaxLevels];
    float bcast[4];
    int found[2];
    int part[8][10];     // per-warp exact integer partial sums of the 10 chains of b
    int bound[8];        // per-warp sum |term|
    int fast_iters;      // iterations whose chains were provably exact integers (statistics)
};

#define VF_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

// Optional phase timing (build with -DVF_LK_TIMING): thread 0 of every CTA accumulates clock64() deltas per phase and
// adds them to vf_lk_timing[k] (cycles) / vf_lk_timing[16 + k] (visits).  Measurement aid only; the marks split lane 0
// from its warp, which inflates whatever follows a mark until the next reconvergence.
#ifdef VF_LK_TIMING
__device__ unsigned long long vf_lk_timing[32];
#define LK_T_DECL long long lk_t_last = clock64();
#define LK_T(k) do { if (threadIdx.x == 0) { const long long t__ = clock64(); atomicAdd(&vf_lk_timing[k], (unsigned long long)(t__ - lk_t_last)); \
                     atomicAdd(&vf_lk_timing[16 + (k)], 1ull); lk_t_last = t__; } } while (0)
#else
#define LK_T_DECL
#define LK_T(k) do { } while (0)
#endif

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float one_a = __fsub_rn(1.f, a), one_b = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, one_b), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, one_b), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// Zero the padding slots of the summation chains once per kernel: the terms never overwrite them, and adding +0.0f
// to a chain value is an exact identity (a chain of integer-valued floats can never be -0.0f).
__device__ __forceinline__ void lk_smem_init(LkSmem& sm) {
    for (int i = threadIdx.x; i < kChainsA * kChain; i += kLkThreads) sm.termsA[i] = 0.f;
    for (int i = threadIdx.x; i < kChainsB * kChain; i += kLkThreads) sm.termsB[i] = 0.f;
    if (threadIdx.x < 80) (&sm.part[0][0])[threadIdx.x] = 0;
    if (threadIdx.x == 0) sm.fast_iters = 0;
}

// Sequential float chain acc = (((0 + t0) + t1) + ...) in the order OpenCV's accumulators see the terms: 27 vector
// loads, 108 dependent FADDs (the latency floor of one LK iteration).
// Not inlined: one copy serves A and b, which keeps the per-iteration code inside the instruction cache.
__device__ __noinline__ float chain_sum(const float* src) {
    const float4* p = reinterpret_cast<const float4*>(src);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kChain / 4; ++i) {
        const float4 v = p[i];
        acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
    }
    return acc;
}

// Stage a TS x TS tile of `img` with origin (x0, y0), reflect-101 outside the image.
// Reflect-101 for tile coordinates: two branch-free folds cover every index within 2n-2 of the image
// (tiles reach at most 26 pixels outside a level, levels are at least 22 pixels wide).
__device__ __forceinline__ int refl101_tile(int i, int n) {
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    return i;
}

// Stage a TS x TS tile of a level with origin (x0, y0), reflect-101 outside the image.  All loads of a thread are
// issued before its first store (one memory round trip per tile).  Not inlined (code size, see lk_track_point).
template <int TS>
__device__ __noinline__ void stage_tile(uint8_t* tile, const LevelRef& L, int x0, int y0) {
    constexpr int PER = (TS * TS + kLkThreads - 1) / kLkThreads;
    uint8_t v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * kLkThreads;
        const int ty = i / TS, tx = i - ty * TS;
        const int gx = refl101_tile(x0 + tx, L.w), gy = refl101_tile(min(y0 + ty, y0 + TS - 1), L.h);
        v[k] = __ldg(L.p + (size_t)gy * L.pitch + gx);
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = threadIdx.x + k * kLkThreads;
        if (i < TS * TS) tile[i] = v[k];
    }
}

// One calcOpticalFlowPyrLK for one point, executed cooperatively by the CTA (all threads pass identical
// arguments and return identical results).  ia / ja select the source and target pyramids in sm.pyr.
// Not inlined: the forward and the reverse pass of a kernel share one copy of this code (the kernel is latency-bound with
// one CTA per SM, so instruction-cache misses are fully exposed; the body must stay well under the 32 KB L1.5 I-cache).
__device__ __noinline__ void lk_track_point(LkSmem& sm, int ia, int ja, float2 prev_pt, float2 init_pt, bool use_init,
                                            int max_level, float2& out_pt, int& out_status, int& iters) {
    const int tid = threadIdx.x;
    const float half = 10.f;
    const float FLT_SCALE = 1.f / (1 << 20);
    int status = 1;
    float outx = init_pt.x, outy = init_pt.y;
    LK_T_DECL

    // The I tiles depend on prev_pt only: fetch them for every level of this pass now, in one burst of loads.
    __syncthreads();   // the previous pass may still be reading the shared buffers
#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        const LevelRef LI = sm.pyr[ia][level];
        const float scale = 1.f / (float)(1 << level);
        const int ipx = __float2int_rd(__fsub_rn(__fmul_rn(prev_pt.x, scale), half));
        const int ipy = __float2int_rd(__fsub_rn(__fmul_rn(prev_pt.y, scale), half));
        if (ipx < -kWin || ipx >= LI.w || ipy < -kWin || ipy >= LI.h) continue;
        stage_tile<kTileI>(sm.tileIs[level], LI, ipx - 1, ipy - 1);
    }
    LK_T(7);

#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        const LevelRef LI = sm.pyr[ia][level], LJ = sm.pyr[ja][level];
        const int cols = LI.w, rows = LI.h, jcols = LJ.w, jrows = LJ.h;
        const float scale = 1.f / (float)(1 << level);
        float ppx = __fmul_rn(prev_pt.x, scale), ppy = __fmul_rn(prev_pt.y, scale);
        float nx, ny;
        if (level == max_level) {
            if (use_init) { nx = __fmul_rn(outx, scale); ny = __fmul_rn(outy, scale); }
            else { nx = ppx; ny = ppy; }
        } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;

        ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -kWin || ipx >= cols || ipy < -kWin || ipy >= rows) {
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        bilinear_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);

        // ---- stage a 32x32 tile of J around the start of the search (the 24x24 tile of I is already resident) ----
        int jx0 = __float2int_rd(__fsub_rn(nx, half)) - kJMargin, jy0 = __float2int_rd(__fsub_rn(ny, half)) - kJMargin;
        const uint8_t* tileI = sm.tileIs[level];
        __syncthreads();   // the previous level may still be reading tileJ / the patch buffers
        stage_tile<kTileJ>(sm.tileJ, LJ, jx0, jy0);
        __syncthreads();
        LK_T(0);
        // ---- Scharr (Ix,Iy) on the 22x22 tap positions; zero outside the image ----
#pragma unroll 1
        for (int i = tid; i < kTileD * kTileD; i += kLkThreads) {
            const int ty = i / kTileD, tx = i - ty * kTileD;
            const int x = ipx + tx, y = ipy + ty;
            short2 d = make_short2(0, 0);
            if (x >= 0 && x < cols && y >= 0 && y < rows) {
                const uint8_t* p = &tileI[(ty + 1) * kTileI + (tx + 1)];
                const int a00 = p[-kTileI - 1], a01 = p[-kTileI], a02 = p[-kTileI + 1];
                const int a10 = p[-1], a12 = p[1];
                const int a20 = p[kTileI - 1], a21 = p[kTileI], a22 = p[kTileI + 1];
                const int t0m = (a00 + a20) * 3 + a10 * 10, t0p = (a02 + a22) * 3 + a12 * 10;
                const int t1m = a20 - a00, t1c = a21 - a01, t1p = a22 - a02;
                d = make_short2((short)(t0p - t0m), (short)((t1m + t1p) * 3 + t1c * 10));
            }
            sm.dtile[i] = d;
        }
        __syncthreads();
        LK_T(1);
        // ---- 21x21 bilinear patch (I in 2^5, derivatives in Scharr units) + terms of A ----
#pragma unroll 1
        for (int i = tid; i < kWinPx; i += kLkThreads) {
            const int y = i / kWin, x = i - y * kWin;
            const uint8_t* s = &tileI[(y + 1) * kTileI + (x + 1)];
            const int ival = VF_DESCALE(s[0] * w00 + s[1] * w01 + s[kTileI] * w10 + s[kTileI + 1] * w11, 9);
            const short2 d00 = sm.dtile[y * kTileD + x], d01 = sm.dtile[y * kTileD + x + 1];
            const short2 d10 = sm.dtile[(y + 1) * kTileD + x], d11 = sm.dtile[(y + 1) * kTileD + x + 1];
            const int ixval = VF_DESCALE(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, 14);
            const int iyval = VF_DESCALE(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, 14);
            sm.Iw[i] = (short)ival;
            sm.dI[i] = make_short2((short)ixval, (short)iyval);
            const float fx = (float)ixval, fy = (float)iyval;   // products are < 2^24: exact in float
            const float t11 = __fmul_rn(fx, fx), t12 = __fmul_rn(fx, fy), t22 = __fmul_rn(fy, fy);
            if (x < 16) {   // SIMD lane l = x & 3 sees columns l, 4+l, 8+l, 12+l of every row in turn
                const int l = x & 3, slot = y * 4 + (x >> 2);
                sm.termsA[(0 + l) * kChain + slot] = t11;
                sm.termsA[(4 + l) * kChain + slot] = t12;
                sm.termsA[(8 + l) * kChain + slot] = t22;
            } else {        // scalar tail: columns 16..20 of every row in turn
                const int slot = y * 5 + (x - 16);
                sm.termsA[12 * kChain + slot] = t11;
                sm.termsA[13 * kChain + slot] = t12;
                sm.termsA[14 * kChain + slot] = t22;
            }
        }
        __syncthreads();
        LK_T(2);
        if (tid < 32) {
            const int lane = tid;
            // Converge the warp before the chain and before the shuffles: a warp that arrives split (e.g. lane 0 after the
            // timing marks) would run the chain once per group and take the WARPSYNC.COLLECTIVE path on every shuffle.
            __syncwarp();
            const float acc = chain_sum(&sm.termsA[min(lane, kChainsA - 1) * kChain]);
            __syncwarp();
            float r[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const float q0 = __shfl_sync(0xffffffffu, acc, m * 4 + 0), q1 = __shfl_sync(0xffffffffu, acc, m * 4 + 1);
                const float q2 = __shfl_sync(0xffffffffu, acc, m * 4 + 2), q3 = __shfl_sync(0xffffffffu, acc, m * 4 + 3);
                const float tail = __shfl_sync(0xffffffffu, acc, 12 + m);
                r[m] = __fadd_rn(tail, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));   // iA += v_reduce_sum(qA)
            }
            if (lane == 0) { sm.bcast[0] = r[0]; sm.bcast[1] = r[1]; sm.bcast[2] = r[2]; }
        }
        __syncthreads();
        LK_T(3);
        const float A11 = __fmul_rn(sm.bcast[0], FLT_SCALE), A12 = __fmul_rn(sm.bcast[1], FLT_SCALE),
                    A22 = __fmul_rn(sm.bcast[2], FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * kWin * kWin));
        if (minEig < 1e-4f || D < FLT_EPSILON) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
#pragma unroll 1
        for (int j = 0; j < 30; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -kWin || inx >= jcols || iny < -kWin || iny >= jrows) {
                if (level == 0) status = 0;
                break;
            }
            ++iters;
            if (inx < jx0 || iny < jy0 || inx + kTileD > jx0 + kTileJ || iny + kTileD > jy0 + kTileJ) {
                // the window left the staged tile: re-centre it (uniform branch)
                jx0 = inx - kJMargin; jy0 = iny - kJMargin;
                __syncthreads();
                stage_tile<kTileJ>(sm.tileJ, LJ, jx0, jy0);
                __syncthreads();
            }
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            LK_T(4);
            const uint8_t* jt = &sm.tileJ[(iny - jy0) * kTileJ + (inx - jx0)];
            auto jdiff = [&](int y, int x) -> int {
                const uint8_t* p = jt + y * kTileJ + x;
                const int v = p[0] * w00 + p[1] * w01 + p[kTileJ] * w10 + p[kTileJ + 1] * w11;
                return VF_DESCALE(v, 9) - (int)sm.Iw[y * kWin + x];
            };
            // terms of b, one work item per thread: threads 0..167 take the 168 column pairs (c, c+4) of the two 8-wide
            // SIMD groups of every row (pmaddwd pair sums), threads 192..244 take 53 pairs of tail pixels (columns 16..20).
            // Besides the float terms of the ordered chains, every chain is also summed as an exact integer together with
            // a bound on sum|term|: when the bound is below 2^24 every partial sum of every float chain is an exactly
            // representable integer, the order of the additions cannot matter, and the chain value IS the integer sum --
            // the 105-deep dependent FADD chain is skipped (bit-identical result).
            int tx = 0, ty = 0, ab = 0;
            if (tid < 168) {
                const int y = tid >> 3, g = (tid >> 2) & 1, c = tid & 3;
                const int xa = 8 * g + c, xb = xa + 4;
                const int da = jdiff(y, xa), db = jdiff(y, xb);
                const short2 pa = sm.dI[y * kWin + xa], pb = sm.dI[y * kWin + xb];
                const int slot = y * 2 + g;
                tx = da * pa.x + db * pb.x; ty = da * pa.y + db * pb.y;             // exact int32 pair sums
                ab = abs(tx) + abs(ty);
                sm.termsB[(2 * c + 0) * kChain + slot] = (float)tx;
                sm.termsB[(2 * c + 1) * kChain + slot] = (float)ty;
            } else if (tid >= 192 && tid < 192 + 53) {
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int t = 2 * (tid - 192) + h2;
                    if (t < kTail) {
                        const int y = t / 5, x = 16 + (t - y * 5);
                        const int d = jdiff(y, x);
                        const short2 pd = sm.dI[y * kWin + x];
                        const int ex = d * pd.x, ey = d * pd.y;
                        tx += ex; ty += ey; ab += abs(ex) + abs(ey);
                        sm.termsB[8 * kChain + t] = (float)ex;
                        sm.termsB[9 * kChain + t] = (float)ey;
                    }
                }
            }
            {
                const int wid = tid >> 5, lane = tid & 31;
                ab = __reduce_add_sync(0xffffffffu, ab);
                if (wid < 6) {          // pair warps: lanes with equal (lane & 3) feed the same two chains
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        tx += __shfl_xor_sync(0xffffffffu, tx, o);
                        ty += __shfl_xor_sync(0xffffffffu, ty, o);
                    }
                    if (lane < 4) { sm.part[wid][2 * lane] = tx; sm.part[wid][2 * lane + 1] = ty; }
                } else {                // tail warps
                    tx = __reduce_add_sync(0xffffffffu, tx);
                    ty = __reduce_add_sync(0xffffffffu, ty);
                    if (lane == 0) { sm.part[wid][8] = tx; sm.part[wid][9] = ty; }
                }
                if (lane == 0) sm.bound[wid] = ab;
            }
            __syncthreads();
            LK_T(5);
            if (tid < 32) {
                const int lane = tid;
                __syncwarp();
                int isum = 0, bnd = 0;
                if (lane < 8) {
#pragma unroll
                    for (int w = 0; w < 6; ++w) isum += sm.part[w][lane];
                } else if (lane < 10) {
                    isum = sm.part[6][lane] + sm.part[7][lane];
                }
#pragma unroll
                for (int w = 0; w < 8; ++w) bnd += sm.bound[w];
                float acc;
                if (bnd < (1 << 24)) acc = (float)isum;                                   // exact: order-independent
                else acc = chain_sum(&sm.termsB[min(lane, kChainsB - 1) * kChain]);       // OpenCV's summation order
                __syncwarp();
                float k[10];
#pragma unroll
                for (int q = 0; q < 10; ++q) k[q] = __shfl_sync(0xffffffffu, acc, q);
                // qb0+qb1 -> (s0,s1,s2,s3); ib1 += (s0+0)+(s2+0); ib2 += (s1+0)+(s3+0)
                const float s0 = __fadd_rn(k[0], k[4]), s1 = __fadd_rn(k[1], k[5]);
                const float s2 = __fadd_rn(k[2], k[6]), s3 = __fadd_rn(k[3], k[7]);
                if (lane == 0) {
                    sm.bcast[0] = __fadd_rn(k[8], __fadd_rn(s0, s2));
                    sm.bcast[1] = __fadd_rn(k[9], __fadd_rn(s1, s3));
                    if (bnd < (1 << 24)) ++sm.fast_iters;
                }
            }
            __syncthreads();
            LK_T(6);
            const float b1 = __fmul_rn(sm.bcast[0], FLT_SCALE), b2 = __fmul_rn(sm.bcast[1], FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            const double dd = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy));
            if (dd <= 0.01 * 0.01) break;
            if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {   // the caller passes an err vector: final window test (lkpyramid.cpp)
            const int ix = __float2int_rd(__fsub_rn(outx, half)), iy = __float2int_rd(__fsub_rn(outy, half));
            if (ix < -kWin || ix >= jcols || iy < -kWin || iy >= jrows) status = 0;
        }
    }
    out_pt = make_float2(outx, outy);
    out_status = status;
}

// Fill sm.pyr[slot] with stream s of a pyramid (threads 0..kMaxLevels-1 of the given quarter).
__device__ __forceinline__ void load_pyr(LkSmem& sm, int slot, const PyrView& pv, const size_t* stride, int s) {
    const int l = threadIdx.x - slot * kMaxLevels;
    if (l >= 0 && l < kMaxLevels) {
        LevelRef r;
        const bool ok = l < pv.n_levels;
        r.p = ok ? pv.lvl[l] + (stride ? (size_t)s * stride[l] : 0) : nullptr;
        r.w = ok ? pv.w[l] : 0; r.h = ok ? pv.h[l] : 0; r.pitch = ok ? pv.pitch[l] : 0;
        sm.pyr[slot][l] = r;
    }
}

__device__ __forceinline__ bool in_border(float2 p, int rows, int cols) {   // feature_tracker.cpp:20-25
    const int x = __float2int_rn(p.x), y = __float2int_rn(p.y);
    return 1 <= x && x < cols - 1 && 1 <= y && y < rows - 1;
}
__device__ __forceinline__ double pt_distance(float2 a, float2 b) {         // feature_tracker.cpp:457-462
    const double dx = (double)__fsub_rn(a.x, b.x), dy = (double)__fsub_rn(a.y, b.y);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// ---- temporal: prev -> cur (maxLevel L) then cur -> prev (maxLevel 1, initial flow), :37-67 ----
__global__ void __launch_bounds__(kLkThreads) lk_temporal_kernel(const DevBatch* __restrict__ bp, int parity) {
    __shared__ LkSmem sm;
    const DevBatch& b = *bp;
    const int s = blockIdx.y, i = blockIdx.x, cap = b.cap;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    if (i >= n_prev) return;
    lk_smem_init(sm);
    load_pyr(sm, 0, b.left[parity ^ 1], b.lvl_stream_stride, s);   // 0 = previous left
    load_pyr(sm, 1, b.left[parity], b.lvl_stream_stride, s);       // 1 = current left
    __syncthreads();
    const float2 p0 = b.kps[parity ^ 1][(size_t)s * cap + i];
    float2 fwd, rev = p0;
    int st = 0, rst = 0, iters = 0;
    const int L = min(b.lk_max_level, b.left[parity].n_levels - 1);
    lk_track_point(sm, 0, 1, p0, p0, false, L, fwd, st, iters);
    int ok = st;
    if (b.flow_back) {
        // the reference also runs the reverse pass on points whose forward status is already 0 and then
        // discards the result (:55-59); the device skips that work, the outcome is the same.
        if (st) lk_track_point(sm, 1, 0, fwd, p0, true, min(1, L), rev, rst, iters);
        ok = st && rst && pt_distance(p0, rev) <= 0.5;
    }
    if (ok && !in_border(fwd, b.H, b.W)) ok = 0;
    if (threadIdx.x == 0) {
        const size_t o = (size_t)s * cap + i;
        b.fwd_pts[o] = fwd; b.fwd_st[o] = (uint8_t)st;
        b.rev_pts[o] = rev; b.rev_st[o] = (uint8_t)rst;
        b.temporal_st[o] = (uint8_t)ok;
        b.lk_iters[o] = iters | (sm.fast_iters << 16);   // low 16 bits: iterations, high: of which on the exact-integer path
    }
}

// velocity of ComputeKpsVelocity (:323-363): (cur - prev) as float, / dt as double, stored as float.
__device__ __forceinline__ float2 velocity(float2 cur, float2 prev, double dt) {
    return make_float2((float)((double)__fsub_rn(cur.x, prev.x) / dt), (float)((double)__fsub_rn(cur.y, prev.y) / dt));
}

// ---- stereo: left -> right then right -> left (both maxLevel L), status (:112-135), then
//      RemoveDistortion + ComputeKpsVelocity for both cameras and the packed featureFrame record ----
__global__ void __launch_bounds__(kLkThreads) lk_stereo_kernel(const DevBatch* __restrict__ bp, int parity) {
    __shared__ LkSmem sm;
    const DevBatch& b = *bp;
    const int s = blockIdx.y, i = blockIdx.x, cap = b.cap, tid = threadIdx.x;
    const int* cnt_prev = b.counters[parity ^ 1] + s * C_COUNT;
    const int n = b.counters[parity][s * C_COUNT + C_N_TOTAL];
    if (i >= n) return;
    const size_t o = (size_t)s * cap + i;
    lk_smem_init(sm);
    load_pyr(sm, 0, b.left[parity], b.lvl_stream_stride, s);   // 0 = current left
    load_pyr(sm, 1, b.right, b.lvl_stream_stride, s);          // 1 = current right
    __syncthreads();
    const float2 p0 = b.kps[parity][o];
    const int id = b.ids[parity][o];
    float2 rp, back = p0;
    int st = 0, rst = 0, iters = 0;
    const int L = min(b.lk_max_level, b.right.n_levels - 1);
    lk_track_point(sm, 0, 1, p0, p0, false, L, rp, st, iters);
    int ok = st;
    if (b.flow_back) {
        if (st) lk_track_point(sm, 1, 0, rp, rp, false, L, back, rst, iters);
        ok = st && rst && in_border(rp, b.H, b.W) && pt_distance(p0, back) <= 0.5;
    }
    // id lookup in the previous frame's maps (previous_left/right_un_distortion_ids_kps_map_)
    if (tid < 2) sm.found[tid] = -1;
    __syncthreads();
    const int n_prev = cnt_prev[C_N_TOTAL];
    const int* prev_ids = b.ids[parity ^ 1] + (size_t)s * cap;
    for (int j = tid; j < n_prev; j += kLkThreads)
        if (prev_ids[j] == id) { sm.found[0] = j; if (b.right_ok[parity ^ 1][(size_t)s * cap + j]) sm.found[1] = j; }
    __syncthreads();
    if (tid == 0) {
        const double dt = b.times[parity][s] - b.times[parity ^ 1][s];
        vf_feature rec;
        rec.id = id; rec.track_cnt = b.cnt[parity][o]; rec.has_right = ok;
        const float2 un = undistort_point(b.cam[0], p0);
        float2 v = make_float2(0.f, 0.f);
        if (sm.found[0] >= 0) v = velocity(un, b.un_left[parity ^ 1][(size_t)s * cap + sm.found[0]], dt);
        rec.x = un.x; rec.y = un.y; rec.u = p0.x; rec.v = p0.y; rec.vx = v.x; rec.vy = v.y;
        rec.xr = rec.yr = rec.ur = rec.vr = rec.vxr = rec.vyr = 0.f;
        float2 unr = make_float2(0.f, 0.f);
        if (ok) {
            unr = undistort_point(b.cam[1], rp);
            float2 vr = make_float2(0.f, 0.f);
            if (sm.found[1] >= 0) vr = velocity(unr, b.un_right[parity ^ 1][(size_t)s * cap + sm.found[1]], dt);
            rec.xr = unr.x; rec.yr = unr.y; rec.ur = rp.x; rec.vr = rp.y; rec.vxr = vr.x; rec.vyr = vr.y;
        }
        b.out[o] = rec;
        b.un_left[parity][o] = un;
        b.un_right[parity][o] = unr;
        b.right_ok[parity][o] = (uint8_t)ok;
        b.lr_pts[o] = rp; b.lr_st[o] = (uint8_t)st; b.rl_pts[o] = back; b.rl_st[o] = (uint8_t)rst;
        b.lk_iters_stereo[o] = iters | (sm.fast_iters << 16);
        if (ok) atomicAdd(&b.counters[parity][s * C_COUNT + C_N_RIGHT], 1);
    }
}

// ---- stand-alone calcOpticalFlowPyrLK (vf_op_lk) ----
__global__ void __launch_bounds__(kLkThreads) lk_op_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts,
                                                           float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                           int* __restrict__ iters_out, int n, int max_level, int use_init) {
    __shared__ LkSmem sm;
    const int i = blockIdx.x;
    if (i >= n) return;
    lk_smem_init(sm);
    load_pyr(sm, 0, prev, nullptr, 0);
    load_pyr(sm, 1, next, nullptr, 0);
    __syncthreads();
    float2 out;
    int st = 0, iters = 0;
    const int L = min(max_level, min(prev.n_levels, next.n_levels) - 1);
    lk_track_point(sm, 0, 1, prev_pts[i], next_pts[i], use_init != 0, L, out, st, iters);
    if (threadIdx.x == 0) {
        next_pts[i] = out; status[i] = (uint8_t)st;
        if (iters_out) iters_out[i] = iters;
    }
}

__global__ void undistort_kernel(CamParams cam, const float2* __restrict__ uv, float2* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = undistort_point(cam, uv[i]);
}
void launch_undistort(const CamParams& cam, const float2* uv, float2* out, int n, cudaStream_t stream) {
    if (n <= 0) return;
    undistort_kernel<<<(n + 127) / 128, 128, 0, stream>>>(cam, uv, out, n);
}

void launch_lk_temporal(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream) {
    dim3 grid(h.cap, h.n_streams);
    lk_temporal_kernel<<<grid, kLkThreads, 0, stream>>>(d_batch, parity);
}
void launch_lk_stereo(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream) {
    dim3 grid(h.cap, h.n_streams);
    lk_stereo_kernel<<<grid, kLkThreads, 0, stream>>>(d_batch, parity);
}
void launch_lk_op(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status,
                  int* iters, int n, int max_level, int use_initial_flow, cudaStream_t stream) {
    if (n <= 0) return;
    lk_op_kernel<<<n, kLkThreads, 0, stream>>>(prev, next, prev_pts, next_pts, status, iters, n, max_level, use_initial_flow);
}

}  // namespace vf

#ifdef VF_LK_TIMING
// measurement aid (timing build only): read and clear the per-phase cycle counters
extern "C" int vf_debug_lk_timing(unsigned long long* out16) {
    unsigned long long zero[32] = {0};
    if (cudaMemcpyFromSymbol(out16, vf::vf_lk_timing, sizeof(zero)) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(vf::vf_lk_timing, zero, sizeof(zero)) != cudaSuccess;
}
#endif
