// Iterative pyramidal Lucas-Kanade for sm_100a: forward + reverse (flow_back) temporal tracking and
// left->right->left stereo tracking, one CTA per feature, one launch per pair of LK calls.
//
// Replaces the four cv::calcOpticalFlowPyrLK calls of the reference
// (vins/estimator/feature_tracker.cpp:41-42, 50-53, 117-119, 125-127) and the status logic around them
// (:55-67, :128-133), plus the right camera's RemoveDistortion / ComputeKpsVelocity (:310-363) in the stereo epilogue.
//
// Arithmetic = OpenCV's LKTrackerInvoker (modules/video/src/lkpyramid.cpp), restated in oracle/vf_oracle.c:
//   * 21x21 bilinear patch of I and of its Scharr derivative in W_BITS=14 fixed point (integers);
//   * A11,A12,A22 and b1,b2 are FLOAT sums accumulated in the exact order OpenCV's 128-bit SIMD loop uses
//     (4 / 8 interleaved lane accumulators over columns 0..15, a scalar tail over columns 16..20, then the
//     v_reduce_sum tree), so positions and status flags are bit-identical to the OpenCV CPU path.
//
// The path is a dependent chain (levels x Gauss-Newton iterations), so the kernel is organised around latency:
//   * everything that depends only on the source point -- the 24x24 uint8 tiles of I on EVERY level, their Scharr
//     tiles, the interpolated patches Iw / dI and the 2x2 matrix A -- is computed in one prologue for all levels at
//     once (4 block barriers in total); the level loop is left with the J tile and the iterations;
//   * the float sums are integer-valued.  A chain whose terms satisfy sum|t| < 2^24 has only exactly representable
//     partial sums, so its value is the integer sum and the order of the additions is irrelevant: such chains are
//     reduced with REDUX (one instruction per warp) instead of 84..105 dependent FADDs.  One iteration then costs ONE
//     block barrier: every thread contributes two pixels, each warp publishes (sum_x, sum_y, sum|.|) and every
//     thread folds the 8 partials itself.  Only when the bound fails are the terms spilled to shared memory and
//     each chain examined on its own (exact integer value if ITS bound holds, else the ordered float chain);
//   * the Scharr derivative is never materialised in HBM: the kernel reads only the uint8 pyramids.
// All float expressions use explicit round-to-nearest intrinsics (no FMA contraction).
#include <float.h>

#include "vf_camera.cuh"
#include "vf_common.cuh"
#include "vf_kernels.h"
#include "vf_lk_common.cuh"

namespace vf {

#ifndef VF_LK_THREADS
#define VF_LK_THREADS 128                    // (A/B aid: 256 threads measured no faster -- the iteration is bound by its dependent chain)
#endif
constexpr int kLkThreads = VF_LK_THREADS;    // 4 warps: one per SM sub-partition; a second CTA (the other LK kernel of a
                                             // pipelined frame pair) interleaves with it instead of queueing behind it
constexpr int kLkWarps = kLkThreads / 32;
constexpr int kLkItems = 256;                // work items of a pass are numbered 0..255; thread t owns items t, t + kLkThreads, ...
constexpr int kIpt = kLkItems / kLkThreads;  // items per thread
constexpr int kTail = 105;                   // 21 rows x 5 tail columns
// Per-level quantities that depend on the source point only (computed once per pass by the prologue).
struct alignas(16) LevelGeo {
    int w00, w01, w10, w11;        // bilinear weights of the source point
    int ipx, ipy, w, h;            // top-left tap of the window; level size
    float A11, A12, A22, Dinv;     // spatial gradient matrix (scaled by 2^-20) and 1 / det
    int valid;                     // window intersects the padded level (lkpyramid.cpp bounds test)
    int skip;                      // minEig < 1e-4 or det < FLT_EPSILON: the level is skipped (status 0 on level 0)
    int pad[2];
};

constexpr int kChainA = 108;                 // slots of one summation chain of A (84 / 105 terms, zero padded)
constexpr int kChainsA = 15;                 // 3 matrices x (4 SIMD-lane accumulators + 1 scalar tail)
constexpr size_t kTermsABytes = sizeof(float) * kChainsA * kChainA;   // staged terms of A per level (optional, see lk_plan_staged)

// Per-level state of one pass, in dynamic shared memory (one block per pyramid level the pass uses).
struct LevelBlock {
    alignas(16) short2 dI[kWinPx + 3];               // interpolated (Ix, Iy)
    alignas(16) short Iw[kWinPx + 7];                // interpolated I, 2^5 fixed point
    alignas(16) short2 dtile[kTileD * kTileD];       // Scharr (Ix, Iy) at the 22x22 tap positions
    alignas(16) uint8_t tileI[kTileI * kTileI];      // 24x24 source pixels
    LevelGeo geo;
};

struct LkSmem {
    LevelRef pyr[2][kMaxLevels];
    alignas(16) uint8_t tileJ[2][kTileJ * kTileJ];
    alignas(16) int4 part[2][kLkWarps];   // per-warp (sum tx, sum ty, sum |t|) of an iteration, double-buffered
    alignas(16) float termsB[10 * kChainA];   // terms of b in accumulation order (chain-major, zero padded), spilled
                                              // only when the global exactness bound fails
    alignas(16) float chainB[12];
    alignas(16) int4 part2[kLkWarps][5];  // per-warp (sum x, sum y, sum |x|, sum |y|) of the four SIMD column chains and of the tail (lk_chain_sums)
    int found[2];
    int termsA_off;                       // byte offset (from the start of this struct) of the staged terms of A,
                                          // [level][kChainsA * kChainA] behind the level blocks, or 0: the chains then form their
                                          // terms on the fly (launches with many CTAs, see lk_plan_staged).  An offset rather than
                                          // a pointer, so that the accesses stay shared-memory instructions
#ifdef VF_LK_TIMING
    unsigned long long tacc[16], tcnt[16];
    int tbase;
#endif
    alignas(16) LevelBlock lv[1];         // [n_levels of the launch]
};

static size_t lk_smem_bytes(int n_levels, bool staged) {
    const size_t nl = (size_t)(n_levels > 1 ? n_levels : 1);
    return sizeof(LkSmem) + sizeof(LevelBlock) * (nl - 1) + (staged ? kTermsABytes * nl : 0);
}
// Staging the terms of A in accumulation order makes the chain phase of the prologue a pure load + add loop (0.8 us per
// kernel at 752x480, where the launch is one wave of 150 CTAs and only latency counts) but costs 6.5 KB of shared memory per
// level.  When the launch has more CTAs than fit at once and dropping the staging area at least doubles the resident CTAs
// per SM (six pyramid levels: 2 -> 5), the chains form their terms on the fly instead.
static bool lk_plan_staged(int n_levels, long long n_ctas) {
    const size_t sm_bytes = 228 * 1024, per_cta = 1024;
    const int regs_limit = 65536 / (96 * kLkThreads);   // 5 CTAs per SM by registers
    int staged = (int)(sm_bytes / (lk_smem_bytes(n_levels, true) + per_cta)), compact = (int)(sm_bytes / (lk_smem_bytes(n_levels, false) + per_cta));
    staged = staged < regs_limit ? staged : regs_limit; compact = compact < regs_limit ? compact : regs_limit;
    if (staged < 1) return false;
    return n_ctas <= (long long)staged * kNumSms || compact < 2 * staged;
}

#ifndef VF_LK_ABLATE
#define VF_LK_ABLATE 0
#endif
// Optional phase timing (build with -DVF_LK_TIMING): every thread reads the clock at a mark (no divergence), thread 0
// accumulates per-phase cycles in shared memory and flushes them to vf_lk_timing[k] / vf_lk_timing[16 + k] at the end.
#ifdef VF_LK_TIMING
__device__ unsigned long long vf_lk_timing[32];
__device__ unsigned long long vf_lk_cta_times[2][256][4];   // [kernel][cta]{start ns, end ns, smid, iters}
__device__ __forceinline__ unsigned long long vf_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int vf_smid() { unsigned int r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#define LK_T_DECL long long lk_t_last = clock64();
#define LK_T(k) do { const long long t__ = clock64(); if (threadIdx.x == 0) { sm.tacc[sm.tbase + (k)] += (unsigned long long)(t__ - lk_t_last); sm.tcnt[sm.tbase + (k)] += 1; } \
                     lk_t_last = t__; } while (0)
#else
#define LK_T_DECL
#define LK_T(k) do { } while (0)
#endif
#ifdef VF_LK_TIMING_FINE
#define LK_TF(k) LK_T(k)
#else
#define LK_TF(k) do { } while (0)
#endif

// Warp-level value of one summation chain whose `n` integer terms are produced by term(k), k = 0..n-1 in OpenCV's
// accumulation order.  Exact integer path when sum|t| < 2^24, else the ordered float chain (all lanes run it
// redundantly; they are in lock-step anyway).
template <typename F>
__device__ __forceinline__ float chain_value(int n, F term) {
    const int lane = threadIdx.x & 31;
    int s = 0;
    unsigned int a = 0u;     // sum|t| <= 105 * 2^25 < 2^32: unsigned, so it cannot wrap into a small value
    for (int k = lane; k < n; k += 32) {
        const int t = term(k);
        s += t;
        a += (unsigned int)abs(t);
    }
    s = __reduce_add_sync(0xffffffffu, s);
    a = __reduce_add_sync(0xffffffffu, a);
    if (a < (unsigned int)kExact) return (float)s;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < n; ++k) acc = __fadd_rn(acc, (float)term(k));
    return acc;
}

extern __shared__ __align__(16) unsigned char lk_smem_raw[];
__device__ __forceinline__ float* lk_terms_a(LkSmem&, int off) { return reinterpret_cast<float*>(lk_smem_raw + off); }   // LkSmem sits at offset 0

// Cold path of an iteration (the global exactness bound failed, ~8 % of the iterations): spill the terms in OpenCV's
// accumulation order and run the 10 ordered float chains, one lane per chain.  Returns (ib1, ib2).  Out of line on purpose.
struct LkTerms { int ex0[kIpt], ey0[kIpt], ex1[kIpt], ey1[kIpt]; };   // by value: an array passed by address would live in local memory
__device__ __noinline__ float2 lk_spill_chains(LkSmem& sm, const LkTerms tm) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int (&ex0)[kIpt] = tm.ex0; const int (&ey0)[kIpt] = tm.ey0; const int (&ex1)[kIpt] = tm.ex1; const int (&ey1)[kIpt] = tm.ey1;
#pragma unroll
    for (int q = 0; q < kIpt; ++q) {
        const int it = tid + q * kLkThreads;
        if (it < 168) {
            const int y = it >> 3, g = (it >> 2) & 1, c = it & 3, slot = y * 2 + g;
            sm.termsB[(2 * c + 0) * kChainA + slot] = (float)(ex0[q] + ex1[q]);      // v_cvt_f32 of the pmaddwd pair sum
            sm.termsB[(2 * c + 1) * kChainA + slot] = (float)(ey0[q] + ey1[q]);
        } else if (it >= 192 && it < 192 + 53) {
            const int t = 2 * (it - 192);
            sm.termsB[8 * kChainA + t] = (float)ex0[q]; sm.termsB[9 * kChainA + t] = (float)ey0[q];
            if (t + 1 < kTail) { sm.termsB[8 * kChainA + t + 1] = (float)ex1[q]; sm.termsB[9 * kChainA + t + 1] = (float)ey1[q]; }
        }
    }
    __syncthreads();
    if (wid == 0) {
        const float4* src = reinterpret_cast<const float4*>(sm.termsB + min(lane, 9) * kChainA);
        float acc = 0.f;
        if (lane < 8) {           // SIMD-lane chains: 42 terms (11 vector loads, zero padded)
#pragma unroll
            for (int i = 0; i < 11; ++i) {
                const float4 v = src[i];
                acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
            }
        } else {                  // scalar tails: 105 terms
#pragma unroll
            for (int i = 0; i < kChainA / 4; ++i) {
                const float4 v = src[i];
                acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
            }
        }
        if (lane < 10) sm.chainB[lane] = acc;
    }
    __syncthreads();
    const float4 k0 = *reinterpret_cast<const float4*>(sm.chainB), k1 = *reinterpret_cast<const float4*>(sm.chainB + 4);
    const float2 k2 = *reinterpret_cast<const float2*>(sm.chainB + 8);
    // qb0+qb1 -> (s0,s1,s2,s3); ib1 += (s0+0)+(s2+0); ib2 += (s1+0)+(s3+0)
    const float s0 = __fadd_rn(k0.x, k1.x), s1 = __fadd_rn(k0.y, k1.y);
    const float s2 = __fadd_rn(k0.z, k1.z), s3 = __fadd_rn(k0.w, k1.w);
    return make_float2(__fadd_rn(k2.x, __fadd_rn(s0, s2)), __fadd_rn(k2.y, __fadd_rn(s1, s3)));
}

// Second exactness test of an iteration whose global bound failed, CHAIN BY CHAIN.  A chain of OpenCV's sums of b adds
// integer terms t_1 .. t_n (pmaddwd pair sums for the eight SIMD-lane chains, single products for the two scalar tails) into a
// float that starts at +0.  If sum |t_i| < 2^24 over the terms of THAT chain, every term and every partial sum is an integer of
// magnitude below 2^24, hence a float, hence every addition of the chain is exact and the chain's value is the integer sum --
// whatever the other chains do.  A chain holds a tenth (SIMD lanes: 42 pair sums) to a quarter (tails: 105 products) of the
// window's terms, so this test still passes when the whole window's sum is several times 2^24: on the benchmark sequence most
// of the iterations the global bound turns away (they are what the slowest features of a launch spend their time in).
// The ten chain values are then combined with the very float additions of lkpyramid.cpp (those need not be exact).
// Pair items belong to the chain of their column offset c = item & 3 = lane & 3 (items are numbered tid, tid + 128), tail items
// to the tails.  .z = 0 if a chain fails its bound: the caller then runs the ordered chains (lk_spill_chains).
struct LkChainIn { int px, py; unsigned int apx, apy; int tx, ty; unsigned int atx, aty; };
__device__ __noinline__ float4 lk_chain_sums(LkSmem& sm, const LkChainIn in) {   // (ib1, ib2, passed, -): by value, see LkTerms
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int cx = in.px, cy = in.py;
    unsigned int ax = in.apx, ay = in.apy;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {                 // lanes of equal column offset: 8 per warp
        cx += __shfl_xor_sync(0xffffffffu, cx, o); cy += __shfl_xor_sync(0xffffffffu, cy, o);
        ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o);
    }
    const int tx = __reduce_add_sync(0xffffffffu, in.tx), ty = __reduce_add_sync(0xffffffffu, in.ty);
    const unsigned int atx = __reduce_add_sync(0xffffffffu, in.atx), aty = __reduce_add_sync(0xffffffffu, in.aty);
    if (lane < 4) sm.part2[wid][lane] = make_int4(cx, cy, (int)ax, (int)ay);
    if (lane == 4) sm.part2[wid][4] = make_int4(tx, ty, (int)atx, (int)aty);
    __syncthreads();
    int sx[5], sy[5];
    unsigned int worst = 0u;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        unsigned int abx = 0u, aby = 0u;
        sx[c] = 0; sy[c] = 0;
#pragma unroll
        for (int w = 0; w < kLkWarps; ++w) {
            const int4 q = sm.part2[w][c];
            sx[c] += q.x; sy[c] += q.y; abx += (unsigned int)q.z; aby += (unsigned int)q.w;   // <= 105 * 2^25: no wrap
        }
        worst = max(worst, max(abx, aby));
    }
    if (worst >= (unsigned int)kExact) return make_float4(0.f, 0.f, 0.f, 0.f);
    // qb0 + qb1 -> (s0, s1, s2, s3); ib1 += s0 + s2; ib2 += s1 + s3  (as lk_spill_chains)
    const float s0 = __fadd_rn((float)sx[0], (float)sx[2]), s2 = __fadd_rn((float)sx[1], (float)sx[3]);
    const float s1 = __fadd_rn((float)sy[0], (float)sy[2]), s3 = __fadd_rn((float)sy[1], (float)sy[3]);
    return make_float4(__fadd_rn((float)sx[4], __fadd_rn(s0, s2)), __fadd_rn((float)sy[4], __fadd_rn(s1, s3)), 1.f, 0.f);
}

// Cold path of an iteration: the window left the staged tile of J, one warp re-stages it around the new origin.  Out of line.
__device__ __noinline__ void lk_restage_tile(LkSmem& sm, int ja, int level, int jbuf, int jx0, int jy0) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();    // everybody is done with the current tile
    if (wid == kLkWarps - 2) {
        uint32_t v[kTileJ / 4];
        load_row_words<kTileJ / 4>(sm.pyr[ja][level], jx0, jy0 + lane, v);
        uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[jbuf] + lane * kTileJ);
        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
        dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
}

// One calcOpticalFlowPyrLK for one point, executed cooperatively by the CTA (all threads pass identical
// arguments and return identical results).  ia / ja select the source and target pyramids in sm.pyr.
// Not inlined: the two passes of a kernel share one copy of this code.
// STAGED: the terms of A are staged in shared memory (sm.termsA_off != 0); two instantiations, so that neither variant of
// the prologue pays for the other (a launch runs one of them).
template <bool STAGED>
__device__ __noinline__ LkResult lk_track_point_impl(LkSmem& sm, int ia, int ja, float2 prev_pt, float2 init_pt, bool use_init,
                                                int max_level, int* iters_io) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float half = 10.f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int nlv = max_level + 1;
    int status = 1, iters = 0, fast = 0;
    float outx = init_pt.x, outy = init_pt.y;
    const int tA_off = sm.termsA_off;
    LK_T_DECL

    // ================= prologue: everything that depends on the source point only, all levels at once =================
    __syncthreads();   // the previous pass may still be reading the shared buffers
    // ---- P1: one item per tile row.  Items 0..24*nlv-1: rows of the 24x24 tiles of I (level = item / 24); items 192..223:
    //      rows of the top level's 32x32 tile of J (it only depends on the start point) ----
    int jx0, jy0, jbuf = 0;
    {
        const float scale = pow2_neg(max_level);
        const float sx = use_init ? init_pt.x : prev_pt.x, sy = use_init ? init_pt.y : prev_pt.y;
        const LevelRef LT = sm.pyr[ja][max_level];
        jx0 = clamp_tile_origin(__float2int_rd(__fsub_rn(__fmul_rn(sx, scale), half)) - kJMargin, LT.w);
        jy0 = clamp_tile_origin(__float2int_rd(__fsub_rn(__fmul_rn(sy, scale), half)) - kJMargin, LT.h);
    }
#pragma unroll
    for (int q = 0; q < kIpt; ++q) {
        const int it = tid + q * kLkThreads;
        if (it < kTileI * nlv) {
            const int level = it / kTileI, row = it - level * kTileI;
            const LevelRef LI = sm.pyr[ia][level];
            const float scale = pow2_neg(level);
            const float ppx = __fsub_rn(__fmul_rn(prev_pt.x, scale), half), ppy = __fsub_rn(__fmul_rn(prev_pt.y, scale), half);
            const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
            const int valid = !(ipx < -kWin || ipx >= LI.w || ipy < -kWin || ipy >= LI.h);
            if (row == 0) {
                LevelGeo& g = sm.lv[level].geo;
                g.ipx = ipx; g.ipy = ipy; g.w = LI.w; g.h = LI.h; g.valid = valid;
                bilinear_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), g.w00, g.w01, g.w10, g.w11);
            }
            if (valid) {
                uint32_t v[kTileI / 4];
                load_row_words<kTileI / 4>(LI, ipx - 1, ipy - 1 + row, v);
                uint32_t* dst = reinterpret_cast<uint32_t*>(sm.lv[level].tileI + row * kTileI);
#pragma unroll
                for (int k = 0; k < kTileI / 4; ++k) dst[k] = v[k];
            }
        } else if (it >= 192 && it < 224) {
            const int row = it - 192;
            uint32_t v[kTileJ / 4];
            load_row_words<kTileJ / 4>(sm.pyr[ja][max_level], jx0, jy0 + row, v);
            uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[0] + row * kTileJ);
            dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
            dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
    LK_T(0);
    // ---- P2: Scharr (Ix,Iy) on the 22x22 tap positions of every level; zero outside the image.
    //      A thread owns the same few tile positions on every level (level-invariant index arithmetic); per level it
    //      first issues the loads of all its positions, then computes, then stores -- the shared-memory stores would
    //      otherwise fence the next position's loads behind them ----
    {
        constexpr int R2 = (kTileD * kTileD + kLkThreads - 1) / kLkThreads;
        int toff[R2], txy[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) {
            const int e = min(tid + r * kLkThreads, kTileD * kTileD - 1);
            const int ty = e / kTileD, tx = e - ty * kTileD;
            toff[r] = (ty + 1) * kTileI + (tx + 1);
            txy[r] = tx | (ty << 8);
        }
#pragma unroll 1
        for (int level = 0; level < nlv; ++level) {
            LevelBlock& B = sm.lv[level];
            const int4 q = *reinterpret_cast<const int4*>(&B.geo.ipx);     // ipx, ipy, w, h
            int a[R2][8];
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const uint8_t* p = B.tileI + toff[r];                      // always readable; stale for skipped levels
                a[r][0] = p[-kTileI - 1]; a[r][1] = p[-kTileI]; a[r][2] = p[-kTileI + 1];
                a[r][3] = p[-1]; a[r][4] = p[1];
                a[r][5] = p[kTileI - 1]; a[r][6] = p[kTileI]; a[r][7] = p[kTileI + 1];
            }
            short2 d[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const int x = q.x + (txy[r] & 0xff), y = q.y + (txy[r] >> 8);
                const bool inside = (unsigned)x < (unsigned)q.z && (unsigned)y < (unsigned)q.w;
                const int t0m = (a[r][0] + a[r][5]) * 3 + a[r][3] * 10, t0p = (a[r][2] + a[r][7]) * 3 + a[r][4] * 10;
                const int t1m = a[r][5] - a[r][0], t1c = a[r][6] - a[r][1], t1p = a[r][7] - a[r][2];
                d[r] = make_short2((short)(inside ? t0p - t0m : 0), (short)(inside ? (t1m + t1p) * 3 + t1c * 10 : 0));
            }
#pragma unroll
            for (int r = 0; r < R2; ++r)
                if (tid + r * kLkThreads < kTileD * kTileD) B.dtile[tid + r * kLkThreads] = d[r];
        }
    }
    __syncthreads();
    LK_T(1);
    // ---- P3: 21x21 bilinear patches (I in 2^5, derivatives in Scharr units) of every level ----
    {
        constexpr int R3 = (kWinPx + kLkThreads - 1) / kLkThreads;
        int ioff[R3], doff[R3], aoff[R3];
#pragma unroll
        for (int r = 0; r < R3; ++r) {
            const int e = min(tid + r * kLkThreads, kWinPx - 1);
            const int y = e / kWin, x = e - y * kWin;
            ioff[r] = (y + 1) * kTileI + (x + 1); doff[r] = y * kTileD + x;
            aoff[r] = x < 16 ? (x & 3) * kChainA + y * 4 + (x >> 2) : 4 * kChainA + y * 5 + (x - 16);   // slot of the staged terms
        }
#pragma unroll 1
        for (int level = 0; level < nlv; ++level) {
            LevelBlock& B = sm.lv[level];
            const int4 w = *reinterpret_cast<const int4*>(&B.geo.w00);
            int sv[R3][4];
            short2 dv[R3][4];
#pragma unroll
            for (int r = 0; r < R3; ++r) {                                 // loads of all positions first (see P2)
                const uint8_t* s = B.tileI + ioff[r];
                sv[r][0] = s[0]; sv[r][1] = s[1]; sv[r][2] = s[kTileI]; sv[r][3] = s[kTileI + 1];
                const short2* dt = B.dtile + doff[r];
                dv[r][0] = dt[0]; dv[r][1] = dt[1]; dv[r][2] = dt[kTileD]; dv[r][3] = dt[kTileD + 1];
            }
#pragma unroll
            for (int r = 0; r < R3; ++r) {
                const int e = tid + r * kLkThreads;
                const int ival = VF_DESCALE(sv[r][0] * w.x + sv[r][1] * w.y + sv[r][2] * w.z + sv[r][3] * w.w, 9);
                const int ixval = VF_DESCALE(dv[r][0].x * w.x + dv[r][1].x * w.y + dv[r][2].x * w.z + dv[r][3].x * w.w, 14);
                const int iyval = VF_DESCALE(dv[r][0].y * w.x + dv[r][1].y * w.y + dv[r][2].y * w.z + dv[r][3].y * w.w, 14);
                if (e < kWinPx) {
                    B.Iw[e] = (short)ival;
                    B.dI[e] = make_short2((short)ixval, (short)iyval);
                    if (STAGED) {                                      // |ixval|, |iyval| <= 4080: products < 2^24, exact in float
                        float* ta = lk_terms_a(sm, tA_off) + level * (kChainsA * kChainA) + aoff[r];
                        ta[0] = (float)(ixval * ixval);
                        ta[5 * kChainA] = (float)(ixval * iyval);
                        ta[10 * kChainA] = (float)(iyval * iyval);
                    }
                }
            }
        }
    }
    __syncthreads();
    LK_T(2);
    // ---- P4: the 15 ordered float chains of A on every level, one LANE per chain (acc = ((0 + t0) + t1) + ...), in
    //      OpenCV's accumulation order: SIMD lane l = x & 3 sees columns l, 4+l, 8+l, 12+l of every row in turn, the scalar
    //      tail sees columns 16..20 of every row.  Either from the terms P3 staged in that order, or (launches with many CTAs,
    //      lk_plan_staged) formed on the fly from the interpolated derivatives: every lane then takes five steps per row and
    //      the fifth step of a four-column chain adds +0 (an exact identity). ----
    for (int level = wid; level < nlv; level += kLkWarps) {
        LevelBlock& B = sm.lv[level];
        const int ch = min(lane, kChainsA - 1), m = ch / 5, c = ch - 5 * m;   // matrix entry (xx, xy, yy), chain
        float acc = 0.f;
        if (STAGED) {               // staged: 27 vector loads and 108 dependent FADDs per lane (the padding slots hold +0)
            const float4* src = reinterpret_cast<const float4*>(lk_terms_a(sm, tA_off) + level * (kChainsA * kChainA) + ch * kChainA);
#pragma unroll
            for (int i = 0; i < kChainA / 4; ++i) {
                const float4 v = src[i];
                acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
            }
        } else {
            const int xstep = c < 4 ? 4 : 1;
            const short2* dIp = B.dI + (c < 4 ? c : 16);
#pragma unroll 3
            for (int y = 0; y < kWin; ++y) {
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const short2 d = dIp[y * kWin + t * xstep];               // columns c + 16 <= 19 and 16 + 4 = 20: in range
                    const int fa = m < 2 ? d.x : d.y, fb = m == 0 ? d.x : d.y;
                    acc = __fadd_rn(acc, (c < 4 && t == 4) ? 0.f : (float)(fa * fb));
                }
            }
        }
        // iA += v_reduce_sum(qA): tail + ((q0 + q2) + (q1 + q3)); then the 2x2 system of the level
        float r[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const float q0 = __shfl_sync(0xffffffffu, acc, 5 * m + 0), q1 = __shfl_sync(0xffffffffu, acc, 5 * m + 1);
            const float q2 = __shfl_sync(0xffffffffu, acc, 5 * m + 2), q3 = __shfl_sync(0xffffffffu, acc, 5 * m + 3);
            const float tl = __shfl_sync(0xffffffffu, acc, 5 * m + 4);
            r[m] = __fadd_rn(tl, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));
        }
        if (lane == 0) {
            const float A11 = __fmul_rn(r[0], FLT_SCALE), A12 = __fmul_rn(r[1], FLT_SCALE), A22 = __fmul_rn(r[2], FLT_SCALE);
            const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
            const float dA = __fsub_rn(A11, A22);
            const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
            const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * kWin * kWin));
            LevelGeo& g = B.geo;
            g.A11 = A11; g.A12 = A12; g.A22 = A22; g.Dinv = __fdiv_rn(1.f, D);
            g.skip = (minEig < 1e-4f || D < FLT_EPSILON) ? 1 : 0;
        }
    }
    __syncthreads();
    LK_T(3);

    // ================= level loop: J tile + iterations =================
    // Work items of an iteration: items 0..167 take the 168 column pairs (c, c+4) of the two 8-wide SIMD groups of every
    // row (pmaddwd pair sums), items 192..244 take 53 pairs of consecutive tail pixels; a thread owns kIpt items.
    int e0[kIpt], e1[kIpt], o0[kIpt], o1[kIpt];
    bool is_pair[kIpt];
#pragma unroll
    for (int q = 0; q < kIpt; ++q) {
        const int it = tid + q * kLkThreads;
        e0[q] = -1; e1[q] = -1;
        is_pair[q] = it < 168;
        if (it < 168) {
            const int y = it >> 3, g = (it >> 2) & 1, c = it & 3;
            e0[q] = y * kWin + 8 * g + c; e1[q] = e0[q] + 4;
        } else if (it >= 192 && it < 192 + 53) {
            const int t = 2 * (it - 192);
            e0[q] = (t / 5) * kWin + 16 + (t - (t / 5) * 5);
            if (t + 1 < kTail) e1[q] = ((t + 1) / 5) * kWin + 16 + ((t + 1) - ((t + 1) / 5) * 5);
        }
        o0[q] = e0[q] < 0 ? 0 : (e0[q] / kWin) * kTileJ + (e0[q] - (e0[q] / kWin) * kWin);
        o1[q] = e1[q] < 0 ? 0 : (e1[q] / kWin) * kTileJ + (e1[q] - (e1[q] / kWin) * kWin);
    }
    bool tile_ready = true;      // the top level's J tile was staged by the prologue
    int pbuf = 0;

#pragma unroll 1
    for (int level = max_level; level >= 0; --level) {
        const LevelRef LJ = sm.pyr[ja][level];
        const LevelBlock& B = sm.lv[level];
        const int jcols = LJ.w, jrows = LJ.h;
        const float scale = pow2_neg(level);
        float nx, ny;
        if (level == max_level) {
            if (use_init) { nx = __fmul_rn(outx, scale); ny = __fmul_rn(outy, scale); }
            else { nx = __fmul_rn(prev_pt.x, scale); ny = __fmul_rn(prev_pt.y, scale); }
        } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;
        const int4 gflags = *reinterpret_cast<const int4*>(&B.geo.valid);     // valid, skip
        if (!gflags.x || gflags.y) {   // window outside the level, or minEig / det below threshold (lkpyramid.cpp)
            if (level == 0) status = 0;
            tile_ready = false;
            continue;
        }
        const float4 gA = *reinterpret_cast<const float4*>(&B.geo.A11);
        const float A11 = gA.x, A12 = gA.y, A22 = gA.z, D = gA.w;
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        // this thread's two source pixels stay in registers for the whole level
        int I0[kIpt], I1[kIpt];
        short2 g0[kIpt], g1[kIpt];
#pragma unroll
        for (int q = 0; q < kIpt; ++q) {
            I0[q] = B.Iw[max(e0[q], 0)]; I1[q] = B.Iw[max(e1[q], 0)];
            g0[q] = e0[q] >= 0 ? B.dI[e0[q]] : make_short2(0, 0); g1[q] = e1[q] >= 0 ? B.dI[e1[q]] : make_short2(0, 0);
        }
        LK_T(11);
        {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (!tile_ready || inx < jx0 || iny < jy0 || inx + kTileD > jx0 + kTileJ || iny + kTileD > jy0 + kTileJ) {
                // no usable prefetched tile: stage it now (one warp, one row per lane)
                jx0 = clamp_tile_origin(inx - kJMargin, jcols); jy0 = clamp_tile_origin(iny - kJMargin, jrows);
                jbuf ^= 1;
                if (wid == kLkWarps - 1) {
                    uint32_t v[kTileJ / 4];
                    load_row_words<kTileJ / 4>(LJ, jx0, jy0 + lane, v);
                    uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[jbuf] + lane * kTileJ);
                    dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                }
                __syncthreads();
            }
        }
        // prefetch the next level's J tile around the current estimate: the last warp issues the loads now and parks the
        // words in registers until the iterations of this level are done
        uint32_t pf[kTileJ / 4 + 1];
        int px0 = 0, py0 = 0;
        if (level > 0) {
            const LevelRef LN = sm.pyr[ja][level - 1];
            px0 = clamp_tile_origin(__float2int_rd(__fadd_rn(__fmul_rn(nx, 2.f), half)) - kJMargin, LN.w);   // 2*(nx + half) - half
            py0 = clamp_tile_origin(__float2int_rd(__fadd_rn(__fmul_rn(ny, 2.f), half)) - kJMargin, LN.h);
            if (wid == kLkWarps - 1) issue_row_words<kTileJ / 4>(sm.pyr[ja][level - 1], px0, py0 + lane, pf);
        }
        LK_T(4);
        const uint8_t* tj = sm.tileJ[jbuf];
        float pdx = 0.f, pdy = 0.f;
        // window origins for which the iteration needs no special handling: inside the level's search range
        // (lkpyramid.cpp bounds test) AND covered by the staged tile.  Everything else takes the cold branch.
        int lox = max(-kWin, jx0), hix = min(jcols - 1, jx0 + kTileJ - kTileD);
        int loy = max(-kWin, jy0), hiy = min(jrows - 1, jy0 + kTileJ - kTileD);
#pragma unroll 1
        for (int j = 0; ; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < lox || inx > hix || iny < loy || iny > hiy) {
                if (inx < -kWin || inx >= jcols || iny < -kWin || iny >= jrows) {
                    if (level == 0) status = 0;
                    break;
                }
                // the window left the staged tile: re-centre it (uniform branch)
                jx0 = inx - kJMargin; jy0 = iny - kJMargin;
                lk_restage_tile(sm, ja, level, jbuf, jx0, jy0);
                lox = max(-kWin, jx0); hix = min(jcols - 1, jx0 + kTileJ - kTileD);
                loy = max(-kWin, jy0); hiy = min(jrows - 1, jy0 + kTileJ - kTileD);
            }
            ++iters;
            int w00, w01, w10, w11;
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const uint8_t* jt = tj + (iny - jy0) * kTileJ + (inx - jx0);
            // exact int32 products (idle items own zero gradients); pair items form OpenCV's pmaddwd pair sums, tail items keep
            // single terms.  tx / ty: what this thread adds to b1 / b2; abx / aby: its sum |term| per
            // component, clamped so that the totals over all contributing items stay below 2^32.  (The products stay in
            // registers for the cold paths below: forming them again there measured slower, frame span 104.6 -> 108.7 us.)
            int ex0[kIpt], ey0[kIpt], ex1[kIpt], ey1[kIpt];
            int tx = 0, ty = 0, abx = 0, aby = 0;
#pragma unroll
            for (int q = 0; q < kIpt; ++q) {
                const uint8_t* p0 = jt + o0[q];
                const uint8_t* p1 = jt + o1[q];
                const int d0 = VF_DESCALE(p0[0] * w00 + p0[1] * w01 + p0[kTileJ] * w10 + p0[kTileJ + 1] * w11, 9) - I0[q];
                const int d1 = VF_DESCALE(p1[0] * w00 + p1[1] * w01 + p1[kTileJ] * w10 + p1[kTileJ + 1] * w11, 9) - I1[q];
                ex0[q] = d0 * g0[q].x; ey0[q] = d0 * g0[q].y; ex1[q] = d1 * g1[q].x; ey1[q] = d1 * g1[q].y;
                const int sx_ = ex0[q] + ex1[q], sy_ = ey0[q] + ey1[q];
                tx += sx_; ty += sy_;
                abx += min(is_pair[q] ? abs(sx_) : abs(ex0[q]) + abs(ex1[q]), kExact);
                aby += min(is_pair[q] ? abs(sy_) : abs(ey0[q]) + abs(ey1[q]), kExact);
            }
            LK_TF(6);
            const int sx = __reduce_add_sync(0xffffffffu, tx);
            const int sy = __reduce_add_sync(0xffffffffu, ty);
            const int sax = __reduce_add_sync(0xffffffffu, abx);
            const int say = __reduce_add_sync(0xffffffffu, aby);
            // every warp publishes its three partial sums (double-buffered: the slot of iteration j is rewritten in j+2,
            // after the barrier of j+1); one barrier; every thread folds the kLkWarps partials itself
            if (lane == 0) sm.part[pbuf][wid] = make_int4(sx, sy, sax, say);
            __syncthreads();
            int bx = 0, by = 0, bz = 0, bw = 0;
#pragma unroll
            for (int w = 0; w < kLkWarps; ++w) {
                const int4 q4 = sm.part[pbuf][w];
                bx += q4.x; by += q4.y; bz += q4.z; bw += q4.w;
            }
            pbuf ^= 1;
            // b1 only adds x-terms, b2 only y-terms: each sum has its own exactness bound
            const unsigned int bnd = max((unsigned int)bz, (unsigned int)bw);
            LK_TF(7);
            float ib1, ib2;
            if (bnd < (unsigned int)kExact) {
                // every partial sum of every chain and of their combination is an exactly representable integer
                ib1 = (float)bx; ib2 = (float)by;
                ++fast;
            } else {
                // (cold, out of line: keeps the hot loop compact in the instruction caches)  First the exactness test chain by
                // chain; the ordered chains only if one of the ten fails it.
                LkChainIn ci = {0, 0, 0u, 0u, 0, 0, 0u, 0u};
#pragma unroll
                for (int q = 0; q < kIpt; ++q) {
                    const int sx_ = ex0[q] + ex1[q], sy_ = ey0[q] + ey1[q];
                    if (is_pair[q]) { ci.px += sx_; ci.py += sy_; ci.apx += (unsigned int)abs(sx_); ci.apy += (unsigned int)abs(sy_); }
                    else {
                        ci.tx += sx_; ci.ty += sy_;
                        ci.atx += (unsigned int)(abs(ex0[q]) + abs(ex1[q])); ci.aty += (unsigned int)(abs(ey0[q]) + abs(ey1[q]));
                    }
                }
                const float4 cs = lk_chain_sums(sm, ci);
                float2 ib = make_float2(cs.x, cs.y);
                if (cs.z != 0.f) ++fast;
                else {
                    LkTerms tm;
#pragma unroll
                    for (int q = 0; q < kIpt; ++q) { tm.ex0[q] = ex0[q]; tm.ey0[q] = ey0[q]; tm.ex1[q] = ex1[q]; tm.ey1[q] = ey1[q]; }
                    ib = lk_spill_chains(sm, tm);
                }
                ib1 = ib.x; ib2 = ib.y;
                LK_TF(8);
            }
            const float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            // The three ways out of the loop (lkpyramid.cpp, in this order), folded into ONE branch on the hot path:
            //   delta.ddot(delta) <= eps^2 (in double; a float estimate settles all but borderline cases);
            //   j > 0 and |delta + prevDelta| < 0.01 on both axes (std::abs on a float widened to double:
            //       < 0.01  <=>  <= 0.01f, the largest float below 0.01) -> step back by delta/2;
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
            LK_TF(9);
        }
        LK_T(5);
        tile_ready = false;
        if (level > 0) {   // publish the prefetched tile of the next level
            jbuf ^= 1;
            if (wid == kLkWarps - 1) {
                const int sh = 8 * (px0 & 3);
                uint32_t v[kTileJ / 4];
#pragma unroll
                for (int k = 0; k < kTileJ / 4; ++k) v[k] = __funnelshift_r(pf[k], pf[k + 1], sh);
                uint4* dst = reinterpret_cast<uint4*>(sm.tileJ[jbuf] + lane * kTileJ);
                dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
            }
            jx0 = px0; jy0 = py0;
            tile_ready = true;
            __syncthreads();
        }
        LK_T(10);
        LK_T(10);
        if (status && level == 0) {   // the caller passes an err vector: final window test (lkpyramid.cpp)
            const int ix = __float2int_rd(__fsub_rn(outx, half)), iy = __float2int_rd(__fsub_rn(outy, half));
            if (ix < -kWin || ix >= jcols || iy < -kWin || iy >= jrows) status = 0;
        }
    }
    if (tid == 0) {
        iters_io[0] += iters;
        iters_io[1] += fast;
    }
    LkResult res;
    res.x = outx; res.y = outy; res.status = status;
    return res;
}

__device__ __forceinline__ LkResult lk_track_point(LkSmem& sm, int ia, int ja, float2 prev_pt, float2 init_pt, bool use_init,
                                                   int max_level, int* iters_io) {
    return sm.termsA_off ? lk_track_point_impl<true>(sm, ia, ja, prev_pt, init_pt, use_init, max_level, iters_io)
                         : lk_track_point_impl<false>(sm, ia, ja, prev_pt, init_pt, use_init, max_level, iters_io);
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

// Zero the staging area of the b chains once per kernel: the terms never overwrite its padding slots, and adding +0.0f
// to a chain value is an exact identity (a chain of integer-valued floats that starts at +0 can never be -0.0f).
__device__ __forceinline__ void lk_smem_init(LkSmem& sm, int* counters2, int n_levels, bool staged) {
    const int off = staged ? (int)(sizeof(LkSmem) + sizeof(LevelBlock) * (size_t)(n_levels > 1 ? n_levels - 1 : 0)) : 0;
    if (threadIdx.x == 0) sm.termsA_off = off;
    if (staged) {
        float* tA = lk_terms_a(sm, off);                        // slots 84..107 of every chain (the tails' real terms 84..104 are written every pass)
        constexpr int kPad = kChainA - 84;
        for (int i = threadIdx.x; i < n_levels * kChainsA * kPad; i += kLkThreads) {
            const int level = i / (kChainsA * kPad), r = i - level * (kChainsA * kPad);
            tA[level * (kChainsA * kChainA) + (r / kPad) * kChainA + 84 + (r % kPad)] = 0.f;
        }
    }
    for (int i = threadIdx.x; i < 10 * kChainA; i += kLkThreads) sm.termsB[i] = 0.f;
    if (threadIdx.x == 0) { counters2[0] = 0; counters2[1] = 0; }
#ifdef VF_LK_TIMING
    if (threadIdx.x < 16) { sm.tacc[threadIdx.x] = 0; sm.tcnt[threadIdx.x] = 0; }
    if (threadIdx.x == 0) sm.tbase = 0;
#endif
    (void)sm;
}

__device__ __forceinline__ void lk_timing_flush(LkSmem& sm) {
#ifdef VF_LK_TIMING
    __syncthreads();
    if (threadIdx.x < 16) {
        atomicAdd(&vf_lk_timing[threadIdx.x], sm.tacc[threadIdx.x]);
        atomicAdd(&vf_lk_timing[16 + threadIdx.x], sm.tcnt[threadIdx.x]);
    }
#else
    (void)sm;
#endif
}

// ---- temporal: prev -> cur (maxLevel L) then cur -> prev (maxLevel 1, initial flow), :37-67 ----
__device__ __forceinline__ void lk_temporal_body(const DevBatch& b, int parity_arg, int cur, int prev) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;   // frame parity; trace slot (frame % 6)
    LkSmem& sm = *reinterpret_cast<LkSmem*>(lk_smem_raw);
    __shared__ int stat[2];
    TraceScope trace(b.trace, TR_LK_TEMPORAL + 16 * tphase);
    const int s = blockIdx.y, i = blockIdx.x, cap = b.cap;
    const int n_prev = b.counters[parity ^ 1][s * C_COUNT + C_N_TOTAL];
    const float2 p0 = b.kps[parity ^ 1][(size_t)s * cap + i];   // i < cap: always readable; in flight beside the count
#ifdef VF_LK_TIMING
    const unsigned long long cta_t0 = vf_globaltimer();
#endif
    if (i >= n_prev) return;
    lk_smem_init(sm, stat, b.n_levels, (parity_arg & 2) != 0);
    load_pyr(sm, 0, b.left[prev], b.lvl_stream_stride, s);   // 0 = previous left
    load_pyr(sm, 1, b.left[cur], b.lvl_stream_stride, s);    // 1 = current left
    __syncthreads();
    const int L = min(b.lk_max_level, b.left[cur].n_levels - 1);
#ifdef VF_LK_TIMING_REPEAT
    {   // cold-vs-warm experiment: the same forward pass twice, the second one is accounted under phases 8..13
        lk_track_point(sm, 0, 1, p0, p0, false, L, stat);
        __syncthreads();
        if (threadIdx.x == 0) { sm.tbase = 8; stat[0] = 0; stat[1] = 0; }
        __syncthreads();
    }
#endif
    const LkResult f = lk_track_point(sm, 0, 1, p0, p0, false, L, stat);
#ifdef VF_LK_TIMING_REPEAT
    __syncthreads();
    if (threadIdx.x == 0) sm.tbase = 0;
    __syncthreads();
#endif
    const float2 fwd = make_float2(f.x, f.y);
    float2 rev = p0;
    int st = f.status, rst = 0;
    int ok = st;
    if (b.flow_back) {
        // the reference also runs the reverse pass on points whose forward status is already 0 and then
        // discards the result (:55-59); the device skips that work, the outcome is the same.
        if (st) {
            const LkResult r = lk_track_point(sm, 1, 0, fwd, p0, true, min(1, L), stat);
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
        b.lk_iters[o] = stat[0] | (stat[1] << 16);   // low 16 bits: iterations, high: of which on the exact-integer path
#ifdef VF_LK_TIMING
        if (i < 256) { vf_lk_cta_times[0][i][0] = cta_t0; vf_lk_cta_times[0][i][1] = vf_globaltimer(); vf_lk_cta_times[0][i][2] = vf_smid(); vf_lk_cta_times[0][i][3] = stat[0]; }
#endif
    }
    lk_timing_flush(sm);
}

// ---- stereo: left -> right then right -> left (both maxLevel L), status (:112-135), then
//      RemoveDistortion + ComputeKpsVelocity for the right camera and the right half of the featureFrame record ----
__device__ __forceinline__ void lk_stereo_body(const DevBatch& b, int parity_arg, int cur, const double* __restrict__ dt_host) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;   // frame parity; trace slot (frame % 6)
    LkSmem& sm = *reinterpret_cast<LkSmem*>(lk_smem_raw);
    __shared__ int stat[2];
    TraceScope trace(b.trace, TR_LK_STEREO + 16 * tphase);
    const int s = blockIdx.y, i = blockIdx.x, cap = b.cap, tid = threadIdx.x;
    const int* cnt_prev = b.counters[parity ^ 1] + s * C_COUNT;
    const int n = b.counters[parity][s * C_COUNT + C_N_TOTAL];
    // everything the CTA needs from HBM goes out in one batch, beside the count (i < cap: always readable)
    const size_t o = (size_t)s * cap + i;
    const float2 p0 = b.kps[parity][o];
    const int pidx = b.prev_idx[parity][o];
#ifdef VF_LK_TIMING
    const unsigned long long cta_t0 = vf_globaltimer();
#endif
    if (i >= n) return;
    lk_smem_init(sm, stat, b.n_levels, (parity_arg & 2) != 0);
    load_pyr(sm, 0, b.left[cur], b.lvl_stream_stride, s);   // 0 = current left
    load_pyr(sm, 1, b.right[parity], b.lvl_stream_stride, s);          // 1 = current right
    if (tid == 0) sm.found[1] = -1;
    __syncthreads();
    // previous_right_un_distortion_ids_kps_map_.find(id): select_tracked recorded where the feature sat in the previous
    // frame's arrays (-1: new corner); it is in the right map iff it had a right observation there
    if (tid == 0 && pidx >= 0 && b.right_ok[parity ^ 1][(size_t)s * cap + pidx]) sm.found[1] = pidx;
    const int L = min(b.lk_max_level, b.right[parity].n_levels - 1);
    const LkResult f = lk_track_point(sm, 0, 1, p0, p0, false, L, stat);
    const float2 rp = make_float2(f.x, f.y);
    float2 back = p0;
    int st = f.status, rst = 0;
    int ok = st;
    if (b.flow_back) {
        if (st) {
            const LkResult r = lk_track_point(sm, 1, 0, rp, rp, false, L, stat);
            back = make_float2(r.x, r.y); rst = r.status;
        }
        ok = st && rst && in_border(rp, b.H, b.W) && pt_distance(p0, back) <= 0.5;
    }
    __syncthreads();
    if (tid == 0) {
        vf_feature* rec = &b.out[tphase % 3][o];     // mapped pinned host memory (ring of three frames): the record goes straight to the caller's side
        rec->has_right = ok;
        float2 unr = make_float2(0.f, 0.f), vr = make_float2(0.f, 0.f);
        if (ok) {
            unr = undistort_point(b.cam[1], rp);
            if (sm.found[1] >= 0) vr = velocity(unr, b.un_right[parity ^ 1][(size_t)s * cap + sm.found[1]], dt_host[s]);
        }
        rec->xr = unr.x; rec->yr = unr.y; rec->ur = ok ? rp.x : 0.f; rec->vr = ok ? rp.y : 0.f; rec->vxr = vr.x; rec->vyr = vr.y;
        b.un_right[parity][o] = unr;
        b.right_ok[parity][o] = (uint8_t)ok;
        b.lr_pts[o] = rp; b.lr_st[o] = (uint8_t)st; b.rl_pts[o] = back; b.rl_st[o] = (uint8_t)rst;
        b.lk_iters_stereo[o] = stat[0] | (stat[1] << 16);
        if (ok) atomicAdd(&b.counters[parity][s * C_COUNT + C_N_RIGHT], 1);
#ifdef VF_LK_TIMING
        if (i < 256) { vf_lk_cta_times[1][i][0] = cta_t0; vf_lk_cta_times[1][i][1] = vf_globaltimer(); vf_lk_cta_times[1][i][2] = vf_smid(); vf_lk_cta_times[1][i][3] = stat[0]; }
#endif
    }
    lk_timing_flush(sm);
}

// ONE entry point for both pairs of LK calls (bit 2 of parity_arg: 0 = temporal, 1 = stereo).  A device function is linked into
// every kernel that calls it, so two kernels would each carry their own 80 KB copy of the pass -- and in overlapped frames the
// temporal launch of frame k+1 and the stereo launch of frame k share every SM: two copies competed for the 32 KB of L1.5
// instruction cache (`no_instruction` was already the second stall reason of either kernel alone).  One kernel, one copy.
#ifdef VF_LK_MINB     // A/B aid: resident CTAs per SM the register budget is sized for (5 caps the kernel at 96 registers, 1 lifts every cap)
#define VF_LK_BOUNDS __launch_bounds__(kLkThreads, VF_LK_MINB)
#else
#define VF_LK_BOUNDS __launch_bounds__(kLkThreads)
#endif
__global__ void VF_LK_BOUNDS lk_pair_kernel(const __grid_constant__ DevBatch b, int parity_arg, int cur, int prev,
                                                             const double* __restrict__ dt_host) {
    pdl_wait(); pdl_release();
    if (parity_arg & 4) lk_stereo_body(b, parity_arg & ~4, cur, dt_host);
    else lk_temporal_body(b, parity_arg, cur, prev);
}

// ---- left camera: RemoveDistortion + ComputeKpsVelocity (:103-104, :310-363) and the left half of the record.
//      Depends on current_kps_ / ids_ only, so it runs beside the stereo LK on the side stream. ----
__global__ void __launch_bounds__(128) undistort_left_kernel(const __grid_constant__ DevBatch b, int parity_arg, const double* __restrict__ dt_host, int* __restrict__ n_host) {
    const int parity = parity_arg & 1, tphase = parity_arg >> 4;   // frame parity; trace slot (frame % 6)
    TraceScope trace(b.trace, TR_UNDISTORT_LEFT + 16 * tphase);
    const int s = blockIdx.y, cap = b.cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = b.counters[parity][s * C_COUNT + C_N_TOTAL];
    if (blockIdx.x == 0 && threadIdx.x == 0) n_host[s] = n;      // feature count of the frame, straight to pinned host memory
    if (i >= n) return;
    const size_t o = (size_t)s * cap + i;
    const float2 p0 = b.kps[parity][o];
    const int id = b.ids[parity][o];
    const float2 un = undistort_point(b.cam[0], p0);
    // previous_left_un_distortion_ids_kps_map_.find(id): every feature of the previous frame is in that map, and
    // select_tracked recorded where this one sat in the previous frame's arrays (-1: new corner)
    const int found = b.prev_idx[parity][o];
    float2 v = make_float2(0.f, 0.f);
    if (found >= 0) v = velocity(un, b.un_left[parity ^ 1][(size_t)s * cap + found], dt_host[s]);   // dt: zero-copy read of the host's value
    vf_feature* rec = &b.out[tphase % 3][o];
    rec->id = id; rec->track_cnt = b.cnt[parity][o];
    rec->x = un.x; rec->y = un.y; rec->u = p0.x; rec->v = p0.y; rec->vx = v.x; rec->vy = v.y;
    b.un_left[parity][o] = un;
}

// ---- stand-alone calcOpticalFlowPyrLK (vf_op_lk) ----
__global__ void __launch_bounds__(kLkThreads) lk_op_kernel(PyrView prev, PyrView next, const float2* __restrict__ prev_pts,
                                                           float2* __restrict__ next_pts, uint8_t* __restrict__ status,
                                                           int* __restrict__ iters_out, int n, int max_level, int use_init) {
    LkSmem& sm = *reinterpret_cast<LkSmem*>(lk_smem_raw);
    __shared__ int stat[2];
    const int i = blockIdx.x;
    if (i >= n) return;
    const int L = min(max_level, min(prev.n_levels, next.n_levels) - 1);
    lk_smem_init(sm, stat, L + 1, (use_init & 2) != 0);
    load_pyr(sm, 0, prev, nullptr, 0);
    load_pyr(sm, 1, next, nullptr, 0);
    __syncthreads();
    const LkResult r = lk_track_point(sm, 0, 1, prev_pts[i], next_pts[i], (use_init & 1) != 0, L, stat);
    __syncthreads();
    if (threadIdx.x == 0) {
        next_pts[i] = make_float2(r.x, r.y); status[i] = (uint8_t)r.status;
        if (iters_out) iters_out[i] = stat[0];
    }
    lk_timing_flush(sm);
}

__global__ void undistort_kernel(CamParams cam, const float2* __restrict__ uv, float2* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = undistort_point(cam, uv[i]);
}
void launch_undistort(const CamParams& cam, const float2* uv, float2* out, int n, cudaStream_t stream) {
    if (n <= 0) return;
    undistort_kernel<<<(n + 127) / 128, 128, 0, stream>>>(cam, uv, out, n);
}

// The LK kernels keep all per-level state of a pass in (dynamic) shared memory: opt in above 48 KB once per device.
cudaError_t lk_prepare() {
    const int bytes = (int)lk_smem_bytes(kMaxLevels, true);
    cudaError_t e = cudaFuncSetAttribute(lk_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lk_op_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    if ((e = lk_warp_prepare()) != cudaSuccess) return e;
    if ((e = lk_generic_prepare()) != cudaSuccess) return e;
    // same carveout as every other kernel of the library (see corners_prepare)
    if ((e = cudaFuncSetAttribute(lk_pair_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lk_op_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(undistort_left_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(undistort_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout);
}

void launch_lk_temporal(const DevBatch* d_batch, const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream) {
    if (h.lk_win != kWin) { launch_lk_temporal_generic(h, parity, cur, prev, stream); return; }   // vf_lk_generic.cu
    if (lk_use_warp_kernel((long long)h.cap * h.n_streams)) { launch_lk_temporal_warp(h, parity, cur, prev, stream); return; }
    dim3 grid(h.cap, h.n_streams);
    const bool staged = lk_plan_staged(h.n_levels, (long long)h.cap * h.n_streams);   // bit 1 of the parity argument
    launch_kernel(lk_pair_kernel, grid, dim3(kLkThreads), lk_smem_bytes(h.n_levels, staged), stream, h, parity | (staged ? 2 : 0), cur, prev,
                  (const double*)nullptr);
}
void launch_lk_stereo(const DevBatch* d_batch, const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream) {
    if (h.lk_win != kWin) { launch_lk_stereo_generic(h, parity, cur, dt_host, stream); return; }
    if (lk_use_warp_kernel((long long)h.cap * h.n_streams)) { launch_lk_stereo_warp(h, parity, cur, dt_host, stream); return; }
    dim3 grid(h.cap, h.n_streams);
    const bool staged = lk_plan_staged(h.n_levels, (long long)h.cap * h.n_streams);
    launch_kernel(lk_pair_kernel, grid, dim3(kLkThreads), lk_smem_bytes(h.n_levels, staged), stream, h, parity | (staged ? 2 : 0) | 4, cur, 0, dt_host);
}
void launch_undistort_left(const DevBatch* d_batch, const DevBatch& h, int parity, const double* dt_host, int* n_host, cudaStream_t stream) {
    dim3 grid((h.cap + 127) / 128, h.n_streams);
    undistort_left_kernel<<<grid, 128, 0, stream>>>(h, parity, dt_host, n_host);
}
void launch_lk_op(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status,
                  int* iters, int n, int max_level, int use_initial_flow, cudaStream_t stream) {
    if (n <= 0) return;
    if (lk_use_warp_kernel(n)) { launch_lk_op_warp(prev, next, prev_pts, next_pts, status, iters, n, max_level, use_initial_flow, stream); return; }
    const int nlv = (max_level < prev.n_levels - 1 ? max_level : prev.n_levels - 1) + 1;
    const bool staged = lk_plan_staged(nlv, n);                                     // bit 1 of the flow flag
    lk_op_kernel<<<n, kLkThreads, lk_smem_bytes(nlv, staged), stream>>>(prev, next, prev_pts, next_pts, status, iters, n, max_level,
                                                                        (use_initial_flow ? 1 : 0) | (staged ? 2 : 0));
}

}  // namespace vf

#ifdef VF_LK_TIMING
// measurement aid (timing build only): read and clear the per-phase cycle counters
extern "C" int vf_debug_lk_timing(unsigned long long* out32) {
    unsigned long long zero[32] = {0};
    if (cudaMemcpyFromSymbol(out32, vf::vf_lk_timing, sizeof(zero)) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(vf::vf_lk_timing, zero, sizeof(zero)) != cudaSuccess;
}
extern "C" int vf_debug_lk_cta_times(unsigned long long* out) {   // [2][256][4]
    return cudaMemcpyFromSymbol(out, vf::vf_lk_cta_times, sizeof(unsigned long long) * 2 * 256 * 4) != cudaSuccess;
}
#endif
