// Pieces shared by the two Lucas-Kanade kernels (vf_lk.cu: one 4-warp CTA per feature, shortest latency per launch of a
// single stream; vf_lk_warp.cu: one warp per feature, no block barrier): tile geometry, the pyramid level descriptor, aligned
// row fetches, OpenCV's bilinear weights and the status helpers of feature_tracker.cpp.
#pragma once
#include <float.h>

#include "vf_common.cuh"

namespace vf {

constexpr int kTileI = 24;                   // staged I tile (window + bilinear tap + Scharr ring)
constexpr int kTileD = 22;                   // derivative tile (window + bilinear tap)
constexpr int kJMargin = 5;                  // staged J tile = 22x22 taps + this margin on every side
constexpr int kTileJ = 22 + 2 * kJMargin;    // 32
constexpr int kExact = 1 << 24;              // below this every integer is a float

struct LevelRef {
    const uint8_t* p;
    int w, h, pitch;
};

#define VF_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))


// 1.f / (1 << level), exactly (OpenCV: (float)(1. / (1 << level))), without an IEEE division on the level chain
__device__ __forceinline__ float pow2_neg(int level) { return __int_as_float((127 - level) << 23); }

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float one_a = __fsub_rn(1.f, a), one_b = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, one_b), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, one_b), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(one_a, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

// One row of NW*4 pixels of a level starting at column x0 of row y, as NW packed words.  The levels carry a REFLECT_101
// border in HBM (vf_common.cuh: kPyrPad), so any row a window can touch is fetched with NW (+1 when unaligned) aligned
// 32-bit loads and realigned with funnel shifts.  Callers keep -26 <= x0, x0 + 4*NW <= w + 26 (+-3 for the alignment)
// and -26 <= y < h + 26.
template <int NW>
__device__ __forceinline__ void load_row_words(const LevelRef& L, int x0, int y, uint32_t (&out)[NW]) {
    const uint8_t* row = L.p + (ptrdiff_t)y * L.pitch;
    const int m = x0 & 3;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + (x0 - m));   // level origin and pitch are 16-byte aligned
    uint32_t w[NW + 1];
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = __ldg(wp + k);
    w[NW] = m ? __ldg(wp + NW) : 0u;
#pragma unroll
    for (int k = 0; k < NW; ++k) out[k] = __funnelshift_r(w[k], w[k + 1], 8 * m);
}
// The same fetch in two halves for a prefetch: issue_row_words only issues the loads (the raw words stay in flight while
// the caller does other work -- realigning them at once would stall the warp on the memory round trip), align_row_word
// realigns one word when the data is finally consumed.  Reads one word more than load_row_words (inside the border).
template <int NW>
__device__ __forceinline__ void issue_row_words(const LevelRef& L, int x0, int y, uint32_t (&raw)[NW + 1]) {
    const uint8_t* row = L.p + (ptrdiff_t)y * L.pitch;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + (x0 - (x0 & 3)));
#pragma unroll
    for (int k = 0; k <= NW; ++k) raw[k] = __ldg(wp + k);
}

// Origin of a speculative (prefetched) J tile, clamped so that the 32x32 tile stays inside the filled border.
__device__ __forceinline__ int clamp_tile_origin(int o, int n) { return min(max(o, -(kPyrPadFill - 2)), n + kPyrPadFill - 2 - kTileJ); }

struct LkResult {
    float x, y;
    int status;
};

__device__ __forceinline__ bool in_border(float2 p, int rows, int cols) {   // feature_tracker.cpp:20-25
    const int x = __float2int_rn(p.x), y = __float2int_rn(p.y);
    return 1 <= x && x < cols - 1 && 1 <= y && y < rows - 1;
}
__device__ __forceinline__ double pt_distance(float2 a, float2 b) {         // feature_tracker.cpp:457-462
    const double dx = (double)__fsub_rn(a.x, b.x), dy = (double)__fsub_rn(a.y, b.y);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}



// velocity of ComputeKpsVelocity (:323-363): (cur - prev) as float, / dt as double, stored as float.
__device__ __forceinline__ float2 velocity(float2 cur, float2 prev, double dt) {
    return make_float2((float)((double)__fsub_rn(cur.x, prev.x) / dt), (float)((double)__fsub_rn(cur.y, prev.y) / dt));
}


}  // namespace vf
