// uint8 Gaussian pyramid (pyrDown) and Scharr derivative planes for sm_100a.
//
// What cv::buildOpticalFlowPyramid / cv::pyrDown compute for the reference's four calcOpticalFlowPyrLK calls
// (vins/estimator/feature_tracker.cpp:41,50,117,125):
//   dst(x,y) = ( sum_{i,j in [-2,2]} k[i]k[j] src(refl101(2x+j), refl101(2y+i)) + 128 ) >> 8,  k = [1 4 6 4 1]
// Integer, bit-exact.
#include <type_traits>

#include "vf_common.cuh"
#include "vf_kernels.h"

namespace vf {

// ------------------------------------------------------------------------------------------------------
// Fused multi-level pyramid: ONE launch builds NL consecutive levels.  A CTA owns a TW x TH tile of the last
// level and recomputes the halo it needs of every level in between from a single staged tile of the input
// level (85 x 85 pixels for NL = 3, TW = TH = 8), so levels 1..NL never travel through HBM between stages and
// the dependent launch chain of the frame shrinks from NL launches to one.  Each level is still written out
// (by the CTA that owns the pixels) because LK reads every level.
// Regions are kept in in-image coordinates and consumers reflect their tap index (REFLECT_101) into the
// region, which is exactly cv::pyrDown's border rule on every intermediate level.
// ------------------------------------------------------------------------------------------------------
struct FusedPyrArgs {
    const uint8_t* src; int sw, sh, spitch; size_t s_stride;          // input level
    uint8_t* dst[3]; int dw[3], dh[3], dpitch[3]; size_t d_stride[3];  // NL output levels
    unsigned long long* trace; int trace_id;
};

template <int NL, int TW, int TH>
__global__ void __launch_bounds__(256) pyr_fused_kernel(FusedPyrArgs a) {
    constexpr int NT = 256;
    pdl_release();                                   // the border kernel behind this one may be scheduled (it waits for this grid)
    TraceScope trace(a.trace, a.trace_id);
    // maximal region extents per relative level k (k = 0 input ... NL top)
    constexpr int RW3 = TW, RH3 = TH;
    constexpr int RW2 = 2 * RW3 + 3, RH2 = 2 * RH3 + 3;
    constexpr int RW1 = NL >= 2 ? 2 * RW2 + 3 : RW2, RH1 = NL >= 2 ? 2 * RH2 + 3 : RH2;
    constexpr int RW0 = NL >= 3 ? 2 * RW1 + 3 : RW1, RH0 = NL >= 3 ? 2 * RH1 + 3 : RH1;
    constexpr int P0 = (RW0 + 15 + 15) / 16 * 16;                      // input tile pitch (16-byte aligned chunks)
    constexpr int PW = NL >= 3 ? RW1 : (NL >= 2 ? RW2 : RW3);          // widest level-1 region
    constexpr int PI = (PW + 3) / 4 * 4;                               // pitch of intermediate regions
    __shared__ __align__(16) uint8_t buf0[RH0 * P0];
    __shared__ __align__(16) uint8_t bufA[(NL >= 3 ? RH1 : (NL >= 2 ? RH2 : RH3)) * PI];
    __shared__ __align__(16) uint8_t bufB[(NL >= 3 ? RH2 : RH3) * PI];
    constexpr int PW2 = (PW + 1) / 2;                                   // horizontal sums, two columns per 32-bit word
    __shared__ uint32_t tmp[RH0 * PW2];

    const int tid = threadIdx.x, img = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    int lox[NL + 1], hix[NL + 1], loy[NL + 1], hiy[NL + 1], lw[NL + 1], lh[NL + 1];
    lw[0] = a.sw; lh[0] = a.sh;
#pragma unroll
    for (int k = 1; k <= NL; ++k) { lw[k] = a.dw[k - 1]; lh[k] = a.dh[k - 1]; }
    lox[NL] = tx0; hix[NL] = min(tx0 + TW, lw[NL]) - 1;
    loy[NL] = ty0; hiy[NL] = min(ty0 + TH, lh[NL]) - 1;
#pragma unroll
    for (int k = NL - 1; k >= 0; --k) {
        lox[k] = max(0, 2 * lox[k + 1] - 2); hix[k] = min(lw[k] - 1, 2 * hix[k + 1] + 2);
        loy[k] = max(0, 2 * loy[k + 1] - 2); hiy[k] = min(lh[k] - 1, 2 * hiy[k + 1] + 2);
    }
    // ---- stage the input region with aligned 16-byte loads (row pitch and base are multiples of 16) ----
    const uint8_t* src = a.src + (size_t)img * a.s_stride;
    const int ax0 = lox[0] & ~15;
    const int nchunk = (hix[0] - ax0) / 16 + 1, nrow0 = hiy[0] - loy[0] + 1;
    for (int i = tid; i < nrow0 * nchunk; i += NT) {
        const int r = i / nchunk, c = i - r * nchunk;
        *reinterpret_cast<uint4*>(&buf0[r * P0 + 16 * c]) =
            *reinterpret_cast<const uint4*>(src + (size_t)(loy[0] + r) * a.spitch + ax0 + 16 * c);
    }
    __syncthreads();
    if (blockIdx.z == 0 && a.trace_id == TR_PYR_LEFT) trace_phase(a.trace, 16);

    const uint8_t* in = buf0;
    int in_pitch = P0, in_ox = ax0;
    auto transition = [&](auto kc) {
        constexpr int k = decltype(kc)::value;
        uint8_t* out = (k & 1) ? bufB : bufA;
        const int ow = hix[k + 1] - lox[k + 1] + 1, oh = hiy[k + 1] - loy[k + 1] + 1;
        const int ih = hiy[k] - loy[k] + 1;
        // Both passes work on PAIRS of output columns (two 16-bit sums per 32-bit word: a sum is at most 16 * 255 after
        // the first pass and 256 * 255 after the second, so the lanes never carry into each other) and index their work
        // with compile-time extents (no runtime division); the REFLECT_101 index arithmetic only runs for taps that
        // actually leave the image.
        // horizontal [1 4 6 4 1] at even columns of every staged input row
        constexpr int OW2 = ((k == 0 ? (NL >= 3 ? RW1 : (NL >= 2 ? RW2 : RW3)) : (k == 1 ? (NL >= 3 ? RW2 : RW3) : RW3)) + 1) / 2;
        const int wk = lw[k], hk = lh[k];
        for (int i = tid; i < ih * OW2; i += NT) {
            const int r = i / OW2, c = 2 * (i - r * OW2);
            if (c >= ow) continue;
            const int gx = 2 * (lox[k + 1] + c);
            const uint8_t* row = in + r * in_pitch - in_ox;
            const bool has1 = c + 1 < ow;
            int b0, b1, b2, b3, b4, b5 = 0, b6 = 0;
            if (gx >= 2 && gx + 4 < wk) {
                b0 = row[gx - 2]; b1 = row[gx - 1]; b2 = row[gx]; b3 = row[gx + 1]; b4 = row[gx + 2];
                if (has1) { b5 = row[gx + 3]; b6 = row[gx + 4]; }
            } else {
                b0 = row[refl101(gx - 2, wk)]; b1 = row[refl101(gx - 1, wk)]; b2 = row[gx];
                b3 = row[refl101(gx + 1, wk)]; b4 = row[refl101(gx + 2, wk)];
                if (has1) { b5 = row[refl101(gx + 3, wk)]; b6 = row[refl101(gx + 4, wk)]; }
            }
            const uint32_t v0 = (uint32_t)(b0 + b4 + 4 * (b1 + b3) + 6 * b2);
            const uint32_t v1 = (uint32_t)(b2 + b6 + 4 * (b3 + b5) + 6 * b4);
            tmp[r * PW2 + (c >> 1)] = v0 | (v1 << 16);
        }
        __syncthreads();
        // vertical pass -> region of level k+1 (kept in shared memory for the next transition)
        for (int i = tid; i < oh * OW2; i += NT) {
            const int r = i / OW2, c2 = i - r * OW2;
            if (2 * c2 >= ow || r >= oh) continue;
            const int gy = 2 * (loy[k + 1] + r);
            int r0, r1, r3, r4;
            const int r2 = gy - loy[k];
            if (gy >= 2 && gy + 2 < hk) { r0 = r2 - 2; r1 = r2 - 1; r3 = r2 + 1; r4 = r2 + 2; }
            else {
                r0 = refl101(gy - 2, hk) - loy[k]; r1 = refl101(gy - 1, hk) - loy[k];
                r3 = refl101(gy + 1, hk) - loy[k]; r4 = refl101(gy + 2, hk) - loy[k];
            }
            const uint32_t v = tmp[r0 * PW2 + c2] + tmp[r4 * PW2 + c2] + 4u * (tmp[r1 * PW2 + c2] + tmp[r3 * PW2 + c2]) + 6u * tmp[r2 * PW2 + c2];
            const uint32_t q = (v + 0x00800080u) >> 8;                                   // (sum + 128) >> 8 in both lanes
            *reinterpret_cast<uint16_t*>(&out[r * PI + 2 * c2]) = (uint16_t)((q & 0xffu) | ((q >> 8) & 0xff00u));   // PI is even
        }
        __syncthreads();
        // write the part of level k+1 this CTA owns: tile scaled by 2^(NL-1-k), 8 pixels per store
        {
            const int sh_ = NL - 1 - k;
            const int ox0 = tx0 << sh_, oy0 = ty0 << sh_;
            const int ox1 = min((tx0 + TW) << sh_, lw[k + 1]), oy1 = min((ty0 + TH) << sh_, lh[k + 1]);
            const int n8 = (ox1 - ox0 + 7) / 8, nr = oy1 - oy0;
            uint8_t* dst = a.dst[k] + (size_t)img * a.d_stride[k];
            for (int i = tid; i < nr * n8; i += NT) {
                const int r = i / n8, c8 = i - r * n8;
                const int gx = ox0 + 8 * c8, gy = oy0 + r;
                const uint8_t* p = out + (gy - loy[k + 1]) * PI + (gx - lox[k + 1]);
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (gx + j < ox1) lo |= (uint32_t)p[j] << (8 * j);
                    if (gx + 4 + j < ox1) hi |= (uint32_t)p[4 + j] << (8 * j);
                }
                *reinterpret_cast<uint2*>(dst + (size_t)gy * a.dpitch[k] + gx) = make_uint2(lo, hi);   // pitch is a multiple of 16
            }
        }
        in = out; in_pitch = PI; in_ox = lox[k + 1];
        if (blockIdx.z == 0 && a.trace_id == TR_PYR_LEFT) trace_phase(a.trace, 17 + k);
        // the next transition overwrites tmp and the other ping-pong buffer only after every thread passed the barriers above
    };
    transition(std::integral_constant<int, 0>{});
    if constexpr (NL >= 2) transition(std::integral_constant<int, 1>{});
    if constexpr (NL >= 3) transition(std::integral_constant<int, 2>{});
}

// Builds levels first+1 .. first+count of a pyramid (count <= 3) in one launch.
static void launch_pyr_fused(const PyrView& pv, const size_t* stride, int first, int count, int n_images, cudaStream_t stream,
                             const Level0Src* src0, unsigned long long* trace, int trace_id) {
    FusedPyrArgs a;
    a.trace = trace; a.trace_id = trace_id;
    a.src = pv.lvl[first]; a.sw = pv.w[first]; a.sh = pv.h[first]; a.spitch = pv.pitch[first]; a.s_stride = stride ? stride[first] : 0;
    if (first == 0 && src0) { a.src = src0->p; a.spitch = src0->pitch; a.s_stride = src0->stride; }   // level 0 still sits in the staging buffer
    for (int k = 0; k < 3; ++k) {
        const int l = first + 1 + (k < count ? k : count - 1);
        a.dst[k] = pv.lvl[l]; a.dw[k] = pv.w[l]; a.dh[k] = pv.h[l]; a.dpitch[k] = pv.pitch[l]; a.d_stride[k] = stride ? stride[l] : 0;
    }
    const int tw = pv.w[first + count], th = pv.h[first + count];
    if (count == 3) {
        dim3 grid((tw + 7) / 8, (th + 3) / 4, n_images);
        pyr_fused_kernel<3, 8, 4><<<grid, 256, 0, stream>>>(a);
    } else if (count == 2) {
        dim3 grid((tw + 15) / 16, (th + 7) / 8, n_images);
        pyr_fused_kernel<2, 16, 8><<<grid, 256, 0, stream>>>(a);
    } else {
        dim3 grid((tw + 31) / 32, (th + 7) / 8, n_images);
        pyr_fused_kernel<1, 32, 8><<<grid, 256, 0, stream>>>(a);
    }
}

// ------------------------------------------------------------------------------------------------------
// REFLECT_101 border of every level (kPyrPadFill pixels on each side), one 32-bit word per thread:
//   rows -F .. h+F-1 :  the left strip (columns -32 .. -1) and the right strip (from the last aligned column to w+31)
//   rows -F .. -1 and h .. h+F-1 :  the words in between.
// Runs after the level interiors are complete (level 0: the uploaded image; levels 1..: pyr_fused_kernel).
// ------------------------------------------------------------------------------------------------------
struct PadArgs {
    uint8_t* lvl[kMaxLevels]; int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels]; size_t stride[kMaxLevels];
    int first_item[kMaxLevels + 1];   // prefix sum of work items (words) per level
    int n_levels;
    Level0Src src0;                   // p != nullptr: level 0 (interior AND border) is written from this staging image
    unsigned long long* trace; int trace_id;
};

__host__ __device__ inline int pad_items(int w, int h, int* strip_words) {
    const int c0 = w & ~3, n_right = (w + kPyrPad - c0) / 4;
    *strip_words = kPyrPad / 4 + n_right;
    return (h + 2 * kPyrPadFill) * (*strip_words) + 2 * kPyrPadFill * (c0 / 4);
}

__global__ void __launch_bounds__(256) pad_levels_kernel(PadArgs a) {
    pdl_wait(); pdl_release();
    TraceScope trace(a.trace, a.trace_id);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
    if (idx >= a.first_item[a.n_levels]) return;
    int l = 0;
    while (idx >= a.first_item[l + 1]) ++l;
    int j = idx - a.first_item[l];
    const int w = a.w[l], h = a.h[l], pitch = a.pitch[l], F = kPyrPadFill;
    uint8_t* base = a.lvl[l] + (size_t)img * a.stride[l];
    if (l == 0 && a.src0.p) {
        // whole padded extent of level 0 from the staging image: rows -F..h+F-1, columns -32..pitch-33, one 16-byte chunk
        // per thread (both row bases and the chunk columns are 16-byte aligned; the chunks tile the padded pitch exactly)
        const int cpr = pitch / 16;
        const int row = j / cpr - F, col = (j - (row + F) * cpr) * 16 - kPyrPad;
        const uint8_t* srow = a.src0.p + (size_t)img * a.src0.stride + (size_t)refl101(row, h) * a.src0.pitch;
        uint4 v;
        if (col >= 0 && col + 15 < w) {
            v = *reinterpret_cast<const uint4*>(srow + col);
        } else {
            uint32_t q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                q[k] = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) q[k] |= (uint32_t)srow[refl101(col + 4 * k + t, w)] << (8 * t);
            }
            v = make_uint4(q[0], q[1], q[2], q[3]);
        }
        *reinterpret_cast<uint4*>(base + (ptrdiff_t)row * pitch + col) = v;
        return;
    }
    const int c0 = w & ~3;
    int strip;
    pad_items(w, h, &strip);
    int row, col;
    if (j < (h + 2 * F) * strip) {
        row = j / strip - F;
        const int k = j - (row + F) * strip;
        col = k < kPyrPad / 4 ? -kPyrPad + 4 * k : c0 + 4 * (k - kPyrPad / 4);
    } else {
        j -= (h + 2 * F) * strip;
        const int per = c0 / 4, rr = j / per;
        col = 4 * (j - rr * per);
        row = rr < F ? rr - F : h + (rr - F);
    }
    const uint8_t* src = base + (size_t)refl101(row, h) * pitch;
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) v |= (uint32_t)src[refl101(col + q, w)] << (8 * q);
    *reinterpret_cast<uint32_t*>(base + (ptrdiff_t)row * pitch + col) = v;   // interior origin and pitch are 16-byte aligned
}

static void launch_pad_levels(const PyrView& pv, const size_t* stride, int n_images, cudaStream_t stream, const Level0Src* src0,
                              unsigned long long* trace, int trace_id) {
    PadArgs a;
    a.trace = trace; a.trace_id = trace_id;
    a.n_levels = pv.n_levels;
    a.src0.p = nullptr; a.src0.pitch = 0; a.src0.stride = 0;
    if (src0) a.src0 = *src0;
    a.first_item[0] = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        const bool ok = l < pv.n_levels;
        a.lvl[l] = ok ? pv.lvl[l] : nullptr; a.w[l] = ok ? pv.w[l] : 0; a.h[l] = ok ? pv.h[l] : 0; a.pitch[l] = ok ? pv.pitch[l] : 0;
        a.stride[l] = (ok && stride) ? stride[l] : 0;
        int strip;
        int items = ok ? pad_items(pv.w[l], pv.h[l], &strip) : 0;
        if (l == 0 && src0) items = (pv.h[0] + 2 * kPyrPadFill) * (pv.pitch[0] / 16);
        a.first_item[l + 1] = a.first_item[l] + items;
    }
    dim3 grid((a.first_item[pv.n_levels] + 255) / 256, n_images);
    launch_kernel(pad_levels_kernel, grid, dim3(256), 0, stream, a);   // behind pyr_fused: programmatic launch when hinted
}

static int g_pyramid_mode = 0;   // 0 = by size, 1 = fused kernel only, 2 = streaming (TMA) kernel for every transition
void set_pyramid_kernel_mode(int mode) { g_pyramid_mode = mode; }

// A transition is a throughput problem (streaming kernel, vf_pyramid_tma.cu) once its source level holds this many pixels
// over the whole batch; below, the fused kernel's single launch for three levels wins on latency.
static bool want_streaming(const PyrView& pv, int level, int n_images) {
    if (g_pyramid_mode == 1 || !pyramid_tma_available()) return false;
    if (g_pyramid_mode == 2) return true;
    return (long long)pv.w[level] * pv.h[level] * n_images >= 600000;
}

// All levels 1..n_levels-1 of a pyramid from its level 0, then the REFLECT_101 borders of all levels in one more launch.
// Large source levels (4K frames, batches of streams) go through the streaming kernel, one transition per launch; the
// rest through the fused kernel, up to three levels per launch.  With src0 the image still sits in an unpadded staging
// buffer (one linear upload): the first pyramid launch reads it there, and level 0 is moved into its padded home either by
// the streaming kernel itself or, on the fused path, by the border launch.  Returns the launch count.
int launch_pyramid(const PyrView& pv, const size_t* stride, int n_images, cudaStream_t stream, const Level0Src* src0,
                   unsigned long long* trace, int trace_id) {
    int first = 0, left = pv.n_levels - 1, n = 0;
    bool home_done = false;
    const bool hint = pdl_hint();
    while (left > 0 && want_streaming(pv, first, n_images)) {
        const bool staged = first == 0 && src0;
        const uint8_t* src = staged ? src0->p : pv.lvl[first];
        const int spitch = staged ? src0->pitch : pv.pitch[first];
        const size_t sstride = staged ? src0->stride : (stride ? stride[first] : 0);
        pdl_hint() = hint && n > 0;                    // only behind a kernel of this pyramid
        const bool ok = launch_pyr_tma(src, pv.w[first], pv.h[first], spitch, sstride, pv.lvl[first + 1], pv.w[first + 1], pv.h[first + 1],
                                       pv.pitch[first + 1], stride ? stride[first + 1] : 0, staged ? pv.lvl[0] : nullptr, pv.pitch[0],
                                       stride ? stride[0] : 0, n_images, stream, trace, trace_id);
        pdl_hint() = hint;
        if (!ok) break;
        if (staged) home_done = true;
        ++first; --left; ++n;
    }
    while (left > 0) {
        const int cnt = (left == 4) ? 2 : (left < 3 ? left : 3);
        launch_pyr_fused(pv, stride, first, cnt, n_images, stream, home_done ? nullptr : src0, trace, trace_id);
        first += cnt; left -= cnt; ++n;
    }
    launch_pad_levels(pv, stride, n_images, stream, home_done ? nullptr : src0, trace, trace_id + 1);
    return n + 1;
}

// calcScharrDeriv (video/lkpyramid.cpp) as a stand-alone plane kernel: int16 interleaved (Ix, Iy).
// The tracker itself never materialises these planes (lk_track computes the same integers from the
// staged uint8 tile); this kernel backs vf_op_scharr / the parity tests.
__global__ void scharr_s16x2_kernel(const uint8_t* __restrict__ src, int w, int h, int pitch, short2* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
    const uint8_t* r0 = src + (size_t)refl101(y - 1, h) * pitch;
    const uint8_t* r1 = src + (size_t)y * pitch;
    const uint8_t* r2 = src + (size_t)refl101(y + 1, h) * pitch;
    const int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10, t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
    const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
    dst[(size_t)y * w + x] = make_short2((short)(t0p - t0m), (short)((t1m + t1p) * 3 + t1c * 10));
}

void launch_scharr(const uint8_t* src, int w, int h, int pitch, int16_t* dst, cudaStream_t stream) {
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    scharr_s16x2_kernel<<<grid, block, 0, stream>>>(src, w, h, pitch, reinterpret_cast<short2*>(dst));
}


// Every kernel of the library asks for the same shared-memory carveout: kernels with different carveouts cannot be
// co-resident on an SM (the split is reconfigured only when the SM drains), which serialises the streams of a frame.
cudaError_t pyramid_prepare() {
    cudaError_t e = cudaSuccess;
    if ((e = cudaFuncSetAttribute(pyr_fused_kernel<3, 8, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(pyr_fused_kernel<2, 16, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(pyr_fused_kernel<1, 32, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(pad_levels_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(scharr_s16x2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout)) != cudaSuccess) return e;
    return pyramid_tma_prepare();
}

}  // namespace vf
