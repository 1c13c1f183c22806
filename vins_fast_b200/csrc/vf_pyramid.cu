// uint8 Gaussian pyramid (pyrDown) and Scharr derivative planes for sm_100a.
//
// pyrdown_u8: what cv::buildOpticalFlowPyramid / cv::pyrDown compute for the reference's four
// calcOpticalFlowPyrLK calls (vins/estimator/feature_tracker.cpp:41,50,117,125):
//   dst(x,y) = ( sum_{i,j in [-2,2]} k[i]k[j] src(refl101(2x+j), refl101(2y+i)) + 128 ) >> 8,  k = [1 4 6 4 1]
// Integer, bit-exact.  HBM-bound streaming stencil: one CTA stages a (2*TH+3) x (2*TW+4) input tile in
// shared memory (16-byte loads in the interior, reflect-101 indexing on the frame edge), runs the
// horizontal 5-tap into a uint16 tile, then each thread runs the vertical 5-tap for PX consecutive
// outputs and writes them with one PX-byte vector store (PX = 16 -> 16-byte HBM stores).
#include "vf_common.cuh"
#include "vf_kernels.h"

namespace vf {

template <int TW, int TH, int PX>
__global__ void __launch_bounds__((TW / PX) * TH)
pyrdown_u8_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, size_t s_img_stride,
                  uint8_t* __restrict__ dst, int dw, int dh, int dpitch, size_t d_img_stride) {
    constexpr int IW = 2 * TW + 4;          // input tile width  (cols 2*x0-2 .. 2*x0+2*TW+1)
    constexpr int IH = 2 * TH + 3;          // input tile height (rows 2*y0-2 .. 2*y0+2*TH)
    constexpr int IWP = (IW + 15) / 16 * 16 + 16;   // smem pitch, allows 16-byte aligned vector fill
    constexpr int NT = (TW / PX) * TH;
    __shared__ __align__(16) uint8_t tile[IH][IWP];
    __shared__ __align__(16) uint16_t hs[IH][TW];

    const int img = blockIdx.z;
    src += (size_t)img * s_img_stride;
    dst += (size_t)img * d_img_stride;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    const int ix0 = 2 * x0 - 2, iy0 = 2 * y0 - 2;

    // ---- stage the input tile.  tile[r][c + SH] = src(iy0 + r, ix0 + c); SH aligns 16-byte chunks ----
    // ix0 = 2*x0-2 with x0 a multiple of TW (>= 64): (ix0 + 2) is 16-byte aligned, so shift by 2... we
    // instead align on the global side: chunk k covers global columns [gx, gx+16), gx = (ix0 & ~15) + 16k.
    const int gbase = ix0 & ~15;            // floor to 16 (works for negative ix0 too: two's complement)
    const int shift = ix0 - gbase;          // 0..15, tile column of ix0 inside the chunked row
    constexpr int CHUNKS = IWP / 16;
    for (int i = tid; i < IH * CHUNKS; i += NT) {
        const int r = i / CHUNKS, k = i - r * CHUNKS;
        const int gy = refl101(iy0 + r, sh);
        const int gx = gbase + 16 * k;
        const uint8_t* row = src + (size_t)gy * spitch;
        uint4 v;
        if (gx >= 0 && gx + 16 <= sw) {
            v = *reinterpret_cast<const uint4*>(row + gx);    // pitch and gx are multiples of 16
        } else {
            uint8_t b[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) b[j] = row[refl101(gx + j, sw)];
            v = *reinterpret_cast<uint4*>(b);
        }
        *reinterpret_cast<uint4*>(&tile[r][16 * k]) = v;
    }
    __syncthreads();

    // ---- horizontal [1 4 6 4 1] at even columns ----
    for (int i = tid; i < IH * TW; i += NT) {
        const int r = i / TW, x = i - r * TW;
        const uint8_t* p = &tile[r][shift + 2 * x];
        hs[r][x] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();

    // ---- vertical pass, PX outputs per thread, one vector store ----
    const int ty = tid / (TW / PX), tx = (tid - ty * (TW / PX)) * PX;
    const int oy = y0 + ty, ox = x0 + tx;
    if (oy >= dh || ox >= dw) return;
    uint8_t o[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        const int r = 2 * ty;
        int v = hs[r][tx + j] + 4 * hs[r + 1][tx + j] + 6 * hs[r + 2][tx + j] + 4 * hs[r + 3][tx + j] + hs[r + 4][tx + j];
        o[j] = (uint8_t)((v + 128) >> 8);
    }
    uint8_t* out = dst + (size_t)oy * dpitch + ox;
    if (ox + PX <= dpitch) {               // rows are padded to a 128-byte pitch: the vector store always fits
        if (PX == 16) *reinterpret_cast<uint4*>(out) = *reinterpret_cast<uint4*>(o);
        else if (PX == 8) *reinterpret_cast<uint2*>(out) = *reinterpret_cast<uint2*>(o);
        else { for (int j = 0; j < PX; ++j) out[j] = o[j]; }
    } else {
        for (int j = 0; j < PX && ox + j < dw; ++j) out[j] = o[j];
    }
}

void launch_pyrdown(const uint8_t* src, int sw, int sh, int spitch, size_t s_img_stride, uint8_t* dst, int dw, int dh,
                    int dpitch, size_t d_img_stride, int n_images, cudaStream_t stream) {
    if (dw >= 1024) {
        constexpr int TW = 256, TH = 16, PX = 16;
        dim3 grid((dw + TW - 1) / TW, (dh + TH - 1) / TH, n_images);
        pyrdown_u8_kernel<TW, TH, PX><<<grid, (TW / PX) * TH, 0, stream>>>(src, sw, sh, spitch, s_img_stride, dst, dw, dh,
                                                                         dpitch, d_img_stride);
    } else {
        constexpr int TW = 128, TH = 8, PX = 16;   // 64 threads: more CTAs for the small EuRoC levels
        dim3 grid((dw + TW - 1) / TW, (dh + TH - 1) / TH, n_images);
        pyrdown_u8_kernel<TW, TH, PX><<<grid, (TW / PX) * TH, 0, stream>>>(src, sw, sh, spitch, s_img_stride, dst, dw, dh,
                                                                         dpitch, d_img_stride);
    }
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

}  // namespace vf
