// Streaming pyrDown for large frames and batches of streams (sm_100a): one level transition per launch, the input tile
// brought into shared memory by the TMA engine (cp.async.bulk.tensor.3d + mbarrier), both filter passes on packed 16-bit
// lanes, 16-byte stores.  Same integers as cv::pyrDown (see vf_pyramid.cu):
//   dst(x,y) = ( sum_{i,j in [-2,2]} k[i]k[j] src(refl101(2x+j), refl101(2y+i)) + 128 ) >> 8,  k = [1 4 6 4 1].
//
// The fused kernel of vf_pyramid.cu builds three levels per launch for the shortest LATENCY on a 752x480 frame, at the price
// of recomputed halos (3.5x read amplification) and byte-wise shared-memory taps; at 3840x2160, or with 64 streams in the
// batch, the pyramid is a THROUGHPUT problem (read N_px, write N_px (S_L - 1) -- plus, in this layout, the
// move of level 0 from the upload buffer into its padded home).  Here:
//   * a CTA owns a 96 x 16 tile of the output level; ONE TMA instruction fetches its 224 x 35 input region (the tensor map
//     covers the interior of the source level, so everything outside the image arrives as zeros and only CTAs on an image
//     edge patch the two halo columns / rows by REFLECT_101 index reflection);
//   * vertical pass: a thread owns one 4-pixel word column of the tile and rolls down the rows -- every input word is read
//     from shared memory once, split once into its even / odd bytes (two 16-bit lanes each), and the five-row window lives
//     in registers; the column sums (<= 16 * 255) go back to shared memory as separate even / odd planes;
//   * horizontal pass: out[j] = E[j-1] + E[j+1] + 4 (O[j-1] + O[j]) + 6 E[j] on those planes, two outputs per 32-bit
//     operation (<= 256 * 255 < 2^16 per lane), (sum + 128) >> 8, sixteen output pixels per thread -> one 16-byte store;
//   * the first transition also moves level 0 from the unpadded upload buffer into its padded home (16-byte chunks straight
//     from the staged tile), so the border kernel no longer re-reads and re-writes the whole image.
// Thirteen-odd CTAs of 128 threads and 15 KB fit an SM, which is what keeps enough TMA transactions in flight.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>

#include "vf_common.cuh"
#include "vf_kernels.h"

namespace vf {

namespace {

constexpr int kOW = 96, kOH = 16;                 // output tile
constexpr int kBoxW = 224, kBoxH = 2 * kOH + 3;   // input box: columns 2*ox0 - 16 .. 2*ox0 + 207, rows 2*oy0 - 2 .. 2*oy0 + 32
constexpr int kColWords = 50;                     // word columns the vertical pass covers: box columns 12 .. 211
constexpr int kFirstWord = 3;                     // = box column 12 (the first tap, 2*ox0 - 2, is box column 14)
constexpr int kPlane = 104;                       // 16-bit entries per plane row (100 used), even so that rows stay word aligned
constexpr int kThreads = 128;

struct TmaPyrArgs {
    CUtensorMap map;                  // source level interior: {w, h, S} uint8, strides {pitch, stream stride}
    int sw, sh;                       // source size
    uint8_t* dst; int dw, dh, dpitch; size_t d_stride;      // destination level (interior origin, padded pitch)
    uint8_t* home; int hpitch; size_t h_stride;             // first transition: padded home of the SOURCE level (else null)
    unsigned long long* trace; int trace_id;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kThreads) pyr_tma_kernel(const __grid_constant__ TmaPyrArgs a) {
    __shared__ __align__(128) uint8_t tile[kBoxH * kBoxW];
    __shared__ __align__(16) uint16_t planeE[kOH * kPlane], planeO[kOH * kPlane];
    __shared__ __align__(8) unsigned long long bar;
    pdl_wait(); pdl_release();
    TraceScope trace(a.trace, a.trace_id);
    const int tid = threadIdx.x, img = blockIdx.z;
    const int ox0 = blockIdx.x * kOW, oy0 = blockIdx.y * kOH;
    const int gx0 = 2 * ox0 - 16, gy0 = 2 * oy0 - 2;          // image coordinates of box element (0, 0)
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kBoxH * kBoxW) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(&a.map), "r"(gx0), "r"(gy0), "r"(img), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();                                           // the barrier is initialised before anybody polls it
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    // ---- REFLECT_101 for the taps that left the image (zeros from the TMA): edge tiles only ----
    const int w = a.sw, h = a.sh;
    const bool edge_x = gx0 + 12 < 0 || gx0 + 212 > w, edge_y = gy0 < 0 || gy0 + kBoxH > h;
    if (edge_x) {
        for (int i = tid; i < kBoxH * 4; i += kThreads) {      // columns -2, -1, w, w + 1 of every staged row
            const int r = i >> 2, k = i & 3;
            const int x = k < 2 ? k - 2 : w + (k - 2);
            const int c = x - gx0, cs = refl101(x, w) - gx0;
            if (c >= 0 && c < kBoxW && cs >= 0 && cs < kBoxW) tile[r * kBoxW + c] = tile[r * kBoxW + cs];
        }
        __syncthreads();
    }
    if (edge_y) {
        for (int i = tid; i < 4 * (kBoxW / 4); i += kThreads) {   // rows -2, -1, h, h + 1, whole words (patched columns included)
            const int k = i / (kBoxW / 4), cw = i - k * (kBoxW / 4);
            const int y = k < 2 ? k - 2 : h + (k - 2);
            const int r = y - gy0, rs = refl101(y, h) - gy0;
            if (r >= 0 && r < kBoxH && rs >= 0 && rs < kBoxH)
                reinterpret_cast<uint32_t*>(tile)[r * (kBoxW / 4) + cw] = reinterpret_cast<const uint32_t*>(tile)[rs * (kBoxW / 4) + cw];
        }
        __syncthreads();
    }
    // ---- first transition: the staged core (192 x 32 source pixels) goes to the padded home of the source level ----
    if (a.home) {
        uint8_t* hb = a.home + (size_t)img * a.h_stride;
        for (int i = tid; i < 2 * kOH * 12; i += kThreads) {
            const int r = i / 12, c16 = i - r * 12;
            const int y = 2 * oy0 + r, x = 2 * ox0 + 16 * c16;
            if (y < h && x < w)                                  // chunks past the right edge spill into the border region, refilled later
                *reinterpret_cast<uint4*>(hb + (size_t)y * a.hpitch + x) = *reinterpret_cast<const uint4*>(&tile[(r + 2) * kBoxW + 16 + 16 * c16]);
        }
    }
    // ---- vertical [1 4 6 4 1] at even rows: thread = (word column, half of the output rows), five-row window in registers ----
    if (tid < 2 * kColWords) {
        const int half = tid / kColWords, cw = tid - half * kColWords;
        const uint32_t* col = reinterpret_cast<const uint32_t*>(tile) + (kFirstWord + cw);
        const int r0 = 16 * half;                                // box row of the first tap of output row 8 * half
        uint32_t e[5], o[5];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = col[(r0 + k) * (kBoxW / 4)];
            e[k + 2] = v & 0x00ff00ffu; o[k + 2] = (v >> 8) & 0x00ff00ffu;
        }
#pragma unroll
        for (int j = 0; j < kOH / 2; ++j) {
            e[0] = e[2]; e[1] = e[3]; e[2] = e[4]; o[0] = o[2]; o[1] = o[3]; o[2] = o[4];
#pragma unroll
            for (int k = 3; k < 5; ++k) {
                const uint32_t v = col[(r0 + 2 * j + k) * (kBoxW / 4)];
                e[k] = v & 0x00ff00ffu; o[k] = (v >> 8) & 0x00ff00ffu;
            }
            const uint32_t se = e[0] + e[4] + 4u * (e[1] + e[3]) + 6u * e[2];     // lanes: columns 4cw+12 (low), 4cw+14 (high)
            const uint32_t so = o[0] + o[4] + 4u * (o[1] + o[3]) + 6u * o[2];     // lanes: columns 4cw+13, 4cw+15
            const int row = 8 * half + j;
            reinterpret_cast<uint32_t*>(planeE)[row * (kPlane / 2) + cw] = se;    // planeE[row][2cw], [2cw+1] = even columns 12 + 4cw, 14 + 4cw
            reinterpret_cast<uint32_t*>(planeO)[row * (kPlane / 2) + cw] = so;
        }
    }
    __syncthreads();
    // ---- horizontal pass + store: thread = 16 output pixels of one row.  Output x = ox0 + j reads even column 16 + 2j =
    //      planeE index j + 2 (centre), j + 1 and j + 3 (the +-2 taps), planeO index j + 1 and j + 2 (the +-1 taps) ----
    if (tid < kOH * (kOW / 16)) {
        const int row = tid / (kOW / 16), g = tid - row * (kOW / 16);
        const int oy = oy0 + row, ox = ox0 + 16 * g;
        if (oy < a.dh && ox < a.dw) {
            const uint32_t* pe = reinterpret_cast<const uint32_t*>(planeE) + row * (kPlane / 2) + 8 * g;   // word k = E[16g + 2k], E[16g + 2k + 1]
            const uint32_t* po = reinterpret_cast<const uint32_t*>(planeO) + row * (kPlane / 2) + 8 * g;
            uint32_t E[10], O[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) { E[k] = pe[k]; O[k] = po[k]; }
            uint32_t q[4];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // outputs j = 16g + 2k (low lane), + 1 (high lane): centre E[j + 2] = word k + 1
                const uint32_t c = E[k + 1];
                const uint32_t em = __funnelshift_r(E[k], E[k + 1], 16);          // E[j + 1], E[j + 2]
                const uint32_t ep = __funnelshift_r(E[k + 1], E[k + 2], 16);      // E[j + 3], E[j + 4]
                const uint32_t om = __funnelshift_r(O[k], O[k + 1], 16);          // O[j + 1], O[j + 2]
                const uint32_t op = O[k + 1];                                     // O[j + 2], O[j + 3]
                const uint32_t s = em + ep + 4u * (om + op) + 6u * c;             // <= 256 * 255 per lane: no carry between lanes
                const uint32_t t = ((s & 0xffffu) + 128u) >> 8 | ((((s >> 16) + 128u) >> 8) << 8);   // two result bytes
                if (k & 1) q[k >> 1] |= t << 16; else q[k >> 1] = t;
            }
            uint8_t* d = a.dst + (size_t)img * a.d_stride + (size_t)oy * a.dpitch + ox;
            if (ox + 16 <= a.dw) {
                *reinterpret_cast<uint4*>(d) = make_uint4(q[0], q[1], q[2], q[3]);                  // origin and pitch are 16-byte aligned
            } else {                                             // last group of a row whose width is not a multiple of 16
                for (int t = 0; ox + t < a.dw; ++t) d[t] = (uint8_t)(q[t >> 2] >> (8 * (t & 3)));
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

bool encode_level(CUtensorMap* map, const uint8_t* base, int w, int h, int pitch, size_t stream_stride, int n_images) {
    EncodeTiledFn fn = encoder();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_images};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(stream_stride ? stream_stride : (size_t)pitch * h)};
    const cuuint32_t box[3] = {kBoxW, kBoxH, 1}, estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool pyramid_tma_available() { return encoder() != nullptr; }

// One level transition src -> dst.  `home`: padded home of the source level when the source is the unpadded upload buffer.
// Returns false (nothing launched) when the tensor map cannot be encoded; the caller then takes the fused kernel.
bool launch_pyr_tma(const uint8_t* src, int sw, int sh, int spitch, size_t s_stride, uint8_t* dst, int dw, int dh, int dpitch, size_t d_stride,
                    uint8_t* home, int hpitch, size_t h_stride, int n_images, cudaStream_t stream, unsigned long long* trace, int trace_id) {
    TmaPyrArgs a;
    if (((uintptr_t)src & 15) || (spitch & 15) || (s_stride & 15)) return false;
    if (!encode_level(&a.map, src, sw, sh, spitch, s_stride, n_images)) return false;
    a.sw = sw; a.sh = sh; a.dst = dst; a.dw = dw; a.dh = dh; a.dpitch = dpitch; a.d_stride = d_stride;
    a.home = home; a.hpitch = hpitch; a.h_stride = h_stride; a.trace = trace; a.trace_id = trace_id;
    dim3 grid((dw + kOW - 1) / kOW, (dh + kOH - 1) / kOH, n_images);
    return launch_kernel(pyr_tma_kernel, grid, dim3(kThreads), 0, stream, a) == cudaSuccess;
}

cudaError_t pyramid_tma_prepare() {
    return cudaFuncSetAttribute(pyr_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kCarveout);
}

}  // namespace vf
