// Shared device/host definitions of the B200-native front-end (sm_100a).
// Data layout in HBM and the per-stream state are described in DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vins_fast_b200.h"

namespace vf {

constexpr int kMaxLevels = 8;        // pyramid levels 0..7 (BASELINE configs use maxLevel 3..5)
constexpr int kWin = 21;             // LK window (reference: cv::Size(21,21), feature_tracker.cpp:42)
constexpr int kWinPx = kWin * kWin;  // 441
constexpr int kNumSms = 148;         // B200
constexpr int kCarveout = 100;       // preferred shared-memory carveout (percent) requested by EVERY kernel of the library

// One image pyramid resident in HBM.  Every level is stored with a REFLECT_101 border of kPyrPad pixels on all four
// sides (what cv::buildOpticalFlowPyramid keeps around each level for the LK window, lkpyramid.cpp), so the LK kernels
// fetch window tiles with plain aligned 32-bit loads wherever the window lies.  lvl[l] points at pixel (0,0) of the
// level (interior origin, 16-byte aligned); pitch[l] is the padded row pitch (multiple of 16).  The innermost
// kPyrPadFill border pixels are filled (pad_levels_kernel); an LK tile reaches at most 26 pixels outside a level.
constexpr int kPyrPad = 32;
constexpr int kPyrPadFill = 28;
struct PyrView {
    uint8_t* lvl[kMaxLevels];
    int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
    int n_levels;
};

struct CamParams {   // camodocal liftProjective inputs (CataCamera.cc:320-323 inverse K is derived on device)
    int model;
    double xi, k1, k2, p1, p2, fx, fy, cx, cy;
};

// Sortable key of a float (monotone for all finite values, negative included).
__host__ __device__ __forceinline__ uint32_t float_key(float v) {
#if defined(__CUDA_ARCH__)
    uint32_t b = __float_as_uint(v);
#else
    union { float f; uint32_t u; } c; c.f = v; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

__host__ __device__ __forceinline__ int refl101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

// Corner-candidate bucketing over the top `bits` bits of the sortable key: 13 bits (sign, 8 exponent bits,
// 4 mantissa bits -> 16 buckets per octave) up to 1 Mpx, 16 bits (128 per octave) above, so that a bucket
// normally holds far fewer keys than one selection chunk (1024).
constexpr int kMaxBucketBits = 16;
__host__ __device__ __forceinline__ int bucket_of(uint32_t key, int bits) { return (int)(key >> (32 - bits)); }
// Every bucket counter (histogram, scatter cursor, start offset) exists kHistRep times; a candidate uses the copy its pixel
// selects.  The local maxima of a frame fall into a few hundred neighbouring buckets -- a few dozen 128-byte lines -- and
// same-line atomics serialise in L2 (3840x2160: 277 k atomics on ~35 lines cost the response kernel 30 us and the scatter 50);
// the copies spread them over eight times as many lines.
constexpr int kHistRep = 8;
__host__ __device__ __forceinline__ int hist_rep(uint32_t pix) { return (int)((pix >> 4) & (uint32_t)(kHistRep - 1)); }
// index of (bucket, copy) in hist / bucket_fill (ascending buckets)
__host__ __device__ __forceinline__ size_t hist_slot(int bucket, uint32_t pix) { return (size_t)bucket * kHistRep + hist_rep(pix); }

// Per-stream counters kept on the device (int32 each).
enum Counter {
    C_N_PREV = 0,       // features carried in from the previous frame (previous_kps_.size())
    C_N_TRACKED = 1,    // survivors of the temporal LK + flow-back + InBorder
    C_N_KEPT = 2,       // survivors of SetFeatureExtractorMask
    C_N_NEW = 3,        // corners added by goodFeaturesToTrack
    C_N_TOTAL = 4,      // current_kps_.size()
    C_N_RIGHT = 5,      // ids_right_.size()
    C_N_LOCALMAX = 6,   // 3x3 local maxima of the response (threshold- and mask-independent)
    C_LANDMARK_ID = 7,  // landmark_id_ (unsigned, wraps)
    C_N_PREV_LEFT_UN = 8,   // size of previous_left_un_distortion_ids_kps_map_
    C_N_PREV_RIGHT_UN = 9,  // size of previous_right_un_distortion_ids_kps_map_
    C_MAXKEY = 10,      // masked max of the response, as sortable key (0 = no unmasked pixel)
    C_N_NZ_BUCKETS = 11,    // non-empty candidate buckets
    C_N_SURV = 12,          // local maxima outside the mask (survivor list of mask_and_max)
    C_COUNT = 16
};

// Measurement aid: when a trace buffer is attached to the batch (vf_batch_set_trace), thread 0 of every CTA records the
// kernel's start (first CTA) and last end on the GPU's nanosecond timer (the latest launch that used the slot wins): [id + 16 * (frame % 6)][0] = min start, [..][1] = max end.
enum TraceId { TR_CLEAR = 0, TR_PYR_LEFT, TR_PAD_LEFT, TR_LK_TEMPORAL, TR_SELECT, TR_RESPONSE, TR_SCAN, TR_SCATTER, TR_MASK,
               TR_GFTT, TR_PYR_RIGHT, TR_PAD_RIGHT, TR_LK_STEREO, TR_UNDISTORT_LEFT, TR_COUNT };
// Programmatic dependent launch (sm_90+).  A kernel launched with the programmatic-serialization attribute (launch_kernel
// below, while vf::pdl_hint() is on) may be scheduled while its predecessor in the stream still runs: its CTAs sit at
// pdl_wait() until that grid has completed and its writes are visible, so the launch latency of every kernel on the frame's
// dependent chain hides behind the kernel before it.  A kernel lets its successor in as soon as it is itself past its
// wait (pdl_release).  Both are no-ops for a kernel launched without the attribute.
bool& pdl_hint();   // host, vf_api.cu: set around the launches whose predecessor in the stream is one of our kernels
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at = {};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = pdl_hint() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
struct TraceScope {
    unsigned long long* slot;
    __device__ __forceinline__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ TraceScope(unsigned long long* buf, int id) : slot(buf ? buf + 2 * id : nullptr) {
        // first CTA; min over the launches that share the id within a frame (a pyramid is several launches)
        if (slot && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) atomicMin(slot, now());
    }
    __device__ __forceinline__ ~TraceScope() { if (slot && threadIdx.x == 0) atomicMax(slot + 1, now()); }
};
// ad-hoc phase stamps inside a single-CTA kernel (slots 224..255 of the trace buffer)
__device__ __forceinline__ void trace_phase(unsigned long long* buf, int k) {
    if (buf && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) buf[224 + k] = TraceScope::now();
}
#endif

}  // namespace vf
