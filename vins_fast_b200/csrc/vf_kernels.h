// Host-callable launchers of the sm_100a kernels (one translation unit per stage).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vf_common.cuh"

namespace vf {

// Device-resident description of a batch of S independent streams (one copy in device memory, one on the host).
// Per-stream arrays are laid out [S][cap] (or [S][H][W]); "parity" double-buffers what one frame hands to the next.
struct DevBatch {
    int n_streams, W, H, cap, min_dist, flow_back, lk_max_level, n_levels;
    int lk_win;                      // LK window width (odd, 5..31); 21 = the reference's literal, served by the tuned kernels
    int bucket_bits, n_buckets;      // candidate bucketing resolution (vf_common.cuh: bucket_of)
    size_t cand_cap;                 // per-stream capacity of cand / cand_sorted (power of two >= W*H)
    double quality;
    CamParams cam[2];

    PyrView left[3];                 // left pyramids, ring of three: frame k builds slot k % 3 while frame k-1 may still track from
                                     // slot (k-2) % 3 into slot (k-1) % 3; stream 0 base
    PyrView right[2];                // right pyramids, slot = frame parity (the stereo LK of frame k may still read its
                                     // pyramid while frame k+1 builds the next one)
    size_t lvl_stream_stride[kMaxLevels];   // bytes between consecutive streams of one level
    const uint8_t* stage_left[2];    // [parity][S][H][stage_pitch] the uploaded left image, unpadded (read by the corner response)
    int stage_pitch;
    size_t stage_stride;

    float2* kps[2];                  // [parity][S][cap] current_kps_ at the end of the frame
    int* ids[2];                     // [parity][S][cap] ids_
    int* cnt[2];                     // [parity][S][cap] track_count_
    int* prev_idx[2];                // [parity][S][cap] index of the feature in the PREVIOUS frame's arrays, -1 for a new corner:
                                     // what the reference looks up through its id -> point maps (ComputeKpsVelocity)
    float2* un_left[2];              // [parity][S][cap] current_un_distortion_kps_ (keyed by ids[parity])
    float2* un_right[2];             // [parity][S][cap] right undistorted point where right_ok
    uint8_t* right_ok[2];            // [parity][S][cap] 1 = id has a right observation this frame
    int* counters[2];                // [parity][S][C_COUNT]
    double* times[2];                // (unused by the kernels: the frame interval reaches them through mapped host memory)

    // per-frame intermediates (kept for the debug getters)
    float2* fwd_pts; uint8_t* fwd_st; float2* rev_pts; uint8_t* rev_st; uint8_t* temporal_st;   // [S][cap]
    float2* lr_pts; uint8_t* lr_st; float2* rl_pts; uint8_t* rl_st;                              // [S][cap]
    int* sort_perm;                  // [S][cap]
    int* lk_iters;                   // [S][cap] inner LK iterations of the temporal passes (indexed like previous_kps_)
    int* lk_iters_stereo;            // [S][cap] inner LK iterations of the stereo passes (indexed like current_kps_)

    // Corner stage (clear -> response -> scan -> scatter): depends on the uploaded image only, so it is double-buffered by
    // frame parity and runs AHEAD of the frame's chain, while the previous frame is still on it.
    float* eig[2];                   // [parity][S][H][W] cornerMinEigenVal of the current left image
    uint8_t* lmflag[2];              // [parity][S][H][W] 1 = 3x3 local maximum of the response (candidate of goodFeaturesToTrack)
    unsigned long long* cand[2];     // [parity][S][cand_cap] local maxima, key = float_key(eig) << 32 | raster index (unordered);
                                     //               reused as sort scratch by gftt_select once scattered
    unsigned long long* cand_sorted[2]; // [parity][S][cand_cap] same keys grouped by bucket, strongest bucket first
    unsigned int* hist[2];           // [parity][S][n_buckets][kHistRep] candidate counts; a candidate uses the copy its pixel selects (vf_common.cuh: hist_slot)
    unsigned int* bucket_start[2];   // [parity][S][n_buckets * kHistRep + 8] (one entry past the end + padding that keeps every stream 16-byte aligned) start of (bucket d, copy) in cand_sorted at index 8 d + 7 - copy, d = n_buckets-1-bucket (descending): bucket d spans [8 d], [8 (d + 1)])
    unsigned int* bucket_fill[2];    // [parity][S][n_buckets][kHistRep] scatter cursors (indexed like hist)
    int* bucket_nz[2];               // [parity][S][n_buckets] the non-empty d, ascending (counter C_N_NZ_BUCKETS)
    int* ccnt[2];                    // [parity][S][C_COUNT] the corner stage's own counters (C_N_LOCALMAX, C_N_NZ_BUCKETS)
    uint8_t* mask;                   // [S][H][W] SetFeatureExtractorMask
    unsigned long long* surv;        // [S][cand_cap] candidates outside the mask (unordered), written by mask_and_max

    vf_feature* out[3];              // [frame % 3][S][cap] packed result records, in MAPPED PINNED HOST memory (zero-copy download);
                                     // a ring of three so that the host may collect frame k-2 while k-1 and k are in flight
    unsigned long long* trace;       // optional [TR_COUNT][2] kernel start / end stamps (measurement aid), else null
};

// ---- vf_pyramid.cu ----
struct Level0Src {   // unpadded staging image(s): rows 16-byte aligned
    const uint8_t* p; int pitch; size_t stride;
};
int launch_pyramid(const PyrView& pv, const size_t* stride, int n_images, cudaStream_t stream, const Level0Src* src0 = nullptr,
                   unsigned long long* trace = nullptr, int trace_id = 0);
void launch_scharr(const uint8_t* src, int w, int h, int pitch, int16_t* dst, cudaStream_t stream);
void set_pyramid_kernel_mode(int mode);   // 0 by size, 1 fused kernel only, 2 streaming (TMA) kernel everywhere (parity tests)

// ---- vf_pyramid_tma.cu ----
bool pyramid_tma_available();
cudaError_t pyramid_tma_prepare();
bool launch_pyr_tma(const uint8_t* src, int sw, int sh, int spitch, size_t s_stride, uint8_t* dst, int dw, int dh, int dpitch, size_t d_stride,
                    uint8_t* home, int hpitch, size_t h_stride, int n_images, cudaStream_t stream, unsigned long long* trace, int trace_id);

// ---- vf_lk.cu ----
// `cur` / `prev`: slots of the current and of the previous left pyramid in DevBatch::left
void launch_lk_temporal(const DevBatch* d_batch, const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream);
// dt_host: [S] time since the previous frame, n_host: [S] feature counts -- both in mapped pinned host memory
void launch_lk_stereo(const DevBatch* d_batch, const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream);
void launch_undistort_left(const DevBatch* d_batch, const DevBatch& h, int parity, const double* dt_host, int* n_host, cudaStream_t stream);
cudaError_t corners_prepare();
cudaError_t pyramid_prepare();
cudaError_t lk_prepare();   // opt in to the LK kernels' dynamic shared memory (once per device, outside stream capture)
void launch_lk_op(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status,
                  int* iters, int n, int max_level, int use_initial_flow, cudaStream_t stream);

// ---- vf_lk_warp.cu: the same passes with one warp per feature ----
void set_lk_kernel_mode(int mode);        // 0 = warp-level passes (COOP / BATCH by size), 1 = vf_lk.cu kernels, 2 = BATCH, 3 = COOP (A/B, parity tests)
bool lk_use_warp_kernel(long long n_features);
cudaError_t lk_warp_prepare();
void launch_lk_temporal_warp(const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream);
void launch_lk_stereo_warp(const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream);
void launch_lk_op_warp(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status, int* iters, int n,
                       int max_level, int use_initial_flow, cudaStream_t stream);

// ---- vf_lk_generic.cu: the same four calls for a window width given at run time (lk_win != 21) ----
bool lk_window_supported(int win);        // odd widths 5..31
cudaError_t lk_generic_prepare();
void launch_lk_temporal_generic(const DevBatch& h, int parity, int cur, int prev, cudaStream_t stream);
void launch_lk_stereo_generic(const DevBatch& h, int parity, int cur, const double* dt_host, cudaStream_t stream);
void launch_lk_op_generic(const PyrView& prev, const PyrView& next, const float2* prev_pts, float2* next_pts, uint8_t* status, int* iters,
                          int n, int max_level, int use_initial_flow, int win, cudaStream_t stream);

// ---- vf_corners.cu ----
void set_response_kernel_mode(int mode);   // 0 auto, 1 tiles, 2 strips (parity tests)
void launch_response_nms(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_bucket_scan(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_bucket_scatter(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_mask_and_max(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_gftt_select(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_clear_frame(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_masked_max_from_mask(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
void launch_undistort(const CamParams& cam, const float2* uv, float2* out, int n, cudaStream_t stream);

// ---- vf_select.cu ----
void launch_select_tracked(const DevBatch* d_batch, const DevBatch& h, int parity, cudaStream_t stream);
size_t select_tracked_smem_bytes(int cap);
cudaError_t select_tracked_prepare(int cap);

}  // namespace vf
