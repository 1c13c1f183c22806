// C ABI of the B200-native stereo front-end (include/vins_fast_b200.h): handle management, HBM layout,
// per-frame orchestration (streams, events, CUDA graph), debug getters and the stand-alone stage operators.
//
// One frame (= FeatureTracker::TrackImage, vins/estimator/feature_tracker.cpp:27-217) on the device:
//
//   main stream                                   side A                         side B
//   ----------------------------------------     ---------------------------    ------------------------
//   H2D times, H2D/D2D left image                                                H2D/D2D right image
//   [graph: left chain]                                                          pyrdown right L1..Ln
//     clear_frame, pyrdown left L1..Ln
//     lk_temporal  (fwd + rev LK, status)         response_nms (eig + local max)
//     select_tracked (compaction, std::sort        bucket_scan
//        order, min_dist keep)                     bucket_scatter
//     mask_and_max   <- joins side A
//     gftt_select    (top-up corners, new ids)
//   lk_stereo (L->R, R->L, undistort, velocity,   <- joins side B
//              packed featureFrame records)
//   D2H records + counters
//
// Only `parity` (frame index & 1) changes between launches, so the left chain is captured once per parity into a
// CUDA graph and replayed.  Everything a frame hands to the next one (pyramid, points, ids, track counts,
// undistorted points, landmark id, time) lives in HBM, double-buffered by parity; nothing but the two images
// goes up and nothing but the <= max_cnt records comes down.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges around the host-side stages of a frame (free when no tool is attached)

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <string>
#include <vector>

#include "vf_common.cuh"
#include "vf_kernels.h"

using namespace vf;

static thread_local std::string g_last_error;

extern "C" const char* vf_last_global_error(void) { return g_last_error.c_str(); }
extern "C" void vf_internal_set_error(const char* msg) { g_last_error = msg ? msg : ""; }

struct vf_tracker {
    vf_batch* batch;
    int stream_index;
    bool owns_batch;
};

struct vf_batch {
    vf_config cfg;
    int S = 0, device = 0;
    cudaStream_t main = nullptr, side_a = nullptr, side_b = nullptr, side_c = nullptr, side_d = nullptr, side_e = nullptr;
    bool own_main = false, serialized = false;
    bool one_graph_sync = false;           // synchronous frames replay SEQ_FRAME_SYNC (one graph) instead of four
    bool corner_after_pyr = false;         // ... whose corner stage starts behind the left pyramid instead of beside it
    bool chain_direct = false;             // overlapped frames issue the chain's four kernels directly instead of replaying a graph
    double host_wait_us = 0, host_issue_us = 0;   // host time spent waiting for frame k-2 / issuing frame k (cumulative)
    cudaEvent_t ev_input = nullptr;        // device-input ordering: recorded on the caller's stream / handed in by the caller
    cudaEvent_t ev_user_input = nullptr;   // vf_batch_set_input_event: the next frame's uploads wait for it (not owned)
    cudaEvent_t ev_fork = nullptr, ev_right = nullptr, ev_done[4] = {nullptr, nullptr, nullptr, nullptr}, ev_img = nullptr,
                ev_sel = nullptr, ev_un = nullptr, ev_pyrl = nullptr, ev_scat = nullptr, ev_rimg = nullptr, ev_fork2 = nullptr, ev_pyr_sync = nullptr,
                ev_chain[2] = {nullptr, nullptr};
    uint8_t* stage[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // unpadded upload targets [camera][parity]: [S][H][stage_pitch]
    int stage_pitch = 0;
    DevBatch h;               // host copy of the device descriptor
    DevBatch* d = nullptr;    // device copy
    std::vector<void*> allocs;
    vf_feature* pin_out[3] = {nullptr, nullptr, nullptr};   // mapped pinned: records, ring of three (slot = frame % 3)
    int* pin_n[3] = {nullptr, nullptr, nullptr};            // mapped pinned: [S] feature counts, same ring
    int* dev_n[3] = {nullptr, nullptr, nullptr};            // their device addresses
    double* pin_dt = nullptr;                      // pinned ring [6][S] (slot = frame % 6): time since the previous frame; dev_dt = its device copy
    double* dev_dt = nullptr;
    std::vector<double> prev_times;
    long frame = 0;
    long known_done = -1;                          // newest frame index a fetch has seen complete (frames complete in order)
    int last_parity = -1;
    cudaGraphExec_t graph[8][6] = {};
    int pyr_launches = 1;   // one per frame % 6 (parity x pyramid ring slot)
    int launches = 0;
    std::string err;
    std::vector<vf_tracker> views;
    size_t bytes_allocated = 0;
    // optional per-stage timing (vf_batch_set_profiling): CUDA events around every kernel, plain launches only
    bool profiling = false;
    cudaEvent_t st_beg[16] = {nullptr}, st_end[16] = {nullptr};
    bool st_used[16] = {false};
    float stage_ms[16] = {0};
};

enum Stage { ST_H2D_LEFT = 0, ST_PYR_LEFT, ST_LK_TEMPORAL, ST_SELECT_TRACKED, ST_RESPONSE_NMS, ST_BUCKET_SCAN, ST_BUCKET_SCATTER,
             ST_MASK_AND_MAX, ST_GFTT_SELECT, ST_H2D_RIGHT, ST_PYR_RIGHT, ST_LK_STEREO, ST_D2H, ST_CLEAR, ST_UNDISTORT_LEFT, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"h2d_left", "pyrdown_left", "lk_temporal", "select_tracked", "response_nms", "bucket_scan",
                                            "bucket_scatter", "mask_and_max", "gftt_select", "h2d_right", "pyrdown_right", "lk_stereo",
                                            "d2h", "clear_frame", "undistort_left"};

struct StageTimer {   // records begin/end events on `st` around a scope when profiling is on
    vf_batch* b; int id; cudaStream_t st;
    StageTimer(vf_batch* b_, int id_, cudaStream_t st_) : b(b_), id(id_), st(st_) {
        if (b->profiling) { cudaEventRecord(b->st_beg[id], st); b->st_used[id] = true; }
    }
    ~StageTimer() { if (b->profiling) cudaEventRecord(b->st_end[id], st); }
};

namespace {

struct NvtxRange {   // RAII range in the "vins_fast_b200" view of nsys / ncu timelines
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (dev >= 0 && dev != prev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define VF_CUDA(bt, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            char buf__[512];                                                                                \
            snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            if (bt) (bt)->err = buf__;                                                                      \
            g_last_error = buf__;                                                                           \
            return VF_ERR_CUDA;                                                                             \
        }                                                                                                   \
    } while (0)

vf_status fail(vf_batch* b, vf_status st, const std::string& msg) {
    if (b) b->err = msg;
    g_last_error = msg;
    return st;
}

int effective_levels(int w, int h, int max_level, int win) {   // buildOpticalFlowPyramid's early stop
    int n = 1;
    for (int l = 0; l <= max_level; ++l) {
        n = l + 1;
        const int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win || nh <= win) break;
        w = nw; h = nh;
    }
    return n;
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

template <typename T>
vf_status dalloc(vf_batch* b, T** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    VF_CUDA(b, cudaMalloc(&p, bytes));
    VF_CUDA(b, cudaMemset(p, 0, bytes));
    b->allocs.push_back(p);
    b->bytes_allocated += bytes;
    *out = reinterpret_cast<T*>(p);
    return VF_OK;
}

vf_status alloc_pyramid(vf_batch* b, PyrView& pv, int W, int H, int n_levels, int S) {
    memset(&pv, 0, sizeof(pv));
    pv.n_levels = n_levels;
    int w = W, h = H;
    for (int l = 0; l < n_levels; ++l) {
        pv.w[l] = w; pv.h[l] = h; pv.pitch[l] = align_up(w + 2 * kPyrPad, 16);
        b->h.lvl_stream_stride[l] = (size_t)pv.pitch[l] * (h + 2 * kPyrPad);
        uint8_t* base = nullptr;
        vf_status st = dalloc(b, &base, b->h.lvl_stream_stride[l] * S);
        if (st != VF_OK) return st;
        pv.lvl[l] = base + (size_t)kPyrPad * pv.pitch[l] + kPyrPad;   // interior origin of stream 0
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
    return VF_OK;
}

vf_status validate_config(const vf_config* c, int n_streams) {
    if (!c) return fail(nullptr, VF_ERR_INVALID_ARG, "config is null");
    if (n_streams < 1 || n_streams > 4096) return fail(nullptr, VF_ERR_INVALID_ARG, "n_streams must be in [1, 4096]");
    if (c->width < 64 || c->height < 64 || c->width > 16384 || c->height > 16384)
        return fail(nullptr, VF_ERR_INVALID_ARG, "image size must be within [64, 16384] in both dimensions");
    if (c->max_cnt < 1 || c->max_cnt > 4096) return fail(nullptr, VF_ERR_INVALID_ARG, "max_cnt must be in [1, 4096]");
    if (c->min_dist < 0 || c->min_dist > 1024) return fail(nullptr, VF_ERR_INVALID_ARG, "min_dist must be in [0, 1024]");
    if (!lk_window_supported(c->lk_win))
        return fail(nullptr, VF_ERR_INVALID_ARG, "lk_win must be odd and in [5, 31] (the reference's cv::Size(21, 21) is the default)");
    if (c->lk_max_level < 0 || c->lk_max_level >= kMaxLevels)
        return fail(nullptr, VF_ERR_INVALID_ARG, "lk_max_level must be in [0, 7]");
    if (!(c->quality_level > 0.0 && c->quality_level <= 1.0))
        return fail(nullptr, VF_ERR_INVALID_ARG, "quality_level must be in (0, 1]");
    for (int k = 0; k < 2; ++k) {
        if (c->cam[k].model != VF_CAM_MEI && c->cam[k].model != VF_CAM_PINHOLE)
            return fail(nullptr, VF_ERR_INVALID_ARG, "camera model must be VF_CAM_MEI or VF_CAM_PINHOLE");
        if (c->cam[k].fx == 0.0 || c->cam[k].fy == 0.0)
            return fail(nullptr, VF_ERR_INVALID_ARG, "camera focal lengths must be non-zero");
    }
    return VF_OK;
}

CamParams to_cam(const vf_camera& c) {
    CamParams p;
    p.model = c.model; p.xi = c.xi; p.k1 = c.k1; p.k2 = c.k2; p.p1 = c.p1; p.p2 = c.p2;
    p.fx = c.fx; p.fy = c.fy; p.cx = c.cx; p.cy = c.cy;
    return p;
}

cudaError_t sync_all(vf_batch* b) {   // every stream a frame is spread over
    cudaError_t e = cudaSuccess;
    for (cudaStream_t st : {b->main, b->side_a, b->side_b, b->side_c, b->side_d, b->side_e})
        if (st) { const cudaError_t r = cudaStreamSynchronize(st); if (r != cudaSuccess) e = r; }
    return e;
}

void destroy_batch(vf_batch* b) {
    if (!b) return;
    DeviceGuard g(b->device);
    sync_all(b);
    for (int w = 0; w < 8; ++w) for (int p = 0; p < 6; ++p) if (b->graph[w][p]) cudaGraphExecDestroy(b->graph[w][p]);
    for (void* p : b->allocs) cudaFree(p);
    if (b->d) cudaFree(b->d);
    for (int p = 0; p < 3; ++p) {
        if (b->pin_out[p]) cudaFreeHost(b->pin_out[p]);
        if (b->pin_n[p]) cudaFreeHost(b->pin_n[p]);
    }
    for (int p = 0; p < 4; ++p) if (b->ev_done[p]) cudaEventDestroy(b->ev_done[p]);
    if (b->pin_dt) cudaFreeHost(b->pin_dt);
    if (b->ev_right) cudaEventDestroy(b->ev_right);
    if (b->ev_input) cudaEventDestroy(b->ev_input);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_img) cudaEventDestroy(b->ev_img);
    if (b->ev_sel) cudaEventDestroy(b->ev_sel);
    if (b->ev_un) cudaEventDestroy(b->ev_un);
    if (b->ev_pyrl) cudaEventDestroy(b->ev_pyrl);
    if (b->ev_scat) cudaEventDestroy(b->ev_scat);
    if (b->ev_rimg) cudaEventDestroy(b->ev_rimg);
    if (b->ev_fork2) cudaEventDestroy(b->ev_fork2);
    if (b->ev_pyr_sync) cudaEventDestroy(b->ev_pyr_sync);
    for (int p = 0; p < 2; ++p) if (b->ev_chain[p]) cudaEventDestroy(b->ev_chain[p]);
    for (int k = 0; k < 16; ++k) { if (b->st_beg[k]) cudaEventDestroy(b->st_beg[k]); if (b->st_end[k]) cudaEventDestroy(b->st_end[k]); }
    if (b->side_a && !b->serialized) cudaStreamDestroy(b->side_a);
    if (b->side_b && !b->serialized) cudaStreamDestroy(b->side_b);
    if (b->side_c && !b->serialized) cudaStreamDestroy(b->side_c);
    if (b->side_d && !b->serialized) cudaStreamDestroy(b->side_d);
    if (b->side_e && !b->serialized) cudaStreamDestroy(b->side_e);
    if (b->own_main && b->main) cudaStreamDestroy(b->main);
    delete b;
}

vf_status reset_state(vf_batch* b) {
    DeviceGuard g(b->device);
    VF_CUDA(b, sync_all(b));
    const size_t S = b->S;
    for (int p = 0; p < 2; ++p) {
        VF_CUDA(b, cudaMemsetAsync(b->h.counters[p], 0, sizeof(int) * S * C_COUNT, b->main));
        VF_CUDA(b, cudaMemsetAsync(b->h.times[p], 0, sizeof(double) * S, b->main));
        VF_CUDA(b, cudaMemsetAsync(b->h.right_ok[p], 0, S * b->h.cap, b->main));
    }
    VF_CUDA(b, cudaStreamSynchronize(b->main));
    b->frame = 0;
    b->known_done = -1;
    b->last_parity = -1;
    std::fill(b->prev_times.begin(), b->prev_times.end(), 0.0);
    return VF_OK;
}

// The kernel sequences of a frame.  Each can be issued directly or captured into a graph (one executable per frame % 6:
// parity, slots of the left pyramid ring, the frame-interval slot and the trace slot are baked into the launches).
//   SEQ_CORNER     side A: per-frame accumulators, response + local maxima, candidate buckets.  Reads the uploaded image
//                  only and writes parity-buffered outputs, so it may run ahead of the frame's chain.
//   SEQ_CHAIN      main: the frame-to-frame chain proper, four kernels in a row with no join in between -- every
//                  cross-stream input (left pyramid, corner stage) is waited for once, ahead of the first kernel.
//   SEQ_HEAD_SYNC  main, synchronous frames: left pyramid + temporal LK back to back (behind the left upload).
//   SEQ_REST_SYNC  main, synchronous frames: what follows the temporal LK, with the stereo pass appended and the left
//                  undistortion on a branch beside it.
//   SEQ_PYR_LEFT / SEQ_PYR_RIGHT   side B / side D: pyramid + borders of one image (overlapped frames; right: both plans).
//   SEQ_TAIL       side C, overlapped frames: stereo pass with the left undistortion on a branch (side E) beside it.
// (A plain launch of one of these kernels costs the host 4-4.5 us -- the batch descriptor travels by value -- a graph
//  launch 3.5 us whatever it holds.)
//   SEQ_FRAME_SYNC a whole synchronous frame as ONE graph: head, corner stage (side A), right pyramid (side D, behind an
//                  EXTERNAL wait for the right upload's event) and the rest, joined inside the graph -- no graph boundary
//                  between lk_temporal and select_tracked.
enum { SEQ_CORNER = 0, SEQ_CHAIN = 1, SEQ_REST_SYNC = 2, SEQ_HEAD_SYNC = 3, SEQ_PYR_LEFT = 4, SEQ_PYR_RIGHT = 5, SEQ_TAIL = 6, SEQ_FRAME_SYNC = 7,
       SEQ_COUNT = 8 };

}  // namespace
namespace vf { bool& pdl_hint() { static thread_local bool on = false; return on; } }
namespace {
// Turns the hint on for the launches of its scope: kernels that directly follow another kernel of ours in their stream.
// Off while profiling (the stage timers' events sit between the kernels) and when VF_NO_PDL is set (A/B measurements).
struct PdlScope {
    bool prev;
    PdlScope(const vf_batch* b, bool on) : prev(pdl_hint()) {
        static const bool disabled = getenv("VF_NO_PDL") != nullptr;
        pdl_hint() = on && !b->profiling && !disabled;
    }
    ~PdlScope() { pdl_hint() = prev; }
};

vf_status issue_sequence(vf_batch* b, int which, int parity_, int cur, int prev, int phase) {
    const DevBatch& h = b->h;
    const int parity = parity_ | (phase << 4);   // kernels take the frame parity in bit 0 and the trace slot above it
    const double* dev_dt = b->dev_dt + (size_t)phase * b->S;
    if (which == SEQ_FRAME_SYNC) {
        // fork: the corner stage (reads the staged left image only) and the right pyramid (its upload is outside the graph:
        // an external event wait) run beside the head of the chain
        VF_CUDA(b, cudaEventRecord(b->ev_fork2, b->main));
        VF_CUDA(b, cudaStreamWaitEvent(b->side_d, b->ev_fork2, 0));
        vf_status st = issue_sequence(b, SEQ_HEAD_SYNC, parity_, cur, prev, phase);   // records ev_pyr_sync behind the pyramid
        if (st != VF_OK) return st;
        // corner stage: beside the head of the chain, or (corner_after_pyr) behind the left pyramid, so that the pyramid -- on
        // the critical path -- has the SMs to itself; the response then runs beside the temporal LK, which leaves most issue slots free
        VF_CUDA(b, cudaStreamWaitEvent(b->side_a, b->corner_after_pyr ? b->ev_pyr_sync : b->ev_fork2, 0));
        if ((st = issue_sequence(b, SEQ_CORNER, parity_, cur, prev, phase)) != VF_OK) return st;
        VF_CUDA(b, cudaEventRecord(b->ev_scat, b->side_a));
        {   // the right upload's event is recorded outside the graph: an external wait node (the flag only exists under capture)
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            VF_CUDA(b, cudaStreamIsCapturing(b->side_d, &cs));
            VF_CUDA(b, cudaStreamWaitEvent(b->side_d, b->ev_rimg, cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : 0));
        }
        if ((st = issue_sequence(b, SEQ_PYR_RIGHT, parity_, cur, prev, phase)) != VF_OK) return st;
        VF_CUDA(b, cudaEventRecord(b->ev_right, b->side_d));
        // select_tracked only needs lk_temporal: it follows it directly (programmatic edge); the joins come behind it
        { StageTimer t(b, ST_SELECT_TRACKED, b->main); PdlScope pdl(b, true); launch_select_tracked(b->d, h, parity, b->main); }
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_scat, 0));
        { StageTimer t(b, ST_MASK_AND_MAX, b->main); launch_mask_and_max(b->d, h, parity, b->main); }
        { StageTimer t(b, ST_GFTT_SELECT, b->main); PdlScope pdl(b, true); launch_gftt_select(b->d, h, parity, b->main); }
        VF_CUDA(b, cudaEventRecord(b->ev_sel, b->main));
        VF_CUDA(b, cudaStreamWaitEvent(b->side_e, b->ev_sel, 0));
        { StageTimer t(b, ST_UNDISTORT_LEFT, b->side_e); launch_undistort_left(b->d, h, parity, dev_dt, b->dev_n[phase % 3], b->side_e); }
        VF_CUDA(b, cudaEventRecord(b->ev_un, b->side_e));
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_right, 0));
        { StageTimer t(b, ST_LK_STEREO, b->main); launch_lk_stereo(b->d, h, parity, cur, dev_dt, b->main); }
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_un, 0));
    } else if (which == SEQ_HEAD_SYNC) {
        {
            StageTimer t(b, ST_PYR_LEFT, b->main);
            PdlScope pdl(b, true);   // the border kernel follows the pyramid kernel
            const Level0Src src = {b->stage[0][parity_], b->stage_pitch, (size_t)b->stage_pitch * h.H};
            b->pyr_launches = launch_pyramid(h.left[cur], h.lvl_stream_stride, b->S, b->main, &src, h.trace, TR_PYR_LEFT + 16 * phase);
        }
        VF_CUDA(b, cudaEventRecord(b->ev_pyr_sync, b->main));
        { StageTimer t(b, ST_LK_TEMPORAL, b->main); PdlScope pdl(b, true); launch_lk_temporal(b->d, h, parity, cur, prev, b->main); }
    } else if (which == SEQ_PYR_LEFT || which == SEQ_PYR_RIGHT) {
        const int cam = which == SEQ_PYR_RIGHT;
        cudaStream_t st = cam ? b->side_d : b->side_b;
        StageTimer t(b, cam ? ST_PYR_RIGHT : ST_PYR_LEFT, st);
        PdlScope pdl(b, true);
        const Level0Src src = {b->stage[cam][parity_], b->stage_pitch, (size_t)b->stage_pitch * h.H};
        b->pyr_launches = launch_pyramid(cam ? h.right[parity_] : h.left[cur], h.lvl_stream_stride, b->S, st, &src, h.trace,
                                         (cam ? TR_PYR_RIGHT : TR_PYR_LEFT) + 16 * phase);
    } else if (which == SEQ_TAIL) {
        VF_CUDA(b, cudaEventRecord(b->ev_fork, b->side_c));
        VF_CUDA(b, cudaStreamWaitEvent(b->side_e, b->ev_fork, 0));
        { StageTimer t(b, ST_UNDISTORT_LEFT, b->side_e); launch_undistort_left(b->d, h, parity, dev_dt, b->dev_n[phase % 3], b->side_e); }
        VF_CUDA(b, cudaEventRecord(b->ev_un, b->side_e));
        { StageTimer t(b, ST_LK_STEREO, b->side_c); launch_lk_stereo(b->d, h, parity, cur, dev_dt, b->side_c); }
        VF_CUDA(b, cudaStreamWaitEvent(b->side_c, b->ev_un, 0));
    } else if (which == SEQ_CORNER) {
        { StageTimer t(b, ST_CLEAR, b->side_a); launch_clear_frame(b->d, h, parity, b->side_a); }
        { StageTimer t(b, ST_RESPONSE_NMS, b->side_a); launch_response_nms(b->d, h, parity, b->side_a); }
        { StageTimer t(b, ST_BUCKET_SCAN, b->side_a); launch_bucket_scan(b->d, h, parity, b->side_a); }
        { StageTimer t(b, ST_BUCKET_SCATTER, b->side_a); launch_bucket_scatter(b->d, h, parity, b->side_a); }
    } else {
        if (which == SEQ_CHAIN) { StageTimer t(b, ST_LK_TEMPORAL, b->main); launch_lk_temporal(b->d, h, parity, cur, prev, b->main); }
        { StageTimer t(b, ST_SELECT_TRACKED, b->main); PdlScope pdl(b, which == SEQ_CHAIN); launch_select_tracked(b->d, h, parity, b->main); }
        { StageTimer t(b, ST_MASK_AND_MAX, b->main); PdlScope pdl(b, true); launch_mask_and_max(b->d, h, parity, b->main); }
        { StageTimer t(b, ST_GFTT_SELECT, b->main); PdlScope pdl(b, true); launch_gftt_select(b->d, h, parity, b->main); }
        if (which == SEQ_REST_SYNC) {
            VF_CUDA(b, cudaEventRecord(b->ev_sel, b->main));
            VF_CUDA(b, cudaStreamWaitEvent(b->side_e, b->ev_sel, 0));
            { StageTimer t(b, ST_UNDISTORT_LEFT, b->side_e); launch_undistort_left(b->d, h, parity, dev_dt, b->dev_n[phase % 3], b->side_e); }
            VF_CUDA(b, cudaEventRecord(b->ev_un, b->side_e));
            { StageTimer t(b, ST_LK_STEREO, b->main); PdlScope pdl(b, true); launch_lk_stereo(b->d, h, parity, cur, dev_dt, b->main); }
            VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_un, 0));
        }
    }
    VF_CUDA(b, cudaGetLastError());
    return VF_OK;
}

vf_status build_graph(vf_batch* b, int phase, int which) {
    cudaGraph_t graph = nullptr;
    const int parity = phase & 1, cur = phase % 3, prev = (phase + 2) % 3;
    cudaStream_t st_cap = which == SEQ_CORNER ? b->side_a : which == SEQ_PYR_LEFT ? b->side_b : which == SEQ_PYR_RIGHT ? b->side_d
                          : which == SEQ_TAIL ? b->side_c : b->main;
    VF_CUDA(b, cudaStreamBeginCapture(st_cap, cudaStreamCaptureModeThreadLocal));
    vf_status st = issue_sequence(b, which, parity, cur, prev, phase);
    cudaError_t e = cudaStreamEndCapture(st_cap, &graph);
    if (st != VF_OK) { if (graph) cudaGraphDestroy(graph); return st; }
    VF_CUDA(b, e);
    e = cudaGraphInstantiate(&b->graph[which][phase], graph, 0);
    cudaGraphDestroy(graph);
    VF_CUDA(b, e);
    VF_CUDA(b, cudaGraphUpload(b->graph[which][phase], st_cap));   // keep the executable resident on the device: shorter launch latency
    return VF_OK;
}

// Enqueue one frame.  img pointers: host or device (kind says which); image s of camera c at img[c] + s*img_stride
// when `contiguous`, else at imgs[c][s].
vf_status enqueue_frame(vf_batch* b, const double* times, const uint8_t* const* img0, const uint8_t* const* img1,
                        const uint8_t* base0, const uint8_t* base1, size_t image_stride, size_t stride0, size_t stride1,
                        cudaMemcpyKind kind, bool overlap) {
    NvtxRange nvtx_frame(overlap ? "vf.enqueue_frame(overlapped)" : "vf.enqueue_frame(synchronous)");
    DeviceGuard g(b->device);
    if (!g.ok) return fail(b, VF_ERR_CUDA, "cudaSetDevice failed");
    const DevBatch& h = b->h;
    const int S = b->S, W = h.W, H = h.H, parity = (int)(b->frame & 1);
    const int phase = (int)(b->frame % 6), cur = phase % 3, prev = (phase + 2) % 3;   // slots of the left pyramid ring
    if (!times) return fail(b, VF_ERR_INVALID_ARG, "times is null");
    // Buffers are double-buffered by parity: everything frame k overwrites was last used by frame k-2, whose completion
    // (stereo pass + download) the GPU streams wait for below.  The host itself only throttles at four frames ahead, so
    // that issuing frame k+1 never has to wait for the GPU to drain (the completion events form a ring of four).
    const int slot = (int)(b->frame & 3);
    const auto t_wait0 = std::chrono::steady_clock::now();
    if (b->frame - 4 > b->known_done) VF_CUDA(b, cudaEventSynchronize(b->ev_done[slot]));   // frame k-4 (frames complete in order)
    const auto t_wait1 = std::chrono::steady_clock::now();
    b->host_wait_us += std::chrono::duration<double, std::micro>(t_wait1 - t_wait0).count();
    if (stride0 < (size_t)W || stride1 < (size_t)W) return fail(b, VF_ERR_INVALID_ARG, "row stride is smaller than the image width");
    for (int s = 0; s < S; ++s) {
        const uint8_t* a = base0 ? base0 + (size_t)s * image_stride : (img0 ? img0[s] : nullptr);
        const uint8_t* c = base1 ? base1 + (size_t)s * image_stride : (img1 ? img1[s] : nullptr);
        // the reference asserts both images are non-empty (stereo only): feature_tracker.cpp:30
        if (!a || !c) return fail(b, VF_ERR_INVALID_ARG, "img0/img1 must both be non-null (stereo is mandatory)");
    }
    // GPU-side waits for the frames whose buffers this one reuses -- skipped when the host has already seen those frames
    // complete (synchronous use: every wait in front of the left upload is on the critical path)
    const bool done2 = b->known_done >= b->frame - 2, done3 = b->known_done >= b->frame - 3;
    if (b->frame >= 2) {
        // the left upload + pyramid only reuse what the CHAIN of frame k-2 read (staging image) and what the stereo pass of
        // frame k-3 read (slot k % 3 of the pyramid ring): they may run while frame k-2's stereo pass is still going
        if (!done2) VF_CUDA(b, cudaStreamWaitEvent(b->side_b, b->ev_chain[parity], 0));
        if (b->frame >= 3 && !done3) VF_CUDA(b, cudaStreamWaitEvent(b->side_b, b->ev_done[(slot + 1) & 3], 0));
    }
    double* pdt = b->pin_dt + (size_t)phase * S;                              // current_time_ - previous_time_ (feature_tracker.cpp:338)
    for (int s = 0; s < S; ++s) { pdt[s] = times[s] - b->prev_times[s]; b->prev_times[s] = times[s]; }
    for (int k = 0; k < ST_COUNT; ++k) b->st_used[k] = false;
    // All uploads of the frame go through side B, in the order the frame needs them: times, left image, right image.
    // The staging buffers are double-buffered by parity, so with asynchronous enqueues the upload of frame k+1 overlaps
    // the chain of frame k; the main stream only waits for the left image.
    // the images go up as ONE linear copy each into unpadded staging buffers (a pitched 2-D copy of 480 short rows is
    // several times slower over PCIe); the pyramid launches read them there and move level 0 into its padded home
    const size_t SP = (size_t)b->stage_pitch, simg = SP * H;
    // Device images are read by copies on the library's internal streams.  They are ordered behind (a) the event the caller
    // handed in with vf_batch_set_input_event, if any, and (b) when the frame work runs on a caller-supplied stream
    // (vf_config.cuda_stream), everything already queued on that stream -- a debayer / rectify kernel there that produces
    // the images cannot race with the copy.  With a library-owned stream and no event the images must be complete when the
    // call is made (include/vins_fast_b200.h); waiting for the own chain stream here would serialise the upload of frame
    // k+1 behind the chain of frame k.
    if (kind == cudaMemcpyDeviceToDevice) {
        cudaEvent_t in = b->ev_user_input;
        b->ev_user_input = nullptr;
        if (!in && !b->own_main) { VF_CUDA(b, cudaEventRecord(b->ev_input, b->main)); in = b->ev_input; }
        if (in) {
            VF_CUDA(b, cudaStreamWaitEvent(b->side_b, in, 0));
            VF_CUDA(b, cudaStreamWaitEvent(b->side_d, in, 0));
        }
    }
    auto upload = [&](int cam, const uint8_t* const* imgs, const uint8_t* base, size_t stride, cudaStream_t st) -> cudaError_t {
        NvtxRange r(cam ? "vf.upload_right" : "vf.upload_left");
        // images back to back at a fixed distance: ONE copy command for the whole batch (a copy command costs the host ~4 us;
        // per-image commands would dominate a frame of a large batch)
        if (base && S > 1 && stride == SP) {
            uint8_t* dst = b->stage[cam][parity];
            if (image_stride == simg) return cudaMemcpyAsync(dst, base, simg * (size_t)S, kind, st);
            if (image_stride > simg) return cudaMemcpy2DAsync(dst, simg, base, image_stride, simg, (size_t)S, kind, st);
        }
        for (int s = 0; s < S; ++s) {
            const uint8_t* a = base ? base + (size_t)s * image_stride : imgs[s];
            uint8_t* dst = b->stage[cam][parity] + (size_t)s * simg;
            cudaError_t e = (stride == SP) ? cudaMemcpyAsync(dst, a, simg, kind, st)
                                           : cudaMemcpy2DAsync(dst, SP, a, stride, W, H, kind, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    // the left upload is the first command of the frame to reach the GPU: everything on the critical path waits for it
    { StageTimer t(b, ST_H2D_LEFT, b->side_b); VF_CUDA(b, upload(0, img0, base0, stride0, b->side_b)); }
    VF_CUDA(b, cudaEventRecord(b->ev_img, b->side_b));
    if (b->frame >= 2 && !done2) {
        // frame k-2 (same parity) must be complete on the GPU before the chain and the right pyramid overwrite its buffers
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_done[(slot + 2) & 3], 0));
        VF_CUDA(b, cudaStreamWaitEvent(b->side_d, b->ev_done[(slot + 2) & 3], 0));
    }
    // right image on side D; its upload queues behind the left one so that the left image (which the frame's critical path
    // waits for) has the link to itself.  A synchronous frame issues it now (the stereo pass is part of its chain); an
    // overlapped one issues it after the chain -- host order is how soon the chain reaches the GPU.
    auto issue_right = [&]() -> vf_status {
        VF_CUDA(b, cudaStreamWaitEvent(b->side_d, b->ev_img, 0));
        // frame interval(s) -> device, ahead of the right image (both consumers wait for the right pyramid's event)
        VF_CUDA(b, cudaMemcpyAsync(b->dev_dt + (size_t)phase * S, pdt, sizeof(double) * (size_t)S, cudaMemcpyHostToDevice, b->side_d));
        { StageTimer t(b, ST_H2D_RIGHT, b->side_d); VF_CUDA(b, upload(1, img1, base1, stride1, b->side_d)); }
        return VF_OK;
    };
    const bool use_graph = b->cfg.use_cuda_graph && !b->profiling;
    static const char* const kSeqNames[SEQ_COUNT] = {"vf.corner_stage", "vf.chain", "vf.rest_sync", "vf.head_sync", "vf.pyramid_left",
                                                     "vf.pyramid_right", "vf.stereo_tail", "vf.frame_sync"};
    auto run = [&](int which, cudaStream_t st) -> vf_status {   // one of the captured kernel sequences, as a graph or directly
        NvtxRange r(kSeqNames[which]);
        if (use_graph && !(which == SEQ_CHAIN && b->chain_direct)) {
            if (!b->graph[which][phase]) {
                vf_status r = build_graph(b, phase, which);
                if (r != VF_OK) return r;
            }
            VF_CUDA(b, cudaGraphLaunch(b->graph[which][phase], st));
            return VF_OK;
        }
        return issue_sequence(b, which, parity, cur, prev, phase);
    };
    auto issue_right_pyramid = [&]() -> vf_status {
        vf_status st = run(SEQ_PYR_RIGHT, b->side_d);
        if (st != VF_OK) return st;
        VF_CUDA(b, cudaEventRecord(b->ev_right, b->side_d));
        return VF_OK;
    };
    // left pyramid: an overlapped frame builds it on side B right behind its upload (it does not depend on the previous
    // frame, so it is built while that one is still on the chain); a synchronous frame builds it on main, back to back with
    // the temporal LK
    if (overlap) {
        vf_status st = run(SEQ_PYR_LEFT, b->side_b);
        if (st != VF_OK) return st;
        VF_CUDA(b, cudaEventRecord(b->ev_pyrl, b->side_b));
    }
    // corner stage on side A right behind the left upload (the staging image is all it reads)
    auto issue_corner = [&](bool join_left_pyramid) -> vf_status {
        VF_CUDA(b, cudaStreamWaitEvent(b->side_a, b->ev_img, 0));
        vf_status st = run(SEQ_CORNER, b->side_a);
        if (st != VF_OK) return st;
        if (join_left_pyramid) VF_CUDA(b, cudaStreamWaitEvent(b->side_a, b->ev_pyrl, 0));   // ev_scat then stands for both
        VF_CUDA(b, cudaEventRecord(b->ev_scat, b->side_a));
        return VF_OK;
    };
    if (overlap) {
        // Asynchronous enqueue: the corner stage and the left pyramid ran while the previous frame was on the chain, so the
        // chain joins them up front (ONE event: what sits between two chains on the main stream is on the critical path of
        // the frame rate) and is four kernels in a row; the stereo pass gets its own stream and overlaps the next chain.
        { vf_status st = issue_corner(true); if (st != VF_OK) return st; }
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_scat, 0));
        { vf_status st = run(SEQ_CHAIN, b->main); if (st != VF_OK) return st; }
        VF_CUDA(b, cudaEventRecord(b->ev_chain[parity], b->main));
        { vf_status st = issue_right(); if (st != VF_OK) return st; }
        { vf_status st = issue_right_pyramid(); if (st != VF_OK) return st; }
        // stereo pass on side C with the left camera's undistortion + velocity on a branch beside it (side E); both need
        // current_kps_ / ids_ (the chain's event), the stereo pass the right pyramid, both the frame interval that travels with it
        VF_CUDA(b, cudaStreamWaitEvent(b->side_c, b->ev_chain[parity], 0));
        VF_CUDA(b, cudaStreamWaitEvent(b->side_c, b->ev_right, 0));
        { vf_status st = run(SEQ_TAIL, b->side_c); if (st != VF_OK) return st; }
        VF_CUDA(b, cudaGetLastError());
        VF_CUDA(b, cudaEventRecord(b->ev_done[slot], b->side_c));
    } else {
        // Synchronous frame: nothing to overlap with, so the temporal LK starts as soon as the left pyramid stands (the
        // corner stage runs beside it and is joined before select_tracked) and the stereo pass follows the corner
        // selection inside the same graph, the left undistortion on a branch beside it.
        // host order = how soon each piece reaches the GPU: the head of the critical path first (it must be queued before
        // the left upload completes), then the right image (its upload queues behind the left one anyway)
        // the stereo pass / left undistortion read what the previous frame's wrote (velocity): if that frame was an
        // overlapped one they ran on other streams -- unless the host has already seen it complete
        const bool wait_prev = b->frame >= 1 && b->known_done < b->frame - 1;
        if (b->one_graph_sync && use_graph) {
            // ONE graph for the whole frame.  The right upload is issued first (it queues behind the left one on the link) so
            // that its event exists when the graph is launched: the graph's right-pyramid branch waits for it as an external
            // event, everything else starts with the left image.
            { vf_status st = issue_right(); if (st != VF_OK) return st; }
            VF_CUDA(b, cudaEventRecord(b->ev_rimg, b->side_d));
            VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_img, 0));
            if (wait_prev) VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_done[(slot + 3) & 3], 0));
            { vf_status st = run(SEQ_FRAME_SYNC, b->main); if (st != VF_OK) return st; }
        } else {
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_img, 0));
        { vf_status st = run(SEQ_HEAD_SYNC, b->main); if (st != VF_OK) return st; }
        { vf_status st = issue_right(); if (st != VF_OK) return st; }
        { vf_status st = issue_corner(false); if (st != VF_OK) return st; }
        // ONE join in front of the rest of the frame (every wait the main stream meets there sits on the critical path): the
        // right pyramid's stream also waits for the corner stage before it records its event
        { vf_status st = run(SEQ_PYR_RIGHT, b->side_d); if (st != VF_OK) return st; }
        VF_CUDA(b, cudaStreamWaitEvent(b->side_d, b->ev_scat, 0));
        VF_CUDA(b, cudaEventRecord(b->ev_right, b->side_d));
        VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_right, 0));
        if (wait_prev) VF_CUDA(b, cudaStreamWaitEvent(b->main, b->ev_done[(slot + 3) & 3], 0));
        { vf_status st = run(SEQ_REST_SYNC, b->main); if (st != VF_OK) return st; }
        }
        VF_CUDA(b, cudaGetLastError());
        VF_CUDA(b, cudaEventRecord(b->ev_sel, b->main));
        VF_CUDA(b, cudaEventRecord(b->ev_chain[parity], b->main));
        VF_CUDA(b, cudaEventRecord(b->ev_done[slot], b->main));
    }
    int n_launch = 2 * b->pyr_launches;
    n_launch += 4 /* corner stage */ + 4 /* chain */ + 2 /* stereo, left undistortion */;
    // no download commands: the records and the feature counts were written to mapped pinned host memory by the kernels
    b->launches = n_launch;
    b->last_parity = parity;
    b->frame++;
    b->host_issue_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_wait1).count();
    return VF_OK;
}

vf_status fetch_frame(vf_batch* b, vf_feature* out, int32_t* n_out, int lag = 0) {
    NvtxRange nvtx_fetch("vf.fetch_frame");
    if (lag < 0 || lag > 2) return fail(b, VF_ERR_INVALID_ARG, "lag must be 0 (last enqueued frame), 1 or 2 (the ones before)");
    if (b->last_parity < 0 || b->frame <= lag) return fail(b, VF_ERR_STATE, "that frame has not been enqueued");
    DeviceGuard g(b->device);
    const int p = (int)((b->frame - 1 - lag) % 3), S = b->S, cap = b->h.cap;   // the records live in a ring of three frames
    VF_CUDA(b, cudaEventSynchronize(b->ev_done[(int)((b->frame - 1 - lag) & 3)]));
    if (b->frame - 1 - lag > b->known_done) b->known_done = b->frame - 1 - lag;
    if (b->profiling) {
        VF_CUDA(b, sync_all(b));
        for (int k = 0; k < ST_COUNT; ++k) {
            b->stage_ms[k] = 0.f;
            if (b->st_used[k]) cudaEventElapsedTime(&b->stage_ms[k], b->st_beg[k], b->st_end[k]);
        }
    }
    for (int s = 0; s < S; ++s) {
        const int n = b->pin_n[p][s];
        if (n_out) n_out[s] = n;
        if (out) memcpy(out + (size_t)s * cap, b->pin_out[p] + (size_t)s * cap, sizeof(vf_feature) * (size_t)n);
    }
    return VF_OK;
}

// Captures + instantiates + uploads every graph (both plans, all six phases) and replays each once on blank images.
vf_status prebuild_and_warm(vf_batch* b) {
    for (int which = 0; which < SEQ_COUNT; ++which)
        for (int phase = 0; phase < 6; ++phase)
            if (!b->graph[which][phase]) { vf_status st = build_graph(b, phase, which); if (st != VF_OK) return st; }
    const size_t simg = (size_t)b->stage_pitch * b->h.H, bytes = simg * (size_t)b->S;
    if (bytes > ((size_t)1 << 30) || getenv("VF_NO_WARM")) return VF_OK;   // graphs are ready; skip the blank frames for huge batches
    uint8_t* blank = nullptr;
    VF_CUDA(b, cudaMalloc(&blank, bytes));
    cudaError_t e = cudaMemset(blank, 0, bytes);
    std::vector<double> t0(b->S, 0.0);
    vf_status st = e == cudaSuccess ? VF_OK : VF_ERR_CUDA;
    for (int pass = 0; pass < 2 && st == VF_OK; ++pass) {   // overlapped plan, then synchronous plan
        for (int k = 0; k < 6 && st == VF_OK; ++k) {
            st = enqueue_frame(b, t0.data(), nullptr, nullptr, blank, blank, simg, (size_t)b->stage_pitch, (size_t)b->stage_pitch,
                               cudaMemcpyDeviceToDevice, pass == 0);
            if (st == VF_OK && pass == 1) st = fetch_frame(b, nullptr, nullptr);
        }
        if (st == VF_OK) st = fetch_frame(b, nullptr, nullptr);
        if (st == VF_OK) st = reset_state(b);
    }
    sync_all(b);
    cudaFree(blank);
    if (st == VF_OK && e != cudaSuccess) return fail(b, VF_ERR_CUDA, "cudaMemset of the warm-up image failed");
    b->host_wait_us = b->host_issue_us = 0;
    return st;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
extern "C" {

void vf_config_default(vf_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = 752; cfg->height = 480;
    cfg->max_cnt = 200; cfg->min_dist = 15; cfg->flow_back = 1;   // euroc_stero.yaml:44-51
    cfg->lk_max_level = 3; cfg->lk_win = 21; cfg->quality_level = 0.01;   // feature_tracker.cpp:42,90
    for (int k = 0; k < 2; ++k) {
        cfg->cam[k].model = VF_CAM_PINHOLE; cfg->cam[k].image_width = 752; cfg->cam[k].image_height = 480;
        cfg->cam[k].fx = cfg->cam[k].fy = 460.0; cfg->cam[k].cx = 376.0; cfg->cam[k].cy = 240.0;
    }
    cfg->device = -1;
    cfg->use_cuda_graph = 1;
    cfg->cuda_stream = nullptr;
}

int32_t vf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

vf_status vf_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(nullptr, VF_ERR_INVALID_ARG, "out is null");
    VF_CUDA((vf_batch*)nullptr, cudaMallocHost(out, bytes ? bytes : 1));
    return VF_OK;
}
void vf_host_free(void* p) { if (p) cudaFreeHost(p); }

vf_status vf_batch_create(const vf_config* cfg, int32_t n_streams, vf_batch** out) {
    if (!out) return fail(nullptr, VF_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    vf_status st = validate_config(cfg, n_streams);
    if (st != VF_OK) return st;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, VF_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    int dev = cfg->device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, VF_ERR_CUDA, "cudaGetDevice failed"); }
    if (dev >= ndev) return fail(nullptr, VF_ERR_INVALID_ARG, "device ordinal out of range");
    DeviceGuard g(dev);
    if (!g.ok) return fail(nullptr, VF_ERR_CUDA, "cudaSetDevice failed");

    vf_batch* b = new vf_batch();
    b->cfg = *cfg; b->S = n_streams; b->device = dev;
    auto bail = [&](vf_status s) { std::string e = b->err; destroy_batch(b); g_last_error = e; return s; };
#define TRY(x) do { vf_status s__ = (x); if (s__ != VF_OK) return bail(s__); } while (0)
#define TRYC(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { b->err = std::string(#x) + ": " + cudaGetErrorString(e__); return bail(VF_ERR_CUDA); } } while (0)

    // the frame-to-frame chain (main) outranks the stereo pass (side C), which outranks the work that only has to be
    // ready by the time the chain needs it (corner response on side A, right pyramid on side B)
    int prio_lo = 0, prio_hi = 0;
    TRYC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));     // numerically lower = higher priority
    int prio_mid = prio_hi < prio_lo ? prio_hi + 1 : prio_hi;
    if (const char* e = getenv("VF_PRIO")) {   // measurement aid: 0 = one priority for all streams, 2 = stereo as high as main
        if (atoi(e) == 0) { prio_hi = prio_lo; prio_mid = prio_lo; }
        if (atoi(e) == 2) prio_mid = prio_hi;
    }
    if (cfg->cuda_stream) { b->main = (cudaStream_t)cfg->cuda_stream; b->own_main = false; }
    else { TRYC(cudaStreamCreateWithPriority(&b->main, cudaStreamNonBlocking, prio_hi)); b->own_main = true; }
    if (getenv("VF_SERIALIZE")) {   // measurement aid: one stream, no overlap between the stages of a frame
        b->side_a = b->main; b->side_b = b->main; b->side_c = b->main; b->side_d = b->main; b->side_e = b->main; b->serialized = true;
    } else {
        TRYC(cudaStreamCreateWithPriority(&b->side_a, cudaStreamNonBlocking, prio_lo));
        TRYC(cudaStreamCreateWithPriority(&b->side_b, cudaStreamNonBlocking, prio_hi));    // left pyramid: the chain waits for it
        TRYC(cudaStreamCreateWithPriority(&b->side_d, cudaStreamNonBlocking, prio_lo));
        TRYC(cudaStreamCreateWithPriority(&b->side_e, cudaStreamNonBlocking, prio_lo));
        TRYC(cudaStreamCreateWithPriority(&b->side_c, cudaStreamNonBlocking, prio_mid));
    }
    TRYC(cudaEventCreateWithFlags(&b->ev_right, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_input, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_img, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_sel, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_un, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_pyrl, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_scat, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_rimg, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_fork2, cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&b->ev_pyr_sync, cudaEventDisableTiming));
    for (int p = 0; p < 2; ++p) TRYC(cudaEventCreateWithFlags(&b->ev_chain[p], cudaEventDisableTiming));
    for (int p = 0; p < 4; ++p) TRYC(cudaEventCreateWithFlags(&b->ev_done[p], cudaEventDisableTiming));

    DevBatch& h = b->h;
    memset(&h, 0, sizeof(h));
    const int S = n_streams, W = cfg->width, H = cfg->height, cap = cfg->max_cnt;
    h.n_streams = S; h.W = W; h.H = H; h.cap = cap; h.min_dist = cfg->min_dist; h.flow_back = cfg->flow_back ? 1 : 0;
    h.lk_max_level = cfg->lk_max_level; h.lk_win = cfg->lk_win; h.quality = cfg->quality_level;
    h.n_levels = effective_levels(W, H, cfg->lk_max_level, cfg->lk_win);
    h.bucket_bits = ((size_t)W * H <= (1u << 20)) ? 13 : kMaxBucketBits;
    h.n_buckets = 1 << h.bucket_bits;
    h.cand_cap = 1;
    while (h.cand_cap < (size_t)W * H) h.cand_cap <<= 1;
    h.cam[0] = to_cam(cfg->cam[0]); h.cam[1] = to_cam(cfg->cam[1]);
    for (int p = 0; p < 3; ++p) TRY(alloc_pyramid(b, h.left[p], W, H, h.n_levels, S));
    for (int p = 0; p < 2; ++p) TRY(alloc_pyramid(b, h.right[p], W, H, h.n_levels, S));
    b->stage_pitch = align_up(W, 16);
    for (int c = 0; c < 2; ++c)
        for (int p = 0; p < 2; ++p) TRY(dalloc(b, &b->stage[c][p], (size_t)S * b->stage_pitch * H));
    if (getenv("VF_TRACE")) {   // measurement aid: per-kernel start / end stamps (read with vf_batch_trace_get)
        TRY(dalloc(b, &h.trace, (size_t)256));
        unsigned long long init[256];
        for (int k = 0; k < 128; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
        TRYC(cudaMemcpy(h.trace, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    h.stage_left[0] = b->stage[0][0]; h.stage_left[1] = b->stage[0][1]; h.stage_pitch = b->stage_pitch; h.stage_stride = (size_t)b->stage_pitch * H;
    const size_t SC = (size_t)S * cap, SP = (size_t)S * W * H;
    for (int p = 0; p < 2; ++p) {
        TRY(dalloc(b, &h.kps[p], SC)); TRY(dalloc(b, &h.ids[p], SC)); TRY(dalloc(b, &h.cnt[p], SC)); TRY(dalloc(b, &h.prev_idx[p], SC));
        TRY(dalloc(b, &h.un_left[p], SC)); TRY(dalloc(b, &h.un_right[p], SC)); TRY(dalloc(b, &h.right_ok[p], SC));
        TRY(dalloc(b, &h.counters[p], (size_t)S * C_COUNT)); TRY(dalloc(b, &h.times[p], (size_t)S));
    }
    TRY(dalloc(b, &h.fwd_pts, SC)); TRY(dalloc(b, &h.fwd_st, SC)); TRY(dalloc(b, &h.rev_pts, SC)); TRY(dalloc(b, &h.rev_st, SC));
    TRY(dalloc(b, &h.temporal_st, SC)); TRY(dalloc(b, &h.lr_pts, SC)); TRY(dalloc(b, &h.lr_st, SC)); TRY(dalloc(b, &h.rl_pts, SC));
    TRY(dalloc(b, &h.rl_st, SC)); TRY(dalloc(b, &h.sort_perm, SC)); TRY(dalloc(b, &h.lk_iters, SC)); TRY(dalloc(b, &h.lk_iters_stereo, SC));
    TRY(dalloc(b, &h.mask, SP));
    TRY(dalloc(b, &h.surv, (size_t)S * h.cand_cap));
    for (int p = 0; p < 2; ++p) {   // corner stage, double-buffered: frame k+1's response runs while frame k is on the chain
        TRY(dalloc(b, &h.eig[p], SP)); TRY(dalloc(b, &h.lmflag[p], SP));
        TRY(dalloc(b, &h.cand[p], (size_t)S * h.cand_cap)); TRY(dalloc(b, &h.cand_sorted[p], (size_t)S * h.cand_cap));
        // every bucket has kHistRep counters / cursors / start offsets (vf_common.cuh)
        TRY(dalloc(b, &h.hist[p], (size_t)S * h.n_buckets * kHistRep)); TRY(dalloc(b, &h.bucket_start[p], (size_t)S * (h.n_buckets * kHistRep + 8)));
        TRY(dalloc(b, &h.bucket_fill[p], (size_t)S * h.n_buckets * kHistRep)); TRY(dalloc(b, &h.bucket_nz[p], (size_t)S * h.n_buckets));
        TRY(dalloc(b, &h.ccnt[p], (size_t)S * C_COUNT));
    }
    // result records, feature counts and frame intervals live in mapped pinned host memory: the kernels write / read them
    // in place, no copy commands at either end of a frame
    for (int p = 0; p < 3; ++p) {
        TRYC(cudaHostAlloc(&b->pin_out[p], sizeof(vf_feature) * SC, cudaHostAllocMapped));
        memset(b->pin_out[p], 0, sizeof(vf_feature) * SC);
        TRYC(cudaHostGetDevicePointer((void**)&h.out[p], b->pin_out[p], 0));
        TRYC(cudaHostAlloc(&b->pin_n[p], sizeof(int) * (size_t)S, cudaHostAllocMapped));
        memset(b->pin_n[p], 0, sizeof(int) * (size_t)S);
        TRYC(cudaHostGetDevicePointer((void**)&b->dev_n[p], b->pin_n[p], 0));
    }
    TRYC(cudaHostAlloc(&b->pin_dt, sizeof(double) * (size_t)S * 6, cudaHostAllocMapped));
    TRY(dalloc(b, &b->dev_dt, (size_t)S * 6));   // device copy: a read of host memory at the end of every LK CTA costs a PCIe round trip
    b->prev_times.assign(S, 0.0);
    TRYC(cudaMalloc(&b->d, sizeof(DevBatch)));
    TRYC(cudaMemcpy(b->d, &h, sizeof(DevBatch), cudaMemcpyHostToDevice));

    // synchronous frames as one graph (VF_SYNC_ONE_GRAPH=0: the four-graph plan, for A/B runs)
    b->one_graph_sync = !(getenv("VF_SYNC_ONE_GRAPH") != nullptr && atoi(getenv("VF_SYNC_ONE_GRAPH")) == 0) && !b->serialized;
    b->corner_after_pyr = !(getenv("VF_CORNER_AFTER_PYR") != nullptr && atoi(getenv("VF_CORNER_AFTER_PYR")) == 0);   // =0: beside the pyramid (A/B)
    // A graph that follows another graph in the same stream starts 5-9 us after it ends; a kernel that follows a kernel ~1 us.
    // The chain is the one sequence whose predecessor in its stream (the previous frame's chain) sits on the critical path.
    b->chain_direct = getenv("VF_CHAIN_DIRECT") != nullptr && atoi(getenv("VF_CHAIN_DIRECT")) != 0;
    TRYC(select_tracked_prepare(cap));
    TRYC(lk_prepare());
    TRYC(corners_prepare());
    TRYC(pyramid_prepare());
    TRYC(cudaDeviceSynchronize());
    b->views.resize(S);
    for (int s = 0; s < S; ++s) { b->views[s].batch = b; b->views[s].stream_index = s; b->views[s].owns_batch = false; }
    // The first frame must cost what the thousandth does (the reference's front-end has no warm-up cliff,
    // vins/estimator/estimator.cpp:143-158 just calls TrackImage): every graph of both plans is captured, instantiated and
    // uploaded here, then each one is replayed once on blank images (loads the kernels, sizes the local-memory pool, walks
    // the launch path) and the tracker state is put back to "no frame seen yet".
    if (cfg->use_cuda_graph) TRY(prebuild_and_warm(b));
#undef TRY
#undef TRYC
    *out = b;
    return VF_OK;
}

void vf_batch_destroy(vf_batch* b) { destroy_batch(b); }

// Per-stage device times of the last fetched frame (CUDA events on the launching streams).  Profiling mode issues
// plain launches (no graph replay) so that events can sit between the kernels; it is a measurement aid, off by default.
vf_status vf_batch_set_profiling(vf_batch* b, int32_t enable) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    DeviceGuard g(b->device);
    if (enable && !b->st_beg[0])
        for (int k = 0; k < ST_COUNT; ++k) { VF_CUDA(b, cudaEventCreate(&b->st_beg[k])); VF_CUDA(b, cudaEventCreate(&b->st_end[k])); }
    VF_CUDA(b, sync_all(b));
    b->profiling = enable != 0;
    return VF_OK;
}
int32_t vf_batch_stage_count(void) { return ST_COUNT; }
const char* vf_batch_stage_name(int32_t i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
vf_status vf_batch_stage_times(vf_batch* b, float* ms_out, int32_t cap) {
    if (!b || !ms_out || cap < ST_COUNT) return fail(b, VF_ERR_INVALID_ARG, "bad argument");
    for (int k = 0; k < ST_COUNT; ++k) ms_out[k] = b->stage_ms[k];
    return VF_OK;
}
// Measurement aid (VF_TRACE=1 in the environment at creation): copies the [TR_COUNT][2] start / end stamps (ns) of the kernels
// launched since the last call and re-arms the buffer.  Returns the number of kernels, 0 when tracing is off.
int32_t vf_batch_trace_get(vf_batch* b, unsigned long long* out, int32_t cap) {
    if (!b || !b->h.trace || !out || cap < 256) return 0;
    DeviceGuard g(b->device);
    if (sync_all(b) != cudaSuccess) return 0;
    if (cudaMemcpy(out, b->h.trace, sizeof(unsigned long long) * 256, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    unsigned long long init[256];
    for (int k = 0; k < 128; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
    cudaMemcpy(b->h.trace, init, sizeof(init), cudaMemcpyHostToDevice);
    return TR_COUNT;
}

vf_status vf_batch_reset(vf_batch* b) { return b ? reset_state(b) : fail(nullptr, VF_ERR_INVALID_ARG, "batch is null"); }
int32_t vf_batch_size(const vf_batch* b) { return b ? b->S : 0; }
const char* vf_batch_last_error(const vf_batch* b) { return b ? b->err.c_str() : g_last_error.c_str(); }
void* vf_batch_cuda_stream(const vf_batch* b) { return b ? (void*)b->main : nullptr; }
int32_t vf_batch_launches_per_frame(const vf_batch* b) { return b ? b->launches : 0; }
// Cumulative host microseconds spent inside the enqueue calls: [0] waiting for the frame two back, [1] issuing the frame.
void vf_batch_host_times(const vf_batch* b, double* out2) { if (b && out2) { out2[0] = b->host_wait_us; out2[1] = b->host_issue_us; } }
vf_tracker* vf_batch_stream(vf_batch* b, int32_t s) { return (b && s >= 0 && s < b->S) ? &b->views[s] : nullptr; }

vf_status vf_batch_track(vf_batch* b, const double* times, const uint8_t* const* img0, size_t stride0,
                         const uint8_t* const* img1, size_t stride1, vf_feature* out, int32_t* n_out) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    if (!img0 || !img1) return fail(b, VF_ERR_INVALID_ARG, "img0/img1 must both be non-null (stereo is mandatory)");
    vf_status st = enqueue_frame(b, times, img0, img1, nullptr, nullptr, 0, stride0, stride1, cudaMemcpyHostToDevice, false);
    if (st != VF_OK) return st;
    return fetch_frame(b, out, n_out);
}

static vf_status enqueue_device(vf_batch* b, const double* times, const uint8_t* d_img0, const uint8_t* d_img1,
                                size_t row_stride, size_t image_stride_bytes, bool overlap) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    if (!d_img0 || !d_img1) return fail(b, VF_ERR_INVALID_ARG, "img0/img1 must both be non-null (stereo is mandatory)");
    if (b->S > 1 && image_stride_bytes < row_stride * (size_t)b->h.H)
        return fail(b, VF_ERR_INVALID_ARG, "image_stride_bytes is smaller than one image");
    return enqueue_frame(b, times, nullptr, nullptr, d_img0, d_img1, image_stride_bytes, row_stride, row_stride,
                         cudaMemcpyDeviceToDevice, overlap);
}

vf_status vf_batch_enqueue_device(vf_batch* b, const double* times, const uint8_t* d_img0, const uint8_t* d_img1,
                                  size_t row_stride, size_t image_stride_bytes) {
    return enqueue_device(b, times, d_img0, d_img1, row_stride, image_stride_bytes, true);
}

vf_status vf_batch_set_input_event(vf_batch* b, void* cuda_event) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    b->ev_user_input = (cudaEvent_t)cuda_event;
    return VF_OK;
}

vf_status vf_batch_fetch(vf_batch* b, vf_feature* out, int32_t* n_out) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    return fetch_frame(b, out, n_out);
}

vf_status vf_batch_fetch_lagged(vf_batch* b, int32_t lag, vf_feature* out, int32_t* n_out) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    return fetch_frame(b, out, n_out, lag);
}

vf_status vf_batch_enqueue(vf_batch* b, const double* times, const uint8_t* const* img0, size_t stride0,
                           const uint8_t* const* img1, size_t stride1) {
    if (!b) return fail(nullptr, VF_ERR_INVALID_ARG, "batch is null");
    if (!img0 || !img1) return fail(b, VF_ERR_INVALID_ARG, "img0/img1 must both be non-null (stereo is mandatory)");
    return enqueue_frame(b, times, img0, img1, nullptr, nullptr, 0, stride0, stride1, cudaMemcpyHostToDevice, true);
}

vf_status vf_batch_track_device(vf_batch* b, const double* times, const uint8_t* d_img0, const uint8_t* d_img1,
                                size_t row_stride, size_t image_stride_bytes, vf_feature* out, int32_t* n_out) {
    vf_status st = enqueue_device(b, times, d_img0, d_img1, row_stride, image_stride_bytes, false);
    if (st != VF_OK) return st;
    return fetch_frame(b, out, n_out);
}

// ---- single stream = batch of one ----
vf_status vf_tracker_create(const vf_config* cfg, vf_tracker** out) {
    if (!out) return fail(nullptr, VF_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    vf_batch* b = nullptr;
    vf_status st = vf_batch_create(cfg, 1, &b);
    if (st != VF_OK) return st;
    vf_tracker* t = new vf_tracker();
    t->batch = b; t->stream_index = 0; t->owns_batch = true;
    *out = t;
    return VF_OK;
}

void vf_tracker_destroy(vf_tracker* t) {
    if (!t) return;
    if (t->owns_batch) { destroy_batch(t->batch); delete t; }
}

vf_status vf_tracker_reset(vf_tracker* t) {
    if (!t) return fail(nullptr, VF_ERR_INVALID_ARG, "tracker is null");
    return reset_state(t->batch);
}

const char* vf_last_error(const vf_tracker* t) { return t ? t->batch->err.c_str() : g_last_error.c_str(); }

vf_status vf_tracker_track(vf_tracker* t, double time_s, const uint8_t* img0, size_t stride0, const uint8_t* img1,
                           size_t stride1, vf_feature* out, int32_t* n_out) {
    if (!t) return fail(nullptr, VF_ERR_INVALID_ARG, "tracker is null");
    if (t->batch->S != 1) return fail(t->batch, VF_ERR_INVALID_ARG, "use vf_batch_track on a multi-stream batch");
    if (!out || !n_out) return fail(t->batch, VF_ERR_INVALID_ARG, "out / n_out is null");
    const uint8_t* a[1] = {img0};
    const uint8_t* c[1] = {img1};
    return vf_batch_track(t->batch, &time_s, a, stride0, c, stride1, out, n_out);
}

vf_status vf_tracker_track_device(vf_tracker* t, double time_s, const uint8_t* d_img0, size_t stride0, const uint8_t* d_img1,
                                  size_t stride1, vf_feature* out, int32_t* n_out) {
    if (!t) return fail(nullptr, VF_ERR_INVALID_ARG, "tracker is null");
    if (t->batch->S != 1) return fail(t->batch, VF_ERR_INVALID_ARG, "use vf_batch_track_device on a multi-stream batch");
    if (!out || !n_out) return fail(t->batch, VF_ERR_INVALID_ARG, "out / n_out is null");
    if (stride0 != stride1) return fail(t->batch, VF_ERR_INVALID_ARG, "device images must share one row stride");
    return vf_batch_track_device(t->batch, &time_s, d_img0, d_img1, stride0, stride0 * (size_t)t->batch->h.H, out, n_out);
}

// ---- debug getters ----
vf_status vf_tracker_debug_get(vf_tracker* t, int32_t what, int32_t arg, void* buf, size_t cap_bytes, size_t* size_out) {
    if (!t || !size_out) return fail(nullptr, VF_ERR_INVALID_ARG, "tracker / size_out is null");
    vf_batch* b = t->batch;
    if (b->last_parity < 0) return fail(b, VF_ERR_STATE, "debug getters are valid after the first frame");
    DeviceGuard g(b->device);
    VF_CUDA(b, sync_all(b));
    const DevBatch& h = b->h;
    const int s = t->stream_index, p = b->last_parity, cap = h.cap;
    int cnt_now[C_COUNT];
    VF_CUDA(b, cudaMemcpy(cnt_now, h.counters[p] + s * C_COUNT, sizeof(cnt_now), cudaMemcpyDeviceToHost));
    VF_CUDA(b, cudaMemcpy(&cnt_now[C_N_LOCALMAX], h.ccnt[p] + s * C_COUNT + C_N_LOCALMAX, sizeof(int), cudaMemcpyDeviceToHost));
    const size_t so = (size_t)s * cap;
    const void* src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case VF_DBG_PYR_LEFT:
        case VF_DBG_PYR_RIGHT: {
            const PyrView& pv = (what == VF_DBG_PYR_LEFT) ? h.left[(int)((b->frame - 1) % 3)] : h.right[p];
            if (arg < 0 || arg >= pv.n_levels) return fail(b, VF_ERR_INVALID_ARG, "pyramid level out of range");
            bytes = (size_t)pv.w[arg] * pv.h[arg];
            *size_out = bytes;
            if (!buf || cap_bytes < bytes) return fail(b, VF_ERR_INVALID_ARG, "buffer too small");
            VF_CUDA(b, cudaMemcpy2D(buf, pv.w[arg], pv.lvl[arg] + (size_t)s * h.lvl_stream_stride[arg], pv.pitch[arg], pv.w[arg],
                                    pv.h[arg], cudaMemcpyDeviceToHost));
            return VF_OK;
        }
        case VF_DBG_EIG: src = h.eig[p] + (size_t)s * h.W * h.H; bytes = sizeof(float) * (size_t)h.W * h.H; break;
        case VF_DBG_MASK: src = h.mask + (size_t)s * h.W * h.H; bytes = (size_t)h.W * h.H; break;
        case VF_DBG_FWD_PTS: src = h.fwd_pts + so; bytes = sizeof(float2) * (size_t)cnt_now[C_N_PREV]; break;
        case VF_DBG_FWD_STATUS: src = h.fwd_st + so; bytes = (size_t)cnt_now[C_N_PREV]; break;
        case VF_DBG_REV_PTS: src = h.rev_pts + so; bytes = sizeof(float2) * (size_t)cnt_now[C_N_PREV]; break;
        case VF_DBG_REV_STATUS: src = h.rev_st + so; bytes = (size_t)cnt_now[C_N_PREV]; break;
        case VF_DBG_TEMPORAL_STATUS: src = h.temporal_st + so; bytes = (size_t)cnt_now[C_N_PREV]; break;
        case VF_DBG_SORT_PERM: src = h.sort_perm + so; bytes = sizeof(int) * (size_t)cnt_now[C_N_TRACKED]; break;
        case VF_DBG_LR_PTS: src = h.lr_pts + so; bytes = sizeof(float2) * (size_t)cnt_now[C_N_TOTAL]; break;
        case VF_DBG_LR_STATUS: src = h.lr_st + so; bytes = (size_t)cnt_now[C_N_TOTAL]; break;
        case VF_DBG_RL_PTS: src = h.rl_pts + so; bytes = sizeof(float2) * (size_t)cnt_now[C_N_TOTAL]; break;
        case VF_DBG_RL_STATUS: src = h.rl_st + so; bytes = (size_t)cnt_now[C_N_TOTAL]; break;
        case VF_DBG_LK_ITERS: src = (arg == 0 ? h.lk_iters : h.lk_iters_stereo) + so;
            bytes = sizeof(int) * (size_t)cnt_now[arg == 0 ? C_N_PREV : C_N_TOTAL]; break;
        case VF_DBG_COUNTS: {
            int v[8] = {cnt_now[C_N_PREV], cnt_now[C_N_TRACKED], cnt_now[C_N_KEPT], cnt_now[C_N_NEW],
                        cnt_now[C_N_TOTAL], cnt_now[C_N_RIGHT], cnt_now[C_N_LOCALMAX], cnt_now[C_LANDMARK_ID]};
            *size_out = sizeof(v);
            if (!buf || cap_bytes < sizeof(v)) return fail(b, VF_ERR_INVALID_ARG, "buffer too small");
            memcpy(buf, v, sizeof(v));
            return VF_OK;
        }
        default: return fail(b, VF_ERR_INVALID_ARG, "unknown debug item");
    }
    *size_out = bytes;
    if (bytes == 0) return VF_OK;
    if (!buf || cap_bytes < bytes) return fail(b, VF_ERR_INVALID_ARG, "buffer too small");
    VF_CUDA(b, cudaMemcpy(buf, src, bytes, cudaMemcpyDeviceToHost));
    return VF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Stand-alone stage operators on host buffers.  Each mirrors one OpenCV / camodocal routine the reference calls,
// runs the same kernels the tracker uses, and exists so the parity tests can pin every stage separately.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct TempPyr {
    PyrView pv;
    std::vector<void*> allocs;
    ~TempPyr() { for (void* p : allocs) cudaFree(p); }
};

vf_status temp_pyramid(TempPyr& tp, const uint8_t* img, int w, int h, int max_level, cudaStream_t st, int win = kWin) {
    memset(&tp.pv, 0, sizeof(tp.pv));
    const int n = effective_levels(w, h, max_level, win);
    tp.pv.n_levels = n;
    int lw = w, lh = h;
    for (int l = 0; l < n; ++l) {
        tp.pv.w[l] = lw; tp.pv.h[l] = lh; tp.pv.pitch[l] = align_up(lw + 2 * kPyrPad, 16);
        void* p = nullptr;
        VF_CUDA((vf_batch*)nullptr, cudaMalloc(&p, (size_t)tp.pv.pitch[l] * (lh + 2 * kPyrPad)));
        tp.allocs.push_back(p);
        tp.pv.lvl[l] = (uint8_t*)p + (size_t)kPyrPad * tp.pv.pitch[l] + kPyrPad;
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy2DAsync(tp.pv.lvl[0], tp.pv.pitch[0], img, w, w, h, cudaMemcpyHostToDevice, st));
    launch_pyramid(tp.pv, nullptr, 1, st);
    VF_CUDA((vf_batch*)nullptr, cudaGetLastError());
    return VF_OK;
}

vf_status op_device(int device, DeviceGuard** g) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, VF_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device >= ndev) return fail(nullptr, VF_ERR_INVALID_ARG, "device ordinal out of range");
    *g = new DeviceGuard(device);
    return VF_OK;
}

struct GuardHolder {
    DeviceGuard* g = nullptr;
    ~GuardHolder() { delete g; }
};

vf_status temp_batch(int w, int h, int cap, int min_dist, double quality, int device, vf_batch** out) {
    vf_config cfg;
    vf_config_default(&cfg);
    cfg.width = w; cfg.height = h; cfg.max_cnt = cap; cfg.min_dist = min_dist; cfg.quality_level = quality;
    cfg.device = device; cfg.use_cuda_graph = 0;
    return vf_batch_create(&cfg, 1, out);
}

struct BatchHolder {
    vf_batch* b = nullptr;
    ~BatchHolder() { destroy_batch(b); }
};

}  // namespace

extern "C" {

vf_status vf_op_build_pyramid(const uint8_t* img, int32_t w, int32_t h, int32_t max_level, int32_t device,
                              uint8_t* levels_out, int32_t* ws, int32_t* hs, int32_t* n_levels) {
    if (!img || !levels_out || !ws || !hs || !n_levels) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    if (w < 1 || h < 1 || max_level < 0 || max_level >= kMaxLevels) return fail(nullptr, VF_ERR_INVALID_ARG, "bad size / level");
    GuardHolder gh;
    vf_status st = op_device(device, &gh.g);
    if (st != VF_OK) return st;
    TempPyr tp;
    st = temp_pyramid(tp, img, w, h, max_level, 0);
    if (st != VF_OK) return st;
    size_t off = 0;
    for (int l = 0; l < tp.pv.n_levels; ++l) {
        VF_CUDA((vf_batch*)nullptr, cudaMemcpy2D(levels_out + off, tp.pv.w[l], tp.pv.lvl[l], tp.pv.pitch[l], tp.pv.w[l],
                                                tp.pv.h[l], cudaMemcpyDeviceToHost));
        ws[l] = tp.pv.w[l]; hs[l] = tp.pv.h[l];
        off += (size_t)tp.pv.w[l] * tp.pv.h[l];
    }
    *n_levels = tp.pv.n_levels;
    return VF_OK;
}

vf_status vf_op_scharr(const uint8_t* img, int32_t w, int32_t h, int32_t device, int16_t* out) {
    if (!img || !out) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    if (w < 1 || h < 1) return fail(nullptr, VF_ERR_INVALID_ARG, "bad size");
    GuardHolder gh;
    vf_status st = op_device(device, &gh.g);
    if (st != VF_OK) return st;
    TempPyr tp;
    st = temp_pyramid(tp, img, w, h, 0, 0);
    if (st != VF_OK) return st;
    int16_t* d = nullptr;
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&d, sizeof(int16_t) * 2 * (size_t)w * h));
    tp.allocs.push_back(d);
    launch_scharr(tp.pv.lvl[0], w, h, tp.pv.pitch[0], d, 0);
    VF_CUDA((vf_batch*)nullptr, cudaGetLastError());
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(out, d, sizeof(int16_t) * 2 * (size_t)w * h, cudaMemcpyDeviceToHost));
    return VF_OK;
}

vf_status vf_op_lk_win(const uint8_t* prev, const uint8_t* next, int32_t w, int32_t h, const float* prev_pts, float* next_pts,
                       uint8_t* status, int32_t n, int32_t max_level, int32_t win, int32_t use_initial_flow, int32_t device,
                       int32_t* iters_out) {
    if (!prev || !next || (n > 0 && (!prev_pts || !next_pts || !status))) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    if (w < 1 || h < 1 || n < 0 || max_level < 0 || max_level >= kMaxLevels) return fail(nullptr, VF_ERR_INVALID_ARG, "bad size / level");
    if (!lk_window_supported(win)) return fail(nullptr, VF_ERR_INVALID_ARG, "win must be odd and in [5, 31]");
    if (win != kWin && (w <= win || h <= win)) return fail(nullptr, VF_ERR_INVALID_ARG, "the image must be larger than the window");
    if (n == 0) return VF_OK;   // calcOpticalFlowPyrLK with no points returns empty outputs
    GuardHolder gh;
    vf_status st = op_device(device, &gh.g);
    if (st != VF_OK) return st;
    TempPyr a, c;
    st = temp_pyramid(a, prev, w, h, max_level, 0, win);
    if (st != VF_OK) return st;
    st = temp_pyramid(c, next, w, h, max_level, 0, win);
    if (st != VF_OK) return st;
    float2 *dp = nullptr, *dn = nullptr;
    uint8_t* ds = nullptr;
    int* di = nullptr;
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&dp, sizeof(float2) * n)); a.allocs.push_back(dp);
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&dn, sizeof(float2) * n)); a.allocs.push_back(dn);
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&ds, n)); a.allocs.push_back(ds);
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&di, sizeof(int) * n)); a.allocs.push_back(di);
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(dp, prev_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(dn, next_pts, sizeof(float2) * n, cudaMemcpyHostToDevice));
    VF_CUDA((vf_batch*)nullptr, lk_prepare());
    VF_CUDA((vf_batch*)nullptr, pyramid_prepare());
    if (win == kWin) launch_lk_op(a.pv, c.pv, dp, dn, ds, di, n, max_level, use_initial_flow, 0);
    else launch_lk_op_generic(a.pv, c.pv, dp, dn, ds, di, n, max_level, use_initial_flow, win, 0);
    VF_CUDA((vf_batch*)nullptr, cudaGetLastError());
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(next_pts, dn, sizeof(float2) * n, cudaMemcpyDeviceToHost));
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(status, ds, n, cudaMemcpyDeviceToHost));
    if (iters_out) VF_CUDA((vf_batch*)nullptr, cudaMemcpy(iters_out, di, sizeof(int) * n, cudaMemcpyDeviceToHost));
    return VF_OK;
}

vf_status vf_op_lk(const uint8_t* prev, const uint8_t* next, int32_t w, int32_t h, const float* prev_pts, float* next_pts,
                   uint8_t* status, int32_t n, int32_t max_level, int32_t use_initial_flow, int32_t device, int32_t* iters_out) {
    return vf_op_lk_win(prev, next, w, h, prev_pts, next_pts, status, n, max_level, kWin, use_initial_flow, device, iters_out);
}

vf_status vf_debug_set_lk_kernel(int32_t mode) {
    if (mode < 0 || mode > 3) return fail(nullptr, VF_ERR_INVALID_ARG, "LK kernel mode must be 0 .. 3");
    set_lk_kernel_mode(mode);
    return VF_OK;
}

vf_status vf_debug_set_pyramid_kernel(int32_t mode) {
    if (mode < 0 || mode > 2) return fail(nullptr, VF_ERR_INVALID_ARG, "pyramid kernel mode must be 0, 1 or 2");
    set_pyramid_kernel_mode(mode);
    return VF_OK;
}

vf_status vf_debug_set_response_kernel(int32_t mode) {
    if (mode < 0 || mode > 2) return fail(nullptr, VF_ERR_INVALID_ARG, "response kernel mode must be 0, 1 or 2");
    set_response_kernel_mode(mode);
    return VF_OK;
}

vf_status vf_op_min_eigen(const uint8_t* img, int32_t w, int32_t h, int32_t device, float* out) {
    if (!img || !out) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    BatchHolder bh;
    vf_status st = temp_batch(w, h, 1, 1, 0.01, device, &bh.b);
    if (st != VF_OK) return st;
    vf_batch* b = bh.b;
    DeviceGuard g(b->device);
    VF_CUDA(b, cudaMemcpy2DAsync(b->stage[0][0], b->stage_pitch, img, w, w, h, cudaMemcpyHostToDevice, b->main));
    launch_clear_frame(b->d, b->h, 0, b->main);
    launch_response_nms(b->d, b->h, 0, b->main);
    VF_CUDA(b, cudaGetLastError());
    VF_CUDA(b, cudaMemcpyAsync(out, b->h.eig[0], sizeof(float) * (size_t)w * h, cudaMemcpyDeviceToHost, b->main));
    VF_CUDA(b, cudaStreamSynchronize(b->main));
    return VF_OK;
}

vf_status vf_op_good_features(const uint8_t* img, int32_t w, int32_t h, int32_t max_corners, double quality, int32_t min_dist,
                              const uint8_t* mask, int32_t device, float* out_xy, int32_t* n_out) {
    if (!img || !out_xy || !n_out) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    if (max_corners < 1) return fail(nullptr, VF_ERR_INVALID_ARG, "max_corners must be >= 1");
    BatchHolder bh;
    vf_status st = temp_batch(w, h, max_corners, min_dist, quality, device, &bh.b);
    if (st != VF_OK) return st;
    vf_batch* b = bh.b;
    DeviceGuard g(b->device);
    const size_t npx = (size_t)w * h;
    VF_CUDA(b, cudaMemcpy2DAsync(b->stage[0][0], b->stage_pitch, img, w, w, h, cudaMemcpyHostToDevice, b->main));
    launch_clear_frame(b->d, b->h, 0, b->main);
    launch_response_nms(b->d, b->h, 0, b->main);
    launch_bucket_scan(b->d, b->h, 0, b->main);
    launch_bucket_scatter(b->d, b->h, 0, b->main);
    if (mask) VF_CUDA(b, cudaMemcpyAsync(b->h.mask, mask, npx, cudaMemcpyHostToDevice, b->main));
    else VF_CUDA(b, cudaMemsetAsync(b->h.mask, 255, npx, b->main));
    VF_CUDA(b, cudaMemsetAsync(b->h.counters[0], 0, sizeof(int) * C_COUNT, b->main));   // what select_tracked does in a frame
    launch_masked_max_from_mask(b->d, b->h, 0, b->main);
    launch_gftt_select(b->d, b->h, 0, b->main);   // C_N_KEPT == 0 -> tops up to max_corners
    VF_CUDA(b, cudaGetLastError());
    int counters[C_COUNT];
    VF_CUDA(b, cudaMemcpyAsync(counters, b->h.counters[0], sizeof(counters), cudaMemcpyDeviceToHost, b->main));
    VF_CUDA(b, cudaStreamSynchronize(b->main));
    const int n = counters[C_N_NEW];
    if (n > 0) VF_CUDA(b, cudaMemcpy(out_xy, b->h.kps[0], sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost));
    *n_out = n;
    return VF_OK;
}

vf_status vf_op_set_mask(const float* pts, const int32_t* track_cnt, int32_t n, int32_t w, int32_t h, int32_t min_dist,
                         int32_t device, int32_t* perm_out, int32_t* keep_out, int32_t* n_keep, uint8_t* mask_out) {
    if (n < 0 || (n > 0 && (!pts || !track_cnt)) || !n_keep) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    for (int i = 0; i < n; ++i) {   // mask.at<uchar>(pt) of the reference needs the rounded point inside the image
        const long x = lrintf(pts[2 * i]), y = lrintf(pts[2 * i + 1]);
        if (x < 0 || x >= w || y < 0 || y >= h) return fail(nullptr, VF_ERR_INVALID_ARG, "point outside the image");
    }
    BatchHolder bh;
    vf_status st = temp_batch(w, h, n > 0 ? n : 1, min_dist, 0.01, device, &bh.b);
    if (st != VF_OK) return st;
    vf_batch* b = bh.b;
    DeviceGuard g(b->device);
    const DevBatch& hb = b->h;
    // frame parity 0 consumes the "previous" state of parity 1
    std::vector<int> ids(n > 0 ? n : 1), cnt(n > 0 ? n : 1);
    std::vector<uint8_t> ones(n > 0 ? n : 1, 1);
    for (int i = 0; i < n; ++i) { ids[i] = i; cnt[i] = track_cnt[i] - 1; }   // the kernel applies ++track_count
    int counters[C_COUNT] = {0};
    counters[C_N_TOTAL] = n;
    VF_CUDA(b, cudaMemcpy(hb.counters[1], counters, sizeof(counters), cudaMemcpyHostToDevice));
    if (n > 0) {
        VF_CUDA(b, cudaMemcpy(hb.fwd_pts, pts, sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice));
        VF_CUDA(b, cudaMemcpy(hb.temporal_st, ones.data(), (size_t)n, cudaMemcpyHostToDevice));
        VF_CUDA(b, cudaMemcpy(hb.ids[1], ids.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
        VF_CUDA(b, cudaMemcpy(hb.cnt[1], cnt.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    }
    launch_clear_frame(b->d, hb, 0, b->main);
    launch_select_tracked(b->d, hb, 0, b->main);
    launch_mask_and_max(b->d, hb, 0, b->main);
    VF_CUDA(b, cudaGetLastError());
    VF_CUDA(b, cudaStreamSynchronize(b->main));
    VF_CUDA(b, cudaMemcpy(counters, hb.counters[0], sizeof(counters), cudaMemcpyDeviceToHost));
    const int nk = counters[C_N_KEPT];
    *n_keep = nk;
    if (perm_out && n > 0) VF_CUDA(b, cudaMemcpy(perm_out, hb.sort_perm, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    if (keep_out && nk > 0) VF_CUDA(b, cudaMemcpy(keep_out, hb.ids[0], sizeof(int) * (size_t)nk, cudaMemcpyDeviceToHost));
    if (mask_out) VF_CUDA(b, cudaMemcpy(mask_out, hb.mask, (size_t)w * h, cudaMemcpyDeviceToHost));
    return VF_OK;
}

vf_status vf_op_undistort(const vf_camera* cam, const float* uv, int32_t n, int32_t device, float* out_xy) {
    if (!cam || n < 0 || (n > 0 && (!uv || !out_xy))) return fail(nullptr, VF_ERR_INVALID_ARG, "null argument");
    if (cam->model != VF_CAM_MEI && cam->model != VF_CAM_PINHOLE) return fail(nullptr, VF_ERR_INVALID_ARG, "unsupported camera model");
    if (n == 0) return VF_OK;
    GuardHolder gh;
    vf_status st = op_device(device, &gh.g);
    if (st != VF_OK) return st;
    TempPyr tp;   // only used as an allocation holder
    float2 *d_in = nullptr, *d_out = nullptr;
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&d_in, sizeof(float2) * n)); tp.allocs.push_back(d_in);
    VF_CUDA((vf_batch*)nullptr, cudaMalloc(&d_out, sizeof(float2) * n)); tp.allocs.push_back(d_out);
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(d_in, uv, sizeof(float2) * n, cudaMemcpyHostToDevice));
    launch_undistort(to_cam(*cam), d_in, d_out, n, 0);
    VF_CUDA((vf_batch*)nullptr, cudaGetLastError());
    VF_CUDA((vf_batch*)nullptr, cudaMemcpy(out_xy, d_out, sizeof(float2) * n, cudaMemcpyDeviceToHost));
    return VF_OK;
}

}  // extern "C"
