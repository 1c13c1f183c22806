/* vins_fast_b200 -- C ABI of the B200-native stereo visual front-end.
 *
 * This is the drop-in boundary for ONE hot path of weihaoysgs/vins-fast:
 *     estimator::FeatureTracker::TrackImage(t, img0, img1)      vins/estimator/feature_tracker.cpp:27-217
 * and the pieces it is made of.  Every entry point below cites the reference interface it replaces.
 * Plain pointers and sizes only (no torch / OpenCV / Eigen types).  The C++ class with the
 * reference's API (include/vins_fast_b200/feature_tracker.hpp, header-only) and the Python mirror
 * (vins_fast_b200/tracker.py) are thin layers over these calls.
 *
 * Conventions: all functions return vf_status (0 = VF_OK); nothing aborts or throws.  The caller owns
 * every host buffer; the library owns all device memory.  A handle is one sequence of frames (one
 * camera rig); handles are independent and may be driven from different host threads, but one handle
 * is not thread-safe (same as the reference: one caller thread, estimator.cpp:152).
 * There is NO CPU fallback: without a CUDA device every compute call fails with VF_ERR_CUDA.
 */
#ifndef VINS_FAST_B200_H
#define VINS_FAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VF_VERSION 100   /* 0.1.0 */

typedef enum {
    VF_OK = 0,
    VF_ERR_INVALID_ARG = 1,   /* null pointer, bad size/stride, unsupported parameter               */
    VF_ERR_CUDA = 2,          /* CUDA runtime error (no device, OOM, launch failure): see last_error */
    VF_ERR_IO = 3,            /* YAML file missing / unreadable                                      */
    VF_ERR_PARSE = 4,         /* YAML key missing or malformed, unknown camera model                 */
    VF_ERR_STATE = 5          /* call order violated (e.g. debug getter before the first frame)      */
} vf_status;

/* camodocal model types on the hot path (CameraFactory.cc:96-136: model_type, case-insensitive). */
enum { VF_CAM_MEI = 0, VF_CAM_PINHOLE = 1 };

/* One camera: what CataCamera::Parameters / PinholeCamera::Parameters::readFromYamlFile read
 * (CataCamera.cc:161-202, PinholeCamera.cc:145-183).  For MEI fx,fy,cx,cy = gamma1,gamma2,u0,v0. */
typedef struct {
    int32_t model;
    int32_t image_width, image_height;   /* informational (as in the YAML) */
    double xi, k1, k2, p1, p2, fx, fy, cx, cy;
} vf_camera;

/* Tracker configuration.  max_cnt / min_dist / flow_back are the reference's YAML keys
 * (feature_tracker.cpp:10-12, euroc_stero.yaml:44-51).  lk_max_level / lk_win / quality_level are the
 * reference's hard-coded call arguments (feature_tracker.cpp:42,90: Size(21,21), 3, 0.01), exposed as
 * optional keys whose defaults equal the hard-coded values. */
typedef struct {
    int32_t width, height;        /* image size (euroc_stero.yaml:18-19 image_width / image_height) */
    int32_t max_cnt;              /* MAX_CNT_FEATURES_PER_FRAME                                      */
    int32_t min_dist;             /* MIN_DIST                                                        */
    int32_t flow_back;            /* OPTICAL_FLOW_BACK                                               */
    int32_t lk_max_level;         /* default 3   (maxLevel of the forward and stereo LK calls)       */
    int32_t lk_win;               /* default 21; any odd width in [5, 31].  The reference passes the literal cv::Size(21, 21) at all
                                   * four call sites (feature_tracker.cpp:42,51,119,127) and has no key for it.  OpenCV's float
                                   * summation ORDER -- which the LK kernels reproduce to stay bit-identical -- is a function of
                                   * the window width (8-wide SIMD groups over the first (win / 8) * 8 columns + a scalar tail per
                                   * row, lkpyramid.cpp): width 21 runs the tuned kernels built around its chain layout
                                   * (csrc/vf_lk.cu, vf_lk_warp.cu), every other width the run-time-width kernels of
                                   * csrc/vf_lk_generic.cu (same results as OpenCV for that width, less tuned).  Even widths and
                                   * widths outside [5, 31] return VF_ERR_INVALID_ARG                                          */
    double quality_level;         /* default 0.01                                                    */
    vf_camera cam[2];             /* [0] left, [1] right  (ReadIntrinsicParameter order, :301-308)   */
    int32_t device;               /* CUDA device ordinal, -1 = current device                        */
    int32_t use_cuda_graph;       /* 1 = replay the frame as a captured CUDA graph (default), 0 = plain launches */
    void* cuda_stream;            /* optional cudaStream_t the frame-to-frame chain runs on (0 = library-owned stream).  With a
                                   * caller-supplied stream, device-resident input images are read only after everything
                                   * already queued on that stream (see vf_batch_set_input_event)                          */
} vf_config;

/* One tracked feature = one entry of the featureFrame map (feature_tracker.cpp:173-213):
 *   map[id] = { (0,[x,y,1,u,v,vx,vy]) [, (1,[xr,yr,1,ur,vr,vxr,vyr])] }.
 * Records come back in ids_ order; right observations exist only where has_right != 0 and appear in the
 * same relative order as ids_right_. */
typedef struct {
    int32_t id;          /* ids_[i]                                  */
    int32_t track_cnt;   /* track_count_[i]                          */
    int32_t has_right;   /* 1 if the stereo match survived (:128-133) */
    float x, y;          /* current_un_distortion_kps_[i]            */
    float u, v;          /* current_kps_[i]                          */
    float vx, vy;        /* pts_velocity[i]                          */
    float xr, yr, ur, vr, vxr, vyr;   /* right camera counterparts    */
} vf_feature;

typedef struct vf_tracker vf_tracker;   /* one stream of stereo frames                         */
typedef struct vf_batch vf_batch;       /* S independent streams advanced in lock-step on one GPU */

/* ---- configuration (replaces common::Setting + camodocal YAML readers) --------------------------- */
void vf_config_default(vf_config* cfg);
/* Reads the tracker keys of the estimator YAML (OpenCV %YAML:1.0 dialect): max_cnt, min_dist, flow_back,
 * image_width, image_height [, lk_max_level, lk_win, quality_level] -- parameter.hpp:36-53,
 * feature_tracker.cpp:8-18.  Missing mandatory key -> VF_ERR_PARSE (the reference would LOG(FATAL)). */
vf_status vf_config_load_yaml(const char* estimator_yaml, vf_config* cfg);
/* CameraFactory::generateCameraFromYamlFile (CameraFactory.cc:96-189) for MEI and PINHOLE. */
vf_status vf_camera_load_yaml(const char* calib_yaml, vf_camera* cam);
/* common::Setting::Get<T>(key) (vins/common/parameter.hpp:36-53): the raw scalar text of `key` (or "section.key")
 * of an OpenCV-dialect YAML file, NUL-terminated into value_out[cap].  Missing key -> VF_ERR_PARSE. */
vf_status vf_yaml_get(const char* yaml_path, const char* key, char* value_out, size_t cap);
const char* vf_last_global_error(void);   /* message of the last failing call that had no handle */

/* ---- one stream: replaces FeatureTracker ctor/dtor and TrackImage --------------------------------- */
vf_status vf_tracker_create(const vf_config* cfg, vf_tracker** out);          /* feature_tracker.cpp:8-18, 301-308 */
void vf_tracker_destroy(vf_tracker* t);
vf_status vf_tracker_reset(vf_tracker* t);                                    /* back to the state after create   */
/* TrackImage(t, img0, img1) (feature_tracker.cpp:27-217).  img0/img1: CV_8UC1, `height` rows of `width`
 * pixels, row strides in bytes (>= width).  out: capacity >= max_cnt.  Copies the images to the device
 * inside the call and keeps nothing of the caller's memory. */
vf_status vf_tracker_track(vf_tracker* t, double time_s, const uint8_t* img0, size_t stride0,
                           const uint8_t* img1, size_t stride1, vf_feature* out, int32_t* n_out);
/* Same, images already resident on the device (device pointers, same layout).  Device images are copied by the
 * library's own streams: they must be COMPLETE when the call is made -- or be ordered by vf_batch_set_input_event, or be
 * produced on the caller-supplied vf_config.cuda_stream -- and must stay unmodified until the frame has been fetched. */
vf_status vf_tracker_track_device(vf_tracker* t, double time_s, const uint8_t* d_img0, size_t stride0,
                                  const uint8_t* d_img1, size_t stride1, vf_feature* out, int32_t* n_out);
const char* vf_last_error(const vf_tracker* t);

/* ---- S independent streams on one GPU (multi-stream throughput, BASELINE config 4) ---------------- */
vf_status vf_batch_create(const vf_config* cfg, int32_t n_streams, vf_batch** out);
void vf_batch_destroy(vf_batch* b);
vf_status vf_batch_reset(vf_batch* b);
int32_t vf_batch_size(const vf_batch* b);
/* One frame for every stream.  times[s]; img0[s]/img1[s] host pointers (strides shared); out is
 * n_streams * max_cnt records (stream s at out + s*max_cnt); n_out[s]. */
vf_status vf_batch_track(vf_batch* b, const double* times, const uint8_t* const* img0, size_t stride0,
                         const uint8_t* const* img1, size_t stride1, vf_feature* out, int32_t* n_out);
/* Device-resident variant: d_img0/d_img1 each hold n_streams images back to back
 * (image s at base + s*image_stride_bytes). */
vf_status vf_batch_track_device(vf_batch* b, const double* times, const uint8_t* d_img0, const uint8_t* d_img1,
                                size_t row_stride, size_t image_stride_bytes, vf_feature* out, int32_t* n_out);
/* Asynchronous halves of vf_batch_track_device: enqueue the frame, then wait for / fetch its result. */
vf_status vf_batch_enqueue_device(vf_batch* b, const double* times, const uint8_t* d_img0, const uint8_t* d_img1,
                                  size_t row_stride, size_t image_stride_bytes);
vf_status vf_batch_fetch(vf_batch* b, vf_feature* out, int32_t* n_out);
/* Orders the device images of the NEXT vf_batch_enqueue_device / vf_*_track_device call behind `cuda_event` (a cudaEvent_t the
 * caller recorded after the work that produces them, on any stream).  One-shot; the event is not owned by the library. */
vf_status vf_batch_set_input_event(vf_batch* b, void* cuda_event);
/* Pipelined front-end (Estimator::FrontendTracker pushes results into feature_buf_ and the back-end consumes them later,
 * estimator.cpp:143-158): vf_batch_enqueue starts frame k from HOST images and returns at once (the host runs at most
 * four frames ahead of the GPU); vf_batch_fetch_lagged(b, lag, ...) then returns frame k-lag, lag = 0 (= vf_batch_fetch),
 * 1 or 2.  With lag 1 the stereo pass and hand-over of frame k-1 overlap the temporal chain of frame k; with lag 2 the
 * host never waits for a running frame, so the upload and pyramid of frame k+1 are ready when the chain of frame k ends
 * (highest throughput, one more frame of latency).  Results are bit-identical to the synchronous call.  The records of a
 * frame stay readable until three more frames have been enqueued.  The host images of a frame must stay valid until that
 * frame has been fetched (or a later frame has). */
vf_status vf_batch_enqueue(vf_batch* b, const double* times, const uint8_t* const* img0, size_t stride0,
                           const uint8_t* const* img1, size_t stride1);
vf_status vf_batch_fetch_lagged(vf_batch* b, int32_t lag, vf_feature* out, int32_t* n_out);
const char* vf_batch_last_error(const vf_batch* b);
void* vf_batch_cuda_stream(const vf_batch* b);        /* cudaStream_t the frame work is ordered on      */
int32_t vf_batch_launches_per_frame(const vf_batch* b); /* kernels launched by the last frame            */
/* Cumulative host microseconds inside the enqueue calls: out2[0] waiting for the frame two back, out2[1] issuing work. */
void vf_batch_host_times(const vf_batch* b, double* out2);
vf_tracker* vf_batch_stream(vf_batch* b, int32_t s);  /* borrowed view of stream s (for debug getters)  */
/* Measurement aid: CUDA events around every kernel / copy of a frame (plain launches instead of the graph replay).
 * vf_batch_stage_times fills ms_out[vf_batch_stage_count()] with the device times of the last fetched frame. */
vf_status vf_batch_set_profiling(vf_batch* b, int32_t enable);
int32_t vf_batch_stage_count(void);
const char* vf_batch_stage_name(int32_t i);
vf_status vf_batch_stage_times(vf_batch* b, float* ms_out, int32_t cap);
/* Measurement aid (set VF_TRACE=1 in the environment before creating the batch): GPU-timer start / end stamps (ns) of every
 * kernel launched since the previous call, out[2*(k + 16*(frame % 6))] = first start, out[.. + 1] = last end of kernel k
 * of that frame (14 kernels, in the order clear, pyr_left, pad_left, lk_temporal, select_tracked, response,
 * bucket_scan, bucket_scatter, mask_and_max, gftt_select, pyr_right, pad_right, lk_stereo, undistort_left);
 * out[224..255] are ad-hoc phase stamps.  `cap` >= 256.
 * Returns the kernel count, 0 when tracing is off. */
int32_t vf_batch_trace_get(vf_batch* b, unsigned long long* out, int32_t cap);

/* ---- stage-level debug getters (parity tests only; valid after a frame) --------------------------- */
enum {
    VF_DBG_PYR_LEFT = 0,     /* arg = level;  uint8  w_l*h_l       current left pyramid level            */
    VF_DBG_PYR_RIGHT = 1,    /* arg = level;  uint8                current right pyramid level           */
    VF_DBG_EIG = 2,          /* float w*h     cornerMinEigenVal map of the current left image            */
    VF_DBG_MASK = 3,         /* uint8 w*h     SetFeatureExtractorMask result (255 / 0)                   */
    VF_DBG_FWD_PTS = 4,      /* float 2*n_prev  forward LK result                                        */
    VF_DBG_FWD_STATUS = 5,   /* uint8 n_prev                                                             */
    VF_DBG_REV_PTS = 6,      /* float 2*n_prev  reverse LK result                                        */
    VF_DBG_REV_STATUS = 7,   /* uint8 n_prev                                                             */
    VF_DBG_TEMPORAL_STATUS = 8, /* uint8 n_prev  after flow-back + InBorder                              */
    VF_DBG_SORT_PERM = 9,    /* int32 n_tracked  std::sort permutation used by the mask                  */
    VF_DBG_LR_PTS = 10,      /* float 2*n     left->right LK result                                      */
    VF_DBG_LR_STATUS = 11,   /* uint8 n                                                                  */
    VF_DBG_RL_PTS = 12,      /* float 2*n     right->left LK result                                      */
    VF_DBG_RL_STATUS = 13,   /* uint8 n                                                                  */
    VF_DBG_COUNTS = 14,      /* int32[8]: n_prev, n_tracked, n_kept, n_new, n_total, n_right, n_local_max, landmark_id */
    VF_DBG_LK_ITERS = 15     /* arg 0: int32 n_prev, temporal passes; arg 1: int32 n, stereo passes: inner LK iterations per feature
                              * (low 16 bits; the high 16 bits count the iterations that took the exact-integer summation path) */
};
/* Copies the item into `buf` (capacity `cap_bytes`), writes the byte size to *size_out. */
vf_status vf_tracker_debug_get(vf_tracker* t, int32_t what, int32_t arg, void* buf, size_t cap_bytes, size_t* size_out);

/* ---- stand-alone stage operators on HOST buffers (parity tests against the OpenCV routines) ------- */
/* buildOpticalFlowPyramid + pyrDown (lkpyramid.cpp / pyramids.cpp; feature_tracker.cpp:41).
 * levels_out: concatenated unpadded levels 0..n-1 (capacity >= 2*w*h); ws/hs: per-level sizes (cap 8). */
vf_status vf_op_build_pyramid(const uint8_t* img, int32_t w, int32_t h, int32_t max_level, int32_t device,
                              uint8_t* levels_out, int32_t* ws, int32_t* hs, int32_t* n_levels);
/* calcScharrDeriv: out int16 interleaved (Ix,Iy), w*h*2. */
vf_status vf_op_scharr(const uint8_t* img, int32_t w, int32_t h, int32_t device, int16_t* out);
/* calcOpticalFlowPyrLK(prev, next, prev_pts, next_pts(in/out), status, winSize 21, max_level, criteria 30/0.01,
 * flags = use_initial_flow ? OPTFLOW_USE_INITIAL_FLOW : 0, minEigThreshold 1e-4). */
vf_status vf_op_lk(const uint8_t* prev, const uint8_t* next, int32_t w, int32_t h, const float* prev_pts,
                   float* next_pts, uint8_t* status, int32_t n, int32_t max_level, int32_t use_initial_flow,
                   int32_t device, int32_t* iters_out);
/* The same call with winSize (win, win), win odd in [5, 31] (the pyramid's early stop follows the window too). */
vf_status vf_op_lk_win(const uint8_t* prev, const uint8_t* next, int32_t w, int32_t h, const float* prev_pts,
                       float* next_pts, uint8_t* status, int32_t n, int32_t max_level, int32_t win,
                       int32_t use_initial_flow, int32_t device, int32_t* iters_out);
/* Which corner-response kernel the library launches (process-wide; parity tests only): 0 = by frame size (default),
 * 1 = tiles (small frames), 2 = strips (large frames / batches).  Both give identical results. */
vf_status vf_debug_set_response_kernel(int32_t mode);
/* Which Lucas-Kanade kernels the library launches (process-wide; A/B measurements and parity tests; takes effect for handles
 * created afterwards): 0 = by launch size (default: one 4-warp CTA per feature with a block barrier per iteration while every
 * feature has an SM to itself, one warp per feature once features queue up per SM); 1 = the former whatever the size; 2 / 3 =
 * warp-level passes with one warp per feature / with a cooperating 4-warp CTA per feature whatever the size.  Identical results. */
vf_status vf_debug_set_lk_kernel(int32_t mode);
/* Which pyramid kernel the library launches (process-wide; parity tests only; takes effect for handles created afterwards):
 * 0 = by size (default), 1 = fused multi-level kernel only, 2 = streaming TMA kernel for every level transition. */
vf_status vf_debug_set_pyramid_kernel(int32_t mode);
/* cornerMinEigenVal(img, 3, ksize 3): out float w*h. */
vf_status vf_op_min_eigen(const uint8_t* img, int32_t w, int32_t h, int32_t device, float* out);
/* goodFeaturesToTrack(img, max_corners, quality, min_dist, mask): out_xy capacity 2*max_corners floats. */
vf_status vf_op_good_features(const uint8_t* img, int32_t w, int32_t h, int32_t max_corners, double quality,
                              int32_t min_dist, const uint8_t* mask, int32_t device, float* out_xy, int32_t* n_out);
/* SetFeatureExtractorMask (feature_tracker.cpp:219-246) on given points/track counts: perm_out = std::sort
 * order (n), keep_out = indices kept in order (cap n), mask_out optional w*h. */
vf_status vf_op_set_mask(const float* pts, const int32_t* track_cnt, int32_t n, int32_t w, int32_t h,
                         int32_t min_dist, int32_t device, int32_t* perm_out, int32_t* keep_out, int32_t* n_keep,
                         uint8_t* mask_out);
/* RemoveDistortion (feature_tracker.cpp:310-321): out_xy float 2*n. */
vf_status vf_op_undistort(const vf_camera* cam, const float* uv, int32_t n, int32_t device, float* out_xy);

/* ---- next data-parallel step after the front-end: landmark triangulation + outlier rejection ---- */
/* The sliding window the estimator keeps on the host (arguments of TriangulatePts / OutliersRejection,
 * feature_manager.cpp:202-203, 379-380): Ps[window][3] = P_wi, Rs[window][9] = R_wi (row-major), camera-to-IMU
 * extrinsics tic[c][3], ric[c][9] for the left (0) and right (1) camera. */
typedef struct {
    int32_t window;
    const double* Ps;
    const double* Rs;
    double tic[2][3];
    double ric[2][9];
} vf_window;
/* FeatureManager::TriangulateOnePoint (feature_manager.cpp:277-293) for n landmarks: pose0 / pose1 = n x 12 (row-major
 * 3x4 T_cw), point0 / point1 = n x 2 normalised coordinates, point_3d = n x 3 (world frame). */
vf_status vf_op_triangulate_one_point(const double* pose0, const double* pose1, const double* point0, const double* point1,
                                      int32_t n, int32_t device, double* point_3d);
/* Landmark table = FeatureManager::features_ flattened: landmark i was first seen in window frame start_frame[i] and has
 * n_obs[i] consecutive observations in rows obs_offset[i] .. obs_offset[i] + n_obs[i] - 1 of `obs` (n_obs_rows rows of 7
 * doubles: point_.x, point_.y, point_.z, track_right_success_ (0 / 1), point_right_.x, .y, .z).
 * FeatureManager::TriangulatePts (feature_manager.cpp:202-275): depth_io[i] = estimated_depth_ (input: <= 0 = not yet
 * triangulated; those are filled from the stereo pair of the first frame, else from the first two frames; a
 * non-positive depth becomes init_depth = INIT_DEPTH, feature_manager.hpp:88). */
vf_status vf_op_triangulate_pts(const vf_window* w, int32_t n, const int32_t* start_frame, const int32_t* n_obs,
                                const int32_t* obs_offset, const double* obs, int32_t n_obs_rows, double init_depth,
                                int32_t device, double* depth_io);
/* FeatureManager::OutliersRejection (feature_manager.cpp:379-429): mean ReProjectionError (:365-377) of landmark i over its
 * observations (landmarks with fewer than 4 are skipped: remove_out 0, ave_err_out 0); remove_out[i] = 1 iff
 * ave_err * focal_length > 3.  ave_err_out may be null. */
vf_status vf_op_outliers_rejection(const vf_window* w, int32_t n, const int32_t* start_frame, const int32_t* n_obs,
                                   const int32_t* obs_offset, const double* obs, int32_t n_obs_rows, const double* depth,
                                   double focal_length, int32_t device, double* ave_err_out, uint8_t* remove_out);

/* Pinned host memory helpers (so callers can hand page-locked images to vf_tracker_track). */
vf_status vf_host_alloc(size_t bytes, void** out);
void vf_host_free(void* p);
int32_t vf_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VINS_FAST_B200_H */
