// Batched landmark triangulation and reprojection-error outlier rejection: the per-feature work
// that follows the front-end in the estimator,
//     FeatureManager::TriangulatePts / TriangulateOnePoint    vins/estimator/feature_manager.cpp:202-293
//     FeatureManager::ReProjectionError / OutliersRejection   vins/estimator/feature_manager.cpp:365-429
// one thread per landmark, float64 throughout (the reference's Eigen types are double).  The sliding-window poses stay on
// the host with the Ceres estimator and are handed in per call (a few hundred bytes); the observation table is the
// estimator's feature_per_frame_ lists flattened.
//
// TriangulateOnePoint: null vector of the 4x4 design matrix.  The reference takes it from Eigen's two-sided JacobiSVD
// (matrixV().rightCols<1>()); here a one-sided (Hestenes) Jacobi SVD orthogonalises the columns of A while accumulating V,
// and the column with the smallest norm is the same singular vector up to sign -- which cancels in x/w, y/w, z/w.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "vf_common.cuh"

extern "C" void vf_internal_set_error(const char* msg);

namespace {

struct WindowDev {
    int window;
    const double* Ps;   // [window][3]
    const double* Rs;   // [window][9] row-major
    double tic[2][3];
    double ric[2][9];
};

__device__ __forceinline__ void mat3_mul(const double* a, const double* b, double* c) {   // c = a b (row-major)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// T_cw = [R^T | -R^T t] with R = Rs[w] ric[c], t = Ps[w] + Rs[w] tic[c]   (feature_manager.cpp:219-231)
__device__ void camera_pose(const WindowDev& W, int frame, int cam, double* pose /*3x4 row-major*/) {
    const double* Rw = W.Rs + 9 * frame;
    const double* Pw = W.Ps + 3 * frame;
    double R[9], t[3];
    mat3_mul(Rw, W.ric[cam], R);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = Pw[i] + (Rw[3 * i] * W.tic[cam][0] + Rw[3 * i + 1] * W.tic[cam][1] + Rw[3 * i + 2] * W.tic[cam][2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) pose[4 * i + j] = R[3 * j + i];
        pose[4 * i + 3] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);
    }
}

// Right singular vector of the smallest singular value of the 4x4 matrix a (row-major), by one-sided Jacobi.
__device__ void null_vector4(double* a, double* v_out) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { alpha += a[4 * i + p] * a[4 * i + p]; beta += a[4 * i + q] * a[4 * i + q]; gamma += a[4 * i + p] * a[4 * i + q]; }
                if (gamma == 0.0 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double ap = a[4 * i + p], aq = a[4 * i + q];
                    a[4 * i + p] = c * ap - s * aq; a[4 * i + q] = s * ap + c * aq;
                    const double vp = v[4 * i + p], vq = v[4 * i + q];
                    v[4 * i + p] = c * vp - s * vq; v[4 * i + q] = s * vp + c * vq;
                }
            }
        }
        if (!rotated) break;
    }
    int best = 0;
    double best_n = 1e300;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double n = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) n += a[4 * i + j] * a[4 * i + j];
        if (n < best_n) { best_n = n; best = j; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v_out[i] = v[4 * i + best];
}

// FeatureManager::TriangulateOnePoint (feature_manager.cpp:277-293)
__device__ void triangulate_one_point(const double* pose0, const double* pose1, double x0, double y0, double x1, double y1, double* p3) {
    double a[16], tp[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a[c] = x0 * pose0[8 + c] - pose0[c];
        a[4 + c] = y0 * pose0[8 + c] - pose0[4 + c];
        a[8 + c] = x1 * pose1[8 + c] - pose1[c];
        a[12 + c] = y1 * pose1[8 + c] - pose1[4 + c];
    }
    null_vector4(a, tp);
    p3[0] = tp[0] / tp[3]; p3[1] = tp[1] / tp[3]; p3[2] = tp[2] / tp[3];
}

__global__ void triangulate_one_point_kernel(const double* __restrict__ pose0, const double* __restrict__ pose1,
                                             const double* __restrict__ pt0, const double* __restrict__ pt1, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[12], b[12], p3[3];
#pragma unroll
    for (int k = 0; k < 12; ++k) { a[k] = pose0[12 * (size_t)i + k]; b[k] = pose1[12 * (size_t)i + k]; }
    triangulate_one_point(a, b, pt0[2 * i], pt0[2 * i + 1], pt1[2 * i], pt1[2 * i + 1], p3);
    out[3 * (size_t)i] = p3[0]; out[3 * (size_t)i + 1] = p3[1]; out[3 * (size_t)i + 2] = p3[2];
}

// FeatureManager::TriangulatePts (feature_manager.cpp:202-275), one landmark per thread.  obs rows are
// {x, y, z, has_right, xr, yr, zr}; depth_io holds estimated_depth_.
__global__ void triangulate_pts_kernel(const __grid_constant__ WindowDev W, int n, const int* __restrict__ start_frame,
                                       const int* __restrict__ n_obs, const int* __restrict__ obs_offset, const double* __restrict__ obs,
                                       double init_depth, double* __restrict__ depth_io) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (depth_io[i] > 0) return;                                               // already triangulated (:210-211)
    const double* o0 = obs + 7 * (size_t)obs_offset[i];
    const int f0 = start_frame[i];
    double left[12], right[12], p3[3];
    double x1, y1;
    if (o0[3] != 0.0) {                                                        // stereo: first left / right observation (:213-245)
        if (f0 < 0 || f0 >= W.window) return;
        camera_pose(W, f0, 0, left);
        camera_pose(W, f0, 1, right);
        x1 = o0[4]; y1 = o0[5];
    } else if (n_obs[i] > 1) {                                                 // two consecutive left observations (:247-275)
        if (f0 < 0 || f0 + 1 >= W.window) return;
        camera_pose(W, f0, 0, left);
        camera_pose(W, f0 + 1, 0, right);
        x1 = o0[7]; y1 = o0[8];
    } else {
        return;
    }
    triangulate_one_point(left, right, o0[0], o0[1], x1, y1, p3);
    const double depth = left[8] * p3[0] + left[9] * p3[1] + left[10] * p3[2] + left[11];
    depth_io[i] = depth > 0 ? depth : init_depth;
}

// FeatureManager::ReProjectionError (feature_manager.cpp:365-377)
__device__ double reprojection_error(const double* Ri, const double* Pi, const double* rici, const double* tici, const double* Rj,
                                     const double* Pj, const double* ricj, const double* ticj, double depth, const double* uvi,
                                     const double* uvj) {
    double a[3], pc[3], pw[3], d[3], e[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] = depth * uvi[k];
#pragma unroll
    for (int r = 0; r < 3; ++r) pc[r] = rici[3 * r] * a[0] + rici[3 * r + 1] * a[1] + rici[3 * r + 2] * a[2] + tici[r];
#pragma unroll
    for (int r = 0; r < 3; ++r) pw[r] = Ri[3 * r] * pc[0] + Ri[3 * r + 1] * pc[1] + Ri[3 * r + 2] * pc[2] + Pi[r];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = pw[k] - Pj[k];
#pragma unroll
    for (int r = 0; r < 3; ++r) e[r] = Rj[r] * d[0] + Rj[3 + r] * d[1] + Rj[6 + r] * d[2] - ticj[r];          // Rj^T d - ticj
#pragma unroll
    for (int r = 0; r < 3; ++r) q[r] = ricj[r] * e[0] + ricj[3 + r] * e[1] + ricj[6 + r] * e[2];                // ricj^T e
    const double rx = q[0] / q[2] - uvj[0], ry = q[1] / q[2] - uvj[1];
    return sqrt(rx * rx + ry * ry);
}

// FeatureManager::OutliersRejection (feature_manager.cpp:379-429), one landmark per thread
__global__ void outliers_kernel(const __grid_constant__ WindowDev W, int n, const int* __restrict__ start_frame,
                                const int* __restrict__ n_obs, const int* __restrict__ obs_offset, const double* __restrict__ obs,
                                const double* __restrict__ depth, double focal_length, double* __restrict__ ave_err_out,
                                unsigned char* __restrict__ remove_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    remove_out[i] = 0;
    ave_err_out[i] = 0.0;
    const int used = n_obs[i];
    if (used < 4) return;                                                       // :388-390
    const int imu_i = start_frame[i];
    if (imu_i < 0 || imu_i + used > W.window) return;
    const double* rows = obs + 7 * (size_t)obs_offset[i];
    const double* pts_i = rows;
    const double dep = depth[i];
    double err = 0;
    int cnt = 0;
    for (int k = 0; k < used; ++k) {
        const int imu_j = imu_i + k;
        const double* o = rows + 7 * k;
        if (imu_i != imu_j) {
            err += reprojection_error(W.Rs + 9 * imu_i, W.Ps + 3 * imu_i, W.ric[0], W.tic[0], W.Rs + 9 * imu_j, W.Ps + 3 * imu_j, W.ric[0],
                                      W.tic[0], dep, pts_i, o);
            ++cnt;
        }
        if (o[3] != 0.0) {                                                      // right observation: both branches of :409-423 are the same call
            err += reprojection_error(W.Rs + 9 * imu_i, W.Ps + 3 * imu_i, W.ric[0], W.tic[0], W.Rs + 9 * imu_j, W.Ps + 3 * imu_j, W.ric[1],
                                      W.tic[1], dep, pts_i, o + 4);
            ++cnt;
        }
    }
    const double ave = err / cnt;
    ave_err_out[i] = ave;
    remove_out[i] = (ave * focal_length > 3) ? 1 : 0;
}

vf_status geo_fail(vf_status st, const std::string& msg) { vf_internal_set_error(msg.c_str()); return st; }

#define GEO_CUDA(call)                                                                                                 \
    do {                                                                                                               \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return geo_fail(VF_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e__)); \
    } while (0)

struct DevBuf {   // frees on scope exit
    std::vector<void*> p;
    ~DevBuf() { for (void* q : p) cudaFree(q); }
    template <typename T> cudaError_t put(T** out, const T* host, size_t count) {
        cudaError_t e = cudaMalloc((void**)out, sizeof(T) * (count ? count : 1));
        if (e != cudaSuccess) return e;
        p.push_back(*out);
        return host && count ? cudaMemcpy(*out, host, sizeof(T) * count, cudaMemcpyHostToDevice) : cudaSuccess;
    }
};

vf_status pick_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return geo_fail(VF_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device >= ndev) return geo_fail(VF_ERR_INVALID_ARG, "device ordinal out of range");
    if (device >= 0) GEO_CUDA(cudaSetDevice(device));
    return VF_OK;
}

vf_status upload_window(const vf_window* w, DevBuf& buf, WindowDev& out) {
    if (!w || !w->Ps || !w->Rs || w->window < 1 || w->window > 4096) return geo_fail(VF_ERR_INVALID_ARG, "window poses missing or window not in [1, 4096]");
    double *dPs = nullptr, *dRs = nullptr;
    GEO_CUDA(buf.put(&dPs, w->Ps, 3 * (size_t)w->window));
    GEO_CUDA(buf.put(&dRs, w->Rs, 9 * (size_t)w->window));
    out.window = w->window; out.Ps = dPs; out.Rs = dRs;
    memcpy(out.tic, w->tic, sizeof(out.tic));
    memcpy(out.ric, w->ric, sizeof(out.ric));
    return VF_OK;
}

vf_status check_table(int32_t n, const int32_t* start_frame, const int32_t* n_obs, const int32_t* obs_offset, const double* obs, int32_t n_rows) {
    if (n < 0 || n_rows < 0 || (n > 0 && (!start_frame || !n_obs || !obs_offset || !obs))) return geo_fail(VF_ERR_INVALID_ARG, "null landmark table");
    for (int i = 0; i < n; ++i)
        if (n_obs[i] < 1 || obs_offset[i] < 0 || (long long)obs_offset[i] + n_obs[i] > n_rows)
            return geo_fail(VF_ERR_INVALID_ARG, "landmark " + std::to_string(i) + ": observation rows outside the table");
    return VF_OK;
}

}  // namespace

extern "C" {

vf_status vf_op_triangulate_one_point(const double* pose0, const double* pose1, const double* point0, const double* point1, int32_t n,
                                      int32_t device, double* point_3d) {
    if (n < 0 || (n > 0 && (!pose0 || !pose1 || !point0 || !point1 || !point_3d))) return geo_fail(VF_ERR_INVALID_ARG, "null argument");
    if (n == 0) return VF_OK;
    vf_status st = pick_device(device);
    if (st != VF_OK) return st;
    DevBuf buf;
    double *a, *b, *p0, *p1, *out;
    GEO_CUDA(buf.put(&a, pose0, 12 * (size_t)n)); GEO_CUDA(buf.put(&b, pose1, 12 * (size_t)n));
    GEO_CUDA(buf.put(&p0, point0, 2 * (size_t)n)); GEO_CUDA(buf.put(&p1, point1, 2 * (size_t)n));
    GEO_CUDA(buf.put(&out, (const double*)nullptr, 3 * (size_t)n));
    triangulate_one_point_kernel<<<(n + 127) / 128, 128>>>(a, b, p0, p1, n, out);
    GEO_CUDA(cudaGetLastError());
    GEO_CUDA(cudaMemcpy(point_3d, out, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    return VF_OK;
}

vf_status vf_op_triangulate_pts(const vf_window* w, int32_t n, const int32_t* start_frame, const int32_t* n_obs, const int32_t* obs_offset,
                                const double* obs, int32_t n_obs_rows, double init_depth, int32_t device, double* depth_io) {
    vf_status st = check_table(n, start_frame, n_obs, obs_offset, obs, n_obs_rows);
    if (st != VF_OK) return st;
    if (n > 0 && !depth_io) return geo_fail(VF_ERR_INVALID_ARG, "depth_io is null");
    if (n == 0) return VF_OK;
    if ((st = pick_device(device)) != VF_OK) return st;
    DevBuf buf;
    WindowDev W;
    if ((st = upload_window(w, buf, W)) != VF_OK) return st;
    int *sf, *no, *oo;
    double *ob, *dep;
    GEO_CUDA(buf.put(&sf, start_frame, (size_t)n)); GEO_CUDA(buf.put(&no, n_obs, (size_t)n)); GEO_CUDA(buf.put(&oo, obs_offset, (size_t)n));
    GEO_CUDA(buf.put(&ob, obs, 7 * (size_t)n_obs_rows + 7));   // one spare row: the two-frame branch reads row 1 only when n_obs > 1
    GEO_CUDA(buf.put(&dep, depth_io, (size_t)n));
    triangulate_pts_kernel<<<(n + 127) / 128, 128>>>(W, n, sf, no, oo, ob, init_depth, dep);
    GEO_CUDA(cudaGetLastError());
    GEO_CUDA(cudaMemcpy(depth_io, dep, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return VF_OK;
}

vf_status vf_op_outliers_rejection(const vf_window* w, int32_t n, const int32_t* start_frame, const int32_t* n_obs, const int32_t* obs_offset,
                                   const double* obs, int32_t n_obs_rows, const double* depth, double focal_length, int32_t device,
                                   double* ave_err_out, uint8_t* remove_out) {
    vf_status st = check_table(n, start_frame, n_obs, obs_offset, obs, n_obs_rows);
    if (st != VF_OK) return st;
    if (n > 0 && (!depth || !remove_out)) return geo_fail(VF_ERR_INVALID_ARG, "depth / remove_out is null");
    if (n == 0) return VF_OK;
    if ((st = pick_device(device)) != VF_OK) return st;
    DevBuf buf;
    WindowDev W;
    if ((st = upload_window(w, buf, W)) != VF_OK) return st;
    int *sf, *no, *oo;
    double *ob, *dep, *ave;
    unsigned char* rm;
    GEO_CUDA(buf.put(&sf, start_frame, (size_t)n)); GEO_CUDA(buf.put(&no, n_obs, (size_t)n)); GEO_CUDA(buf.put(&oo, obs_offset, (size_t)n));
    GEO_CUDA(buf.put(&ob, obs, 7 * (size_t)n_obs_rows)); GEO_CUDA(buf.put(&dep, depth, (size_t)n));
    GEO_CUDA(buf.put(&ave, (const double*)nullptr, (size_t)n)); GEO_CUDA(buf.put(&rm, (const unsigned char*)nullptr, (size_t)n));
    outliers_kernel<<<(n + 127) / 128, 128>>>(W, n, sf, no, oo, ob, dep, focal_length, ave, rm);
    GEO_CUDA(cudaGetLastError());
    if (ave_err_out) GEO_CUDA(cudaMemcpy(ave_err_out, ave, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    GEO_CUDA(cudaMemcpy(remove_out, rm, (size_t)n, cudaMemcpyDeviceToHost));
    return VF_OK;
}

}  // extern "C"
