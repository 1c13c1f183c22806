// camodocal liftProjective + perspective division on the device, float64, one thread per point.
// Restates, expression by expression (same association order, explicit round-to-nearest ops so nvcc cannot
// contract them into FMAs):
//   CataCamera::liftProjective     camera_models/src/camera_models/CataCamera.cc:556-626   (MEI, n = 8 fixed-point steps)
//   CataCamera::distortion         camera_models/src/camera_models/CataCamera.cc:766-782
//   inverse K                      camera_models/src/camera_models/CataCamera.cc:320-323
//   PinholeCamera::liftProjective  camera_models/src/camera_models/PinholeCamera.cc:450-510
//   PinholeCamera::distortion      camera_models/src/camera_models/PinholeCamera.cc:646-662
//   RemoveDistortion (x/z, y/z -> float)   vins/estimator/feature_tracker.cpp:310-321
#pragma once
#include "vf_common.cuh"

namespace vf {

__device__ __forceinline__ double dM(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dA(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dS(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ void cam_distortion(const CamParams& c, double x, double y, double& dx, double& dy) {
    const double mx2 = dM(x, x), my2 = dM(y, y), mxy = dM(x, y);
    const double rho2 = dA(mx2, my2);
    const double rad = dA(dM(c.k1, rho2), dM(dM(c.k2, rho2), rho2));
    // x*rad + 2*p1*mxy + p2*(rho2 + 2*mx2)
    dx = dA(dA(dM(x, rad), dM(dM(2.0, c.p1), mxy)), dM(c.p2, dA(rho2, dM(2.0, mx2))));
    dy = dA(dA(dM(y, rad), dM(dM(2.0, c.p2), mxy)), dM(c.p1, dA(rho2, dM(2.0, my2))));
}

__device__ __forceinline__ float2 undistort_point(const CamParams& c, float2 uv) {
    const double i11 = __ddiv_rn(1.0, c.fx), i13 = __ddiv_rn(-c.cx, c.fx);
    const double i22 = __ddiv_rn(1.0, c.fy), i23 = __ddiv_rn(-c.cy, c.fy);
    const double mx_d = dA(dM(i11, (double)uv.x), i13), my_d = dA(dM(i22, (double)uv.y), i23);
    double mx_u, my_u;
    if (c.k1 == 0.0 && c.k2 == 0.0 && c.p1 == 0.0 && c.p2 == 0.0) {
        mx_u = mx_d; my_u = my_d;
    } else {
        double dx, dy;
        cam_distortion(c, mx_d, my_d, dx, dy);
        mx_u = dS(mx_d, dx); my_u = dS(my_d, dy);
        for (int i = 1; i < 8; ++i) {
            cam_distortion(c, mx_u, my_u, dx, dy);
            mx_u = dS(mx_d, dx); my_u = dS(my_d, dy);
        }
    }
    double z = 1.0;
    if (c.model == VF_CAM_MEI) {
        if (c.xi == 1.0) {
            z = __ddiv_rn(dS(dS(1.0, dM(mx_u, mx_u)), dM(my_u, my_u)), 2.0);
        } else {
            const double rho2 = dA(dM(mx_u, mx_u), dM(my_u, my_u));
            const double den = dA(c.xi, sqrt(dA(1.0, dM(dS(1.0, dM(c.xi, c.xi)), rho2))));
            z = dS(1.0, __ddiv_rn(dM(c.xi, dA(rho2, 1.0)), den));
        }
    }
    return make_float2((float)__ddiv_rn(mx_u, z), (float)__ddiv_rn(my_u, z));
}

}  // namespace vf
