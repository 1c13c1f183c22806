/* TEST INFRASTRUCTURE (oracle) -- see vf_oracle.h for scope, provenance and the parity statement.
 *
 * Compile WITHOUT fused multiply-add and without fast-math so every float expression rounds the way
 * the x86-64 baseline (SSE2/SSE3, no FMA) build of OpenCV rounds:
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -c oracle/vf_oracle.c
 */
#include "vf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int refl101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}
static inline int cv_round(float v) { return (int)lrintf(v); }   /* round-half-even, as cvRound (SSE cvtss2si) */
static inline int cv_floor(float v) { return (int)floorf(v); }

/* ------------------------------------------------------------------------------------------------
 * A2  pyrDown (imgproc/pyramids.cpp): separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8.
 * Reached from feature_tracker.cpp:41,50,117,125 through buildOpticalFlowPyramid.
 * ------------------------------------------------------------------------------------------------ */
void vfo_pyrdown_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* rows = (int*)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; ++y) {
        for (int k = 0; k < 5; ++k) {
            const uint8_t* s = src + (size_t)refl101(2 * y + k - 2, h) * w;
            int* r = rows + (size_t)k * dw;
            for (int x = 0; x < dw; ++x) {
                int c = 2 * x;
                r[x] = s[refl101(c - 2, w)] + 4 * s[refl101(c - 1, w)] + 6 * s[refl101(c, w)] +
                       4 * s[refl101(c + 1, w)] + s[refl101(c + 2, w)];
            }
        }
        for (int x = 0; x < dw; ++x) {
            int v = rows[x] + 4 * rows[dw + x] + 6 * rows[2 * dw + x] + 4 * rows[3 * dw + x] + rows[4 * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(rows);
}

/* buildOpticalFlowPyramid (video/lkpyramid.cpp): level l = pyrDown(level l-1); stop after a level whose
 * successor would be <= winSize in either dimension. */
void vfo_pyramid_build(vfo_pyramid* p, const uint8_t* img, int w, int h, int max_level, int win) {
    memset(p, 0, sizeof(*p));
    if (max_level > VFO_MAX_LEVELS - 1) max_level = VFO_MAX_LEVELS - 1;
    p->w[0] = w; p->h[0] = h;
    p->img[0] = (uint8_t*)malloc((size_t)w * h);
    memcpy(p->img[0], img, (size_t)w * h);
    p->n_levels = 1;
    for (int l = 0; l <= max_level; ++l) {
        if (l > 0) {
            p->w[l] = (p->w[l - 1] + 1) / 2; p->h[l] = (p->h[l - 1] + 1) / 2;
            p->img[l] = (uint8_t*)malloc((size_t)p->w[l] * p->h[l]);
            vfo_pyrdown_u8(p->img[l - 1], p->w[l - 1], p->h[l - 1], p->img[l]);
            p->n_levels = l + 1;
        }
        int nw = (p->w[l] + 1) / 2, nh = (p->h[l] + 1) / 2;
        if (nw <= win || nh <= win) break;
    }
}

void vfo_pyramid_free(vfo_pyramid* p) {
    for (int l = 0; l < VFO_MAX_LEVELS; ++l) { free(p->img[l]); p->img[l] = NULL; }
    p->n_levels = 0;
}

/* ------------------------------------------------------------------------------------------------
 * A3  calcScharrDeriv (video/lkpyramid.cpp): vertical pass t0 = (r-1 + r+1)*3 + r0*10, t1 = r+1 - r-1,
 * horizontal pass Ix = t0[x+1]-t0[x-1], Iy = (t1[x-1]+t1[x+1])*3 + t1[x]*10; reflect-101 inside the image.
 * ------------------------------------------------------------------------------------------------ */
void vfo_scharr_s16(const uint8_t* src, int w, int h, int16_t* dst) {
    int* t0 = (int*)malloc(sizeof(int) * (size_t)w);
    int* t1 = (int*)malloc(sizeof(int) * (size_t)w);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = src + (size_t)refl101(y - 1, h) * w;
        const uint8_t* r1 = src + (size_t)y * w;
        const uint8_t* r2 = src + (size_t)refl101(y + 1, h) * w;
        for (int x = 0; x < w; ++x) {
            t0[x] = (r0[x] + r2[x]) * 3 + r1[x] * 10;
            t1[x] = r2[x] - r0[x];
        }
        for (int x = 0; x < w; ++x) {
            int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
            dst[((size_t)y * w + x) * 2 + 0] = (int16_t)(t0[xp] - t0[xm]);
            dst[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1[xm] + t1[xp]) * 3 + t1[x] * 10);
        }
    }
    free(t0); free(t1);
}

/* ------------------------------------------------------------------------------------------------
 * A4  calcOpticalFlowPyrLK / LKTrackerInvoker (video/lkpyramid.cpp).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int w, h, pad, stride;   /* stride in elements of the padded buffer */
    uint8_t* img;            /* (h+2pad) x (w+2pad), REFLECT_101 border  */
    int16_t* der;            /* same geometry x2 (Ix,Iy), ZERO border    */
} lk_level;

static void lk_level_make(lk_level* L, const uint8_t* img, int w, int h, int pad, int with_deriv) {
    L->w = w; L->h = h; L->pad = pad; L->stride = w + 2 * pad;
    L->img = (uint8_t*)malloc((size_t)L->stride * (h + 2 * pad));
    for (int y = -pad; y < h + pad; ++y) {
        const uint8_t* s = img + (size_t)refl101(y, h) * w;
        uint8_t* d = L->img + (size_t)(y + pad) * L->stride;
        for (int x = -pad; x < w + pad; ++x) d[x + pad] = s[refl101(x, w)];
    }
    L->der = NULL;
    if (with_deriv) {
        int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)w * h);
        vfo_scharr_s16(img, w, h, tmp);
        L->der = (int16_t*)calloc((size_t)L->stride * (h + 2 * pad) * 2, sizeof(int16_t));
        for (int y = 0; y < h; ++y)
            memcpy(L->der + ((size_t)(y + pad) * L->stride + pad) * 2, tmp + (size_t)y * w * 2,
                   sizeof(int16_t) * 2 * (size_t)w);
        free(tmp);
    }
}
static void lk_level_free(lk_level* L) { free(L->img); free(L->der); L->img = NULL; L->der = NULL; }

#define LK_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static inline float reduce4(const float q[4]) {
    /* v_reduce_sum(v_float32x4) of OpenCV's SSE backend: (q0+q2)+(q1+q3) */
    return (q[0] + q[2]) + (q[1] + q[3]);
}

void vfo_lk(const vfo_pyramid* prev, const vfo_pyramid* next, const float* prev_pts, float* next_pts,
            uint8_t* status, int n, int max_level, int win, int use_initial_flow, int max_count,
            double eps, float min_eig_thr, int accum, int* iters_out) {
    if (n <= 0) return;
    if (max_level > prev->n_levels - 1) max_level = prev->n_levels - 1;
    if (max_level > next->n_levels - 1) max_level = next->n_levels - 1;
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    eps *= eps;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int W_BITS = 14;
    int16_t* Iw = (int16_t*)malloc(sizeof(int16_t) * (size_t)win * win);
    int16_t* dIw = (int16_t*)malloc(sizeof(int16_t) * (size_t)win * win * 2);
    for (int i = 0; i < n; ++i) status[i] = 1;
    if (iters_out) memset(iters_out, 0, sizeof(int) * (size_t)n);
    const int simd_w = (accum == 0) ? (win / 8) * 8 : 0;   /* columns covered by the 8-wide SIMD loop */

    for (int level = max_level; level >= 0; --level) {
        lk_level LI, LJ;
        lk_level_make(&LI, prev->img[level], prev->w[level], prev->h[level], win, 1);
        lk_level_make(&LJ, next->img[level], next->w[level], next->h[level], win, 0);
        const int cols = LI.w, rows = LI.h, jcols = LJ.w, jrows = LJ.h;
        for (int p = 0; p < n; ++p) {
            float ppx = prev_pts[2 * p] * (float)(1. / (1 << level));
            float ppy = prev_pts[2 * p + 1] * (float)(1. / (1 << level));
            float nx, ny;
            if (level == max_level) {
                if (use_initial_flow) {
                    nx = next_pts[2 * p] * (float)(1. / (1 << level));
                    ny = next_pts[2 * p + 1] * (float)(1. / (1 << level));
                } else { nx = ppx; ny = ppy; }
            } else { nx = next_pts[2 * p] * 2.f; ny = next_pts[2 * p + 1] * 2.f; }
            next_pts[2 * p] = nx; next_pts[2 * p + 1] = ny;

            ppx -= half; ppy -= half;
            int ipx = cv_floor(ppx), ipy = cv_floor(ppy);
            if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
                if (level == 0) status[p] = 0;
                continue;
            }
            float a = ppx - ipx, b = ppy - ipy;
            int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
            int iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
            int iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
            int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

            float qA11[4] = {0, 0, 0, 0}, qA12[4] = {0, 0, 0, 0}, qA22[4] = {0, 0, 0, 0};
            float iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; ++y) {
                const uint8_t* s = LI.img + (size_t)(y + ipy + LI.pad) * LI.stride + (ipx + LI.pad);
                const int16_t* d = LI.der + ((size_t)(y + ipy + LI.pad) * LI.stride + (ipx + LI.pad)) * 2;
                const int st = LI.stride;
                for (int x = 0; x < win; ++x) {
                    int ival = LK_DESCALE(s[x] * iw00 + s[x + 1] * iw01 + s[x + st] * iw10 + s[x + st + 1] * iw11,
                                          W_BITS - 5);
                    int ixval = LK_DESCALE(d[2 * x] * iw00 + d[2 * x + 2] * iw01 + d[2 * (x + st)] * iw10 +
                                               d[2 * (x + st) + 2] * iw11, W_BITS);
                    int iyval = LK_DESCALE(d[2 * x + 1] * iw00 + d[2 * x + 3] * iw01 + d[2 * (x + st) + 1] * iw10 +
                                               d[2 * (x + st) + 3] * iw11, W_BITS);
                    Iw[y * win + x] = (int16_t)ival;
                    dIw[(y * win + x) * 2] = (int16_t)ixval;
                    dIw[(y * win + x) * 2 + 1] = (int16_t)iyval;
                    float fx = (float)ixval, fy = (float)iyval;
                    if (x < simd_w) {
                        int l = x & 3;
                        qA22[l] = fy * fy + qA22[l];
                        qA12[l] = fx * fy + qA12[l];
                        qA11[l] = fx * fx + qA11[l];
                    } else {
                        iA11 += (float)(ixval * ixval);
                        iA12 += (float)(ixval * iyval);
                        iA22 += (float)(iyval * iyval);
                    }
                }
            }
            if (simd_w) { iA11 += reduce4(qA11); iA12 += reduce4(qA12); iA22 += reduce4(qA22); }
            float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < min_eig_thr || D < FLT_EPSILON) {
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= half; ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_count; ++j) {
                int inx = cv_floor(nx), iny = cv_floor(ny);
                if (inx < -win || inx >= jcols || iny < -win || iny >= jrows) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                if (iters_out) iters_out[p]++;
                a = nx - inx; b = ny - iny;
                iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
                iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0};
                float ib1 = 0, ib2 = 0;
                for (int y = 0; y < win; ++y) {
                    const uint8_t* s = LJ.img + (size_t)(y + iny + LJ.pad) * LJ.stride + (inx + LJ.pad);
                    const int st = LJ.stride;
                    const int16_t* Ip = Iw + y * win;
                    const int16_t* dp = dIw + y * win * 2;
                    int diff[32];
                    for (int x = 0; x < win; ++x)
                        diff[x] = LK_DESCALE(s[x] * iw00 + s[x + 1] * iw01 + s[x + st] * iw10 + s[x + st + 1] * iw11,
                                             W_BITS - 5) - Ip[x];
                    int x = 0;
                    for (; x < simd_w; x += 8) {
                        /* lanes of qb0: (d0*Ix0+d4*Ix4, d0*Iy0+d4*Iy4, d1*Ix1+d5*Ix5, d1*Iy1+d5*Iy5)
                         * lanes of qb1: (d2*Ix2+d6*Ix6, d2*Iy2+d6*Iy6, d3*Ix3+d7*Ix7, d3*Iy3+d7*Iy7)
                         * each lane: exact int32 pair sum (pmaddwd), converted to float, then added. */
                        const int* dd = diff + x;
                        const int16_t* q = dp + 2 * x;
                        qb0[0] += (float)(dd[0] * q[0] + dd[4] * q[8]);
                        qb0[1] += (float)(dd[0] * q[1] + dd[4] * q[9]);
                        qb0[2] += (float)(dd[1] * q[2] + dd[5] * q[10]);
                        qb0[3] += (float)(dd[1] * q[3] + dd[5] * q[11]);
                        qb1[0] += (float)(dd[2] * q[4] + dd[6] * q[12]);
                        qb1[1] += (float)(dd[2] * q[5] + dd[6] * q[13]);
                        qb1[2] += (float)(dd[3] * q[6] + dd[7] * q[14]);
                        qb1[3] += (float)(dd[3] * q[7] + dd[7] * q[15]);
                    }
                    for (; x < win; ++x) {
                        ib1 += (float)(diff[x] * dp[2 * x]);
                        ib2 += (float)(diff[x] * dp[2 * x + 1]);
                    }
                }
                if (simd_w) {
                    float s0 = qb0[0] + qb1[0], s1 = qb0[1] + qb1[1], s2 = qb0[2] + qb1[2], s3 = qb0[3] + qb1[3];
                    float f0[4] = {s0, s2, 0.f, 0.f}, f1[4] = {s1, s3, 0.f, 0.f};
                    ib1 += reduce4(f0);
                    ib2 += reduce4(f1);
                }
                float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
                float dx = (float)((A12 * b2 - A22 * b1) * D);
                float dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx; ny += dy;
                next_pts[2 * p] = nx + half; next_pts[2 * p + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= eps) break;
                if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                    next_pts[2 * p] -= dx * 0.5f; next_pts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[p] && level == 0) {   /* the caller always passes an err vector (feature_tracker.cpp:40-42) */
                float fx = next_pts[2 * p] - half, fy = next_pts[2 * p + 1] - half;
                int ix = cv_floor(fx), iy = cv_floor(fy);
                if (ix < -win || ix >= jcols || iy < -win || iy >= jrows) status[p] = 0;
            }
        }
        lk_level_free(&LI); lk_level_free(&LJ);
    }
    free(Iw); free(dIw);
}

/* ------------------------------------------------------------------------------------------------
 * A6  cornerMinEigenVal (imgproc/corner.cpp) with Sobel 3x3 (scale 1/(4*3*255)), 3x3 unnormalised box
 * (RowSum<float,double> + sliding ColumnSum<double,float>), REFLECT_101, canonical non-FMA order.
 * ------------------------------------------------------------------------------------------------ */
void vfo_min_eigen_val(const uint8_t* img, int w, int h, float* eig) {
    const float k1 = (float)(1.0 / 3060.0), k0 = (float)(2.0 / 3060.0);
    size_t npx = (size_t)w * h;
    float* R = (float*)malloc(sizeof(float) * npx);   /* row pass of Dx: I[x+1]-I[x-1]          */
    float* T = (float*)malloc(sizeof(float) * npx);   /* row pass of Dy: (k1*I[x-1]+k0*I[x])+k1*I[x+1] */
    float* cov = (float*)malloc(sizeof(float) * npx * 3);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = img + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float l = (float)s[refl101(x - 1, w)], c = (float)s[x], r = (float)s[refl101(x + 1, w)];
            R[(size_t)y * w + x] = r - l;
            T[(size_t)y * w + x] = (k1 * l + k0 * c) + k1 * r;
        }
    }
    for (int y = 0; y < h; ++y) {
        int ym = refl101(y - 1, h), yp = refl101(y + 1, h);
        for (int x = 0; x < w; ++x) {
            float dx = k0 * R[(size_t)y * w + x] + k1 * (R[(size_t)ym * w + x] + R[(size_t)yp * w + x]);
            float dy = T[(size_t)yp * w + x] - T[(size_t)ym * w + x];
            float* c = cov + ((size_t)y * w + x) * 3;
            c[0] = dx * dx; c[1] = dx * dy; c[2] = dy * dy;
        }
    }
    /* box 3x3: row sums in double ((a+b)+c), column running sum in double, cast to float */
    double* rs = (double*)malloc(sizeof(double) * npx * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
            for (int k = 0; k < 3; ++k)
                rs[((size_t)y * w + x) * 3 + k] = (double)cov[((size_t)y * w + xm) * 3 + k] +
                                                  (double)cov[((size_t)y * w + x) * 3 + k] +
                                                  (double)cov[((size_t)y * w + xp) * 3 + k];
        }
    double* sum = (double*)calloc((size_t)w * 3, sizeof(double));
    /* prime with rows -1 (== row 1 by reflection) and 0 */
    for (int k = 0; k < w * 3; ++k) {
        sum[k] += rs[(size_t)refl101(-1, h) * w * 3 + k];
        sum[k] += rs[k];
    }
    for (int y = 0; y < h; ++y) {
        const double* sp = rs + (size_t)refl101(y + 1, h) * w * 3;
        const double* sm = rs + (size_t)refl101(y - 1, h) * w * 3;
        for (int x = 0; x < w; ++x) {
            float c3[3];
            for (int k = 0; k < 3; ++k) {
                double s0 = sum[x * 3 + k] + sp[x * 3 + k];
                c3[k] = (float)s0;
                sum[x * 3 + k] = s0 - sm[x * 3 + k];
            }
            float a = c3[0] * 0.5f, b = c3[1], c = c3[2] * 0.5f;
            eig[(size_t)y * w + x] = (float)((a + c) - sqrtf((a - c) * (a - c) + b * b));
        }
    }
    free(R); free(T); free(cov); free(rs); free(sum);
}

/* goodFeaturesToTrack (imgproc/featureselect.cpp). */
typedef struct { float v; int idx; } cand_t;
static int cand_cmp(const void* pa, const void* pb) {   /* greaterThanPtr: value desc, then address desc */
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->idx > b->idx) ? -1 : (a->idx < b->idx) ? 1 : 0;
}

int vfo_good_features(const uint8_t* img, int w, int h, int max_corners, double quality, int min_dist,
                      const uint8_t* mask, const float* eig_in, float* out_xy, int* n_candidates_out) {
    size_t npx = (size_t)w * h;
    float* eig = (float*)malloc(sizeof(float) * npx);
    if (eig_in) memcpy(eig, eig_in, sizeof(float) * npx); else vfo_min_eigen_val(img, w, h, eig);
    double max_val = 0;   /* minMaxLoc with mask: max over mask != 0 (stays 0 if the mask is empty) */
    int any = 0;
    for (size_t i = 0; i < npx; ++i)
        if (!mask || mask[i]) { if (!any || eig[i] > max_val) max_val = eig[i]; any = 1; }
    if (!any) max_val = 0;
    float thr = (float)(max_val * quality);
    for (size_t i = 0; i < npx; ++i) eig[i] = eig[i] > thr ? eig[i] : 0.f;   /* THRESH_TOZERO */
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * npx);
    int nc = 0;
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            size_t i = (size_t)y * w + x;
            float v = eig[i];
            if (v == 0 || (mask && !mask[i])) continue;
            float m = v;   /* dilate 3x3 (border never reached because 1 <= x,y < dim-1) */
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) { float q = eig[i + dy * w + dx]; if (q > m) m = q; }
            if (v == m) { cand[nc].v = v; cand[nc].idx = (int)i; ++nc; }
        }
    if (n_candidates_out) *n_candidates_out = nc;
    qsort(cand, (size_t)nc, sizeof(cand_t), cand_cmp);   /* total order -> unique result */
    int n_out = 0;
    const float md2 = (float)min_dist * (float)min_dist;
    for (int i = 0; i < nc; ++i) {
        int y = cand[i].idx / w, x = cand[i].idx - y * w;
        int good = 1;
        if (min_dist >= 1)
            for (int j = 0; j < n_out; ++j) {
                float dx = x - out_xy[2 * j], dy = y - out_xy[2 * j + 1];
                if (dx * dx + dy * dy < md2) { good = 0; break; }
            }
        if (good) {
            out_xy[2 * n_out] = (float)x; out_xy[2 * n_out + 1] = (float)y;
            ++n_out;
            if (max_corners > 0 && n_out == max_corners) break;
        }
    }
    free(cand); free(eig);
    return n_out;
}

/* A7  cv::circle(mask, c, r, 0, FILLED) == { (x,y) : (x-cx)^2 + (y-cy)^2 <= r^2 } (checked vs cv2 for r <= 120). */
void vfo_mask_disc(uint8_t* mask, int w, int h, int cx, int cy, int r) {
    for (int y = cy - r; y <= cy + r; ++y) {
        if (y < 0 || y >= h) continue;
        for (int x = cx - r; x <= cx + r; ++x) {
            if (x < 0 || x >= w) continue;
            if ((x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r) mask[(size_t)y * w + x] = 0;
        }
    }
}

/* A9  liftProjective (CataCamera.cc:556-626, 766-782; PinholeCamera.cc:450-510, 646-662) + /z. */
static void cam_distortion(const vfo_camera* c, double x, double y, double* dx, double* dy) {
    double mx2 = x * x, my2 = y * y, mxy = x * y;
    double rho2 = mx2 + my2;
    double rad = c->k1 * rho2 + c->k2 * rho2 * rho2;
    *dx = x * rad + 2.0 * c->p1 * mxy + c->p2 * (rho2 + 2.0 * mx2);
    *dy = y * rad + 2.0 * c->p2 * mxy + c->p1 * (rho2 + 2.0 * my2);
}

void vfo_undistort(const vfo_camera* c, double u, double v, double* xo, double* yo) {
    double i11 = 1.0 / c->fx, i13 = -c->cx / c->fx, i22 = 1.0 / c->fy, i23 = -c->cy / c->fy;
    double mx_d = i11 * u + i13, my_d = i22 * v + i23, mx_u, my_u;
    if (c->k1 == 0.0 && c->k2 == 0.0 && c->p1 == 0.0 && c->p2 == 0.0) {
        mx_u = mx_d; my_u = my_d;
    } else {
        double dx, dy;
        cam_distortion(c, mx_d, my_d, &dx, &dy);
        mx_u = mx_d - dx; my_u = my_d - dy;
        for (int i = 1; i < 8; ++i) {
            cam_distortion(c, mx_u, my_u, &dx, &dy);
            mx_u = mx_d - dx; my_u = my_d - dy;
        }
    }
    double z = 1.0;
    if (c->model == 0) {
        if (c->xi == 1.0) z = (1.0 - mx_u * mx_u - my_u * my_u) / 2.0;
        else {
            double rho2 = mx_u * mx_u + my_u * my_u;
            z = 1.0 - c->xi * (rho2 + 1.0) / (c->xi + sqrt(1.0 + (1.0 - c->xi * c->xi) * rho2));
        }
    }
    *xo = mx_u / z; *yo = my_u / z;
}

/* ------------------------------------------------------------------------------------------------
 * Whole TrackImage (feature_tracker.cpp:27-217), same control flow as oracle/cv2_oracle.py.
 * ------------------------------------------------------------------------------------------------ */
struct vfo_tracker {
    int w, h, max_cnt, min_dist, flow_back, lk_max_level, lk_win, lk_accum;
    double quality;
    vfo_camera cam[2];
    uint32_t landmark_id;
    double prev_time, cur_time;
    int has_prev;
    vfo_pyramid prev_pyr;
    int n;                         /* current number of tracked features */
    float* kps;                    /* prev/cur points, 2*max_cnt */
    int* ids; int* cnt;
    int n_prev_left; int* prev_left_ids; float* prev_left_un;
    int n_prev_right; int* prev_right_ids; float* prev_right_un;
};

vfo_tracker* vfo_tracker_create(int w, int h, int max_cnt, int min_dist, int flow_back, int lk_max_level,
                                int lk_win, double quality, const vfo_camera* cam0, const vfo_camera* cam1,
                                int lk_accum) {
    vfo_tracker* t = (vfo_tracker*)calloc(1, sizeof(*t));
    t->w = w; t->h = h; t->max_cnt = max_cnt; t->min_dist = min_dist; t->flow_back = flow_back;
    t->lk_max_level = lk_max_level; t->lk_win = lk_win; t->quality = quality; t->lk_accum = lk_accum;
    t->cam[0] = *cam0; t->cam[1] = *cam1;
    size_t cap = (size_t)(max_cnt > 0 ? max_cnt : 1);
    t->kps = (float*)calloc(cap * 2, sizeof(float));
    t->ids = (int*)calloc(cap, sizeof(int)); t->cnt = (int*)calloc(cap, sizeof(int));
    t->prev_left_ids = (int*)calloc(cap, sizeof(int)); t->prev_left_un = (float*)calloc(cap * 2, sizeof(float));
    t->prev_right_ids = (int*)calloc(cap, sizeof(int)); t->prev_right_un = (float*)calloc(cap * 2, sizeof(float));
    return t;
}

void vfo_tracker_destroy(vfo_tracker* t) {
    if (!t) return;
    vfo_pyramid_free(&t->prev_pyr);
    free(t->kps); free(t->ids); free(t->cnt);
    free(t->prev_left_ids); free(t->prev_left_un); free(t->prev_right_ids); free(t->prev_right_un);
    free(t);
}

static int in_border(float x, float y, int rows, int cols) {   /* feature_tracker.cpp:20-25 */
    int ix = cv_round(x), iy = cv_round(y);
    return 1 <= ix && ix < cols - 1 && 1 <= iy && iy < rows - 1;
}
static double pt_distance(float ax, float ay, float bx, float by) {   /* feature_tracker.cpp:457-462 */
    double dx = ax - bx, dy = ay - by;   /* float subtraction, widened */
    return sqrt(dx * dx + dy * dy);
}
static int find_id(const int* ids, int n, int id) {
    for (int i = 0; i < n; ++i) if (ids[i] == id) return i;
    return -1;
}

int vfo_tracker_track(vfo_tracker* t, double time, const uint8_t* img0, const uint8_t* img1, vfo_feature* out) {
    const int W = t->w, H = t->h, cap = t->max_cnt, win = t->lk_win;
    t->cur_time = time;
    vfo_pyramid cur, right;
    vfo_pyramid_build(&cur, img0, W, H, t->lk_max_level, win);
    vfo_pyramid_build(&right, img1, W, H, t->lk_max_level, win);
    float* cur_kps = (float*)calloc((size_t)cap * 2 + 2, sizeof(float));
    float* tmp_pts = (float*)calloc((size_t)cap * 2 + 2, sizeof(float));
    uint8_t* st = (uint8_t*)calloc((size_t)cap + 1, 1);
    uint8_t* st2 = (uint8_t*)calloc((size_t)cap + 1, 1);
    int n = 0;

    if (t->has_prev && t->n > 0) {   /* feature_tracker.cpp:37-72 */
        n = t->n;
        vfo_lk(&t->prev_pyr, &cur, t->kps, cur_kps, st, n, t->lk_max_level, win, 0, 30, 0.01, 1e-4f, t->lk_accum, NULL);
        if (t->flow_back) {
            memcpy(tmp_pts, t->kps, sizeof(float) * 2 * (size_t)n);
            vfo_lk(&cur, &t->prev_pyr, cur_kps, tmp_pts, st2, n, 1, win, 1, 30, 0.01, 1e-4f, t->lk_accum, NULL);
            for (int i = 0; i < n; ++i)
                if (!(st[i] && st2[i] &&
                      pt_distance(t->kps[2 * i], t->kps[2 * i + 1], tmp_pts[2 * i], tmp_pts[2 * i + 1]) <= 0.5))
                    st[i] = 0;
        }
        for (int i = 0; i < n; ++i)
            if (st[i] && !in_border(cur_kps[2 * i], cur_kps[2 * i + 1], H, W)) st[i] = 0;
        int j = 0;
        for (int i = 0; i < n; ++i)
            if (st[i]) {
                cur_kps[2 * j] = cur_kps[2 * i]; cur_kps[2 * j + 1] = cur_kps[2 * i + 1];
                t->ids[j] = t->ids[i]; t->cnt[j] = t->cnt[i]; ++j;
            }
        n = j;
    }
    for (int i = 0; i < n; ++i) t->cnt[i]++;   /* :76-79 */

    /* SetFeatureExtractorMask :219-246 */
    uint8_t* mask = (uint8_t*)malloc((size_t)W * H);
    memset(mask, 255, (size_t)W * H);
    {
        int* perm = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        vf_oracle_stdsort_perm(t->cnt, n, perm);
        float* k2 = (float*)malloc(sizeof(float) * 2 * (size_t)(n + 1));
        int* i2 = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        int* c2 = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        int m = 0;
        for (int q = 0; q < n; ++q) {
            int j = perm[q];
            int cx = cv_round(cur_kps[2 * j]), cy = cv_round(cur_kps[2 * j + 1]);
            if (mask[(size_t)cy * W + cx] == 255) {
                k2[2 * m] = cur_kps[2 * j]; k2[2 * m + 1] = cur_kps[2 * j + 1];
                i2[m] = t->ids[j]; c2[m] = t->cnt[j]; ++m;
                vfo_mask_disc(mask, W, H, cx, cy, t->min_dist);
            }
        }
        memcpy(cur_kps, k2, sizeof(float) * 2 * (size_t)m);
        memcpy(t->ids, i2, sizeof(int) * (size_t)m); memcpy(t->cnt, c2, sizeof(int) * (size_t)m);
        n = m;
        free(perm); free(k2); free(i2); free(c2);
    }
    int n_new = cap - n;   /* :85-101 */
    if (n_new > 0) {
        float* np = (float*)malloc(sizeof(float) * 2 * (size_t)n_new);
        int got = vfo_good_features(img0, W, H, n_new, t->quality, t->min_dist, mask, NULL, np, NULL);
        for (int i = 0; i < got; ++i) {
            cur_kps[2 * n] = np[2 * i]; cur_kps[2 * n + 1] = np[2 * i + 1];
            t->cnt[n] = 1; t->ids[n] = (int)(t->landmark_id++); ++n;
        }
        free(np);
    }
    free(mask);

    /* RemoveDistortion + ComputeKpsVelocity (left) :103-104, :310-363 */
    float* un = (float*)calloc((size_t)cap * 2 + 2, sizeof(float));
    for (int i = 0; i < n; ++i) {
        double x, y;
        vfo_undistort(&t->cam[0], (double)cur_kps[2 * i], (double)cur_kps[2 * i + 1], &x, &y);
        un[2 * i] = (float)x; un[2 * i + 1] = (float)y;
    }
    const double dt = t->cur_time - t->prev_time;
    for (int i = 0; i < n; ++i) {
        out[i].id = t->ids[i]; out[i].track_cnt = t->cnt[i]; out[i].has_right = 0;
        out[i].x = un[2 * i]; out[i].y = un[2 * i + 1]; out[i].u = cur_kps[2 * i]; out[i].v = cur_kps[2 * i + 1];
        out[i].vx = out[i].vy = 0.f;
        out[i].xr = out[i].yr = out[i].ur = out[i].vr = out[i].vxr = out[i].vyr = 0.f;
        if (t->n_prev_left > 0) {
            int j = find_id(t->prev_left_ids, t->n_prev_left, t->ids[i]);
            if (j >= 0) {
                out[i].vx = (float)((double)(un[2 * i] - t->prev_left_un[2 * j]) / dt);
                out[i].vy = (float)((double)(un[2 * i + 1] - t->prev_left_un[2 * j + 1]) / dt);
            }
        }
    }

    /* stereo :112-142 */
    int n_right = 0;
    int* right_ids = (int*)calloc((size_t)cap + 1, sizeof(int));
    float* right_un = (float*)calloc((size_t)cap * 2 + 2, sizeof(float));
    if (n > 0) {
        float* rk = (float*)calloc((size_t)cap * 2 + 2, sizeof(float));
        vfo_lk(&cur, &right, cur_kps, rk, st, n, t->lk_max_level, win, 0, 30, 0.01, 1e-4f, t->lk_accum, NULL);
        if (t->flow_back) {
            vfo_lk(&right, &cur, rk, tmp_pts, st2, n, t->lk_max_level, win, 0, 30, 0.01, 1e-4f, t->lk_accum, NULL);
            for (int i = 0; i < n; ++i)
                if (!(st[i] && st2[i] && in_border(rk[2 * i], rk[2 * i + 1], H, W) &&
                      pt_distance(cur_kps[2 * i], cur_kps[2 * i + 1], tmp_pts[2 * i], tmp_pts[2 * i + 1]) <= 0.5))
                    st[i] = 0;
        }
        for (int i = 0; i < n; ++i) {
            if (!st[i]) continue;
            double x, y;
            vfo_undistort(&t->cam[1], (double)rk[2 * i], (double)rk[2 * i + 1], &x, &y);
            float fx = (float)x, fy = (float)y;
            out[i].has_right = 1; out[i].ur = rk[2 * i]; out[i].vr = rk[2 * i + 1]; out[i].xr = fx; out[i].yr = fy;
            if (t->n_prev_right > 0) {
                int j = find_id(t->prev_right_ids, t->n_prev_right, t->ids[i]);
                if (j >= 0) {
                    out[i].vxr = (float)((double)(fx - t->prev_right_un[2 * j]) / dt);
                    out[i].vyr = (float)((double)(fy - t->prev_right_un[2 * j + 1]) / dt);
                }
            }
            right_ids[n_right] = t->ids[i]; right_un[2 * n_right] = fx; right_un[2 * n_right + 1] = fy; ++n_right;
        }
        free(rk);
    }

    /* roll :158-168 */
    vfo_pyramid_free(&t->prev_pyr);
    t->prev_pyr = cur;
    vfo_pyramid_free(&right);
    memcpy(t->kps, cur_kps, sizeof(float) * 2 * (size_t)n);
    t->n = n; t->has_prev = 1;
    memcpy(t->prev_left_ids, t->ids, sizeof(int) * (size_t)n);
    memcpy(t->prev_left_un, un, sizeof(float) * 2 * (size_t)n);
    t->n_prev_left = n;
    memcpy(t->prev_right_ids, right_ids, sizeof(int) * (size_t)n_right);
    memcpy(t->prev_right_un, right_un, sizeof(float) * 2 * (size_t)n_right);
    t->n_prev_right = n_right;
    t->prev_time = t->cur_time;
    free(cur_kps); free(tmp_pts); free(st); free(st2); free(un); free(right_ids); free(right_un);
    return n;
}
