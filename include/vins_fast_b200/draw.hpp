// Host-side rendering of the track image, pixel-identical to the OpenCV drawing calls of the reference's
// FeatureTracker::DrawTracker (vins/estimator/feature_tracker.cpp:267-299):
//     cv::hconcat(left, right) ; cv::cvtColor(GRAY2RGB)
//     cv::circle(img, Point2f, 2, Scalar(255 (1 - len), 0, 255 len), 2)      per left feature,  len = min(1, track_count / 20)
//     cv::circle(img, Point2f + (cols, 0), 2, Scalar(0, 255, 0), 2)           per right feature
//     cv::arrowedLine(img, cur, prev, Scalar(0, 255, 0), 1, 8, 0, 0.2)        per feature found in the previous frame
// OpenCV is not vendored in the reference and not installed here, so the rasterisation rules are restated (and checked
// against the cv2 wheel by tests/test_host.py::test_draw_tracker_is_pixel_identical_to_opencv):
//   * Point2f -> Point is cvRound per coordinate (saturate_cast<int>), Scalar -> uchar is cvRound + clamp;
//   * circle(radius 2, thickness 2) with its centre inside the image is the disc dx^2 + dy^2 <= 10, clipped per pixel;
//   * line(thickness 1, LINE_8, shift 0) is cv::LineIterator with leftToRight = true: the end point with the smaller x is
//     the start, err = dx - 2 dy along the major axis, a diagonal step whenever err < 0; a segment with an end point outside
//     the image is first cut by cv::clipLine (truncating divisions) and the iterator then connects the CUT end points;
//   * arrowedLine = the shaft plus two tip segments from cvRound(pt2 + tip (cos, sin)(angle +- pi/4)), tip = 0.2 |pt1 - pt2|,
//     angle = atan2(pt1.y - pt2.y, pt1.x - pt2.x) on the ROUNDED end points.
// No dependency on OpenCV or on the CUDA library: plain C++11 over an interleaved RGB buffer.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstring>

namespace vf_draw {

struct Rgb { unsigned char r, g, b; };

inline int RoundHalfEven(double v) { return (int)std::lrint(v); }          // cvRound
inline unsigned char SaturateU8(double v) { const int i = RoundHalfEven(v); return (unsigned char)(i < 0 ? 0 : i > 255 ? 255 : i); }

struct Canvas {
    unsigned char* data; int width, height; size_t step;                    // 3 bytes per pixel, `step` bytes per row
    void Put(int x, int y, Rgb c) const {
        if (x < 0 || y < 0 || x >= width || y >= height) return;
        unsigned char* p = data + (size_t)y * step + 3 * (size_t)x;
        p[0] = c.r; p[1] = c.g; p[2] = c.b;
    }
};

// hconcat + GRAY2RGB: left | right, every grey value replicated into the three channels
inline void SideBySideGrayToRgb(const unsigned char* left, size_t lstep, const unsigned char* right, size_t rstep, int w, int h,
                                const Canvas& out) {
    for (int y = 0; y < h; ++y) {
        unsigned char* d = out.data + (size_t)y * out.step;
        const unsigned char* a = left + (size_t)y * lstep;
        const unsigned char* b = right + (size_t)y * rstep;
        for (int x = 0; x < w; ++x) { d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = a[x]; }
        d += 3 * (size_t)w;
        for (int x = 0; x < w; ++x) { d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = b[x]; }
    }
}

// cv::circle(img, (fx, fy), 2, colour, 2)
inline void CircleR2T2(const Canvas& img, float fx, float fy, Rgb c) {
    const int cx = (int)std::lrintf(fx), cy = (int)std::lrintf(fy);
    for (int dy = -3; dy <= 3; ++dy)
        for (int dx = -3; dx <= 3; ++dx)
            if (dx * dx + dy * dy <= 10) img.Put(cx + dx, cy + dy, c);
}

// cv::clipLine(Size(w, h), pt1, pt2): Cohen-Sutherland on the end points with OpenCV's truncating divisions (the clipped
// end points, not the pixels, are what the line iterator then connects).  Returns false when nothing is left.
inline bool ClipLine(long long w, long long h, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (x2 - x1) / (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (x2 - x1) / (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// cv::line(img, p1, p2, colour, 1, LINE_8, 0) on integer end points
inline void Line8(const Canvas& img, int x1, int y1, int x2, int y2, Rgb c) {
    if ((unsigned)x1 >= (unsigned)img.width || (unsigned)x2 >= (unsigned)img.width || (unsigned)y1 >= (unsigned)img.height ||
        (unsigned)y2 >= (unsigned)img.height) {                              // LineIterator clips first, then runs Bresenham
        long long a = x1, b = y1, cc = x2, d = y2;
        if (!ClipLine(img.width, img.height, a, b, cc, d)) return;
        x1 = (int)a; y1 = (int)b; x2 = (int)cc; y2 = (int)d;
    }
    int dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }                   // left to right
    int ystep = 1;
    if (dy < 0) { dy = -dy; ystep = -1; }
    const bool steep = dy > dx;
    if (steep) { const int t = dx; dx = dy; dy = t; }
    int err = dx - (dy + dy);
    const int plus = dx + dx, minus = -(dy + dy);
    int x = x1, y = y1;
    for (int i = 0; i <= dx; ++i) {
        img.Put(x, y, c);
        const bool diag = err < 0;
        err += minus + (diag ? plus : 0);
        if (steep) { y += ystep; if (diag) x += 1; }
        else { x += 1; if (diag) y += ystep; }
    }
}

// cv::arrowedLine(img, (x1, y1), (x2, y2), colour, 1, 8, 0, tip_length) -- the arrow head sits at (x2, y2)
inline void ArrowedLine(const Canvas& img, float fx1, float fy1, float fx2, float fy2, Rgb c, double tip_length = 0.2) {
    const int x1 = (int)std::lrintf(fx1), y1 = (int)std::lrintf(fy1), x2 = (int)std::lrintf(fx2), y2 = (int)std::lrintf(fy2);
    const double ddx = (double)x1 - x2, ddy = (double)y1 - y2;
    const double tip = std::sqrt(ddx * ddx + ddy * ddy) * tip_length;
    Line8(img, x1, y1, x2, y2, c);
    const double angle = std::atan2(ddy, ddx), quarter = 3.1415926535897932384626433832795 / 4;
    for (int s = 0; s < 2; ++s) {
        const double a = s == 0 ? angle + quarter : angle - quarter;
        Line8(img, RoundHalfEven(x2 + tip * std::cos(a)), RoundHalfEven(y2 + tip * std::sin(a)), x2, y2, c);
    }
}

// Colour of a left feature by track age (feature_tracker.cpp:276-279): blue = new, red = tracked for >= 20 frames
inline Rgb TrackAgeColour(int track_count) {
    const double len = std::fmin(1.0, 1.0 * track_count / 20);
    Rgb c = {SaturateU8(255 * (1 - len)), 0, SaturateU8(255 * len)};
    return c;
}

// The whole DrawTracker: `out` is h x 2w RGB.  kps / right_kps: (x, y) pairs; prev_xy[2 i .. 2 i + 1] is the previous
// position of left feature i where has_prev[i] != 0 (previous_left_ids_kps_map_.find(ids_[i])).
inline void DrawTrackImage(const unsigned char* left, size_t lstep, const unsigned char* right, size_t rstep, int w, int h,
                           const float* kps, const int* track_count, int n, const float* right_kps, int n_right,
                           const float* prev_xy, const unsigned char* has_prev, const Canvas& out) {
    SideBySideGrayToRgb(left, lstep, right, rstep, w, h, out);
    for (int j = 0; j < n; ++j) CircleR2T2(out, kps[2 * j], kps[2 * j + 1], TrackAgeColour(track_count[j]));
    const Rgb green = {0, 255, 0};
    const float cols = (float)w;
    for (int j = 0; j < n_right; ++j) CircleR2T2(out, right_kps[2 * j] + cols, right_kps[2 * j + 1], green);   // right_kp.x += cols
    for (int i = 0; i < n; ++i)
        if (has_prev[i]) ArrowedLine(out, kps[2 * i], kps[2 * i + 1], prev_xy[2 * i], prev_xy[2 * i + 1], green, 0.2);
}

}  // namespace vf_draw
