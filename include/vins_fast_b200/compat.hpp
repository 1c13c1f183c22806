// Types the reference's FeatureTracker interface is written in (cv::Mat, cv::Point2f, Eigen::Matrix<double,7,1>).
//
// Where the real libraries are installed (the reference's own build: OpenCV 3 + Eigen 3.3.7,
// vins/cmake/packages.cmake:11, vins/thirdparty/eigen) their headers are used and the class in feature_tracker.hpp is a
// source-level drop-in.  Where they are not (this repository's build and test image has neither), the minimal
// stand-ins below provide exactly the members feature_tracker.hpp and its callers touch, so the host layer still
// compiles and is tested.  Define VF_COMPAT_FORCE_SHIMS to use the stand-ins even when the libraries exist.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// Fatal errors of the host layer mirror the reference's assert / LOG(FATAL): message to stderr and std::abort();
// define VINS_FAST_B200_THROW to get a std::runtime_error instead (used by this repository's tests).
#ifdef VINS_FAST_B200_THROW
#define VF_FATAL(msg) throw std::runtime_error(std::string(msg))
#else
#define VF_FATAL(msg) do { std::fprintf(stderr, "FATAL %s:%d: %s\n", __FILE__, __LINE__, std::string(msg).c_str()); std::abort(); } while (0)
#endif

#if !defined(VF_COMPAT_FORCE_SHIMS) && defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#define VF_HAVE_OPENCV 1
#endif
#if __has_include(<Eigen/Core>)
#define VF_HAVE_EIGEN 1
#endif
#endif

#ifdef VF_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_8UC3 16
#endif
namespace cv {

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

// Dense 8-bit matrix with cv::Mat's shallow-copy semantics (shared, reference-counted storage) or a view of
// caller-owned memory (the cv::Mat(rows, cols, type, data, step) constructor).
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;   // bytes per row

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) {
        create(r, c, type);
        const int ch = channels();
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                for (int k = 0; k < ch; ++k) data[y * step + (size_t)x * ch + k] = (unsigned char)s.val[k];
    }
    Mat(int r, int c, int type, void* ext, size_t ext_step = 0) : rows(r), cols(c), data((unsigned char*)ext), type_(type) {
        step = ext_step ? ext_step : (size_t)c * channels();
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * channels();
        store_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step);
        data = store_->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    bool isContinuous() const { return step == (size_t)cols * channels(); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + y * m.step, data + y * step, (size_t)cols * channels());
        return m;
    }
    template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }

private:
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<unsigned char>> store_;
};

}  // namespace cv

// cvRound of OpenCV: round half to even (SSE cvtsd2si / lrint in the default rounding mode)
static inline int cvRound(double v) { return (int)std::lrint(v); }
static inline int cvRound(float v) { return (int)std::lrintf(v); }
#endif  // VF_HAVE_OPENCV

#ifdef VF_HAVE_EIGEN
#include <Eigen/Core>
#else
namespace Eigen {
// Fixed-size column vector: the members the estimator uses on a featureFrame entry (operator(), operator[],
// the comma initialiser of feature_tracker.cpp:189, rows()).
template <typename T, int R, int C>
class Matrix {
    static_assert(C == 1, "the stand-in covers column vectors only");
public:
    Matrix() { for (int i = 0; i < R; ++i) v_[i] = T(0); }
    T& operator()(int i) { return v_[i]; }
    const T& operator()(int i) const { return v_[i]; }
    T& operator[](int i) { return v_[i]; }
    const T& operator[](int i) const { return v_[i]; }
    static int rows() { return R; }
    static int cols() { return 1; }
    const T* data() const { return v_; }
    struct CommaInit {
        Matrix& m; int i;
        CommaInit& operator,(T x) { m.v_[i++] = x; return *this; }
    };
    CommaInit operator<<(T x) { v_[0] = x; return CommaInit{*this, 1}; }
private:
    T v_[R];
};
}  // namespace Eigen
#endif  // VF_HAVE_EIGEN
