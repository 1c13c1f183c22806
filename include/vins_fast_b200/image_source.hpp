// Input side of the front-end without ROS / cv_bridge:
//   * common::ImageMsg + GetImageFromMsg / CopyMono8 -- what common::GetImageFromROSMsg does with a sensor_msgs::Image
//     (vins/common/utils.hpp:17-37): "8UC1" is relabelled mono8, the pixels are copied into a CONTINUOUS CV_8UC1 image
//     (step == cols), here optionally straight into a page-locked slot so the H2D DMA of the frame is asynchronous;
//   * estimator::StereoSync -- the image queues and the 3 ms pairing rule of Estimator::tFrontendProcess
//     (vins/estimator/estimator.cpp:103-141): drop the older head while the stamps differ by more than 0.003 s;
//   * estimator::PinnedImageRing -- page-locked stereo slots (vf_host_alloc) that FrontendPipeline uploads from;
//   * common::ReadPgm / WritePgm / EurocImageList -- raw (binary PGM, "P5") readers and the EuRoC `data.csv` listing
//     (timestamp [ns], filename), so recorded sequences can be replayed without a PNG decoder.
// Host code only; C++11.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../vins_fast_b200.h"
#include "compat.hpp"

namespace common {

/// The fields of sensor_msgs::Image that GetImageFromROSMsg reads (utils.hpp:17-37)
struct ImageMsg {
    double stamp = 0;                 // header.stamp.toSec()
    int width = 0, height = 0;
    size_t step = 0;                  // bytes per row on the wire
    std::string encoding;             // "mono8" or "8UC1" (relabelled mono8, utils.hpp:20-31)
    const unsigned char* data = nullptr;
};

inline bool IsMono8(const ImageMsg& m) { return m.encoding == "mono8" || m.encoding == "8UC1"; }

/// Copies the message's pixels into dst (rows of `dst_step` bytes): the toCvCopy(MONO8) + clone() of the reference in one pass
inline void CopyMono8(const ImageMsg& m, unsigned char* dst, size_t dst_step) {
    if (!m.data || m.width <= 0 || m.height <= 0 || m.step < (size_t)m.width) VF_FATAL("CopyMono8: empty or malformed image message");
    if (!IsMono8(m)) VF_FATAL("CopyMono8: encoding '" + m.encoding + "' is not mono8 / 8UC1 (the front-end consumes MONO8 only)");
    if (m.step == (size_t)m.width && dst_step == (size_t)m.width) { std::memcpy(dst, m.data, (size_t)m.width * m.height); return; }
    for (int y = 0; y < m.height; ++y) std::memcpy(dst + (size_t)y * dst_step, m.data + (size_t)y * m.step, (size_t)m.width);
}

/// common::GetImageFromROSMsg: a continuous CV_8UC1 image owning its pixels
inline cv::Mat GetImageFromMsg(const ImageMsg& m) {
    cv::Mat img(m.height, m.width, CV_8UC1);
    CopyMono8(m, img.data, (size_t)img.step);
    return img;
}

/// Binary PGM ("P5", maxval 255) -> continuous CV_8UC1; false when the file is missing or not such a PGM
inline bool ReadPgm(const std::string& path, cv::Mat& out) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    int w = 0, h = 0, maxv = 0, field = 0, c;
    char magic[3] = {0, 0, 0};
    bool ok = std::fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '5';
    int* dst[3] = {&w, &h, &maxv};
    while (ok && field < 3) {                                   // header integers, '#' comments allowed between them
        c = std::fgetc(f);
        if (c == EOF) { ok = false; break; }
        if (c == '#') { while ((c = std::fgetc(f)) != EOF && c != '\n') {} continue; }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
        if (c < '0' || c > '9') { ok = false; break; }
        int v = 0;
        while (c >= '0' && c <= '9') { v = v * 10 + (c - '0'); c = std::fgetc(f); }
        *dst[field++] = v;                                      // the single whitespace after maxval was consumed by the loop
    }
    ok = ok && w > 0 && h > 0 && maxv == 255;
    if (ok) {
        out = cv::Mat(h, w, CV_8UC1);
        ok = std::fread(out.data, 1, (size_t)w * h, f) == (size_t)w * h;
    }
    std::fclose(f);
    return ok;
}

inline bool WritePgm(const std::string& path, const cv::Mat& img) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P5\n%d %d\n255\n", img.cols, img.rows);
    bool ok = true;
    for (int y = 0; y < img.rows && ok; ++y) ok = std::fwrite(img.ptr<unsigned char>(y), 1, (size_t)img.cols, f) == (size_t)img.cols;
    std::fclose(f);
    return ok;
}

/// EuRoC ASL listing `mav0/camN/data.csv`: "#timestamp [ns],filename" then one "<ns>,<file>" per image
inline bool EurocImageList(const std::string& csv_path, std::vector<std::pair<double, std::string>>& out) {
    std::ifstream in(csv_path.c_str());
    if (!in.is_open()) return false;
    std::string line;
    out.clear();
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        const size_t comma = line.find(',');
        if (comma == std::string::npos) return false;
        const long long ns = std::atoll(line.substr(0, comma).c_str());
        out.emplace_back((double)ns * 1e-9, line.substr(comma + 1));
    }
    return true;
}

}  // namespace common

namespace estimator {

/// img0_buf_ / img1_buf_ and the pairing rule of Estimator::tFrontendProcess (estimator.cpp:103-141)
class StereoSync {
public:
    struct Item { double stamp; cv::Mat image; };
    void Image0Callback(double stamp, const cv::Mat& img) { img0_buf_.push_back(Item{stamp, img}); }
    void Image1Callback(double stamp, const cv::Mat& img) { img1_buf_.push_back(Item{stamp, img}); }
    /// One pass of the loop body: true and a synchronised pair, or false (a head was thrown away, or a queue is empty)
    bool Next(double& time, cv::Mat& image0, cv::Mat& image1) {
        if (img0_buf_.empty() || img1_buf_.empty()) return false;
        const double time0 = img0_buf_.front().stamp, time1 = img1_buf_.front().stamp;
        if (time0 < time1 - 0.003) { img0_buf_.pop_front(); ++thrown0_; return false; }      // "Throw img0"
        if (time0 > time1 + 0.003) { img1_buf_.pop_front(); ++thrown1_; return false; }      // "Throw img1"
        time = time0;
        image0 = img0_buf_.front().image; img0_buf_.pop_front();
        image1 = img1_buf_.front().image; img1_buf_.pop_front();
        return true;
    }
    bool Idle() const { return img0_buf_.empty() || img1_buf_.empty(); }
    int thrown0() const { return thrown0_; }
    int thrown1() const { return thrown1_; }
private:
    std::deque<Item> img0_buf_, img1_buf_;
    int thrown0_ = 0, thrown1_ = 0;
};

/// `slots` stereo pairs of page-locked memory; slot k % slots belongs to frame k.  cv::Mat views, no ownership transfer.
class PinnedImageRing {
public:
    PinnedImageRing() {}
    ~PinnedImageRing() { Release(); }
    PinnedImageRing(const PinnedImageRing&) = delete;
    PinnedImageRing& operator=(const PinnedImageRing&) = delete;
    void Ensure(int width, int height, int slots) {
        if (base_ && width == w_ && height == h_ && slots == slots_) return;
        Release();
        w_ = width; h_ = height; slots_ = slots;
        void* p = nullptr;
        if (vf_host_alloc((size_t)2 * slots * width * height, &p) != VF_OK) VF_FATAL(std::string("vf_host_alloc: ") + vf_last_global_error());
        base_ = (unsigned char*)p;
    }
    cv::Mat Left(long frame) { return cv::Mat(h_, w_, CV_8UC1, Slot(frame, 0)); }
    cv::Mat Right(long frame) { return cv::Mat(h_, w_, CV_8UC1, Slot(frame, 1)); }
    unsigned char* Slot(long frame, int cam) { return base_ + ((size_t)(frame % slots_) * 2 + cam) * (size_t)w_ * h_; }
    int slots() const { return slots_; }
private:
    void Release() { if (base_) vf_host_free(base_); base_ = nullptr; }
    unsigned char* base_ = nullptr;
    int w_ = 0, h_ = 0, slots_ = 0;
};

}  // namespace estimator
