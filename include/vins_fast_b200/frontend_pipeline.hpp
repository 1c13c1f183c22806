// estimator::FrontendPipeline -- the front-end half of the reference's Estimator, pipelined on the GPU.
//
// Mirrors Estimator::FrontendTracker and its hand-off to the back-end (vins/estimator/estimator.cpp:143-158):
//     input_image_cnt_++;  frame_feature = feature_tracker_->TrackImage(t, img0, img1);
//     if (input_image_cnt_ % 2 == 0) feature_buf_.push(make_pair(t, frame_feature));          // under feature_buf_mutex_
// with the one difference the GPU makes worthwhile: Push(t, img0, img1) only ENQUEUES frame k (vf_batch_enqueue) and
// collects frame k-depth (vf_batch_fetch_lagged), so the stereo pass and the result hand-over of a frame overlap the
// temporal chain of the next one.  The back-end thread drains the queue with Pop() exactly like it drains feature_buf_
// (estimator.cpp:179-200): it sees the same (t, featureFrame) pairs, `depth` frames later (1 by default; 2 keeps the GPU
// busy back to back, for replaying recorded sequences).  Results are bit-identical to the synchronous
// FeatureTracker::TrackImage.
//
// Input side: Push copies the two images into a ring of page-locked slots and uploads from there, so
// the H2D copy is a true asynchronous DMA whatever memory the caller's cv::Mat lives in (a cudaMemcpyAsync from pageable
// memory is staged synchronously by the driver and would serialise the "asynchronous" enqueue), and the caller's buffers
// are free again when Push returns -- the same ownership the reference gets from GetImageFromROSMsg's clone()
// (vins/common/utils.hpp:17-37).  Sources that can write in place (a ROS callback, a file reader) skip the copy:
// AcquireLeft() / AcquireRight() hand out the slot of the next frame as cv::Mat views, PushAcquired(t) starts it.
#pragma once

#include <mutex>
#include <queue>
#include <string>
#include <utility>
#include <vector>

#include "feature_tracker.hpp"
#include "image_source.hpp"

namespace estimator {

class FrontendPipeline {
public:
    typedef FeatureTracker::FeatureFrame FeatureFrame;

    /// every_nth: which results go to the back-end queue (the reference keeps every 2nd, estimator.cpp:153)
    FrontendPipeline(int max_cnt, int min_dist, int flow_back, const vf_camera& left, const vf_camera& right, int every_nth = 2,
                     int device = -1, int depth = 1)
        : every_nth_(every_nth < 1 ? 1 : every_nth), depth_(depth < 1 ? 1 : (depth > 2 ? 2 : depth)) {
        vf_config_default(&cfg_);
        cfg_.max_cnt = max_cnt; cfg_.min_dist = min_dist; cfg_.flow_back = flow_back; cfg_.device = device;
        cfg_.cam[0] = left; cfg_.cam[1] = right;
    }
    ~FrontendPipeline() { vf_batch_destroy(batch_); }
    FrontendPipeline(const FrontendPipeline&) = delete;
    FrontendPipeline& operator=(const FrontendPipeline&) = delete;

    /// Estimator::FrontendTracker(t, img0, img1), asynchronous: starts frame k, hands frame k-depth to the queue.
    /// The pixels are copied into the pinned slot of frame k; img0 / img1 may be released as soon as this returns.
    void Push(double t, const cv::Mat& img0, const cv::Mat& img1) {
        if (img0.empty() || img1.empty()) VF_FATAL("FrontendPipeline::Push: !img0.empty() && !img1.empty()");
        if (img0.type() != CV_8UC1 || img1.type() != CV_8UC1 || img0.rows != img1.rows || img0.cols != img1.cols)
            VF_FATAL("FrontendPipeline::Push: both images must be CV_8UC1 of one size");
        Ensure(img0.cols, img0.rows);
        CopyRows(img0, ring_.Slot(input_image_cnt_, 0));
        CopyRows(img1, ring_.Slot(input_image_cnt_, 1));
        PushAcquired(t);
    }
    /// Same from image messages (common::GetImageFromROSMsg, utils.hpp:17-37): mono8 / 8UC1 pixels straight into the slot
    void Push(const common::ImageMsg& msg0, const common::ImageMsg& msg1) {
        if (msg0.width != msg1.width || msg0.height != msg1.height) VF_FATAL("FrontendPipeline::Push: the two messages differ in size");
        Ensure(msg0.width, msg0.height);
        common::CopyMono8(msg0, ring_.Slot(input_image_cnt_, 0), (size_t)msg0.width);
        common::CopyMono8(msg1, ring_.Slot(input_image_cnt_, 1), (size_t)msg1.width);
        PushAcquired(msg0.stamp);
    }
    /// Zero-copy ingestion: continuous CV_8UC1 views of the NEXT frame's page-locked slot (valid until PushAcquired)
    cv::Mat AcquireLeft(int width, int height) { Ensure(width, height); return ring_.Left(input_image_cnt_); }
    cv::Mat AcquireRight(int width, int height) { Ensure(width, height); return ring_.Right(input_image_cnt_); }
    void PushAcquired(double t) {
        if (!batch_) VF_FATAL("FrontendPipeline::PushAcquired: no slot has been acquired");
        times_[input_image_cnt_ % kSlots] = t;
        const uint8_t* p0 = ring_.Slot(input_image_cnt_, 0);
        const uint8_t* p1 = ring_.Slot(input_image_cnt_, 1);
        if (vf_batch_enqueue(batch_, &t, &p0, (size_t)cfg_.width, &p1, (size_t)cfg_.width) != VF_OK)
            VF_FATAL(std::string("vf_batch_enqueue: ") + vf_batch_last_error(batch_));
        input_image_cnt_++;
        if (input_image_cnt_ > depth_) Collect(depth_);
    }

    /// collects the frames that are still in flight (call when the image source has ended)
    void Flush() {
        while (input_image_cnt_ > collected_) Collect((int)(input_image_cnt_ - collected_ - 1));
    }

    /// back-end side: feature_buf_.front() / pop() (estimator.cpp:179-200); false when the queue is empty
    bool Pop(std::pair<double, FeatureFrame>& out) {
        std::unique_lock<std::mutex> lck(feature_buf_mutex_);
        if (feature_buf_.empty()) return false;
        out = std::move(feature_buf_.front());
        feature_buf_.pop();
        return true;
    }
    size_t Pending() {
        std::unique_lock<std::mutex> lck(feature_buf_mutex_);
        return feature_buf_.size();
    }
    long FramesTracked() const { return collected_; }

private:
    void Ensure(int width, int height) {
        if (batch_ && cfg_.width == width && cfg_.height == height) return;
        vf_batch_destroy(batch_);
        batch_ = nullptr;
        cfg_.width = width; cfg_.height = height;
        if (vf_batch_create(&cfg_, 1, &batch_) != VF_OK) VF_FATAL(std::string("vf_batch_create: ") + vf_last_global_error());
        records_.resize((size_t)(cfg_.max_cnt > 0 ? cfg_.max_cnt : 1));
        input_image_cnt_ = 0; collected_ = 0;
        // slot k % 4 is rewritten by frame k + 4: by then frame k has been collected (depth <= 2), so its upload has long completed
        ring_.Ensure(width, height, kSlots);
    }
    static void CopyRows(const cv::Mat& m, unsigned char* dst) {
        if (m.isContinuous()) { std::memcpy(dst, m.data, (size_t)m.cols * m.rows); return; }
        for (int y = 0; y < m.rows; ++y) std::memcpy(dst + (size_t)y * m.cols, m.ptr<unsigned char>(y), (size_t)m.cols);
    }
    // featureFrame of the frame `lag` behind the last enqueued one (feature_tracker.cpp:173-213) -> queue
    void Collect(int lag) {
        int32_t n = 0;
        if (vf_batch_fetch_lagged(batch_, lag, records_.data(), &n) != VF_OK)
            VF_FATAL(std::string("vf_batch_fetch_lagged: ") + vf_batch_last_error(batch_));
        const long index = input_image_cnt_ - 1 - lag;          // 0-based index of the collected frame
        const int slot = (int)(index % kSlots);
        collected_ = index + 1;
        if (collected_ % every_nth_ == 0) {                       // input_image_cnt_ % 2 == 0 in the reference (1-based count)
            FeatureFrame ff = FeatureTracker::ToFeatureFrame(records_.data(), n);
            std::unique_lock<std::mutex> lck(feature_buf_mutex_);
            feature_buf_.push(std::make_pair(times_[slot], std::move(ff)));
        }
    }

    vf_config cfg_;
    vf_batch* batch_ = nullptr;
    int every_nth_, depth_;
    long input_image_cnt_ = 0, collected_ = 0;
    std::vector<vf_feature> records_;
    static constexpr int kSlots = 4;
    PinnedImageRing ring_;
    double times_[kSlots] = {0, 0, 0, 0};
    std::mutex feature_buf_mutex_;
    std::queue<std::pair<double, FeatureFrame>> feature_buf_;
};

}  // namespace estimator
