// estimator::FeatureTracker -- drop-in for the reference's visual front-end class
// (vins/estimator/feature_tracker.hpp:18-107, feature_tracker.cpp), header-only over the C ABI of
// libvinsfast_b200.so (include/vins_fast_b200.h).  The arithmetic of TrackImage runs in hand-written sm_100a kernels;
// this class only marshals: it copies nothing but the two images down and the <= max_cnt records back, and rebuilds
// the reference's member vectors and the returned featureFrame map from those records.
//
// Same public names, argument meaning and fields as the reference:
//   FeatureTracker()                              reads max_cnt / flow_back / min_dist / show_track / frontend_wait_key
//                                                 from common::Setting                     (feature_tracker.cpp:8-18)
//   ReadIntrinsicParameter({cam0.yaml, cam1.yaml})                                         (feature_tracker.cpp:301-308)
//   TrackImage(t, img0, img1) -> map<int, vector<pair<int, Matrix<double,7,1>>>>           (feature_tracker.cpp:27-217)
//   InBorder, PtDistance, ReduceVector, RemoveDistortion, ComputeKpsVelocity, SetFeatureExtractorMask,
//   getDrawTrackResultImg, getCurrentImage, getPreviousImage, public OPTICAL_FLOW_BACK / MAX_CNT_FEATURES_PER_FRAME /
//   MIN_DIST / SHOW_TRACK_RESULT / FRONTEND_WAIT_KEY / draw_track_img_mutex_.
// Differences, all forced by the boundary:
//   * cameras are vf_camera descriptions instead of camodocal::CameraPtr (RemoveDistortion takes a camera index);
//   * a failure inside the library is fatal like the reference's assert / LOG(FATAL) (VF_FATAL: abort, or an exception
//     under VINS_FAST_B200_THROW); there is no CPU fallback;
//   * the image window of DrawTracker (cv::imshow when frontend_wait_key == 0) is not reproduced; the track image is
//     rendered on the host into draw_track_img_result_ only when show_track != 0 (draw.hpp: pixel-identical to the
//     reference's cv::circle / cv::arrowedLine calls).
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../vins_fast_b200.h"
#include "compat.hpp"
#include "draw.hpp"
#ifndef VF_USE_EXTERNAL_SETTING   // define it when the including project brings its own common::Setting (vins/common/parameter.hpp)
#include "parameter.hpp"
#endif

namespace estimator {

class FeatureTracker {
public:
    typedef std::map<int, std::vector<std::pair<int, Eigen::Matrix<double, 7, 1>>>> FeatureFrame;

    /// constructor: tracker keys from common::Setting (feature_tracker.cpp:8-18)
    FeatureTracker() {
        MAX_CNT_FEATURES_PER_FRAME = common::Setting::getSingleton()->Get<int>("max_cnt");
        OPTICAL_FLOW_BACK = common::Setting::getSingleton()->Get<int>("flow_back");
        MIN_DIST = common::Setting::getSingleton()->Get<int>("min_dist");
        SHOW_TRACK_RESULT = common::Setting::getSingleton()->Get<int>("show_track");
        FRONTEND_WAIT_KEY = common::Setting::getSingleton()->Get<int>("frontend_wait_key");
        vf_config_default(&cfg_);
    }
    /// for callers that do not go through the YAML singleton (tests, embedding)
    FeatureTracker(int max_cnt, int min_dist, int flow_back, int show_track = 0) {
        MAX_CNT_FEATURES_PER_FRAME = max_cnt; MIN_DIST = min_dist; OPTICAL_FLOW_BACK = flow_back;
        SHOW_TRACK_RESULT = show_track; FRONTEND_WAIT_KEY = 1;
        vf_config_default(&cfg_);
    }
    ~FeatureTracker() { vf_tracker_destroy(handle_); }
    FeatureTracker(const FeatureTracker&) = delete;
    FeatureTracker& operator=(const FeatureTracker&) = delete;

    /// CUDA device ordinal (-1 = current), LK pyramid depth (the reference's literal 3) and LK window width (its literal
    /// cv::Size(21, 21), feature_tracker.cpp:42,51,119,127; any odd width 5..31); call before the first frame
    void SetDevice(int device) { cfg_.device = device; Release(); }
    void SetLkMaxLevel(int level) { cfg_.lk_max_level = level; Release(); }
    void SetLkWindow(int win) { cfg_.lk_win = win; Release(); }

    /// Read camera intrinsic param: index 0 = left, 1 = right (feature_tracker.cpp:301-308)
    void ReadIntrinsicParameter(const std::vector<std::string>& calib_file_path) {
        for (const auto& path : calib_file_path) {
            vf_camera cam;
            if (vf_camera_load_yaml(path.c_str(), &cam) != VF_OK) VF_FATAL(std::string("ReadIntrinsicParameter: ") + vf_last_global_error());
            cameras_.push_back(cam);
        }
        Release();
    }
    void SetCameras(const vf_camera& left, const vf_camera& right) { cameras_.clear(); cameras_.push_back(left); cameras_.push_back(right); Release(); }

    /// featureFrame of one frame's records (feature_tracker.cpp:173-213): all left observations in ids_ order, then the
    /// right ones, so every id holds [(0, left)] or [(0, left), (1, right)] -- what feature_manager.cpp:36-43 asserts
    static FeatureFrame ToFeatureFrame(const vf_feature* records, int n) {
        FeatureFrame feature_frame;
        for (int i = 0; i < n; ++i) {                                        // :175-193
            const vf_feature& r = records[i];
            Eigen::Matrix<double, 7, 1> xyz_uv_velocity;
            xyz_uv_velocity << (double)r.x, (double)r.y, 1.0, (double)r.u, (double)r.v, (double)r.vx, (double)r.vy;
            feature_frame[r.id].emplace_back(0, xyz_uv_velocity);
        }
        for (int i = 0; i < n; ++i) {                                        // :195-213
            const vf_feature& r = records[i];
            if (!r.has_right) continue;
            Eigen::Matrix<double, 7, 1> xyz_uv_velocity;
            xyz_uv_velocity << (double)r.xr, (double)r.yr, 1.0, (double)r.ur, (double)r.vr, (double)r.vxr, (double)r.vyr;
            feature_frame[r.id].emplace_back(1, xyz_uv_velocity);
        }
        return feature_frame;
    }

    /// Main process image fun (feature_tracker.cpp:27-217)
    FeatureFrame TrackImage(double current_time, const cv::Mat& img0, const cv::Mat& img1 = cv::Mat()) {
        if (img0.empty() || img1.empty()) VF_FATAL("TrackImage: !img0.empty() && !img1.empty()");   // assert, :30
        if (img0.type() != CV_8UC1 || img1.type() != CV_8UC1 || img0.rows != img1.rows || img0.cols != img1.cols)
            VF_FATAL("TrackImage: both images must be CV_8UC1 of one size");
        if (cameras_.size() < 2) VF_FATAL("TrackImage: ReadIntrinsicParameter needs two cameras (stereo)");
        Ensure(img0.cols, img0.rows);
        current_time_ = current_time;
        int32_t n = 0;
        if (vf_tracker_track(handle_, current_time, img0.data, img0.step, img1.data, img1.step, records_.data(), &n) != VF_OK)
            VF_FATAL(std::string("vf_tracker_track: ") + vf_last_error(handle_));
        // the reference's member vectors, rebuilt from the records (ids_ order; right observations in ids_ order)
        previous_left_ids_kps_map_.clear();
        for (size_t i = 0; i < current_kps_.size(); ++i) previous_left_ids_kps_map_[ids_[i]] = current_kps_[i];   // :158-162 of the previous call
        ids_.clear(); track_count_.clear(); current_kps_.clear(); current_un_distortion_kps_.clear(); pts_velocity.clear();
        ids_right_.clear(); current_right_img_kps_.clear(); current_right_un_distortion_kps_.clear(); right_pts_velocity.clear();
        for (int i = 0; i < n; ++i) {
            const vf_feature& r = records_[i];
            ids_.push_back(r.id); track_count_.push_back(r.track_cnt);
            current_kps_.emplace_back(r.u, r.v); current_un_distortion_kps_.emplace_back(r.x, r.y); pts_velocity.emplace_back(r.vx, r.vy);
            if (!r.has_right) continue;
            ids_right_.push_back(r.id);
            current_right_img_kps_.emplace_back(r.ur, r.vr); current_right_un_distortion_kps_.emplace_back(r.xr, r.yr);
            right_pts_velocity.emplace_back(r.vxr, r.vyr);
        }
        FeatureFrame feature_frame = ToFeatureFrame(records_.data(), n);
        previous_img_ = current_img_;
        current_img_ = img0;           // shallow, like the reference (:32); the device copy is what the next frame tracks from
        current_right_img_ = img1;
        if (SHOW_TRACK_RESULT) {
            std::unique_lock<std::mutex> lck(draw_track_img_mutex_);
            DrawTracker();
        }
        previous_time_ = current_time_;
        return feature_frame;
    }

    /// Check the pts in image (feature_tracker.cpp:20-25)
    bool InBorder(const cv::Point2f& pt, int row, int col) {
        const int img_x = cvRound(pt.x), img_y = cvRound(pt.y);
        return 1 <= img_x && img_x < col - 1 && 1 <= img_y && img_y < row - 1;
    }

    /// Remove the distortion (feature_tracker.cpp:310-321) with camera `cam` (0 = left, 1 = right), on the device
    std::vector<cv::Point2f> RemoveDistortion(std::vector<cv::Point2f>& src_pts, int cam) {
        std::vector<cv::Point2f> out(src_pts.size());
        if (src_pts.empty()) return out;
        if (cam < 0 || cam >= (int)cameras_.size()) VF_FATAL("RemoveDistortion: camera index out of range");
        static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "Point2f must be two packed floats");
        if (vf_op_undistort(&cameras_[cam], &src_pts[0].x, (int32_t)src_pts.size(), cfg_.device, &out[0].x) != VF_OK)
            VF_FATAL(std::string("vf_op_undistort: ") + vf_last_global_error());
        return out;
    }

    /// Compute the kps velocity (feature_tracker.cpp:323-363); host restatement of what the stereo kernel's epilogue does
    std::vector<cv::Point2f> ComputeKpsVelocity(std::vector<int>& ids, std::vector<cv::Point2f>& un_distortion_kps,
                                                std::map<int, cv::Point2f>& current_id_pts,
                                                std::map<int, cv::Point2f>& previous_id_pts) {
        assert(ids.size() == un_distortion_kps.size());
        std::vector<cv::Point2f> vel;
        current_id_pts.clear();
        for (unsigned int i = 0; i < ids.size(); i++) current_id_pts.insert(std::make_pair(ids[i], un_distortion_kps[i]));
        const double dt = current_time_ - previous_time_;
        for (unsigned int i = 0; i < un_distortion_kps.size(); i++) {
            std::map<int, cv::Point2f>::iterator it = previous_id_pts.empty() ? previous_id_pts.end() : previous_id_pts.find(ids[i]);
            if (it != previous_id_pts.end())
                vel.emplace_back((float)((un_distortion_kps[i].x - it->second.x) / dt), (float)((un_distortion_kps[i].y - it->second.y) / dt));
            else
                vel.emplace_back(0.f, 0.f);
        }
        return vel;
    }

    /// Set feature extractor mask (feature_tracker.cpp:219-246): the mask the last TrackImage call built on the device
    void SetFeatureExtractorMask(cv::Mat& mask) {
        if (!handle_) VF_FATAL("SetFeatureExtractorMask: no frame has been tracked");
        mask = cv::Mat(cfg_.height, cfg_.width, CV_8UC1, cv::Scalar(255));
        size_t bytes = 0;
        if (vf_tracker_debug_get(handle_, VF_DBG_MASK, 0, mask.data, (size_t)cfg_.width * cfg_.height, &bytes) != VF_OK)
            VF_FATAL(std::string("vf_tracker_debug_get: ") + vf_last_error(handle_));
    }

    /// Get draw tracker result image to show (feature_tracker.cpp:451-455)
    cv::Mat getDrawTrackResultImg() {
        std::unique_lock<std::mutex> lck(draw_track_img_mutex_);
        return draw_track_img_result_;
    }
    cv::Mat getCurrentImage() const { return current_img_; }
    cv::Mat getPreviousImage() const { return previous_img_; }

    /// Compute distance between two cv points (feature_tracker.cpp:457-462)
    double PtDistance(cv::Point2f& pt1, cv::Point2f& pt2) const {
        const double dx = pt1.x - pt2.x, dy = pt1.y - pt2.y;
        return std::sqrt(dx * dx + dy * dy);
    }

    /// Template fun, reduce the vector according the status (feature_tracker.hpp:74-81)
    template <typename T, typename S>
    void ReduceVector(std::vector<T>& v, const std::vector<S>& status) {
        int j = 0;
        for (int i = 0; i < int(v.size()); i++)
            if (status[i]) v[j++] = v[i];
        v.resize(j);
    }

    /// Draw tracker result (feature_tracker.cpp:267-299): left|right side by side as RGB, a circle per left feature coloured
    /// by track age (blue = new, red = >= 20 frames), a green circle per right feature, a green arrow to the previous
    /// position.  Host-side, only run when show_track != 0; the rasterisation is pixel-identical to the reference's
    /// cv::circle / cv::arrowedLine calls (draw.hpp; checked against cv2 by the tests).
    void DrawTracker() {
        const int W = current_img_.cols, H = current_img_.rows, n = (int)current_kps_.size();
        cv::Mat out(H, 2 * W, CV_8UC3);
        std::vector<float> prev_xy(2 * (size_t)n + 2, 0.f);
        std::vector<unsigned char> has_prev((size_t)n + 1, 0);
        for (int i = 0; i < n; ++i) {
            auto iter = previous_left_ids_kps_map_.find(ids_[i]);
            if (iter == previous_left_ids_kps_map_.end()) continue;
            has_prev[i] = 1; prev_xy[2 * i] = iter->second.x; prev_xy[2 * i + 1] = iter->second.y;
        }
        static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "Point2f must be two packed floats");
        const vf_draw::Canvas canvas = {out.data, 2 * W, H, (size_t)out.step};
        vf_draw::DrawTrackImage(current_img_.data, current_img_.step, current_right_img_.data, current_right_img_.step, W, H,
                                n ? &current_kps_[0].x : nullptr, n ? track_count_.data() : nullptr, n,
                                current_right_img_kps_.empty() ? nullptr : &current_right_img_kps_[0].x, (int)current_right_img_kps_.size(),
                                prev_xy.data(), has_prev.data(), canvas);
        draw_track_img_result_ = out;
    }

    // read-only views of the reference's private state (the reference keeps these private; tests use them)
    const std::vector<int>& ids() const { return ids_; }
    const std::vector<int>& ids_right() const { return ids_right_; }
    const std::vector<int>& track_count() const { return track_count_; }
    const std::vector<cv::Point2f>& current_kps() const { return current_kps_; }
    const std::vector<cv::Point2f>& current_right_kps() const { return current_right_img_kps_; }
    const std::vector<vf_feature>& records() const { return records_; }
    vf_tracker* handle() const { return handle_; }

public:
    int OPTICAL_FLOW_BACK;
    int MAX_CNT_FEATURES_PER_FRAME = 0;
    int MIN_DIST;   /// min distance between two features
    int SHOW_TRACK_RESULT;
    int FRONTEND_WAIT_KEY;

public:
    std::mutex draw_track_img_mutex_;

private:
    void Release() { vf_tracker_destroy(handle_); handle_ = nullptr; }
    void Ensure(int width, int height) {
        if (handle_ && cfg_.width == width && cfg_.height == height && cfg_.max_cnt == MAX_CNT_FEATURES_PER_FRAME &&
            cfg_.min_dist == MIN_DIST && cfg_.flow_back == OPTICAL_FLOW_BACK)
            return;
        Release();
        cfg_.width = width; cfg_.height = height;
        cfg_.max_cnt = MAX_CNT_FEATURES_PER_FRAME; cfg_.min_dist = MIN_DIST; cfg_.flow_back = OPTICAL_FLOW_BACK;
        cfg_.cam[0] = cameras_[0]; cfg_.cam[1] = cameras_[1];
        if (vf_tracker_create(&cfg_, &handle_) != VF_OK) VF_FATAL(std::string("vf_tracker_create: ") + vf_last_global_error());
        records_.resize((size_t)std::max(1, cfg_.max_cnt));
    }

    vf_config cfg_;
    vf_tracker* handle_ = nullptr;
    std::vector<vf_feature> records_;
    double current_time_ = 0, previous_time_ = 0;
    std::vector<vf_camera> cameras_;
    std::vector<int> ids_, ids_right_;
    std::vector<int> track_count_;
    std::vector<cv::Point2f> current_kps_, current_right_img_kps_;
    std::vector<cv::Point2f> current_un_distortion_kps_, current_right_un_distortion_kps_;
    std::vector<cv::Point2f> pts_velocity, right_pts_velocity;
    std::map<int, cv::Point2f> previous_left_ids_kps_map_;   /// only used to draw the track image
    cv::Mat current_img_, current_right_img_, previous_img_;
    cv::Mat draw_track_img_result_;
};

}  // namespace estimator
