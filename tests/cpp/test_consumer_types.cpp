// Feeds featureFrames built by this library's host layer (estimator::FeatureTracker::ToFeatureFrame, the conversion
// TrackImage and FrontendPipeline use) into the reference's UNCHANGED consumer types -- estimator/frame_feature.hpp and
// estimator/id_features.hpp, included from the reference tree at compile time (this container only) -- through the calls
// FeatureManager::AddFeatureAndCheckParallax makes on them (vins/estimator/feature_manager.cpp:25-62).
//
//   test_consumer_types <records.bin> <n_frames>
// records.bin: per frame { int32 n; n x vf_feature }.   stdout: per frame "last_track new_landmark long_track right_obs sum".
#define VINS_FAST_B200_THROW 1
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <type_traits>
#include <vector>

#include "vins_fast_b200/feature_tracker.hpp"
#include "estimator/id_features.hpp"        // the reference's own headers, not a copy

#ifndef VF_HAVE_EIGEN
#error this test needs the real Eigen (the reference's vendored copy)
#endif

static_assert(std::is_same<estimator::FeatureTracker::FeatureFrame,
                           std::map<int, std::vector<std::pair<int, Eigen::Matrix<double, 7, 1>>>>>::value,
              "featureFrame must be the very type AddFeatureAndCheckParallax takes");

int main(int argc, char** argv) {
    if (argc != 3) return 64;
    std::FILE* in = std::fopen(argv[1], "rb");
    if (!in) return 65;
    const int n_frames = std::atoi(argv[2]);
    std::list<estimator::IDWithObservedFeatures> features_;
    const double td = 0.0;
    for (int frame_count = 0; frame_count < n_frames; ++frame_count) {
        int32_t n = 0;
        if (std::fread(&n, sizeof(n), 1, in) != 1) return 66;
        std::vector<vf_feature> rec((size_t)n);
        if (n && std::fread(rec.data(), sizeof(vf_feature), (size_t)n, in) != (size_t)n) return 67;
        const estimator::FeatureTracker::FeatureFrame image = estimator::FeatureTracker::ToFeatureFrame(rec.data(), n);
        int last_track_num = 0, new_landmark_num = 0, long_time_track_num = 0, right_obs = 0;
        double sum = 0;
        for (auto& id_pts : image) {
            estimator::ObservedFrameFeature f(id_pts.second[0].second, td);
            if (!(id_pts.second.size() <= 2) || id_pts.second[0].first != 0) return 1;
            if (id_pts.second.size() == 2) {
                f.RightObservation(id_pts.second[1].second);
                if (id_pts.second[1].first != 1) return 2;
                right_obs++;
            }
            if (f.point_.z() != 1.0) return 3;
            const int feature_id = id_pts.first;
            auto it = std::find_if(features_.begin(), features_.end(),
                                   [feature_id](const estimator::IDWithObservedFeatures& o) { return o.feature_id_ == feature_id; });
            if (it == features_.end()) {
                features_.push_back(estimator::IDWithObservedFeatures(feature_id, frame_count));
                features_.back().feature_per_frame_.push_back(f);
                new_landmark_num++;
            } else {
                it->feature_per_frame_.push_back(f);
                last_track_num++;
                if (it->feature_per_frame_.size() >= 4) long_time_track_num++;
            }
            sum += f.point_.x() + f.point_.y() + f.uv_.x() + f.uv_.y() + f.velocity_.x() + f.velocity_.y();
            if (f.track_right_success_) sum += f.point_right_.x() + f.uv_right_.x() + f.velocity_right_.x();
        }
        if ((size_t)(last_track_num + new_landmark_num) != image.size()) return 4;
        std::printf("%d %d %d %d %.17g\n", last_track_num, new_landmark_num, long_time_track_num, right_obs, sum);
    }
    std::fclose(in);
    return 0;
}
