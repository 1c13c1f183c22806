// Drives the header-only C++ drop-in (include/vins_fast_b200/feature_tracker.hpp) the way the reference's Estimator does
// (vins/estimator/estimator.cpp:13-14 construct + ReadIntrinsicParameter, :143-158 FrontendTracker -> TrackImage), checks
// the invariants the consumer asserts on a featureFrame (vins/estimator/feature_manager.cpp:36-43), and dumps every
// featureFrame so the Python test can compare it with the CPU oracle.
//
//   test_feature_tracker run  <estimator.yaml> <cam0.yaml> <cam1.yaml> <frames.bin> <W> <H> <n_frames> <out.bin>
//   test_feature_tracker host <estimator.yaml> <cam0.yaml> <cam1.yaml>      (no GPU: host layer only, TrackImage must fail loudly)
//
// frames.bin: n_frames x { double t; uint8 left[H*W]; uint8 right[H*W] }.
// out.bin:    per frame { int32 n_ids; n_ids x { int32 id; int32 n_obs; n_obs x { int32 camera_id; double v[7] } } }
#define VINS_FAST_B200_THROW 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "vins_fast_b200/feature_tracker.hpp"
#include "vins_fast_b200/frontend_pipeline.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int run_host(const char* est, const char* cam0, const char* cam1) {
    auto setting = common::Setting::getSingleton();
    CHECK(setting->InitParamSetting(est));
    CHECK(setting->is_init_);
    CHECK(common::Setting::Get<int>("max_cnt") == 150);
    CHECK(common::Setting::Get<int>("min_dist") == 30);
    CHECK(common::Setting::Get<int>("flow_back") == 1);
    CHECK(common::Setting::Get<int>("show_track") == 1);
    CHECK(common::Setting::Get<int>("no_such_key") == 0);              // cv::FileNode of a missing key reads as 0
    CHECK(common::Setting::Get<std::string>("cam0_calib") == "cam0_mei.yaml");
    CHECK(common::Setting::Get<double>("td") == 0.5);
    estimator::FeatureTracker ft;
    CHECK(ft.MAX_CNT_FEATURES_PER_FRAME == 150 && ft.MIN_DIST == 30 && ft.OPTICAL_FLOW_BACK == 1 && ft.SHOW_TRACK_RESULT == 1);
    ft.ReadIntrinsicParameter({cam0, cam1});
    cv::Point2f a(10.5f, 20.5f), b(13.5f, 24.5f);
    CHECK(ft.PtDistance(a, b) == 5.0);
    CHECK(ft.InBorder(cv::Point2f(1.f, 1.f), 480, 752) && !ft.InBorder(cv::Point2f(0.4f, 5.f), 480, 752));
    CHECK(ft.InBorder(cv::Point2f(750.5f, 5.f), 480, 752));            // cvRound(750.5) = 750 (half to even) < 751
    CHECK(!ft.InBorder(cv::Point2f(751.0f, 5.f), 480, 752));
    std::vector<int> v = {1, 2, 3, 4};
    std::vector<unsigned char> st = {1, 0, 0, 1};
    ft.ReduceVector(v, st);
    CHECK(v.size() == 2 && v[0] == 1 && v[1] == 4);
    bool threw = false;
    try { ft.TrackImage(0.0, cv::Mat(), cv::Mat()); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);                                                        // assert(!img0.empty() && !img1.empty())
    if (vf_device_count() == 0) {
        cv::Mat img(480, 752, CV_8UC1, cv::Scalar(7));
        threw = false;
        try { ft.TrackImage(0.0, img, img); } catch (const std::runtime_error& e) {
            threw = std::strstr(e.what(), "no CUDA device") != nullptr;  // never a silent CPU fallback
        }
        CHECK(threw);
    }
    threw = false;
    try { common::Setting::getSingleton()->InitParamSetting("/nonexistent/estimator.yaml"); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    std::printf("host ok\n");
    return 0;
}

static int run_frames(int argc, char** argv) {
    CHECK(argc == 10);
    const int W = std::atoi(argv[6]), H = std::atoi(argv[7]), n_frames = std::atoi(argv[8]);
    CHECK(common::Setting::getSingleton()->InitParamSetting(argv[2]));
    estimator::FeatureTracker ft;
    ft.ReadIntrinsicParameter({argv[3], argv[4]});
    std::FILE* in = std::fopen(argv[5], "rb");
    std::FILE* out = std::fopen(argv[9], "wb");
    CHECK(in && out);
    std::vector<unsigned char> left((size_t)W * H), right((size_t)W * H);
    for (int k = 0; k < n_frames; ++k) {
        double t;
        CHECK(std::fread(&t, sizeof(t), 1, in) == 1);
        CHECK(std::fread(left.data(), 1, left.size(), in) == left.size());
        CHECK(std::fread(right.data(), 1, right.size(), in) == right.size());
        cv::Mat img0(H, W, CV_8UC1, left.data()), img1(H, W, CV_8UC1, right.data());
        auto frame_feature = ft.TrackImage(t, img0, img1);
        // what FeatureManager::AddFeatureAndCheckParallax asserts on every entry (feature_manager.cpp:36-43)
        for (auto& id_pts : frame_feature) {
            CHECK(id_pts.second.size() == 1 || id_pts.second.size() == 2);
            CHECK(id_pts.second[0].first == 0);
            CHECK(id_pts.second[0].second(2) == 1.0);
            if (id_pts.second.size() == 2) { CHECK(id_pts.second[1].first == 1); CHECK(id_pts.second[1].second(2) == 1.0); }
        }
        CHECK(frame_feature.size() == ft.ids().size());                 // ids are unique
        CHECK(ft.ids_right().size() <= ft.ids().size());
        if (ft.SHOW_TRACK_RESULT) {
            cv::Mat draw = ft.getDrawTrackResultImg();
            CHECK(draw.rows == H && draw.cols == 2 * W && draw.type() == CV_8UC3);
        }
        cv::Mat mask;
        ft.SetFeatureExtractorMask(mask);
        CHECK(mask.rows == H && mask.cols == W);
        // every new corner (track count 1) was picked where the mask was still 255; every older one zeroed its own centre
        for (size_t i = 0; i < ft.ids().size(); ++i) {
            const cv::Point2f p = ft.current_kps()[i];
            const unsigned char m = mask.at<unsigned char>(cvRound(p.y), cvRound(p.x));
            if (ft.track_count()[i] == 1) CHECK(m == 255); else CHECK(m == 0);
        }
        const int32_t n_ids = (int32_t)frame_feature.size();
        std::fwrite(&n_ids, sizeof(n_ids), 1, out);
        for (auto& id_pts : frame_feature) {
            const int32_t id = id_pts.first, n_obs = (int32_t)id_pts.second.size();
            std::fwrite(&id, sizeof(id), 1, out);
            std::fwrite(&n_obs, sizeof(n_obs), 1, out);
            for (auto& obs : id_pts.second) {
                const int32_t cam = obs.first;
                double v[7];
                for (int q = 0; q < 7; ++q) v[q] = obs.second(q);
                std::fwrite(&cam, sizeof(cam), 1, out);
                std::fwrite(v, sizeof(double), 7, out);
            }
        }
    }
    std::fclose(in);
    std::fclose(out);
    std::printf("run ok: %d frames\n", n_frames);
    return 0;
}

// Estimator::FrontendTracker + feature_buf_ (estimator.cpp:143-158) on the pipelined API: the queue must hold exactly what the
// synchronous TrackImage returns for every 2nd frame, value for value.
//   test_feature_tracker pipe <cam0.yaml> <cam1.yaml> <frames.bin> <W> <H> <n_frames> [depth]
static int run_pipe(int argc, char** argv) {
    CHECK(argc == 8 || argc == 9);
    const int W = std::atoi(argv[5]), H = std::atoi(argv[6]), n_frames = std::atoi(argv[7]);
    vf_camera c0, c1;
    CHECK(vf_camera_load_yaml(argv[2], &c0) == VF_OK && vf_camera_load_yaml(argv[3], &c1) == VF_OK);
    estimator::FeatureTracker ft(150, 30, 1);
    ft.SetCameras(c0, c1);
    const int depth = argc > 8 ? std::atoi(argv[8]) : 1;
    estimator::FrontendPipeline pipe(150, 30, 1, c0, c1, 2, -1, depth);
    std::FILE* in = std::fopen(argv[4], "rb");
    CHECK(in);
    std::vector<std::vector<unsigned char>> left(n_frames, std::vector<unsigned char>((size_t)W * H)), right = left;
    std::vector<double> ts(n_frames);
    std::vector<estimator::FeatureTracker::FeatureFrame> want;
    for (int k = 0; k < n_frames; ++k) {
        CHECK(std::fread(&ts[k], sizeof(double), 1, in) == 1);
        CHECK(std::fread(left[k].data(), 1, left[k].size(), in) == left[k].size());
        CHECK(std::fread(right[k].data(), 1, right[k].size(), in) == right[k].size());
        cv::Mat img0(H, W, CV_8UC1, left[k].data()), img1(H, W, CV_8UC1, right[k].data());
        want.push_back(ft.TrackImage(ts[k], img0, img1));
    }
    std::fclose(in);
    for (int k = 0; k < n_frames; ++k) {
        cv::Mat img0(H, W, CV_8UC1, left[k].data()), img1(H, W, CV_8UC1, right[k].data());
        pipe.Push(ts[k], img0, img1);
    }
    pipe.Flush();
    CHECK(pipe.FramesTracked() == n_frames);
    CHECK((int)pipe.Pending() == n_frames / 2);
    std::pair<double, estimator::FeatureTracker::FeatureFrame> item;
    for (int k = 1; k < n_frames; k += 2) {                 // input_image_cnt_ % 2 == 0  <=>  0-based index odd
        CHECK(pipe.Pop(item));
        CHECK(item.first == ts[k]);
        CHECK(item.second.size() == want[k].size());
        auto a = item.second.begin();
        auto b = want[k].begin();
        for (; a != item.second.end(); ++a, ++b) {
            CHECK(a->first == b->first && a->second.size() == b->second.size());
            for (size_t o = 0; o < a->second.size(); ++o) {
                CHECK(a->second[o].first == b->second[o].first);
                for (int q = 0; q < 7; ++q) CHECK(a->second[o].second(q) == b->second[o].second(q));
            }
        }
    }
    CHECK(!pipe.Pop(item));
    std::printf("pipe ok: %d frames, %d queued\n", n_frames, n_frames / 2);
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 5 && std::strcmp(argv[1], "host") == 0) return run_host(argv[2], argv[3], argv[4]);
        if (argc >= 2 && std::strcmp(argv[1], "run") == 0) return run_frames(argc, argv);
        if (argc >= 2 && std::strcmp(argv[1], "pipe") == 0) return run_pipe(argc, argv);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    std::fprintf(stderr, "usage: see the header comment\n");
    return 64;
}
