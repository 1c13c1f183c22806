// Input side (include/vins_fast_b200/image_source.hpp) without a GPU:
//   test_image_source host <dir>
//     <dir>/cam0/data.csv, <dir>/cam1/data.csv (EuRoC listings) and the PGM files they name were written by the Python test.
//     Reads both listings, feeds the images through StereoSync in arrival order (the 3 ms pairing rule of
//     Estimator::tFrontendProcess, estimator.cpp:103-141) and prints one line per synchronised pair:
//        pair <stamp> <sum of left pixels> <sum of right pixels>
//     then "thrown <n0> <n1>".  Also exercises CopyMono8 / GetImageFromMsg (strided 8UC1 message) and the PGM round trip.
//   test_image_source replay <dir> <cam0.yaml> <cam1.yaml> <out.bin>      (GPU)
//     Replays the same directory through FrontendPipeline (pinned ring, every frame queued) and dumps the featureFrames
//     in the layout of test_feature_tracker's out.bin, preceded by the pair's stamp (double).
#define VINS_FAST_B200_THROW 1
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "vins_fast_b200/frontend_pipeline.hpp"
#include "vins_fast_b200/image_source.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

struct Arrival { double stamp; int cam; std::string file; };

static int load(const std::string& dir, std::vector<Arrival>& arrivals) {
    for (int cam = 0; cam < 2; ++cam) {
        std::vector<std::pair<double, std::string>> list;
        CHECK(common::EurocImageList(dir + "/cam" + std::to_string(cam) + "/data.csv", list));
        for (auto& e : list) arrivals.push_back(Arrival{e.first, cam, dir + "/cam" + std::to_string(cam) + "/data/" + e.second});
    }
    std::stable_sort(arrivals.begin(), arrivals.end(), [](const Arrival& a, const Arrival& b) { return a.stamp < b.stamp; });
    return 0;
}

static unsigned long long pixel_sum(const cv::Mat& m) {
    unsigned long long s = 0;
    for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) s += m.at<unsigned char>(y, x);
    return s;
}

static int run_host(const std::string& dir) {
    std::vector<Arrival> arrivals;
    if (load(dir, arrivals)) return 1;
    estimator::StereoSync sync;
    for (auto& a : arrivals) {
        cv::Mat img;
        CHECK(common::ReadPgm(a.file, img));
        CHECK(img.isContinuous() && img.type() == CV_8UC1);
        if (a.cam == 0) sync.Image0Callback(a.stamp, img); else sync.Image1Callback(a.stamp, img);
        double t; cv::Mat i0, i1;
        while (!sync.Idle())
            if (sync.Next(t, i0, i1)) std::printf("pair %.9f %llu %llu\n", t, pixel_sum(i0), pixel_sum(i1));
    }
    std::printf("thrown %d %d\n", sync.thrown0(), sync.thrown1());
    // GetImageFromROSMsg: a strided "8UC1" message becomes a continuous mono8 image
    std::vector<unsigned char> wire(5 * 16, 0);
    for (int y = 0; y < 5; ++y) for (int x = 0; x < 11; ++x) wire[y * 16 + x] = (unsigned char)(y * 11 + x);
    common::ImageMsg msg;
    msg.width = 11; msg.height = 5; msg.step = 16; msg.encoding = "8UC1"; msg.data = wire.data(); msg.stamp = 1.5;
    cv::Mat img = common::GetImageFromMsg(msg);
    CHECK(img.rows == 5 && img.cols == 11 && img.isContinuous());
    for (int y = 0; y < 5; ++y) for (int x = 0; x < 11; ++x) CHECK(img.at<unsigned char>(y, x) == y * 11 + x);
    msg.encoding = "bgr8";
    bool threw = false;
    try { common::GetImageFromMsg(msg); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    // PGM round trip, with a comment line in the header
    const std::string p = dir + "/roundtrip.pgm";
    CHECK(common::WritePgm(p, img));
    cv::Mat back;
    CHECK(common::ReadPgm(p, back) && back.rows == 5 && back.cols == 11 && std::memcmp(back.data, img.data, 55) == 0);
    CHECK(!common::ReadPgm(dir + "/missing.pgm", back));
    std::printf("host ok\n");
    return 0;
}

static int run_replay(const std::string& dir, const char* cam0, const char* cam1, const char* out_path) {
    std::vector<Arrival> arrivals;
    if (load(dir, arrivals)) return 1;
    vf_camera c0, c1;
    CHECK(vf_camera_load_yaml(cam0, &c0) == VF_OK && vf_camera_load_yaml(cam1, &c1) == VF_OK);
    estimator::FrontendPipeline pipe(150, 30, 1, c0, c1, /*every_nth=*/1, -1, /*depth=*/2);
    estimator::StereoSync sync;
    int pushed = 0;
    for (auto& a : arrivals) {
        cv::Mat img;
        CHECK(common::ReadPgm(a.file, img));
        if (a.cam == 0) sync.Image0Callback(a.stamp, img); else sync.Image1Callback(a.stamp, img);
        double t; cv::Mat i0, i1;
        while (!sync.Idle()) {
            if (!sync.Next(t, i0, i1)) continue;
            if (pushed % 2 == 0) {
                pipe.Push(t, i0, i1);                                    // copy into the pinned slot
            } else {                                                     // a source that fills the slot in place
                cv::Mat l = pipe.AcquireLeft(i0.cols, i0.rows), r = pipe.AcquireRight(i0.cols, i0.rows);
                std::memcpy(l.data, i0.data, (size_t)i0.cols * i0.rows);
                std::memcpy(r.data, i1.data, (size_t)i1.cols * i1.rows);
                pipe.PushAcquired(t);
            }
            std::memset(i0.data, 0, (size_t)i0.cols * i0.rows);          // the caller's buffers are free once Push returns
            std::memset(i1.data, 0, (size_t)i1.cols * i1.rows);
            ++pushed;
        }
    }
    pipe.Flush();
    CHECK(pipe.FramesTracked() == pushed && (int)pipe.Pending() == pushed);
    std::FILE* out = std::fopen(out_path, "wb");
    CHECK(out);
    std::pair<double, estimator::FeatureTracker::FeatureFrame> item;
    while (pipe.Pop(item)) {
        std::fwrite(&item.first, sizeof(double), 1, out);
        const int32_t n_ids = (int32_t)item.second.size();
        std::fwrite(&n_ids, sizeof(n_ids), 1, out);
        for (auto& id_pts : item.second) {
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
    std::fclose(out);
    std::printf("replay ok: %d pairs, thrown %d %d\n", pushed, sync.thrown0(), sync.thrown1());
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc == 3 && std::strcmp(argv[1], "host") == 0) return run_host(argv[2]);
        if (argc == 6 && std::strcmp(argv[1], "replay") == 0) return run_replay(argv[2], argv[3], argv[4], argv[5]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    std::fprintf(stderr, "usage: see the header comment\n");
    return 64;
}
