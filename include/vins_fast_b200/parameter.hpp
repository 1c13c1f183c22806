// common::Setting -- the configuration singleton FeatureTracker's constructor reads
// (vins/common/parameter.hpp:19-79: getSingleton(), InitParamSetting(path), Get<T>(key), is_init_).
// Same names and call pattern; backed by the library's reader for the OpenCV "%YAML:1.0" dialect (vf_yaml_get)
// instead of cv::FileStorage, so it needs neither OpenCV nor glog.
//
// Error behaviour mirrors the reference's LOG(FATAL): message to stderr and std::abort(); define
// VINS_FAST_B200_THROW to get a std::runtime_error instead (used by this repository's tests).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../vins_fast_b200.h"
#include "compat.hpp"

namespace common {

class Setting {
public:
    static std::shared_ptr<Setting> getSingleton() {
        static std::mutex m;
        std::lock_guard<std::mutex> lck(m);
        std::shared_ptr<Setting>& s = instance();
        if (!s) s = std::shared_ptr<Setting>(new Setting());
        return s;
    }

    // parameter.hpp:67-79 -- opens the estimator YAML; LOG(FATAL) when it does not exist
    bool InitParamSetting(const std::string& config_file_path) {
        char probe[8];
        const vf_status st = vf_yaml_get(config_file_path.c_str(), "%", probe, sizeof(probe));
        if (st == VF_ERR_IO) {
            VF_FATAL("parameter file" + config_file_path + "does not exist.");
            return false;
        }
        path_ = config_file_path;
        is_init_ = true;
        return true;
    }

    const std::string& getFile() const { return path_; }

    // parameter.hpp:36-53 -- Get<int>("max_cnt"), Get<std::string>("cam0_calib"), ... ; "section.key" reaches one level down
    template <typename T>
    static T Get(const std::string& key) {
        std::shared_ptr<Setting>& s = instance();
        if (!s) VF_FATAL("singleton_ == nullptr");
        if (!s->is_init_) VF_FATAL("File is not open success");
        char buf[1024];
        if (vf_yaml_get(s->path_.c_str(), key.c_str(), buf, sizeof(buf)) != VF_OK) return T();   // cv::FileNode of a missing key converts to 0 / ""
        return convert<T>(buf);
    }

    bool is_init_ = false;

private:
    Setting() = default;
    static std::shared_ptr<Setting>& instance() {
        static std::shared_ptr<Setting> s;
        return s;
    }
    template <typename T>
    static T convert(const char* text) {
        std::istringstream is(text);
        double d = 0;
        is >> d;                 // cv::FileNode -> int truncates a real the same way
        return static_cast<T>(d);
    }
    std::string path_;
};

template <>
inline std::string Setting::convert<std::string>(const char* text) { return std::string(text); }

}  // namespace common
