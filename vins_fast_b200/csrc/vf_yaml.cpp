// Reader for the OpenCV "%YAML:1.0" dialect the reference's configuration files use, limited to what the
// hot path needs: top-level scalars and one level of nested scalar maps.  `!!opencv-matrix` blocks (the
// body_T_cam extrinsics, read only by the estimator) are skipped.
//
// Replaces, for the tracker keys only:
//   common::Setting::InitParamSetting / Get<T>       vins/common/parameter.hpp:36-79  (cv::FileStorage READ)
//   FeatureTracker::FeatureTracker (max_cnt, flow_back, min_dist, show_track, frontend_wait_key)
//                                                     vins/estimator/feature_tracker.cpp:8-18
//   CameraFactory::generateCameraFromYamlFile         camera_models/src/camera_models/CameraFactory.cc:96-189
//   CataCamera / PinholeCamera Parameters::readFromYamlFile   CataCamera.cc:161-202, PinholeCamera.cc:145-183
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#include "../../include/vins_fast_b200.h"

namespace {

// error text goes to the library-wide "last global error" slot (vf_last_global_error)
struct ErrSlot {
    ErrSlot& operator=(const std::string& m);
    ErrSlot& operator=(const char* m) { return *this = std::string(m); }
};
ErrSlot g_yaml_error;

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

std::string strip_comment(const std::string& line) {
    bool in_quote = false;
    for (size_t i = 0; i < line.size(); ++i) {
        if (line[i] == '"') in_quote = !in_quote;
        if (line[i] == '#' && !in_quote) return line.substr(0, i);
    }
    return line;
}

// "key" and "section.key" -> raw scalar text
bool parse_yaml(const std::string& path, std::map<std::string, std::string>& out, ErrSlot& err) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) { err = "cannot open " + path; return false; }
    std::string line, section;
    int section_indent = -1;
    bool skipping_block = false;
    int skip_indent = 0;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        line = strip_comment(line);
        if (trim(line).empty()) continue;
        if (line[0] == '%') continue;                 // %YAML:1.0
        if (trim(line) == "---") continue;
        int indent = 0;
        while (indent < (int)line.size() && line[indent] == ' ') ++indent;
        if (skipping_block) {                          // inside an !!opencv-matrix (or any nested block we ignore)
            if (indent > skip_indent) continue;
            skipping_block = false;
        }
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;      // continuation of a flow sequence, ignored
        std::string key = trim(line.substr(0, colon));
        std::string val = trim(line.substr(colon + 1));
        if (key.empty()) continue;
        if (indent == 0) { section.clear(); section_indent = -1; }
        if (!val.empty() && val.compare(0, 2, "!!") == 0) { skipping_block = true; skip_indent = indent; continue; }
        if (val.empty()) {                             // start of a nested map
            if (indent == 0) { section = key; section_indent = 0; }
            else { skipping_block = true; skip_indent = indent; }
            continue;
        }
        if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
        if (indent > 0 && !section.empty() && indent > section_indent) out[section + "." + key] = val;
        else if (indent == 0) out[key] = val;
    }
    return true;
}

bool get_double(const std::map<std::string, std::string>& m, const std::string& key, double& v) {
    auto it = m.find(key);
    if (it == m.end()) return false;
    char* end = nullptr;
    v = std::strtod(it->second.c_str(), &end);
    return end != it->second.c_str();
}

bool get_int(const std::map<std::string, std::string>& m, const std::string& key, int32_t& v) {
    double d;
    if (!get_double(m, key, d)) return false;
    v = (int32_t)d;   // cv::FileNode -> int truncates a real the same way
    return true;
}

}  // namespace

extern "C" void vf_internal_set_error(const char* msg);   // vf_api.cu
namespace {
ErrSlot& ErrSlot::operator=(const std::string& m) { vf_internal_set_error(m.c_str()); return *this; }
}

extern "C" vf_status vf_camera_load_yaml(const char* calib_yaml, vf_camera* cam) {
    if (!calib_yaml || !cam) { g_yaml_error = "null argument"; return VF_ERR_INVALID_ARG; }
    std::map<std::string, std::string> m;
    if (!parse_yaml(calib_yaml, m, g_yaml_error)) return VF_ERR_IO;
    std::string model = "mei";                          // CameraFactory default when model_type is absent
    if (m.count("model_type")) model = m["model_type"];
    std::transform(model.begin(), model.end(), model.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    *cam = vf_camera();
    get_int(m, "image_width", cam->image_width);
    get_int(m, "image_height", cam->image_height);
    bool ok = get_double(m, "distortion_parameters.k1", cam->k1) && get_double(m, "distortion_parameters.k2", cam->k2) &&
              get_double(m, "distortion_parameters.p1", cam->p1) && get_double(m, "distortion_parameters.p2", cam->p2);
    if (model == "mei") {
        cam->model = VF_CAM_MEI;
        ok = ok && get_double(m, "mirror_parameters.xi", cam->xi) && get_double(m, "projection_parameters.gamma1", cam->fx) &&
             get_double(m, "projection_parameters.gamma2", cam->fy) && get_double(m, "projection_parameters.u0", cam->cx) &&
             get_double(m, "projection_parameters.v0", cam->cy);
    } else if (model == "pinhole") {
        cam->model = VF_CAM_PINHOLE;
        ok = ok && get_double(m, "projection_parameters.fx", cam->fx) && get_double(m, "projection_parameters.fy", cam->fy) &&
             get_double(m, "projection_parameters.cx", cam->cx) && get_double(m, "projection_parameters.cy", cam->cy);
    } else {
        // kannala_brandt / scaramuzza / pinhole_full exist in camodocal but are not on the reference's configured path
        g_yaml_error = "unsupported camera model_type '" + model + "' (hot path covers MEI and PINHOLE)";
        return VF_ERR_PARSE;
    }
    if (!ok) { g_yaml_error = std::string("missing camera parameter in ") + calib_yaml; return VF_ERR_PARSE; }
    return VF_OK;
}

extern "C" vf_status vf_config_load_yaml(const char* estimator_yaml, vf_config* cfg) {
    if (!estimator_yaml || !cfg) { g_yaml_error = "null argument"; return VF_ERR_INVALID_ARG; }
    std::map<std::string, std::string> m;
    if (!parse_yaml(estimator_yaml, m, g_yaml_error)) return VF_ERR_IO;
    vf_config_default(cfg);
    // the reference reads these with Setting::Get<int> and LOG(FATAL)s when the file is not open; a missing key
    // would silently read 0 there -- here it is an error
    if (!get_int(m, "max_cnt", cfg->max_cnt) || !get_int(m, "min_dist", cfg->min_dist) || !get_int(m, "flow_back", cfg->flow_back)) {
        g_yaml_error = std::string("missing tracker key (max_cnt / min_dist / flow_back) in ") + estimator_yaml;
        return VF_ERR_PARSE;
    }
    get_int(m, "image_width", cfg->width);
    get_int(m, "image_height", cfg->height);
    get_int(m, "lk_max_level", cfg->lk_max_level);       // optional keys; defaults = the reference's literals
    get_int(m, "lk_win", cfg->lk_win);
    get_double(m, "quality_level", cfg->quality_level);
    // cam0_calib / cam1_calib (estimator.cpp:95-99): a listed file must be readable; a YAML without the keys leaves cfg->cam to the caller
    for (int k = 0; k < 2; ++k) {
        const std::string key = k == 0 ? "cam0_calib" : "cam1_calib";
        if (!m.count(key)) continue;
        std::string p = m[key];
        std::ifstream probe(p.c_str());
        if (!probe.is_open()) {   // also try next to the estimator YAML (the shipped paths are absolute /home/weihao/...)
            const std::string dir = std::string(estimator_yaml).substr(0, std::string(estimator_yaml).find_last_of('/') + 1);
            p = dir + p.substr(p.find_last_of('/') + 1);
            probe.open(p.c_str());
        }
        // the reference aborts in CameraFactory::generateCameraFromYamlFile when a listed file cannot be read
        // (CameraFactory.cc:96-110); keeping the default pinhole instead would silently corrupt the undistorted points
        if (!probe.is_open()) {
            g_yaml_error = "calibration file '" + m[key] + "' (key " + key + " of " + estimator_yaml + ") cannot be opened, neither at that "
                           "path nor as " + p;
            return VF_ERR_IO;
        }
        vf_status st = vf_camera_load_yaml(p.c_str(), &cfg->cam[k]);
        if (st != VF_OK) return st;
    }
    return VF_OK;
}

// common::Setting::Get<T>(key) (vins/common/parameter.hpp:36-53): raw scalar text of a top-level key, or of
// "section.key" for one level of nesting.  The C++ mirror (include/vins_fast_b200/parameter.hpp) converts it to T.
extern "C" vf_status vf_yaml_get(const char* yaml_path, const char* key, char* value_out, size_t cap) {
    if (!yaml_path || !key || !value_out || cap == 0) { g_yaml_error = "null argument"; return VF_ERR_INVALID_ARG; }
    std::map<std::string, std::string> m;
    if (!parse_yaml(yaml_path, m, g_yaml_error)) return VF_ERR_IO;
    auto it = m.find(key);
    if (it == m.end()) { g_yaml_error = std::string("key '") + key + "' not found in " + yaml_path; return VF_ERR_PARSE; }
    if (it->second.size() + 1 > cap) { g_yaml_error = "value buffer too small"; return VF_ERR_INVALID_ARG; }
    std::copy(it->second.begin(), it->second.end(), value_out);
    value_out[it->second.size()] = 0;
    return VF_OK;
}
