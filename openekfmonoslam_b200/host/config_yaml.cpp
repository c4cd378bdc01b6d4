// config_yaml.cpp -- reader for the reference's configuration files (OpenCV YAML 1.0 as written in
// experiments/s3/config.yml): a RunConfiguration block names one entry of each section and every value is a quoted
// string (modules/Configuration/ConfigurationManager.cpp:74-111, ConfigurationDefines.h:47-95).  Only the nesting
// "Section: / Name: / Key: "value"" by indentation is needed, so no YAML library is.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "../../include/EKF.h"

namespace {
typedef std::map<std::string, std::string> KV;

std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string unquote(const std::string& s)
{
    std::string t = trim(s);
    if (t.size() >= 2 && (t[0] == '"' || t[0] == '\'')) t = t.substr(1, t.size() - 2);
    return t;
}
}  // namespace

bool ekfbLoadConfig(const char* fileName, ekfb_params* p, int* minMatches, int* maxMapSize)
{
    EkfHostConfig c;
    const bool ok = ekfbLoadConfigFull(fileName, &c);
    *p = c.params;
    if (minMatches) *minMatches = c.policy.min_matches_per_image;
    if (maxMapSize) *maxMapSize = c.policy.max_map_size;
    return ok;
}

bool ekfbLoadConfigFull(const char* fileName, EkfHostConfig* cfg)
{
    ekfb_params* p = &cfg->params;
    *cfg = EkfHostConfig();
    std::ifstream in(fileName);
    if (!in) return false;
    // flat map "Section/Name/Key" -> value, path built from indentation levels 0 / 2 / 4
    KV kv;
    std::string line, path[3];
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '%' || line[0] == '#') continue;
        size_t ind = line.find_first_not_of(' ');
        if (ind == std::string::npos) continue;
        std::string body = trim(line);
        if (body[0] == '#') continue;
        size_t colon = body.find(':');
        if (colon == std::string::npos) continue;
        std::string key = trim(body.substr(0, colon)), val = unquote(body.substr(colon + 1));
        int level = ind >= 4 ? 2 : ind >= 2 ? 1 : 0;
        path[level] = key;
        if (!val.empty()) {
            std::string full = level == 0 ? key : level == 1 ? path[0] + "/" + key : path[0] + "/" + path[1] + "/" + key;
            kv[full] = val;
        }
    }
    const std::string ekfName = kv["RunConfiguration/ExtendedKalmanFilter"], camName = kv["RunConfiguration/CameraCalibration"];
    if (ekfName.empty() || camName.empty()) return false;
    const std::string e = "ExtendedKalmanFilter/" + ekfName + "/", c = "CameraCalibration/" + camName + "/";
    bool ok = true;
    auto num = [&](const std::string& k, bool required) {
        KV::const_iterator it = kv.find(k);
        if (it == kv.end()) {
            if (required) { std::cerr << "ekfbLoadConfig: missing key " << k << std::endl; ok = false; }
            return 0.0;
        }
        return atof(it->second.c_str());
    };
    p->pixels_x = (int)num(c + "PixelsX", true); p->pixels_y = (int)num(c + "PixelsY", true);
    p->fx = num(c + "FX", true); p->fy = num(c + "FY", true); p->k1 = num(c + "K1", true); p->k2 = num(c + "K2", true);
    p->cx = num(c + "CX", true); p->cy = num(c + "CY", true); p->dx = num(c + "DX", true); p->dy = num(c + "DY", true);
    p->pixel_error_x = num(c + "PixelErrorX", true); p->pixel_error_y = num(c + "PixelErrorY", true);
    p->angular_vision_x = num(c + "AngularVisionX", true); p->angular_vision_y = num(c + "AngularVisionY", true);
    p->init_inv_depth_rho = num(e + "InitInvDepthRho", true); p->init_linear_accel_sd = num(e + "InitLinearAccelSD", true);
    p->init_angular_accel_sd = num(e + "InitAngularAccelSD", true); p->linear_accel_sd = num(e + "LinearAccelSD", true);
    p->angular_accel_sd = num(e + "AngularAccelSD", true); p->inverse_depth_rho_sd = num(e + "InverseDepthRhoSD", true);
    p->matching_coef = num(e + "MatchingCompCoefSecondBestVSFirst", true);
    p->ransac_threshold = num(e + "RansacThresholdPredictDistance", true);
    p->ransac_all_inliers_prob = num(e + "RansacAllInliersProbability", true);
    p->ransac_chi2 = num(e + "RansacChi2Threshold", true);
    // map management; optional keys default as in ExtendedKalmanFilterParametersConfiguration (0 / false)
    auto flag = [&](const std::string& k) {
        KV::const_iterator it = kv.find(k);
        return it != kv.end() && (it->second == "true" || it->second == "1");
    };
    cfg->policy.min_matches_per_image = (int)num(e + "MinMatchesPerImage", true);
    cfg->policy.max_map_size = (int)num(e + "MaxMapSize", false);
    cfg->policy.max_map_features_count = (int)num(e + "MaxMapFeaturesCount", false);
    cfg->policy.always_remove_unseen = flag(e + "AlwaysRemoveUnseenMapFeatures") ? 1 : 0;
    cfg->policy.good_feature_matching_percent = num(e + "GoodFeatureMatchingPercent", false);
    cfg->policy.linearity_index_threshold = num(e + "InverseDepthLinearityIndexThreshold", false);
    cfg->mapManagementFrequency = (int)num(e + "MapManagementFrequency", false);
    cfg->detectNewFeaturesImageAreasDivideTimes = (int)num(e + "DetectNewFeaturesImageAreasDivideTimes", false);
    cfg->detectNewFeaturesImageMaskEllipseSize = num(e + "DetectNewFeaturesImageMaskEllipseSize", false);
    return ok;
}
