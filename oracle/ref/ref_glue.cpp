// ref_glue.cpp -- C interface around the REFERENCE's own sources (compiled unmodified from /root/reference against
// oracle/cvshim) so that the oracle can be pinned against the reference's arithmetic.  TEST INFRASTRUCTURE ONLY.
//
// What is the reference's code here: everything under kalmanFilter/modules/{Core,1PointRansacEKF,Gui} that the
// per-frame path executes (EKF::step and the free functions it calls).  What is NOT: OpenCV (cvshim.hpp: Mat
// semantics + the oracle's cv2-pinned inv / eigen / ellipse), the configuration loader (the two parameter structs
// are filled directly) and the feature detector / descriptor extractor (the front end is outside the hot path: an
// injected detector returns the caller's keypoints, filtered by the mask the way OpenCV's detectors do).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../reference/kalmanFilter/modules/Configuration/ConfigurationManager.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/EKF.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/ImageFeaturePrediction.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/ImageFeatureMeasurement.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/Matching.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/StateAndCovariancePrediction.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/MeasurementPrediction.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/1PointRansac.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/Update.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/MapManagement.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/AddMapFeature.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/DetectNewImageFeatures.h"
#include "../../../reference/kalmanFilter/modules/1PointRansacEKF/CommonFunctions.h"
#include "../../../reference/kalmanFilter/modules/Core/EKFMath.h"

#include "../ekf_oracle.h"

// defined in the reference's EKF.cpp (no header declares it)
void rescueOutliers(const VectorFeatureMatch& outlierMatches, const VectorImageFeaturePrediction& outlierMatchFeaturePrediction,
                    const VectorMatd& outlierMatchFeaturePredictionJacobians, VectorFeatureMatch& rescuedMatches,
                    VectorImageFeaturePrediction& rescuedPredictions, VectorMatd& rescuedJacobians);

// ---- ConfigurationManager singleton (the YAML loader is replaced: parameters come from orc_params) ----
IMPLEMENT_SINGLETON_METHODS(ConfigurationManager)
ConfigurationManager::ConfigurationManager()
    : cameraCalibration(NULL), featureDetector(NULL), descriptorExtractor(NULL), ekfParams(NULL), _camCalibConfig(NULL),
      _ekfConfig(NULL), _featureDetectorConfig(NULL), _descriptorExtractorConfig(NULL) {}
ConfigurationManager::~ConfigurationManager() {}
bool ConfigurationManager::loadConfigurationFromFile(const char*) { return true; }

namespace {

struct InjectedDetector : cv::FeatureDetector {
    std::vector<cv::KeyPoint> kps;
    // OpenCV detectors apply the mask as a post-filter: KeyPointsFilter::runByPixelsMask keeps a keypoint iff
    // mask((int)(y + 0.5f), (int)(x + 0.5f)) != 0
    void detect(const cv::Mat& image, std::vector<cv::KeyPoint>& out, const cv::Mat& mask) const
    {
        (void)image;
        out.clear();
        for (size_t i = 0; i < kps.size(); ++i) {
            if (!mask.empty()) {
                const int yy = (int)(kps[i].pt.y + 0.5f), xx = (int)(kps[i].pt.x + 0.5f);
                if (xx < 0 || xx >= mask.cols || yy < 0 || yy >= mask.rows) continue;
                if (mask.at<uchar>(yy, xx) == 0) continue;
            }
            out.push_back(kps[i]);
        }
    }
};

struct InjectedExtractor : cv::DescriptorExtractor {
    const uint8_t* desc;
    InjectedExtractor() : desc(NULL) {}
    void compute(const cv::Mat& image, std::vector<cv::KeyPoint>& kps, cv::Mat& descriptors) const
    {
        (void)image;
        descriptors = cv::Mat((int)kps.size(), 32, CV_8U);
        for (size_t i = 0; i < kps.size(); ++i) std::memcpy(descriptors.ptr<uchar>((int)i), desc + (size_t)kps[i].class_id * 32, 32);
    }
};

}  // namespace

struct ref_filter {
    CameraCalibration calib;
    ExtendedKalmanFilterParameters ekfp;
    InjectedDetector det;
    InjectedExtractor ext;
    EKF* ekf;
    cv::Mat image;  // only its size is used by the reference (mask allocation)
    // per-frame vectors, as in EKF::step
    VectorImageFeaturePrediction preds;
    VectorMatd jacs;
    VectorMapFeature unseen;
    VectorFeatureMatch matches;
    VectorImageFeaturePrediction matchedPreds;
    VectorMatd matchedJacs;
    VectorFeatureMatch inlierMatches, outlierMatches, rescuedMatches;
    VectorImageFeaturePrediction inlierPreds, rescuedPreds, outlierPreds;
    VectorMatd inlierJacs, rescuedJacs, outlierJacs;
};

static void clear_frame(ref_filter* f)
{
    for (size_t i = 0; i < f->preds.size(); ++i) delete f->preds[i];
    for (size_t i = 0; i < f->jacs.size(); ++i) delete f->jacs[i];
    for (size_t i = 0; i < f->matches.size(); ++i) delete f->matches[i];
    for (size_t i = 0; i < f->outlierPreds.size(); ++i) delete f->outlierPreds[i];
    for (size_t i = 0; i < f->outlierJacs.size(); ++i) delete f->outlierJacs[i];
    f->preds.clear(); f->jacs.clear(); f->unseen.clear(); f->matches.clear(); f->matchedPreds.clear(); f->matchedJacs.clear();
    f->inlierMatches.clear(); f->outlierMatches.clear(); f->rescuedMatches.clear(); f->inlierPreds.clear();
    f->rescuedPreds.clear(); f->outlierPreds.clear(); f->inlierJacs.clear(); f->rescuedJacs.clear(); f->outlierJacs.clear();
}

static void inject(ref_filter* f, const float* kp, const uint8_t* desc, int nkp)
{
    f->det.kps.clear();
    for (int i = 0; i < nkp; ++i) {
        cv::KeyPoint k(kp[2 * i], kp[2 * i + 1], 7.f);
        k.class_id = i;
        f->det.kps.push_back(k);
    }
    f->ext.desc = desc;
}

extern "C" {

ref_filter* ref_create(const orc_params* p)
{
    ref_filter* f = new ref_filter();
    CameraCalibration& c = f->calib;
    c.pixelsX = p->pixels_x; c.pixelsY = p->pixels_y; c.fx = p->fx; c.fy = p->fy; c.k1 = p->k1; c.k2 = p->k2;
    c.cx = p->cx; c.cy = p->cy; c.dx = p->dx; c.dy = p->dy; c.pixelErrorX = p->pixel_error_x; c.pixelErrorY = p->pixel_error_y;
    c.angularVisionX = p->angular_vision_x; c.angularVisionY = p->angular_vision_y;
    ExtendedKalmanFilterParameters& e = f->ekfp;
    e.alwaysRemoveUnseenMapFeatures = false; e.maxMapFeaturesCount = 0; e.maxMapSize = 0;
    e.mapManagementFrequency = 0;  // map management (host-side policy) stays off: fixed map
    e.detectNewFeaturesImageAreasDivideTimes = 2; e.minMatchesPerImage = 0; e.reserveFeaturesDepth = 1024;
    e.reserveFeaturesInvDepth = 1024; e.initInvDepthRho = p->init_inv_depth_rho; e.initLinearAccelSD = p->init_linear_accel_sd;
    e.initAngularAccelSD = p->init_angular_accel_sd; e.linearAccelSD = p->linear_accel_sd; e.angularAccelSD = p->angular_accel_sd;
    e.inverseDepthRhoSD = p->inverse_depth_rho_sd; e.detectNewFeaturesImageMaskEllipseSize = 10;
    e.matchingCompCoefSecondBestVSFirst = p->matching_coef; e.goodFeatureMatchingPercent = 0.5;
    e.ransacThresholdPredictDistance = p->ransac_threshold; e.ransacAllInliersProbability = p->ransac_all_inliers_prob;
    e.ransacChi2Threshold = p->ransac_chi2; e.inverseDepthLinearityIndexThreshold = 0.1;
    ConfigurationManager& cm = ConfigurationManager::getInstance();
    cm.cameraCalibration = &f->calib;
    cm.ekfParams = &f->ekfp;
    cm.featureDetector = &f->det;
    cm.descriptorExtractor = &f->ext;
    f->ekf = new EKF("", "");
    f->image = cv::Mat::zeros(p->pixels_y, p->pixels_x, CV_8UC3);
    return f;
}

void ref_destroy(ref_filter* f)
{
    clear_frame(f);
    delete f->ekf;
    delete f;
}

void ref_set_state(ref_filter* f, int32_t n, int32_t N, const double* x, const int32_t* type, const int32_t* off, const double* P,
                   const uint8_t* desc)
{
    State& s = f->ekf->state;
    s.removeAllFeatures();
    for (int i = 0; i < 3; ++i) {
        s.position[i] = x[i];
        s.linearVelocity[i] = x[7 + i];
        s.angularVelocity[i] = x[10 + i];
    }
    s.setOrientation(x + 3);
    for (int i = 0; i < N; ++i) {
        const int dim = type[i] == MAPFEATURE_TYPE_INVERSE_DEPTH ? 6 : 3;
        cv::Mat d(1, 32, CV_8U);
        std::memcpy(d.ptr<uchar>(0), desc + (size_t)i * 32, 32);
        s.addFeature(new MapFeature(x + off[i], dim, off[i], d, (MapFeatureType)type[i]));
    }
    Matd Pm(n, n);
    for (int i = 0; i < n; ++i) std::memcpy(Pm[i], P + (size_t)i * n, sizeof(double) * n);
    f->ekf->stateCovarianceMatrix = Pm;
}

void ref_dims(const ref_filter* f, int32_t* n, int32_t* N)
{
    *n = f->ekf->stateCovarianceMatrix.rows;
    *N = (int32_t)f->ekf->state.mapFeatures.size();
}

void ref_get_state(const ref_filter* f, double* x, double* P)
{
    const State& s = f->ekf->state;
    const Matd& Pm = f->ekf->stateCovarianceMatrix;
    if (x) {
        for (int i = 0; i < 3; ++i) {
            x[i] = s.position[i];
            x[7 + i] = s.linearVelocity[i];
            x[10 + i] = s.angularVelocity[i];
        }
        for (int i = 0; i < 4; ++i) x[3 + i] = s.orientation[i];
        for (size_t i = 0; i < s.mapFeatures.size(); ++i)
            for (int j = 0; j < s.mapFeatures[i]->positionDimension; ++j)
                x[s.mapFeatures[i]->covarianceMatrixPos + j] = s.mapFeatures[i]->position[j];
    }
    if (P)
        for (int i = 0; i < Pm.rows; ++i) std::memcpy(P + (size_t)i * Pm.cols, Pm[i], sizeof(double) * Pm.cols);
}

void ref_get_features(const ref_filter* f, uint8_t* desc, int32_t* tp, int32_t* tm)
{
    const State& s = f->ekf->state;
    for (size_t i = 0; i < s.mapFeatures.size(); ++i) {
        if (desc) std::memcpy(desc + i * 32, s.mapFeatures[i]->descriptor.ptr<uchar>(0), 32);
        if (tp) tp[i] = (int32_t)s.mapFeatures[i]->timesPredicted;
        if (tm) tm[i] = (int32_t)s.mapFeatures[i]->timesMatched;
    }
}

// the reference's add-feature path (AddMapFeature.cpp:293-350)
void ref_add_feature(ref_filter* f, const double* uv, const uint8_t* desc32)
{
    cv::Mat d(1, 32, CV_8U);
    std::memcpy(d.ptr<uchar>(0), desc32, 32);
    ImageFeatureMeasurement m(uv, d);
    addFeatureToStateAndCovariance(&m, f->ekf->state, f->ekf->stateCovarianceMatrix);
}

// ---- map management: the reference's own functions, glued like EKF.cpp:575-592 ----
void ref_set_policy(ref_filter* f, const orc_map_policy* pol, int32_t map_management_frequency)
{
    ExtendedKalmanFilterParameters& e = f->ekfp;
    e.minMatchesPerImage = pol->min_matches_per_image;
    e.maxMapFeaturesCount = pol->max_map_features_count;
    e.maxMapSize = pol->max_map_size;
    e.alwaysRemoveUnseenMapFeatures = pol->always_remove_unseen != 0;
    e.goodFeatureMatchingPercent = pol->good_feature_matching_percent;
    e.inverseDepthLinearityIndexThreshold = pol->linearity_index_threshold;
    e.mapManagementFrequency = map_management_frequency;   // > 0: ref_step runs the reference's whole map management
}

void ref_get_layout(const ref_filter* f, int32_t* type, int32_t* off)
{
    const State& s = f->ekf->state;
    for (size_t i = 0; i < s.mapFeatures.size(); ++i) {
        type[i] = (int32_t)s.mapFeatures[i]->featureType;
        off[i] = s.mapFeatures[i]->covarianceMatrixPos;
    }
}

void ref_set_counters(ref_filter* f, const int32_t* tp, const int32_t* tm)
{
    State& s = f->ekf->state;
    for (size_t i = 0; i < s.mapFeatures.size(); ++i) {
        s.mapFeatures[i]->timesPredicted = tp[i];
        s.mapFeatures[i]->timesMatched = tm[i];
    }
}

// after ref_measure .. ref_update_map_features of the same frame; returns newFeaturesNeededCount
int32_t ref_map_management(ref_filter* f)
{
    ExtendedKalmanFilterParameters* p = &f->ekfp;
    State& state = f->ekf->state;
    Matd& P = f->ekf->stateCovarianceMatrix;
    int needed = p->minMatchesPerImage - static_cast<int>(f->inlierMatches.size() + f->rescuedMatches.size());
    removeBadMapFeatures(state, P);
    if (needed > 0 && (p->alwaysRemoveUnseenMapFeatures ||
                       (p->maxMapFeaturesCount > 0 && state.mapFeatures.size() + needed > p->maxMapFeaturesCount) ||
                       (p->maxMapSize > 0 && P.rows + needed * 6 > p->maxMapSize)))
        removeFeaturesFromStateAndCovariance(f->unseen, state, P);
    convertMapFeaturesInverseDepthToDepth(state, P);
    return needed;
}

int32_t ref_remove_bad(ref_filter* f)
{
    const size_t before = f->ekf->state.mapFeatures.size();
    removeBadMapFeatures(f->ekf->state, f->ekf->stateCovarianceMatrix);
    return (int32_t)(before - f->ekf->state.mapFeatures.size());
}

void ref_convert(ref_filter* f) { convertMapFeaturesInverseDepthToDepth(f->ekf->state, f->ekf->stateCovarianceMatrix); }

// detectNewImageFeatures (DetectNewImageFeatures.cpp:321-419) on the injected keypoints, with the predictions of the
// last ref_measure; consumes libc rand() exactly as the reference does.  out_kp = indices into the injected keypoints.
int32_t ref_detect_new(ref_filter* f, const float* kp, const uint8_t* desc, int32_t nkp, int32_t max_new, double* out_uv,
                       uint8_t* out_desc)
{
    inject(f, kp, desc, nkp);
    VectorImageFeatureMeasurement found;
    detectNewImageFeatures(f->image, f->preds, (uint)max_new, found);
    for (size_t i = 0; i < found.size(); ++i) {
        out_uv[2 * i] = found[i]->imagePos[0];
        out_uv[2 * i + 1] = found[i]->imagePos[1];
        std::memcpy(out_desc + i * 32, found[i]->descriptorData.ptr<uchar>(0), 32);
        delete found[i];
    }
    return (int32_t)found.size();
}

void ref_init(ref_filter* f)
{
    f->ekf->state.removeAllFeatures();
    initState(f->ekf->state);
    initCovariance(f->ekf->stateCovarianceMatrix);
}

// the reference's own EKF::init (EKF.cpp:170-237) on the injected keypoints: initState, initCovariance,
// detectNewImageFeatures (zone-balanced, libc rand) and addFeaturesToStateAndCovariance
void ref_full_init(ref_filter* f, const float* kp, const uint8_t* desc, int32_t nkp)
{
    clear_frame(f);
    f->ekf->state.removeAllFeatures();
    inject(f, kp, desc, nkp);
    f->ekf->init(f->image);
}

// ---- whole frame through the reference's own orchestrator ----
void ref_step(ref_filter* f, const float* kp, const uint8_t* desc, int32_t nkp)
{
    inject(f, kp, desc, nkp);
    f->ekf->step(f->image);
}

// ---- phase by phase: the reference's free functions, glued exactly like EKF::step (EKF.cpp:273-572) ----
void ref_predict(ref_filter* f) { stateAndCovariancePrediction(f->ekf->state, f->ekf->stateCovarianceMatrix); }

void ref_measure(ref_filter* f)
{
    clear_frame(f);
    std::vector<int> none;
    predictCameraMeasurements(f->ekf->state, f->ekf->stateCovarianceMatrix, f->ekf->state.mapFeatures, none, f->preds, f->jacs,
                              f->unseen);
}

void ref_match(ref_filter* f, const float* kp, const uint8_t* desc, int32_t nkp)
{
    inject(f, kp, desc, nkp);
    matchPredictedFeatures(f->image, f->ekf->state.mapFeatures, f->preds, f->matches);
    for (size_t i = 0; i < f->matches.size(); ++i)  // EKF.cpp:368-392
        for (size_t j = 0; j < f->preds.size(); ++j)
            if (f->preds[j]->featureIndex == f->matches[i]->featureIndex) {
                f->matchedPreds.push_back(f->preds[j]);
                f->matchedJacs.push_back(f->jacs[j]);
                break;
            }
}

void ref_ransac(ref_filter* f)
{
    ransac(f->ekf->state, f->ekf->stateCovarianceMatrix, f->matchedPreds, f->matchedJacs, f->matches, f->inlierMatches, f->inlierPreds,
           f->inlierJacs, f->outlierMatches);
}

void ref_update_li(ref_filter* f) { update(f->ekf->state, f->ekf->stateCovarianceMatrix, f->inlierMatches, f->inlierPreds, f->inlierJacs); }

void ref_rescue(ref_filter* f)  // EKF.cpp:448-506
{
    VectorMapFeature outlierFeatures, unseenOutliers;
    std::vector<int> idx;
    for (size_t i = 0; i < f->outlierMatches.size(); ++i) {
        outlierFeatures.push_back(f->ekf->state.mapFeatures[f->outlierMatches[i]->featureIndex]);
        idx.push_back(f->outlierMatches[i]->featureIndex);
    }
    predictCameraMeasurements(f->ekf->state, f->ekf->stateCovarianceMatrix, outlierFeatures, idx, f->outlierPreds, f->outlierJacs,
                              unseenOutliers);
    const size_t np = f->outlierPreds.size(), no = f->outlierMatches.size();
    if (0 < np && np < no) {
        VectorFeatureMatch kept;
        size_t j = 0;
        for (size_t i = 0; i < no && j < np; ++i)
            if (f->outlierMatches[i]->featureIndex == f->outlierPreds[j]->featureIndex) {
                j++;
                kept.push_back(f->outlierMatches[i]);
            }
        f->outlierMatches = kept;
    }
    if (f->outlierMatches.size() && np > 0)
        rescueOutliers(f->outlierMatches, f->outlierPreds, f->outlierJacs, f->rescuedMatches, f->rescuedPreds, f->rescuedJacs);
}

void ref_update_hi(ref_filter* f)
{
    if (f->rescuedMatches.size()) update(f->ekf->state, f->ekf->stateCovarianceMatrix, f->rescuedMatches, f->rescuedPreds, f->rescuedJacs);
}

void ref_update_map_features(ref_filter* f)
{
    VectorFeatureMatch all = f->inlierMatches;
    for (size_t i = 0; i < f->rescuedMatches.size(); ++i) all.push_back(f->rescuedMatches[i]);
    updateMapFeatures(f->preds, all, f->ekf->state);
}

// ---- per-feature results (same layout as the oracle's getters) ----
void ref_get_measure(const ref_filter* f, uint8_t* vis, double* h, double* S, double* Hx, double* Hf)
{
    const State& s = f->ekf->state;
    const size_t N = s.mapFeatures.size();
    std::memset(vis, 0, N);
    for (size_t k = 0; k < f->preds.size(); ++k) {
        const ImageFeaturePrediction* p = f->preds[k];
        const int i = p->featureIndex;
        const MapFeature* mf = s.mapFeatures[i];
        vis[i] = 1;
        h[2 * i] = p->imagePos[0]; h[2 * i + 1] = p->imagePos[1];
        for (int a = 0; a < 4; ++a) S[4 * i + a] = p->covarianceMatrix[a / 2][a % 2];
        const Matd& J = *f->jacs[k];
        for (int r = 0; r < 2; ++r) {
            for (int j = 0; j < 7; ++j) Hx[14 * i + 7 * r + j] = J[r][j];
            for (int j = 0; j < 6; ++j) Hf[12 * i + 6 * r + j] = j < mf->positionDimension ? J[r][mf->covarianceMatrixPos + j] : 0.0;
        }
    }
}

void ref_get_match(const ref_filter* f, uint8_t* matched, double* z, float* dist)
{
    const size_t N = f->ekf->state.mapFeatures.size();
    std::memset(matched, 0, N);
    for (size_t k = 0; k < f->matches.size(); ++k) {
        const int i = f->matches[k]->featureIndex;
        matched[i] = 1;
        z[2 * i] = f->matches[k]->imagePos[0]; z[2 * i + 1] = f->matches[k]->imagePos[1];
        dist[i] = f->matches[k]->distance;
    }
}

void ref_get_sets(const ref_filter* f, uint8_t* inlier, uint8_t* outlier, uint8_t* rescued)
{
    const size_t N = f->ekf->state.mapFeatures.size();
    std::memset(inlier, 0, N); std::memset(outlier, 0, N); std::memset(rescued, 0, N);
    for (size_t k = 0; k < f->inlierMatches.size(); ++k) inlier[f->inlierMatches[k]->featureIndex] = 1;
    for (size_t k = 0; k < f->outlierMatches.size(); ++k) outlier[f->outlierMatches[k]->featureIndex] = 1;
    for (size_t k = 0; k < f->rescuedMatches.size(); ++k) rescued[f->rescuedMatches[k]->featureIndex] = 1;
}

}  // extern "C"
