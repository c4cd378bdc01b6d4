// EKF.h -- drop-in for the reference's public filter class (kalmanFilter/modules/1PointRansacEKF/EKF.h:41-63) whose
// per-frame body runs on a B200 through the C ABI of ekf_b200.h.
//
// Same surface as the reference: EKF(configurationFileName, outputPath), init(image), step(image) and the public data
// members `stateCovarianceMatrix` and `state` that the callers read (kalmanFilter/samples/EKF/main.cpp:91,137,141;
// android/EKFMonoSlam/jni/EKFNative.cpp:177,191-193).  Errors follow the reference: methods return void, problems go
// to std::cerr (E/EKF.cpp:126-127 ignores a failed configuration load as well).
//
// Two things the reference takes from its process-global ConfigurationManager are per-object here:
//   * the front end.  The reference calls cv::FeatureDetector::detect + cv::DescriptorExtractor::compute on the frame
//     (E/Matching.cpp:204-215, E/DetectNewImageFeatures.cpp:343-347); here a FrontEnd object supplied with
//     setFrontEnd() does that on the host.  (STAR / BRIEF need OpenCV; see samples/ekf_main.cpp for the
//     keypoints-from-file front end, the seam the reference's own HandMatching.cpp:37-99 uses.)
//   * parameters (ekfb_params), read from the reference's YAML files by a small parser (config_yaml.cpp).
#ifndef EKFB_EKF_H
#define EKFB_EKF_H

#include <cstdint>
#include <string>
#include <vector>

#include "ekf_b200.h"
#include "ekfb_cv_compat.hpp"

// E/MapFeature.h:39-78
enum MapFeatureType { MAPFEATURE_TYPE_INVALID, MAPFEATURE_TYPE_DEPTH, MAPFEATURE_TYPE_INVERSE_DEPTH };

class MapFeature {
public:
    MapFeatureType featureType;
    double position[6];
    int positionDimension;
    int covarianceMatrixPos;
    unsigned char descriptor[32];
    unsigned int timesPredicted, timesMatched;
};
typedef std::vector<MapFeature*> VectorMapFeature;

// E/State.h:40-81 (same member names; arrays instead of five heap blocks)
class State {
public:
    State();
    ~State();
    void setOrientation(const double* q);
    void removeAllFeatures();
    double position[3];
    double orientation[4];
    double orientationRotationMatrix[9];
    double linearVelocity[3];
    double angularVelocity[3];
    VectorMapFeature mapFeaturesDepth, mapFeaturesInvDepth, mapFeatures;

private:
    State(const State&);
    State& operator=(const State&);
};

struct EkfKeyPoint { float x, y; };

// host front end: keypoints + 32-byte binary descriptors of one frame
class FrontEnd {
public:
    virtual ~FrontEnd() {}
    virtual void detectAndDescribe(const cv::Mat& image, std::vector<EkfKeyPoint>& keypoints,
                                   std::vector<unsigned char>& descriptors /* 32 bytes each */) = 0;
};

class EKF {
public:
    EKF(const char* configurationFileName, const char* outputPath);
    ~EKF();

    void init(const cv::Mat& image);
    void step(const cv::Mat& image);

    // Attributes (E/EKF.h:50-51).  After every step(): state (camera + all feature positions) and the 13x13 camera
    // block of stateCovarianceMatrix are current; syncCovariance() downloads the full matrix when a caller needs it.
    Matd stateCovarianceMatrix;
    State state;

    // ---- additions (no counterpart in the reference's header) ----
    void setFrontEnd(FrontEnd* fe) { _frontEnd = fe; }
    void setDevice(int device) { _device = device; }
    void syncCovariance();
    const ekfb_frame_info& lastFrameInfo() const { return _info; }
    ekfb_handle handle() const { return _h; }
    bool ok() const { return _h != nullptr; }

private:
    void uploadState(const std::vector<double>& P, int n);
    void downloadState();
    int _ekfSteps;
    std::string _strOutputPath;
    ekfb_params _params;
    int _minMatchesPerImage, _maxFeatures, _device;
    bool _configOk;
    FrontEnd* _frontEnd;
    ekfb_handle _h;
    ekfb_frame_info _info;
    std::vector<EkfKeyPoint> _kps;
    std::vector<unsigned char> _desc;
};

// reads the reference's YAML 1.0 configuration (experiments/s3/config.yml, kalmanFilter/samples/EKF/config.yml)
bool ekfbLoadConfig(const char* fileName, ekfb_params* params, int* minMatchesPerImage, int* maxMapSize);

// Host-side addFeatureToStateAndCovariance for one inverse-depth feature (AddMapFeature.cpp:43-350): appends the six
// feature rows to x and grows the row-major n x n covariance P to (n+6) x (n+6); n is updated.
void ekfbAddInverseDepthFeature(const ekfb_params& camera, const double* uv, std::vector<double>& x, std::vector<double>& P, int& n);

#endif
