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

// everything EKF reads from the configuration file: camera + filter parameters (ekfb_params), the map-management policy
// (ekfb_map_policy) and the host-side fields of ExtendedKalmanFilterParameters (.../ExtendedKalmanFilterParameters.h:37-76)
struct EkfHostConfig {
    ekfb_params params;
    ekfb_map_policy policy;
    int mapManagementFrequency;                     // MapManagementFrequency
    int detectNewFeaturesImageAreasDivideTimes;     // DetectNewFeaturesImageAreasDivideTimes
    double detectNewFeaturesImageMaskEllipseSize;   // DetectNewFeaturesImageMaskEllipseSize
};

typedef void (*EkfbDrawFn)(void* user, unsigned char* mask, int W, int H, double x, double y);

// one "Frame k" record of output.yml (E/EKF.cpp:257-268 ... 618-628)
struct EkfbFrameTrace {
    double usPrediction, usMatching, usRansac, usUpdateLI, usRescue, usUpdateHI, usMapManagement;
    int totalMatches, liInliers, hiInliers, invDepthCount, depthCount;
    double state[13];
    double cov[169];
};

class EkfbTraceWriter {
public:
    EkfbTraceWriter();
    ~EkfbTraceWriter();
    bool open(const std::string& path);
    bool isOpen() const { return _f != nullptr; }
    void frame(int step, const EkfbFrameTrace& t);
    void close();

private:
    void* _f;
    EkfbTraceWriter(const EkfbTraceWriter&);
    EkfbTraceWriter& operator=(const EkfbTraceWriter&);
};

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

    // Attributes (E/EKF.h:50-51).  After every step(): state (camera, every feature's position, type, covarianceMatrixPos,
    // descriptor, timesPredicted / timesMatched) and the 13x13 camera block of stateCovarianceMatrix are current.
    // DIFFERENCE FROM THE REFERENCE: the rest of stateCovarianceMatrix lives on the GPU; syncCovariance() downloads the
    // full n x n matrix when a caller needs it (the two shipped callers read only the camera block).
    Matd stateCovarianceMatrix;
    State state;

    // ---- additions (no counterpart in the reference's header) ----
    void setFrontEnd(FrontEnd* fe) { _frontEnd = fe; }
    // detector + descriptor on the GPU instead of a host FrontEnd object: FAST-9/16 (threshold) + 3x3 NMS + BRIEF-style
    // descriptor (ekfb_set_image / ekfb_detect_keypoints).  init / step then need the frame itself: CV_8UC1, CV_8UC3 (BGR,
    // desktop: FileSequenceImageGenerator.cpp:82) or CV_8UC4 (Android: EKFNative.cpp:134-137), converted to grey on the
    // device like cv::cvtColor(COLOR_BGR2GRAY).
    void useDeviceFrontEnd(int fastThreshold) { _deviceFrontEnd = true; _fastThreshold = fastThreshold; }
    void setDevice(int device) { _device = device; }
    void syncCovariance();
    const ekfb_frame_info& lastFrameInfo() const { return _info; }
    ekfb_handle handle() const { return _h; }
    // healthy: a device handle exists and the last init / step completed without an error (lastStatus() == EKFB_OK)
    bool ok() const { return _h != nullptr && _lastStatus == 0; }
    // ekfb status of the last init / step: EKFB_OK, or the code of the call that failed (CUDA error, capacity, wrong frame
    // size ...), or EKFB_ERR_NUMERIC when an innovation covariance of the frame was not positive definite.  A failed frame
    // leaves the step counter where it was and skips map management.
    int lastStatus() const { return _lastStatus; }
    // test hook: after every step() append the frame's per-feature results (numbering before map management) to a binary
    // file: int32 N, then matched[N], inlier[N], rescued[N] (uint8), keypoint index[N] (int32), z[N][2] (double)
    bool dumpFrameSetsTo(const char* path);
    int capacityFeatures() const { return _maxFeatures; }

    const ekfb_map_result& lastMapResult() const { return _mapResult; }
    int lastNewFeatures() const { return _lastAdded; }

private:
    void downloadState();
    void mirrorLayout();
    void refreshMirror(bool layout);
    bool acquireKeypoints(const cv::Mat& image);
    int addNewFeatures(int wanted, bool useDeviceMask);
    bool growCapacity(int minFeatures);
    void fail(int status, const char* where);
    void dumpFrameSets();
    int _ekfSteps;
    std::string _strOutputPath;
    EkfHostConfig _cfg;
    int _maxFeatures, _device;
    bool _configOk;
    FrontEnd* _frontEnd;
    bool _deviceFrontEnd;
    int _fastThreshold;
    ekfb_handle _h;
    ekfb_frame_info _info;
    ekfb_map_result _mapResult;
    int _lastAdded, _featuresBefore;
    std::vector<EkfKeyPoint> _kps;
    std::vector<unsigned char> _desc;
    std::vector<unsigned char> _mask, _stamp;
    int _stampR;
    EkfbTraceWriter _trace;
    void* _logFile;   // FILE*: log.txt (E/EKF.cpp:135-136)
    void writeLogState();
    int _lastStatus;
    void* _setDump;   // FILE*
};

// reads the reference's YAML 1.0 configuration (experiments/s3/config.yml, kalmanFilter/samples/EKF/config.yml)
bool ekfbLoadConfig(const char* fileName, ekfb_params* params, int* minMatchesPerImage, int* maxMapSize);
bool ekfbLoadConfigFull(const char* fileName, EkfHostConfig* config);

// detectNewImageFeatures after the detector call (E/DetectNewImageFeatures.cpp:172-419): keypoints are filtered by
// `mask` (W x H, 0 = masked) the way the detector's mask argument does, then -- if more than maxNew survive -- picked zone by
// zone with libc rand() so that the (2^divideTimes)^2 image zones end up evenly populated; `predXY` are the predicted
// features of the frame (zone occupancy).  After each accepted feature `draw` blacks out its neighbourhood in `mask`.
// Returns the count; outIdx = indices into kpXY in selection order.
int ekfbSelectNewFeatures(int W, int H, int divideTimes, unsigned char* mask, const float* kpXY, int nKp, const double* predXY,
                          int nPred, int maxNew, EkfbDrawFn draw, void* user, int* outIdx);
// blacks out a (2R+1)^2 stamp (0 = ellipse pixel) centred on the truncated pixel of (x, y), clipped to the image
void ekfbStampEllipse(const unsigned char* stamp, int R, unsigned char* mask, int W, int H, double x, double y);

// Host-side addFeatureToStateAndCovariance for one inverse-depth feature (AddMapFeature.cpp:43-350): appends the six
// feature rows to x and grows the row-major n x n covariance P to (n+6) x (n+6); n is updated.
void ekfbAddInverseDepthFeature(const ekfb_params& camera, const double* uv, std::vector<double>& x, std::vector<double>& P, int& n);

#endif
