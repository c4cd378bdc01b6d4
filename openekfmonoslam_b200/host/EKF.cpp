// EKF.cpp -- host side of the drop-in EKF class (include/EKF.h): configuration, initial map, per-frame calls into
// libekf_b200 (include/ekf_b200.h).  Replaces kalmanFilter/modules/1PointRansacEKF/EKF.cpp:124-666; the per-frame
// mathematics is on the GPU, the host keeps the reference's public state object in sync.
#include "../../include/EKF.h"

#include <cmath>
#include <cstring>
#include <iostream>

namespace {
const double kEpsilon = 2.22e-16;  // modules/Core/EKFMath.h:37

void quatToRot(const double* q, double* R)  // modules/Core/EKFMath.cpp:121-141
{
    const double r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = r * r + x * x - y * y - z * z; R[1] = 2 * (x * y - r * z); R[2] = 2 * (z * x + r * y);
    R[3] = 2 * (x * y + r * z); R[4] = r * r - x * x + y * y - z * z; R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (z * x - r * y); R[7] = 2 * (y * z + r * x); R[8] = r * r - x * x - y * y + z * z;
}

// d(R(q) a)/dq, 3x4 (E/CommonFunctions.cpp:87-145)
void dRotDq(const double* q, const double* a, double* J)
{
    const double w = q[0], x = q[1], y = q[2], z = q[3], ax = a[0], ay = a[1], az = a[2];
    J[0] = 2 * (w * ax - z * ay + y * az); J[4] = 2 * (z * ax + w * ay - x * az); J[8] = 2 * (-y * ax + x * ay + w * az);
    J[1] = 2 * (x * ax + y * ay + z * az); J[5] = 2 * (y * ax - x * ay - w * az); J[9] = 2 * (z * ax + w * ay - x * az);
    J[2] = 2 * (-y * ax + x * ay + w * az); J[6] = 2 * (x * ax + y * ay + z * az); J[10] = 2 * (-w * ax + z * ay - y * az);
    J[3] = 2 * (-z * ax - w * ay + x * az); J[7] = 2 * (w * ax - z * ay + y * az); J[11] = 2 * (x * ax + y * ay + z * az);
}

// Append one inverse-depth feature observed at pixel uv to (x, P) (row-major n x n), the reference's
// addFeatureToStateAndCovariance (E/AddMapFeature.cpp:43-350): y = (r, theta, phi, rho0), P grows by 6 rows/cols with
// the initialisation Jacobians.
}  // namespace

void ekfbAddInverseDepthFeature(const ekfb_params& c, const double* uv, std::vector<double>& x, std::vector<double>& P, int& n)
{
    const double px = uv[0] - c.cx, py = uv[1] - c.cy;
    const double mx = c.dx * px, my = c.dy * py, rd = mx * mx + my * my;
    const double dist = 1 + c.k1 * rd + c.k2 * rd * rd;
    const double und[2] = {c.cx + px * dist, c.cy + py * dist};
    const double* q = &x[3];
    double R[9];
    quatToRot(q, R);
    const double gc[3] = {-(c.cx - und[0]) / c.fx, -(c.cy - und[1]) / c.fy, 1.0};
    double gw[3];
    for (int i = 0; i < 3; ++i) gw[i] = R[3 * i] * gc[0] + R[3 * i + 1] * gc[1] + R[3 * i + 2] * gc[2];
    const double xw = gw[0], yw = gw[1], zw = gw[2];
    const double y6[6] = {x[0], x[1], x[2], atan2(xw, zw), atan2(-yw, sqrt(xw * xw + zw * zw)), c.init_inv_depth_rho};
    // Jacobians (E/AddMapFeature.cpp:116-216)
    const double xxzz = xw * xw + zw * zw, sq = sqrt(xxzz), nsq = xxzz + yw * yw;
    const double dth[3] = {zw / xxzz, 0.0, -xw / xxzz};
    const double dph[3] = {xw * yw / (nsq * sq), -sq / nsq, zw * yw / (nsq * sq)};
    double dgw_dq[12];
    dRotDq(q, gc, dgw_dq);
    double J[42] = {0}, JH[18] = {0};
    J[0] = J[8] = J[16] = 1.0;
    for (int i = 0; i < 4; ++i) {
        double a = 0, b = 0;
        for (int k = 0; k < 3; ++k) { a += dth[k] * dgw_dq[k * 4 + i]; b += dph[k] * dgw_dq[k * 4 + i]; }
        J[3 * 7 + 3 + i] = a;
        J[4 * 7 + 3 + i] = b;
    }
    double sub[6];
    for (int j = 0; j < 3; ++j) {
        double a = 0, b = 0;
        for (int k = 0; k < 3; ++k) { a += dth[k] * R[k * 3 + j]; b += dph[k] * R[k * 3 + j]; }
        sub[j] = a; sub[3 + j] = b;
    }
    const double s2[4] = {sub[0] / c.fx, sub[1] / c.fy, sub[3] / c.fx, sub[4] / c.fy};
    const double k12 = c.k1 + 2.0 * c.k2 * rd, k1p = 1.0 + c.k1 * rd + c.k2 * rd * rd;
    const double dx2 = 2.0 * c.dx * c.dx, dy2 = 2.0 * c.dy * c.dy;
    const double dhu[4] = {k1p + px * k12 * (px * dx2), px * k12 * (py * dy2), py * k12 * (px * dx2), py * k12 * (py * dy2) + k1p};
    JH[9] = s2[0] * dhu[0] + s2[1] * dhu[2];  JH[10] = s2[0] * dhu[1] + s2[1] * dhu[3];
    JH[12] = s2[2] * dhu[0] + s2[3] * dhu[2]; JH[13] = s2[2] * dhu[1] + s2[3] * dhu[3];
    JH[17] = 1.0;
    const double noise[3] = {c.pixel_error_x * c.pixel_error_x, c.pixel_error_y * c.pixel_error_y,
                             c.inverse_depth_rho_sd * c.inverse_depth_rho_sd};
    const int m = n + 6;
    std::vector<double> Pn((size_t)m * m, 0.0);
    for (int i = 0; i < n; ++i) std::memcpy(&Pn[(size_t)i * m], &P[(size_t)i * n], sizeof(double) * n);
    for (int r = 0; r < 6; ++r)          // new-vs-previous = J * P[0:7, :]  and its mirror block P[:, 0:7] * J^T
        for (int j = 0; j < n; ++j) {
            double a = 0, b = 0;
            for (int k = 0; k < 7; ++k) { a += J[r * 7 + k] * P[(size_t)k * n + j]; b += P[(size_t)j * n + k] * J[r * 7 + k]; }
            Pn[(size_t)(n + r) * m + j] = a;
            Pn[(size_t)j * m + n + r] = b;
        }
    for (int r = 0; r < 6; ++r)
        for (int s = 0; s < 6; ++s) {
            double a = 0;
            for (int k = 0; k < 7; ++k) a += Pn[(size_t)(n + r) * m + k] * J[s * 7 + k];
            double b = 0;
            for (int k = 0; k < 3; ++k) b += JH[r * 3 + k] * noise[k] * JH[s * 3 + k];
            Pn[(size_t)(n + r) * m + n + s] = a + b;
        }
    P.swap(Pn);
    for (int i = 0; i < 6; ++i) x.push_back(y6[i]);
    n = m;
}

// C hooks for the CPU-side tests of the host logic (tests/test_host_ekf.py): no device calls
extern "C" int ekfb_host_load_config(const char* file, ekfb_params* p, int* minMatches, int* maxMapSize)
{
    return ekfbLoadConfig(file, p, minMatches, maxMapSize) ? 0 : 1;
}
extern "C" void ekfb_host_add_feature(const ekfb_params* p, const double* uv, const double* xIn, const double* PIn, int n,
                                      double* xOut, double* POut)
{
    std::vector<double> x(xIn, xIn + n), P(PIn, PIn + (size_t)n * n);
    ekfbAddInverseDepthFeature(*p, uv, x, P, n);
    std::memcpy(xOut, x.data(), sizeof(double) * n);
    std::memcpy(POut, P.data(), sizeof(double) * (size_t)n * n);
}

State::State()
{
    for (int i = 0; i < 3; ++i) position[i] = linearVelocity[i] = angularVelocity[i] = 0.0;
    const double q0[4] = {0, 0, 0, 0};
    setOrientation(q0);
}
State::~State() { removeAllFeatures(); }
void State::setOrientation(const double* q)  // E/State.cpp:131-139
{
    for (int i = 0; i < 4; ++i) orientation[i] = q[i];
    quatToRot(orientation, orientationRotationMatrix);
}
void State::removeAllFeatures()
{
    for (size_t i = 0; i < mapFeatures.size(); ++i) delete mapFeatures[i];
    mapFeatures.clear(); mapFeaturesDepth.clear(); mapFeaturesInvDepth.clear();
}

EKF::EKF(const char* configurationFileName, const char* outputPath)
    : _ekfSteps(0), _strOutputPath(outputPath ? outputPath : ""), _minMatchesPerImage(0), _maxFeatures(0), _device(0),
      _configOk(false), _frontEnd(nullptr), _h(nullptr)
{
    std::memset(&_params, 0, sizeof(_params));
    std::memset(&_info, 0, sizeof(_info));
    int maxMapSize = 0;
    _configOk = ekfbLoadConfig(configurationFileName, &_params, &_minMatchesPerImage, &maxMapSize);
    if (!_configOk) std::cerr << "EKF: could not load configuration " << configurationFileName << std::endl;
    // capacity in features: MaxMapSize is in rows of the state (E/EKF.cpp:582-584); without it allow 4x the target
    _maxFeatures = maxMapSize > 13 ? (maxMapSize - 13) / 3 : 4 * (_minMatchesPerImage > 0 ? _minMatchesPerImage : 64);
    if (!_strOutputPath.empty())
        std::cerr << "EKF: output traces (output.yml, log.txt, overlays) are not written by this build" << std::endl;
}

EKF::~EKF()
{
    if (_h) ekfb_destroy(_h);
}

void EKF::uploadState(const std::vector<double>& P, int n)
{
    const int N = (int)state.mapFeatures.size();
    std::vector<double> x(n, 0.0);
    std::vector<int32_t> type(N), off(N);
    std::vector<unsigned char> desc((size_t)N * 32);
    for (int i = 0; i < 3; ++i) { x[i] = state.position[i]; x[7 + i] = state.linearVelocity[i]; x[10 + i] = state.angularVelocity[i]; }
    for (int i = 0; i < 4; ++i) x[3 + i] = state.orientation[i];
    for (int i = 0; i < N; ++i) {
        const MapFeature* f = state.mapFeatures[i];
        type[i] = f->featureType; off[i] = f->covarianceMatrixPos;
        for (int j = 0; j < f->positionDimension; ++j) x[f->covarianceMatrixPos + j] = f->position[j];
        std::memcpy(&desc[(size_t)i * 32], f->descriptor, 32);
    }
    if (ekfb_set_state(_h, 0, n, N, x.data(), type.data(), off.data(), P.data(), desc.data()) != EKFB_OK)
        std::cerr << "EKF: " << ekfb_last_error() << std::endl;
}

void EKF::downloadState()
{
    int32_t n = 0, N = 0;
    ekfb_get_dims(_h, 0, &n, &N);
    std::vector<double> x(n);
    if (ekfb_get_state(_h, 0, x.data(), nullptr, 0) != EKFB_OK) { std::cerr << "EKF: " << ekfb_last_error() << std::endl; return; }
    for (int i = 0; i < 3; ++i) { state.position[i] = x[i]; state.linearVelocity[i] = x[7 + i]; state.angularVelocity[i] = x[10 + i]; }
    state.setOrientation(&x[3]);
    for (int i = 0; i < N; ++i) {
        MapFeature* f = state.mapFeatures[i];
        for (int j = 0; j < f->positionDimension; ++j) f->position[j] = x[f->covarianceMatrixPos + j];
    }
    ekfb_record rec;
    if (ekfb_get_records(_h, &rec) == EKFB_OK) {
        _info = rec.info;
        for (int i = 0; i < 13; ++i)
            for (int j = 0; j < 13; ++j) stateCovarianceMatrix[i][j] = rec.P_cam[i * 13 + j];
    }
}

void EKF::syncCovariance()
{
    if (!_h) return;
    int32_t n = 0, N = 0;
    ekfb_get_dims(_h, 0, &n, &N);
    std::vector<double> P((size_t)n * n);
    if (ekfb_get_state(_h, 0, nullptr, P.data(), 0) != EKFB_OK) return;
    if (stateCovarianceMatrix.rows != n) stateCovarianceMatrix = Matd(n, n);
    for (int i = 0; i < n; ++i) std::memcpy(stateCovarianceMatrix[i], &P[(size_t)i * n], sizeof(double) * n);
}

void EKF::init(const cv::Mat& image)  // E/EKF.cpp:170-237
{
    if (!_configOk || !_frontEnd) {
        std::cerr << "EKF::init: no configuration or no front end" << std::endl;
        return;
    }
    if (!_h && ekfb_create(&_params, _device, 1, _maxFeatures, 16384, &_h) != EKFB_OK) {
        std::cerr << "EKF::init: " << ekfb_last_error() << std::endl;
        _h = nullptr;
        return;
    }
    // initState / initCovariance (E/CommonFunctions.cpp:39-80)
    std::vector<double> x(13, 0.0), P(169, 0.0);
    x[3] = 1.0; x[10] = x[11] = x[12] = kEpsilon;
    for (int i = 0; i < 7; ++i) P[i * 13 + i] = kEpsilon;
    for (int i = 0; i < 3; ++i) {
        P[(7 + i) * 13 + 7 + i] = _params.init_linear_accel_sd * _params.init_linear_accel_sd;
        P[(10 + i) * 13 + 10 + i] = _params.init_angular_accel_sd * _params.init_angular_accel_sd;
    }
    state.removeAllFeatures();
    for (int i = 0; i < 3; ++i) { state.position[i] = 0; state.linearVelocity[i] = 0; state.angularVelocity[i] = kEpsilon; }
    state.setOrientation(&x[3]);
    // new features: the front end's keypoints, in its order, up to MinMatchesPerImage (the reference balances them over
    // a 4x4 zone grid with libc rand(), E/DetectNewImageFeatures.cpp:172-419 -- front-end policy, out of the hot path)
    _frontEnd->detectAndDescribe(image, _kps, _desc);
    int want = _minMatchesPerImage > 0 ? _minMatchesPerImage : (int)_kps.size();
    if (want > _maxFeatures) want = _maxFeatures;
    int n = 13;
    for (int i = 0; i < (int)_kps.size() && i < want; ++i) {
        const double uv[2] = {_kps[i].x, _kps[i].y};
        MapFeature* f = new MapFeature();
        f->featureType = MAPFEATURE_TYPE_INVERSE_DEPTH;
        f->positionDimension = 6;
        f->covarianceMatrixPos = n;
        f->timesPredicted = f->timesMatched = 0;
        std::memcpy(f->descriptor, &_desc[(size_t)i * 32], 32);
        ekfbAddInverseDepthFeature(_params, uv, x, P, n);
        std::memcpy(f->position, &x[n - 6], sizeof(double) * 6);
        state.mapFeatures.push_back(f);
        state.mapFeaturesInvDepth.push_back(f);
    }
    stateCovarianceMatrix = Matd(n, n);
    for (int i = 0; i < n; ++i) std::memcpy(stateCovarianceMatrix[i], &P[(size_t)i * n], sizeof(double) * n);
    uploadState(P, n);
}

void EKF::step(const cv::Mat& image)  // E/EKF.cpp:242-666
{
    if (!_h || !_frontEnd) {
        std::cerr << "EKF::step: filter not initialised" << std::endl;
        return;
    }
    _ekfSteps++;
    _frontEnd->detectAndDescribe(image, _kps, _desc);
    std::vector<float> xy(_kps.size() * 2);
    for (size_t i = 0; i < _kps.size(); ++i) { xy[2 * i] = _kps[i].x; xy[2 * i + 1] = _kps[i].y; }
    if (ekfb_set_keypoints(_h, 0, xy.data(), _desc.data(), (int)_kps.size()) != EKFB_OK || ekfb_step(_h) != EKFB_OK) {
        std::cerr << "EKF::step: " << ekfb_last_error() << std::endl;
        return;
    }
    downloadState();
}
