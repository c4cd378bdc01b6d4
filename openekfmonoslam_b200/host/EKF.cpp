// EKF.cpp -- host side of the drop-in EKF class (include/EKF.h): configuration, initial map, per-frame calls into
// libekf_b200 (include/ekf_b200.h).  Replaces kalmanFilter/modules/1PointRansacEKF/EKF.cpp:124-666; the per-frame
// mathematics is on the GPU, the host keeps the reference's public state object in sync.
#include "../../include/EKF.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
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
extern "C" int ekfb_host_load_config_full(const char* file, EkfHostConfig* out) { return ekfbLoadConfigFull(file, out) ? 0 : 1; }
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
    : _ekfSteps(0), _strOutputPath(outputPath ? outputPath : ""), _maxFeatures(0), _device(0), _configOk(false),
      _frontEnd(nullptr), _deviceFrontEnd(false), _fastThreshold(20), _h(nullptr), _lastAdded(0), _featuresBefore(0), _stampR(0),
      _logFile(nullptr), _lastStatus(0), _setDump(nullptr)
{
    std::memset(&_info, 0, sizeof(_info));
    std::memset(&_mapResult, 0, sizeof(_mapResult));
    _configOk = ekfbLoadConfigFull(configurationFileName, &_cfg);
    if (!_configOk) std::cerr << "EKF: could not load configuration " << configurationFileName << std::endl;
    // INITIAL capacity in features: MaxMapSize is in rows of the state (E/EKF.cpp:582-584) and a frame may add up to
    // MinMatchesPerImage features on top of it before the next removal; without a limit start at 4x the target.  The
    // reference's map has no upper bound (E/EKF.cpp:594-611 adds every feature it asked for), so when a frame's new
    // features do not fit the handle is re-created with a larger capacity (growCapacity) -- never truncated.
    const int minM = _cfg.policy.min_matches_per_image > 0 ? _cfg.policy.min_matches_per_image : 64;
    if (_cfg.policy.max_map_features_count > 0) _maxFeatures = _cfg.policy.max_map_features_count + minM;
    else if (_cfg.policy.max_map_size > 13) _maxFeatures = (_cfg.policy.max_map_size - 13) / 3 + minM;
    else _maxFeatures = 4 * minM;
    if (!_strOutputPath.empty()) {
        // E/EKF.cpp:129-143: output.yml and log.txt in outputPath.  The per-frame PNG overlays and videoOutput.mpg are GUI
        // artefacts (modules/Gui) and are not written by this build.
        if (!_trace.open(_strOutputPath + "output.yml"))
            std::cerr << "EKF: cannot write " << _strOutputPath << "output.yml" << std::endl;
        _logFile = std::fopen((_strOutputPath + "log.txt").c_str(), "w");
    }
}

EKF::~EKF()
{
    if (_h) ekfb_destroy(_h);
    if (_setDump) std::fclose((FILE*)_setDump);
    if (_logFile) std::fclose((FILE*)_logFile);
}

// log.txt: the state after init and after every step, in the layout of State::showDetailed (E/State.cpp:229-240,371-399,
// E/MapFeature.cpp:130-145): camera position, quaternion, Euler angles (C/EKFMath.cpp:355-365), velocities, the map size,
// then one line per map feature with its position and its (timesMatched/timesPredicted) counters.  %g = the six significant
// digits of an ostream's default formatting.
void EKF::writeLogState()
{
    FILE* f = (FILE*)_logFile;
    if (!f) return;
    const double* q = state.orientation;
    const double roll = std::atan2(2 * (q[0] * q[1] + q[2] * q[3]), 1 - 2 * (q[1] * q[1] + q[2] * q[2]));
    const double pitch = std::asin(2 * (q[0] * q[2] - q[3] * q[1]));
    const double yaw = std::atan2(2 * (q[0] * q[3] + q[1] * q[2]), 1 - 2 * (q[2] * q[2] + q[3] * q[3]));
    std::fprintf(f, "Posicion de la camara: %g, %g, %g\n", state.position[0], state.position[1], state.position[2]);
    std::fprintf(f, "Orientacion(cuaternions): %g, %g, %g, %g\n", q[0], q[1], q[2], q[3]);
    std::fprintf(f, "Orientacion en angulos eulerianos: %g, %g, %g\n", roll, pitch, yaw);
    std::fprintf(f, "Velocidad lineal (con respecto al mundo): %g, %g, %g\n", state.linearVelocity[0], state.linearVelocity[1],
                 state.linearVelocity[2]);
    std::fprintf(f, "Velocidad angular (con respecto a la camara): %g, %g, %g\n", state.angularVelocity[0], state.angularVelocity[1],
                 state.angularVelocity[2]);
    std::fprintf(f, "Cantidad de features en el mapa: %zu\n\n", state.mapFeatures.size());
    std::fprintf(f, "Map Features (%zu):\n", state.mapFeatures.size());
    for (size_t i = 0; i < state.mapFeatures.size(); ++i) {
        const MapFeature* m = state.mapFeatures[i];
        std::fprintf(f, "%zu: ", i);
        for (int a = 0; a < m->positionDimension; ++a) std::fprintf(f, a ? ", %g" : "%g", m->position[a]);
        std::fprintf(f, " (%u/%u)\n", m->timesMatched, m->timesPredicted);
    }
    std::fflush(f);
}

void EKF::fail(int status, const char* where)
{
    _lastStatus = status ? status : EKFB_ERR_ARG;
    std::cerr << where << ": " << ekfb_last_error() << std::endl;
}

bool EKF::dumpFrameSetsTo(const char* path)
{
    if (_setDump) std::fclose((FILE*)_setDump);
    _setDump = path ? (void*)std::fopen(path, "wb") : nullptr;
    return _setDump != nullptr;
}

void EKF::dumpFrameSets()
{
    int32_t n = 0, N = 0;
    ekfb_get_dims(_h, 0, &n, &N);
    const int M = N > 0 ? N : 1;
    std::vector<unsigned char> matched(M), inl(M), resc(M);
    std::vector<int32_t> kp(M);
    std::vector<double> z(2 * (size_t)M);
    if (ekfb_get_feature_results(_h, 0, nullptr, nullptr, nullptr, nullptr, nullptr, matched.data(), z.data(), kp.data(), nullptr,
                                 inl.data(), nullptr, resc.data()) != EKFB_OK)
        return;
    FILE* f = (FILE*)_setDump;
    std::fwrite(&N, 4, 1, f);
    std::fwrite(matched.data(), 1, N, f); std::fwrite(inl.data(), 1, N, f); std::fwrite(resc.data(), 1, N, f);
    std::fwrite(kp.data(), 4, N, f); std::fwrite(z.data(), 8, 2 * (size_t)N, f);
    std::fflush(f);
}

// The reference's map grows without bound; the device handle has a fixed capacity.  Re-create it with room for at least
// minFeatures features and carry the whole filter over (state, covariance, layout, descriptors, hit counters).
bool EKF::growCapacity(int minFeatures)
{
    int32_t n = 0, N = 0;
    ekfb_get_dims(_h, 0, &n, &N);
    const int newMax = std::max(minFeatures, 2 * _maxFeatures);
    std::vector<double> x(n), P((size_t)n * n);
    const int M = N > 0 ? N : 1;
    std::vector<int32_t> type(M), off(M), tp(M), tm(M);
    std::vector<unsigned char> desc((size_t)M * 32);
    int rc = ekfb_get_state(_h, 0, x.data(), P.data(), 0);
    if (rc == EKFB_OK) rc = ekfb_get_map_snapshot(_h, 0, nullptr, type.data(), off.data(), desc.data(), tp.data(), tm.data(), nullptr);
    ekfb_handle h2 = nullptr;
    if (rc == EKFB_OK) rc = ekfb_create(&_cfg.params, _device, 1, newMax, 16384, &h2);
    if (rc == EKFB_OK) rc = ekfb_set_state(h2, 0, n, N, x.data(), type.data(), off.data(), P.data(), desc.data());
    if (rc == EKFB_OK && N > 0) rc = ekfb_set_hit_counters(h2, 0, tp.data(), tm.data());
    if (rc != EKFB_OK) {
        if (h2) ekfb_destroy(h2);
        fail(rc, "EKF: could not grow the map capacity");
        return false;
    }
    std::cerr << "EKF: map capacity grown from " << _maxFeatures << " to " << newMax << " features" << std::endl;
    ekfb_destroy(_h);
    _h = h2;
    _maxFeatures = newMax;
    return true;
}

// Bring the host mirror in line with the device in ONE round trip (ekfb_get_map_snapshot): camera state and every feature's
// position, the 13x13 block, and -- after features were removed, converted or added -- the feature list itself (what
// State::removeFeatures / addFeature / convertToDepth do to the reference's vectors: E/State.cpp:142-206,
// E/MapManagement.cpp:457-496).  MapFeature objects are reused by index; surplus ones are deleted.
void EKF::downloadState() { refreshMirror(false); }
void EKF::mirrorLayout() { refreshMirror(true); }

void EKF::refreshMirror(bool layout)
{
    int32_t n = 0, N = 0;
    ekfb_get_dims(_h, 0, &n, &N);
    std::vector<double> x(n);
    const int M = N > 0 ? N : 1;
    std::vector<int32_t> type, off, tp, tm;
    std::vector<unsigned char> desc;
    if (layout) { type.resize(M); off.resize(M); tp.resize(M); tm.resize(M); desc.resize((size_t)M * 32); }
    ekfb_record rec;
    if (ekfb_get_map_snapshot(_h, 0, x.data(), layout ? type.data() : nullptr, layout ? off.data() : nullptr,
                              layout ? desc.data() : nullptr, layout ? tp.data() : nullptr, layout ? tm.data() : nullptr, &rec) != EKFB_OK) {
        std::cerr << "EKF: " << ekfb_last_error() << std::endl;
        return;
    }
    if (layout) {
        while ((int)state.mapFeatures.size() > N) { delete state.mapFeatures.back(); state.mapFeatures.pop_back(); }
        while ((int)state.mapFeatures.size() < N) state.mapFeatures.push_back(new MapFeature());
        state.mapFeaturesDepth.clear();
        state.mapFeaturesInvDepth.clear();
        for (int i = 0; i < N; ++i) {
            MapFeature* f = state.mapFeatures[i];
            f->featureType = (MapFeatureType)type[i];
            f->positionDimension = type[i] == MAPFEATURE_TYPE_INVERSE_DEPTH ? 6 : 3;
            f->covarianceMatrixPos = off[i];
            f->timesPredicted = (unsigned)tp[i];
            f->timesMatched = (unsigned)tm[i];
            std::memcpy(f->descriptor, &desc[(size_t)i * 32], 32);
            (type[i] == MAPFEATURE_TYPE_INVERSE_DEPTH ? state.mapFeaturesInvDepth : state.mapFeaturesDepth).push_back(f);
        }
    }
    for (int i = 0; i < 3; ++i) { state.position[i] = x[i]; state.linearVelocity[i] = x[7 + i]; state.angularVelocity[i] = x[10 + i]; }
    state.setOrientation(&x[3]);
    for (int i = 0; i < N && i < (int)state.mapFeatures.size(); ++i) {
        MapFeature* f = state.mapFeatures[i];
        for (int j = 0; j < f->positionDimension; ++j) f->position[j] = x[f->covarianceMatrixPos + j];
    }
    if (stateCovarianceMatrix.rows < 13) stateCovarianceMatrix = Matd(13, 13);
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) stateCovarianceMatrix[i][j] = rec.P_cam[i * 13 + j];
    _info.status = rec.info.status;   // the record is read after the high-innovation update
}

namespace {
struct DrawCtx {
    ekfb_handle h;
    const std::vector<unsigned char>* stamp;
    int R, maxAxes;
    double ellipseSize;
};
// drawUncertaintyEllipse2D(mask, (x, y), diag(size, size), 2 (cols + rows), black, filled) (E/DetectNewImageFeatures.cpp:286-291):
// the pixel set is the same for every centre as long as the ellipse lies inside the image, so it is stamped from the
// copy the device rasterised once; an ellipse that crosses the border is rasterised by the device on the mask itself.
void drawNewFeatureEllipse(void* user, unsigned char* mask, int W, int H, double x, double y)
{
    const DrawCtx* c = (const DrawCtx*)user;
    const int cx = (int)(float)x, cy = (int)(float)y;
    if (cx - c->R >= 0 && cx + c->R < W && cy - c->R >= 0 && cy + c->R < H) {
        ekfbStampEllipse(c->stamp->data(), c->R, mask, W, H, x, y);
    } else {
        const double S[4] = {c->ellipseSize, 0.0, 0.0, c->ellipseSize};
        if (ekfb_raster_ellipse(c->h, W, H, x, y, S, c->maxAxes, 0, mask) != EKFB_OK)
            std::cerr << "EKF: " << ekfb_last_error() << std::endl;
    }
}
}  // namespace

// detectNewImageFeatures + addFeaturesToStateAndCovariance (E/EKF.cpp:196-222 at init, :594-611 per frame) on the
// keypoints of the current frame (_kps / _desc).  Returns the number of features added.
int EKF::addNewFeatures(int wanted, bool useDeviceMask)
{
    const int W = _cfg.params.pixels_x, H = _cfg.params.pixels_y;
    int32_t n = 0, N = 0;
    ekfb_get_dims(_h, 0, &n, &N);
    if (wanted <= 0 || _kps.empty()) return 0;
    _mask.assign((size_t)W * H, 255);
    std::vector<double> pred;
    if (useDeviceMask) {
        // one round trip: this frame's predictions (zone occupancy; indexed by the numbering before the map changed) + the mask
        const int nb = _featuresBefore;
        std::vector<unsigned char> vis(nb > 0 ? nb : 1);
        std::vector<double> h(2 * (size_t)(nb > 0 ? nb : 1));
        if (ekfb_get_new_feature_inputs(_h, 0, nb, vis.data(), h.data(), _mask.data()) != EKFB_OK) {
            std::cerr << "EKF: " << ekfb_last_error() << std::endl;
            return 0;
        }
        for (int i = 0; i < nb; ++i)
            if (vis[i]) { pred.push_back(h[2 * i]); pred.push_back(h[2 * i + 1]); }
    }
    if (_stamp.empty()) {   // the ellipse of E/DetectNewImageFeatures.cpp:221-223, rasterised once by the device
        const double es = _cfg.detectNewFeaturesImageMaskEllipseSize;
        _stampR = (int)(2.0 * std::sqrt(es * 5.9915)) + 2;
        const int D = 2 * _stampR + 1;
        _stamp.assign((size_t)D * D, 255);
        const double S[4] = {es, 0.0, 0.0, es};
        if (ekfb_raster_ellipse(_h, D, D, _stampR, _stampR, S, 2 * (W + H), 0, _stamp.data()) != EKFB_OK) {
            std::cerr << "EKF: " << ekfb_last_error() << std::endl;
            _stamp.clear();
            return 0;
        }
    }
    std::vector<float> xy(_kps.size() * 2);
    for (size_t i = 0; i < _kps.size(); ++i) { xy[2 * i] = _kps[i].x; xy[2 * i + 1] = _kps[i].y; }
    std::vector<int> idx(wanted);
    DrawCtx ctx = {_h, &_stamp, _stampR, 2 * (W + H), _cfg.detectNewFeaturesImageMaskEllipseSize};
    const int k = ekfbSelectNewFeatures(W, H, _cfg.detectNewFeaturesImageAreasDivideTimes, _mask.data(), xy.data(), (int)_kps.size(),
                                        pred.data(), (int)pred.size() / 2, wanted, &drawNewFeatureEllipse, &ctx, idx.data());
    if (k == 0) return 0;
    std::vector<double> uv(2 * k);
    std::vector<unsigned char> dd((size_t)32 * k);
    for (int a = 0; a < k; ++a) {
        uv[2 * a] = _kps[idx[a]].x;
        uv[2 * a + 1] = _kps[idx[a]].y;
        std::memcpy(&dd[(size_t)32 * a], &_desc[(size_t)32 * idx[a]], 32);
    }
    if (N + k > _maxFeatures && !growCapacity(N + k + std::max(_cfg.policy.min_matches_per_image, 16))) return 0;
    const int rcAdd = ekfb_add_features(_h, 0, k, uv.data(), dd.data());
    if (rcAdd != EKFB_OK) {
        fail(rcAdd, "EKF: ekfb_add_features");
        return 0;
    }
    return k;
}

// The frame's front-end output: either from the host FrontEnd object (uploaded with ekfb_set_keypoints) or detected and
// described on the device from the image (then only copied back for the new-feature selection).
bool EKF::acquireKeypoints(const cv::Mat& image)
{
    if (!_deviceFrontEnd) {
        _frontEnd->detectAndDescribe(image, _kps, _desc);
        std::vector<float> xy(_kps.size() * 2);
        for (size_t i = 0; i < _kps.size(); ++i) { xy[2 * i] = _kps[i].x; xy[2 * i + 1] = _kps[i].y; }
        return ekfb_set_keypoints(_h, 0, xy.data(), _desc.data(), (int)_kps.size()) == EKFB_OK;
    }
    const int W = _cfg.params.pixels_x, H = _cfg.params.pixels_y;
    if (image.empty() || image.cols != W || image.rows != H) {
        std::cerr << "EKF: the device front end needs a " << W << "x" << H << " frame" << std::endl;
        return false;
    }
    // the frame goes to the device as it is (grey, BGR or BGRA); the grey conversion of cv::cvtColor runs there
    int32_t n = 0;
    if (ekfb_set_image_color(_h, 0, image.data, (int)image.step, (int)image.elemSize()) != EKFB_OK ||
        ekfb_detect_keypoints(_h, 0, _fastThreshold, &n) != EKFB_OK)
        return false;
    _kps.resize(n);
    _desc.resize((size_t)n * 32);
    return ekfb_get_keypoints(_h, 0, reinterpret_cast<float*>(_kps.data()), _desc.data()) == EKFB_OK;
}

void EKF::init(const cv::Mat& image)  // E/EKF.cpp:170-237
{
    _lastStatus = EKFB_OK;
    if (_logFile) {
        // E/EKF.cpp:172-180: the reference re-seeds rand() from time() whenever it logs, which makes a logged run choose other
        // new features than an unlogged one.  Here that is opt-in (EKFB_LOG_RANDOM_SEED=1): by default a run is
        // reproducible whether or not it writes log.txt, and the line says which seed stream is in use.
        const char* reseed = std::getenv("EKFB_LOG_RANDOM_SEED");
        if (reseed && reseed[0] == '1') {
            const time_t seed = time(nullptr);
            srand(static_cast<unsigned int>(seed));
            std::fprintf((FILE*)_logFile, "Random Seed: %lld\n\n", (long long)seed);
        } else
            std::fprintf((FILE*)_logFile, "Random Seed: (the caller's rand() stream, not re-seeded)\n\n");
        std::fprintf((FILE*)_logFile, "~~~~~~~~~~~~ STEP %d ~~~~~~~~~~~~\n", _ekfSteps);
    }
    if (!_configOk || (!_frontEnd && !_deviceFrontEnd)) {
        std::cerr << "EKF::init: no configuration or no front end" << std::endl;
        _lastStatus = EKFB_ERR_ARG;
        return;
    }
    int rc = EKFB_OK;
    if (!_h && (rc = ekfb_create(&_cfg.params, _device, 1, _maxFeatures, 16384, &_h)) != EKFB_OK) {
        _h = nullptr;
        fail(rc, "EKF::init");
        return;
    }
    // initState / initCovariance (E/CommonFunctions.cpp:39-80)
    std::vector<double> x(13, 0.0), P(169, 0.0);
    x[3] = 1.0; x[10] = x[11] = x[12] = kEpsilon;
    for (int i = 0; i < 7; ++i) P[i * 13 + i] = kEpsilon;
    for (int i = 0; i < 3; ++i) {
        P[(7 + i) * 13 + 7 + i] = _cfg.params.init_linear_accel_sd * _cfg.params.init_linear_accel_sd;
        P[(10 + i) * 13 + 10 + i] = _cfg.params.init_angular_accel_sd * _cfg.params.init_angular_accel_sd;
    }
    state.removeAllFeatures();
    if ((rc = ekfb_set_state(_h, 0, 13, 0, x.data(), nullptr, nullptr, P.data(), nullptr)) != EKFB_OK) {
        fail(rc, "EKF::init");
        return;
    }
    // detectNewImageFeatures(image, noPredictions, MinMatchesPerImage) + addFeaturesToStateAndCovariance, on the device
    if (!acquireKeypoints(image)) {
        fail(EKFB_ERR_ARG, "EKF::init (front end)");
        return;
    }
    _lastAdded = addNewFeatures(_cfg.policy.min_matches_per_image, false);
    refreshMirror(true);
    writeLogState();
}

void EKF::step(const cv::Mat& image)  // E/EKF.cpp:242-666
{
    if (!_h || (!_frontEnd && !_deviceFrontEnd)) {
        std::cerr << "EKF::step: filter not initialised" << std::endl;
        _lastStatus = EKFB_ERR_ARG;
        return;
    }
    _lastStatus = EKFB_OK;
    if (!acquireKeypoints(image)) {
        fail(EKFB_ERR_ARG, "EKF::step (front end)");
        return;
    }
    int rc = EKFB_OK;
    float ms[8] = {0};
    if (_trace.isOpen()) {
        // phase by phase with device timers, the seven intervals the reference writes (E/EKF.cpp:291 ... 618)
        ekfb_timer_record(_h, 0);
        if ((rc = ekfb_predict(_h)) == EKFB_OK) rc = ekfb_measure(_h);
        ekfb_timer_record(_h, 1);
        if (rc == EKFB_OK) rc = ekfb_match(_h);
        ekfb_timer_record(_h, 2);
        if (rc == EKFB_OK) rc = ekfb_ransac(_h);
        ekfb_timer_record(_h, 3);
        if (rc == EKFB_OK) rc = ekfb_update(_h, 0);
        ekfb_timer_record(_h, 4);
        if (rc == EKFB_OK) rc = ekfb_rescue(_h);
        ekfb_timer_record(_h, 5);
        if (rc == EKFB_OK) rc = ekfb_update(_h, 1);
        ekfb_timer_record(_h, 6);
        if (rc == EKFB_OK) rc = ekfb_update_map_features(_h);
    } else {
        rc = ekfb_step(_h);
    }
    if (rc != EKFB_OK) {
        fail(rc, "EKF::step");
        return;
    }
    _ekfSteps++;   // a failed frame does not count (the reference has no failure path: E/EKF.cpp:242)
    if (_logFile) std::fprintf((FILE*)_logFile, "\n\n~~~~~~~~~~~~ STEP %d ~~~~~~~~~~~~\n", _ekfSteps);   // E/EKF.cpp:246-250
    ekfb_peek_frame_info(_h, 0, &_info);   // the counters ekfb_step synchronised; status is refreshed from the record below
    if (_setDump) dumpFrameSets();
    // map management (E/EKF.cpp:572-612): bad / unseen features out, one conversion, new features in
    std::memset(&_mapResult, 0, sizeof(_mapResult));
    _mapResult.converted = -1;
    _lastAdded = 0;
    int numericStatus = 0;
    if (_cfg.mapManagementFrequency > 0 && _ekfSteps % _cfg.mapManagementFrequency == 0) {
        _featuresBefore = (int)state.mapFeatures.size();
        if ((rc = ekfb_map_management(_h, &_cfg.policy, &_mapResult)) != EKFB_OK) {
            fail(rc, "EKF::step (map management)");
            return;
        }
        ekfb_frame_info now;
        ekfb_peek_frame_info(_h, 0, &now);     // ekfb_map_management synchronised the counters: status covers both updates
        numericStatus = now.status;
        if (_mapResult.new_features_needed > 0) _lastAdded = addNewFeatures(_mapResult.new_features_needed, true);
        if (_lastStatus != EKFB_OK) return;
    }
    if (_trace.isOpen()) ekfb_timer_record(_h, 7);
    refreshMirror(true);
    if (numericStatus) _info.status = numericStatus;
    if (_info.status != 0) {
        _lastStatus = _info.status;
        std::cerr << "EKF::step: frame " << _ekfSteps << " finished with status " << _info.status
                  << " (innovation covariance not positive definite: that update was skipped)" << std::endl;
    }
    if (_trace.isOpen()) {
        for (int i = 0; i < 7; ++i) ekfb_timer_elapsed_ms(_h, i, i + 1, &ms[i]);
        EkfbFrameTrace t;
        t.usPrediction = 1e3 * ms[0]; t.usMatching = 1e3 * ms[1]; t.usRansac = 1e3 * ms[2]; t.usUpdateLI = 1e3 * ms[3];
        t.usRescue = 1e3 * ms[4]; t.usUpdateHI = 1e3 * ms[5]; t.usMapManagement = 1e3 * ms[6];
        t.totalMatches = _info.n_matches; t.liInliers = _info.n_inliers; t.hiInliers = _info.n_rescued;
        t.invDepthCount = (int)state.mapFeaturesInvDepth.size(); t.depthCount = (int)state.mapFeaturesDepth.size();
        for (int i = 0; i < 3; ++i) { t.state[i] = state.position[i]; t.state[7 + i] = state.linearVelocity[i]; t.state[10 + i] = state.angularVelocity[i]; }
        for (int i = 0; i < 4; ++i) t.state[3 + i] = state.orientation[i];
        for (int i = 0; i < 13; ++i)
            for (int j = 0; j < 13; ++j) t.cov[i * 13 + j] = stateCovarianceMatrix[i][j];
        _trace.frame(_ekfSteps, t);
    }
    writeLogState();   // E/EKF.cpp:662-665
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


// ---- flat C binding of the class: what a JNI / ctypes / cgo layer binds (android/EKFMonoSlam/jni/EKFNative.cpp:79-218 keeps an
// EKF* in a Java long and calls init / step with the camera frame, then reads state.position / orientation) ----
extern "C" void* ekfb_host_ekf_create(const char* config, const char* outputPath, int device, int fastThreshold)
{
    EKF* e = new EKF(config, outputPath);
    e->setDevice(device);
    if (fastThreshold > 0) e->useDeviceFrontEnd(fastThreshold);
    return e;
}
extern "C" void ekfb_host_ekf_destroy(void* e) { delete (EKF*)e; }
// frames: 8-bit, `channels` = 1 (grey), 3 (BGR) or 4 (BGRA), tightly packed rows.  Both return the ekfb status of the call
// (EKFB_OK = 0; a failed frame reports the code of the failing device call, EKFB_ERR_NUMERIC a non-positive-definite update)
extern "C" int ekfb_host_ekf_init(void* e, const unsigned char* frame, int width, int height, int channels)
{
    cv::Mat img(height, width, channels == 1 ? CV_8UC1 : channels == 3 ? CV_8UC3 : CV_8UC4, const_cast<unsigned char*>(frame));
    ((EKF*)e)->init(img);
    return ((EKF*)e)->handle() ? ((EKF*)e)->lastStatus() : EKFB_ERR_CUDA;
}
extern "C" int ekfb_host_ekf_step(void* e, const unsigned char* frame, int width, int height, int channels)
{
    cv::Mat img(height, width, channels == 1 ? CV_8UC1 : channels == 3 ? CV_8UC3 : CV_8UC4, const_cast<unsigned char*>(frame));
    ((EKF*)e)->step(img);
    return ((EKF*)e)->handle() ? ((EKF*)e)->lastStatus() : EKFB_ERR_CUDA;
}
// x13 = r, q, v, omega; counts = {map features, state dimension, matches, inliers, rescued, removed, converted (-1 none), added}
extern "C" void ekfb_host_ekf_get(void* ep, double* x13, int* counts8)
{
    EKF* e = (EKF*)ep;
    const State& s = e->state;
    for (int i = 0; i < 3; ++i) { x13[i] = s.position[i]; x13[7 + i] = s.linearVelocity[i]; x13[10 + i] = s.angularVelocity[i]; }
    for (int i = 0; i < 4; ++i) x13[3 + i] = s.orientation[i];
    int32_t n = 0, N = 0;
    if (e->handle()) ekfb_get_dims(e->handle(), 0, &n, &N);
    const ekfb_frame_info& fi = e->lastFrameInfo();
    const ekfb_map_result& mr = e->lastMapResult();
    const int c[8] = {(int)s.mapFeatures.size(), n, fi.n_matches, fi.n_inliers, fi.n_rescued, mr.n_removed_bad + mr.n_removed_unseen,
                      mr.converted, e->lastNewFeatures()};
    for (int i = 0; i < 8; ++i) counts8[i] = c[i];
}
