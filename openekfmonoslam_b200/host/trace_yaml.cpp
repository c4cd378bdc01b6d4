// trace_yaml.cpp -- the on-disk trace of a run: output.yml in the layout the reference writes through cv::FileStorage
// (kalmanFilter/modules/1PointRansacEKF/EKF.cpp:129-143,257-268,291,344,410-416,437,513-516,539,618-628 and
// State::write, State.cpp:339-360), so that kalmanFilter/resultReader/main.cpp:82-150 reads runs of this library:
//   Frame k: { Prediction, Matching, Ransac, totalMatches, liInliers, UpdateLI, RescueOutliers, hiInliers, UpdateHI,
//              MapManagement, StateEstimation (1x13), MapFeaturesInvDepthCount, MapFeaturesDepthCount,
//              StateCovarianceMatrixEstimation (13x13) }
// Numbers are formatted as OpenCV's YAML emitter does (integers as "12.", everything else "%.16e").
#include <cmath>
#include <cstdio>

#include "../../include/EKF.h"

namespace {
void putDouble(FILE* f, double v)
{
    if (std::isfinite(v) && v == std::floor(v) && std::fabs(v) < 2147483647.0) std::fprintf(f, "%d.", (int)v);
    else if (std::isnan(v)) std::fprintf(f, ".Nan");
    else if (std::isinf(v)) std::fprintf(f, v < 0 ? "-.Inf" : ".Inf");
    else std::fprintf(f, "%.16e", v);
}

void putMatrix(FILE* f, const char* name, int rows, int cols, const double* data)
{
    std::fprintf(f, "   %s: !!opencv-matrix\n      rows: %d\n      cols: %d\n      dt: d\n      data: [ ", name, rows, cols);
    for (int i = 0; i < rows * cols; ++i) {
        putDouble(f, data[i]);
        if (i + 1 < rows * cols) std::fprintf(f, (i % 3 == 2) ? ",\n          " : ", ");
    }
    std::fprintf(f, " ]\n");
}
}  // namespace

EkfbTraceWriter::EkfbTraceWriter() : _f(nullptr) {}
EkfbTraceWriter::~EkfbTraceWriter() { close(); }

bool EkfbTraceWriter::open(const std::string& path)
{
    close();
    _f = std::fopen(path.c_str(), "w");
    if (!_f) return false;
    std::fprintf((FILE*)_f, "%%YAML:1.0\n");
    return true;
}

void EkfbTraceWriter::close()
{
    if (_f) std::fclose((FILE*)_f);
    _f = nullptr;
}

void EkfbTraceWriter::frame(int step, const EkfbFrameTrace& t)
{
    if (!_f) return;
    FILE* f = (FILE*)_f;
    std::fprintf(f, "Frame %d:\n   # \n   # Running time (microseconds)\n   # \n", step);
    const char* names[7] = {"Prediction", "Matching", "Ransac", "UpdateLI", "RescueOutliers", "UpdateHI", "MapManagement"};
    const double us[7] = {t.usPrediction, t.usMatching, t.usRansac, t.usUpdateLI, t.usRescue, t.usUpdateHI, t.usMapManagement};
    for (int i = 0; i < 7; ++i) {
        std::fprintf(f, "   %s: ", names[i]);
        putDouble(f, us[i]);
        std::fprintf(f, "\n");
        if (i == 2) std::fprintf(f, "   totalMatches: %d\n   liInliers: %d\n", t.totalMatches, t.liInliers);
        if (i == 4) std::fprintf(f, "   hiInliers: %d\n", t.hiInliers);
    }
    std::fprintf(f, "   # \n   # State and Covariance Estimation\n   # \n");
    putMatrix(f, "StateEstimation", 1, 13, t.state);
    std::fprintf(f, "   MapFeaturesInvDepthCount: %d\n   MapFeaturesDepthCount: %d\n", t.invDepthCount, t.depthCount);
    putMatrix(f, "StateCovarianceMatrixEstimation", 13, 13, t.cov);
    std::fflush(f);
}

// C hooks for the CPU-side test
extern "C" void* ekfb_host_trace_open(const char* path)
{
    EkfbTraceWriter* w = new EkfbTraceWriter();
    if (!w->open(path)) { delete w; return nullptr; }
    return w;
}
extern "C" void ekfb_host_trace_frame(void* w, int step, const EkfbFrameTrace* t) { ((EkfbTraceWriter*)w)->frame(step, *t); }
extern "C" void ekfb_host_trace_close(void* w) { delete (EkfbTraceWriter*)w; }
