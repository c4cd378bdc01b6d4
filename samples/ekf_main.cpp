// ekf_main.cpp -- the reference's sample driver (kalmanFilter/samples/EKF/main.cpp:45-177) against the drop-in EKF class:
//     ekf_sample config.yml frames.kpseq [outDir/]
// The reference reads numbered PNG frames and runs STAR + BRIEF on them; OpenCV is not available here, so the frame
// source is a keypoint-sequence file (one record per frame: int32 count, count x (float x, float y), count x 32 bytes),
// i.e. the front-end output stored on disk -- the same seam as the reference's HandMatching.cpp.  Frame 0 initialises
// the filter (EKF::init), every further frame is one EKF::step; the 13-state is printed per frame.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "../include/EKF.h"

class FileFrontEnd : public FrontEnd {
public:
    explicit FileFrontEnd(const char* path) : _in(path, std::ios::binary) {}
    bool good() const { return _in.good(); }
    // returns false at end of sequence
    bool next()
    {
        int32_t count = 0;
        if (!_in.read(reinterpret_cast<char*>(&count), 4) || count < 0) return false;
        _kps.resize(count);
        _desc.resize((size_t)count * 32);
        if (count) {
            _in.read(reinterpret_cast<char*>(_kps.data()), (std::streamsize)count * 8);
            _in.read(reinterpret_cast<char*>(_desc.data()), (std::streamsize)count * 32);
        }
        return (bool)_in;
    }
    void detectAndDescribe(const cv::Mat&, std::vector<EkfKeyPoint>& kps, std::vector<unsigned char>& desc)
    {
        kps = _kps;
        desc = _desc;
    }

private:
    std::ifstream _in;
    std::vector<EkfKeyPoint> _kps;
    std::vector<unsigned char> _desc;
};

int main(int argc, const char* argv[])
{
    if (argc < 3) {
        std::cerr << "usage: ekf_sample config.yml frames.kpseq [outDir/]" << std::endl;
        return 2;
    }
    FileFrontEnd frames(argv[2]);
    if (!frames.good()) {
        std::cerr << "cannot open " << argv[2] << std::endl;
        return 2;
    }
    EKF extendedKalmanFilter(argv[1], argc > 3 ? argv[3] : "");
    extendedKalmanFilter.setFrontEnd(&frames);
    if (const char* dump = std::getenv("EKFB_DUMP_SETS")) extendedKalmanFilter.dumpFrameSetsTo(dump);   // parity tests
    cv::Mat image;  // frames carry no pixels here; the filter only needs the front-end output
    if (!frames.next()) {
        std::cout << "No se puede iniciar Kalman Filter dado que no hay imagenes disponibles." << std::endl;
        return 0;
    }
    extendedKalmanFilter.init(image);
    if (!extendedKalmanFilter.ok()) return 1;
    int stepCount = 0;
    while (frames.next()) {
        extendedKalmanFilter.step(image);
        if (extendedKalmanFilter.lastStatus() != 0) {
            std::cerr << "frame " << stepCount + 1 << ": status " << extendedKalmanFilter.lastStatus() << std::endl;
            if (!extendedKalmanFilter.handle() || extendedKalmanFilter.lastStatus() != EKFB_ERR_NUMERIC) return 1;
        }
        const State& s = extendedKalmanFilter.state;
        const ekfb_frame_info& fi = extendedKalmanFilter.lastFrameInfo();
        int32_t nNow = 0, NNow = 0;   // state dimension after this frame's map management
        ekfb_get_dims(extendedKalmanFilter.handle(), 0, &nNow, &NNow);
        std::printf("STEP %d matches %d inliers %d rescued %d x", ++stepCount, fi.n_matches, fi.n_inliers, fi.n_rescued);
        for (int i = 0; i < 3; ++i) std::printf(" %.17g", s.position[i]);
        for (int i = 0; i < 4; ++i) std::printf(" %.17g", s.orientation[i]);
        for (int i = 0; i < 3; ++i) std::printf(" %.17g", s.linearVelocity[i]);
        for (int i = 0; i < 3; ++i) std::printf(" %.17g", s.angularVelocity[i]);
        std::printf(" P00 %.17g N %d n %d removed %d converted %d added %d\n", extendedKalmanFilter.stateCovarianceMatrix[0][0],
                    (int)s.mapFeatures.size(), nNow, extendedKalmanFilter.lastMapResult().n_removed_bad +
                    extendedKalmanFilter.lastMapResult().n_removed_unseen, extendedKalmanFilter.lastMapResult().converted,
                    extendedKalmanFilter.lastNewFeatures());
    }
    return 0;
}
