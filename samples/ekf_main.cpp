// ekf_main.cpp -- the reference's sample driver (kalmanFilter/samples/EKF/main.cpp:45-177) against the drop-in EKF class:
//     ekf_sample config.yml frames.kpseq [outDir/]        keypoint-sequence file (front-end output stored on disk)
//     ekf_sample config.yml framesDir/   [outDir/]        numbered image files, as the reference's driver reads them
// Image mode (second argument ends in '/'): FileSequenceImageGenerator(framesDir, "", "png", 90, 6550) like main.cpp:50 -- the
// extension and the index range can be changed with EKFB_SEQ_EXT / EKFB_SEQ_BEGIN / EKFB_SEQ_END -- and the detector +
// descriptor run on the GPU (EKF::useDeviceFrontEnd: FAST-9/16 with threshold EKFB_FAST_THRESHOLD, default 20, + the
// repository's 256-bit descriptor; the reference's STAR + BRIEF need OpenCV 2.4).
// Keypoint mode: one record per frame (int32 count, count x (float x, float y), count x 32 bytes) -- the same seam as the
// reference's HandMatching.cpp.  Frame 0 initialises the filter (EKF::init), every further frame is one EKF::step; the
// 13-state is printed per frame.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include <string>

#include "../include/EKF.h"
#include "../include/ImageGenerator.h"

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

static void print_step(EKF& extendedKalmanFilter, int stepCount);

static int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

// the reference's main loop (main.cpp:50-177): generator -> init on the first image -> step on every further one
static int run_image_sequence(const char* config, const char* dir, const char* outDir)
{
    const char* ext = std::getenv("EKFB_SEQ_EXT");
    FileSequenceImageGenerator generator(dir, "", ext ? ext : "png", env_int("EKFB_SEQ_BEGIN", 90), env_int("EKFB_SEQ_END", 6550));
    generator.init();
    EKF extendedKalmanFilter(config, outDir);
    extendedKalmanFilter.useDeviceFrontEnd(env_int("EKFB_FAST_THRESHOLD", 20));
    cv::Mat image = generator.getNextImage();
    if (image.empty()) {
        std::cout << "No se puede iniciar Kalman Filter dado que no hay imagenes disponibles." << std::endl;
        return 0;
    }
    extendedKalmanFilter.init(image);
    if (!extendedKalmanFilter.ok()) return 1;
    int stepCount = 0;
    image = generator.getNextImage();
    while (!image.empty()) {
        extendedKalmanFilter.step(image);
        if (extendedKalmanFilter.lastStatus() != 0 &&
            (!extendedKalmanFilter.handle() || extendedKalmanFilter.lastStatus() != EKFB_ERR_NUMERIC))
            return 1;
        print_step(extendedKalmanFilter, ++stepCount);
        image = generator.getNextImage();
    }
    return 0;
}

int main(int argc, const char* argv[])
{
    if (argc < 3) {
        std::cerr << "usage: ekf_sample config.yml frames.kpseq|framesDir/ [outDir/]" << std::endl;
        return 2;
    }
    const std::string source(argv[2]);
    if (!source.empty() && source[source.size() - 1] == '/') return run_image_sequence(argv[1], argv[2], argc > 3 ? argv[3] : "");
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
        print_step(extendedKalmanFilter, ++stepCount);
    }
    return 0;
}

static void print_step(EKF& extendedKalmanFilter, int stepCount)
{
    const State& s = extendedKalmanFilter.state;
    const ekfb_frame_info& fi = extendedKalmanFilter.lastFrameInfo();
    int32_t nNow = 0, NNow = 0;   // state dimension after this frame's map management
    ekfb_get_dims(extendedKalmanFilter.handle(), 0, &nNow, &NNow);
    std::printf("STEP %d matches %d inliers %d rescued %d x", stepCount, fi.n_matches, fi.n_inliers, fi.n_rescued);
    for (int i = 0; i < 3; ++i) std::printf(" %.17g", s.position[i]);
    for (int i = 0; i < 4; ++i) std::printf(" %.17g", s.orientation[i]);
    for (int i = 0; i < 3; ++i) std::printf(" %.17g", s.linearVelocity[i]);
    for (int i = 0; i < 3; ++i) std::printf(" %.17g", s.angularVelocity[i]);
    std::printf(" P00 %.17g N %d n %d removed %d converted %d added %d\n", extendedKalmanFilter.stateCovarianceMatrix[0][0],
                (int)s.mapFeatures.size(), nNow, extendedKalmanFilter.lastMapResult().n_removed_bad +
                extendedKalmanFilter.lastMapResult().n_removed_unseen, extendedKalmanFilter.lastMapResult().converted,
                extendedKalmanFilter.lastNewFeatures());
}
