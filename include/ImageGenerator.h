// ImageGenerator.h -- the frame sources of the reference's sample driver (kalmanFilter/modules/ImageGenerator/
// ImageGenerator.h:39-49, FileSequenceImageGenerator.h/.cpp): an abstract generator with init() / getNextImage() and the
// numbered-file sequence "<path><prefix>%05d.<ext>" that kalmanFilter/samples/EKF/main.cpp:50 constructs.
//
// The reference decodes with cv::imread; OpenCV is not a dependency here, so the files are decoded by a small reader of its
// own (openekfmonoslam_b200/host/image_generator.cpp): 8-bit non-interlaced PNG (grey, grey + alpha, RGB, RGBA, palette;
// zlib inflates the IDAT stream) and binary PGM / PPM.  Like cv::imread with its default flag, the result is always an
// 8-bit 3-channel BGR image (alpha dropped, grey replicated), which is what EKF::init / EKF::step then receive.
// The camera and video generators of the reference (CameraImageGenerator, VideoFileImageGenerator, SlidingWindow...) wrap
// cv::VideoCapture and are not rebuilt.
#ifndef EKFB_IMAGE_GENERATOR_H
#define EKFB_IMAGE_GENERATOR_H

#include <string>

#include "ekfb_cv_compat.hpp"

class ImageGenerator {
public:
    virtual ~ImageGenerator() {}
    virtual void init() = 0;
    virtual cv::Mat& getNextImage() = 0;   // an empty Mat when the source is exhausted (or a file cannot be read)

protected:
    cv::Mat _image;
};

class FileSequenceImageGenerator : public ImageGenerator {
public:
    FileSequenceImageGenerator();
    FileSequenceImageGenerator(std::string path, std::string filePrefix, std::string fileExtension, int imageBeginIndex, int imageEndIndex);
    ~FileSequenceImageGenerator();
    void init();
    cv::Mat& getNextImage();

private:
    std::string _path, _filePrefix, _fileExtension;
    int _imageBeginIndex, _imageEndIndex, _imageActualIndex;
};

// decodes one PNG / PGM / PPM file into an 8-bit BGR image; false (and an empty Mat) if the file is missing or of a kind the
// reader does not handle (16-bit, interlaced, ...), with the reason on std::cerr
bool ekfbReadImage(const char* fileName, cv::Mat& bgr);

#endif
