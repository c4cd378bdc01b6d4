// ekfb_cv_compat.hpp -- the two OpenCV types that appear in the reference's public EKF interface
// (kalmanFilter/modules/1PointRansacEKF/EKF.h:47-51: `const cv::Mat &image`, `Matd stateCovarianceMatrix`), for builds
// without OpenCV.  Define EKFB_HAVE_OPENCV to compile against the real <opencv2/core/core.hpp> instead.
#ifndef EKFB_CV_COMPAT_HPP
#define EKFB_CV_COMPAT_HPP

#ifdef EKFB_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#else
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_8U
#define CV_8U 0
#define CV_64F 6
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#endif

namespace cv {
typedef unsigned char uchar;

// image / matrix container: rows x cols, element type code, row-major with `step` bytes per row
class Mat {
public:
    int flags, rows, cols;
    size_t step;
    uchar* data;
    Mat() : flags(0), rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : flags(type), rows(r), cols(c), step(0), data(nullptr) { allocate(); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : flags(type), rows(r), cols(c), data((uchar*)ext)
    {
        step = st ? st : (size_t)c * elemSize();
    }
    size_t elemSize() const
    {
        const int depth = flags & 7, cn = (flags >> 3) + 1;
        return (size_t)(depth == CV_64F ? 8 : depth == 5 ? 4 : 1) * cn;
    }
    int type() const { return flags; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    template <typename T> T& at(int i, int j) { return ptr<T>(i)[j]; }
    template <typename T> const T& at(int i, int j) const { return ptr<T>(i)[j]; }

protected:
    std::shared_ptr<std::vector<uchar>> buf_;
    void allocate()
    {
        step = (size_t)cols * elemSize();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)rows * step + 16, 0);
        data = buf_->data();
    }
};

template <typename T> class Mat_ : public Mat {
public:
    Mat_() { flags = sizeof(T) == 8 ? CV_64F : CV_8U; }
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_8U) {}
    T* operator[](int i) { return (T*)(data + (size_t)i * step); }
    const T* operator[](int i) const { return (const T*)(data + (size_t)i * step); }
    T& operator()(int i, int j) { return (*this)[i][j]; }
    const T& operator()(int i, int j) const { return (*this)[i][j]; }
};
}  // namespace cv
#endif  // EKFB_HAVE_OPENCV

typedef cv::Mat_<double> Matd;  // modules/Core/Base.h:66-67

#endif
