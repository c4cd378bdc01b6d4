// cvshim.hpp -- a minimal stand-in for the subset of OpenCV 2.4 that the reference's per-frame EKF sources use,
// so that those sources compile UNMODIFIED, in place under /root/reference, into oracle/_ref/libref.so.
// TEST INFRASTRUCTURE ONLY (same status as the oracle).  It reproduces cv::Mat's header/buffer semantics that the
// reference's results depend on: reference-counted shallow copies, ROI views, copyTo() re-allocating the
// destination header on a size/type mismatch, external-data constructors, the comma initialiser.  Arithmetic
// primitives are the oracle's cv2-pinned restatements (Mat::inv LU / closed forms, cv::eigen Jacobi, cv::ellipse).
#ifndef CVSHIM_HPP
#define CVSHIM_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_FOURCC(a, b, c, d) 0

// the oracle's restatements of the OpenCV primitives (oracle/ekf_oracle.cpp), pinned against cv2 fixtures
extern "C" {
void orc_eigen2x2(const double* A, double* evals, double* evecs_rows);
int orc_invert(const double* A, int32_t n, double* Ainv);
void orc_gemm_acc(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int K, int N);
void orc_fill_ellipse(uint8_t* img, int32_t width, int32_t height, int32_t cx, int32_t cy, int32_t ax_w, int32_t ax_h,
                      double angle_deg);
}

namespace cv {

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return (int)lrint(v); }
template <typename T> static inline T saturate_castf(float v) { return (T)v; }

inline int cvRound(double v) { return (int)lrint(v); }

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    int size() const { return end - start; }
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

template <typename T> struct conv_ {
    template <typename U> static T from(U v) { return (T)v; }
};
template <> struct conv_<int> {  // saturate_cast<int>(float/double) rounds to nearest even
    static int from(int v) { return v; }
    static int from(float v) { return (int)lrintf(v); }
    static int from(double v) { return (int)lrint(v); }
};

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x(conv_<T>::from(p.x)), y(conv_<T>::from(p.y)) {}
    bool operator!=(const Point_& o) const { return x != o.x || y != o.y; }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    template <typename U> Size_(const Size_<U>& s) : width(conv_<T>::from(s.width)), height(conv_<T>::from(s.height)) {}
};
typedef Size_<int> Size;
typedef Size_<int> Size2i;
typedef Size_<float> Size2f;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s) : pt(x, y), size(s), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_;

class Mat {
public:
    int flags;  // type code
    int rows, cols;
    size_t step;  // bytes between rows
    uchar* data;
    std::shared_ptr<std::vector<uchar>> buf;  // null for external data

    Mat() : flags(0), rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : flags(0), rows(0), cols(0), step(0), data(nullptr) { create(r, c, type); }
    Mat(Size s, int type) : flags(0), rows(0), cols(0), step(0), data(nullptr) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : flags(type), rows(r), cols(c), data((uchar*)ext)
    {
        step = st ? st : (size_t)c * elemSizeOf(type);
    }
    Mat(const Mat& m, const Range& rr, const Range& cr) { *this = m; applyRange(rr, cr); }
    template <typename T> Mat(const MatCommaInitializer_<T>& ci);

    static size_t elemSizeOf(int type)
    {
        const int depth = type & 7, cn = (type >> CV_CN_SHIFT) + 1;
        const size_t d = depth == CV_8U ? 1 : depth == CV_32F ? 4 : depth == CV_64F ? 8 : 4;
        return d * cn;
    }
    size_t elemSize() const { return elemSizeOf(flags); }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && flags == type) return;  // cv::Mat::create: nothing to do
        flags = type; rows = r; cols = c;
        step = (size_t)c * elemSizeOf(type);
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 16);
        data = buf->data();
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    void applyRange(const Range& rr, const Range& cr)
    {
        int r0 = rr.start == INT32_MIN ? 0 : rr.start, r1 = rr.end == INT32_MAX ? rows : rr.end;
        int c0 = cr.start == INT32_MIN ? 0 : cr.start, c1 = cr.end == INT32_MAX ? cols : cr.end;
        assert(0 <= r0 && r0 <= r1 && r1 <= rows && 0 <= c0 && c0 <= c1 && c1 <= cols);
        data += (size_t)r0 * step + (size_t)c0 * elemSize();
        rows = r1 - r0; cols = c1 - c0;
    }
    Mat operator()(const Range& rr, const Range& cr) const { return Mat(*this, rr, cr); }
    Mat row(int i) const { return Mat(*this, Range(i, i + 1), Range::all()); }
    Mat col(int j) const { return Mat(*this, Range::all(), Range(j, j + 1)); }
    Mat rowRange(int a, int b) const { return Mat(*this, Range(a, b), Range::all()); }
    Mat colRange(int a, int b) const { return Mat(*this, Range::all(), Range(a, b)); }

    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step))[j]; }
    template <typename T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step))[j]; }

    // cv::Mat::copyTo: dst.create(size, type) -- keeps the destination buffer if size and type already match
    // (so copying into a same-sized ROI view writes in place), otherwise re-allocates the destination header.
    void copyTo(const Mat& dst_) const
    {
        Mat& dst = const_cast<Mat&>(dst_);
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, flags);
        const size_t rowBytes = (size_t)cols * elemSize();
        if (dst.data == data && dst.step == step) return;
        for (int i = 0; i < rows; ++i) std::memmove(dst.data + (size_t)i * dst.step, data + (size_t)i * step, rowBytes);
    }
    Mat clone() const { Mat m; copyTo(m); return m; }
    Mat& setTo(const Scalar& s)
    {
        assert(depth() == CV_8U);
        for (int i = 0; i < rows; ++i) std::memset(data + (size_t)i * step, (int)s.val[0], (size_t)cols * elemSize());
        return *this;
    }
    static Mat zeros(int r, int c, int type)
    {
        Mat m(r, c, type);
        for (int i = 0; i < r; ++i) std::memset(m.data + (size_t)i * m.step, 0, (size_t)c * m.elemSize());
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type)
    {
        Mat m = zeros(r, c, type);
        assert((type & 7) == CV_8U);
        for (int i = 0; i < r; ++i) std::memset(m.data + (size_t)i * m.step, 1, (size_t)c * m.elemSize());
        return m;
    }
};

template <typename T> class Mat_ : public Mat {
public:
    Mat_() { flags = DataDepth<T>::value; }
    Mat_(int r, int c) : Mat(r, c, DataDepth<T>::value) {}
    Mat_(int r, int c, T* ext) : Mat(r, c, DataDepth<T>::value, ext) {}
    Mat_(const Mat& m) : Mat(m)
    {
        if (!m.empty()) assert(m.type() == DataDepth<T>::value);
        flags = DataDepth<T>::value;
    }
    Mat_(const Mat_& m) : Mat(m) {}
    Mat_(const Mat_& m, const Range& rr, const Range& cr) : Mat(m, rr, cr) {}
    Mat_(const MatCommaInitializer_<T>& ci);
    Mat_& operator=(const Mat_& m) { Mat::operator=(m); return *this; }
    Mat_& operator=(const Mat& m) { Mat::operator=(m); return *this; }

    T* operator[](int i) { return (T*)(data + (size_t)i * step); }
    const T* operator[](int i) const { return (const T*)(data + (size_t)i * step); }
    T& operator()(int i, int j) { return (*this)[i][j]; }
    const T& operator()(int i, int j) const { return (*this)[i][j]; }
    Mat_ operator()(const Range& rr, const Range& cr) const { return Mat_(*this, rr, cr); }
    Mat_ row(int i) const { return Mat_(*this, Range(i, i + 1), Range::all()); }
    Mat_ col(int j) const { return Mat_(*this, Range::all(), Range(j, j + 1)); }

    static Mat_ zeros(int r, int c)
    {
        Mat_ m(r, c);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) m[i][j] = T(0);
        return m;
    }
    static Mat_ eye(int r, int c)
    {
        Mat_ m = zeros(r, c);
        for (int i = 0; i < std::min(r, c); ++i) m[i][i] = T(1);
        return m;
    }
    Mat_ t() const
    {
        Mat_ o(cols, rows);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) o[j][i] = (*this)[i][j];
        return o;
    }
    Mat_ inv() const  // DECOMP_LU
    {
        assert(rows == cols);
        Mat_ c(rows, cols), o(rows, cols);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) c[i][j] = (*this)[i][j];
        if (!orc_invert((const double*)c.data, rows, (double*)o.data))
            for (int i = 0; i < rows; ++i)
                for (int j = 0; j < cols; ++j) o[i][j] = 0;
        return o;
    }
};

template <typename T> class MatCommaInitializer_ {
public:
    Mat_<T> m;
    int idx;
    MatCommaInitializer_(const Mat_<T>& m_, T first) : m(m_), idx(0) { put(first); }
    void put(T v) { m[idx / m.cols][idx % m.cols] = v; ++idx; }
    template <typename U> MatCommaInitializer_& operator,(U v) { put((T)v); return *this; }
    operator Mat_<T>() const { return m; }
};
template <typename T, typename U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v)
{
    return MatCommaInitializer_<T>(m, (T)v);
}
template <typename T> inline Mat_<T>::Mat_(const MatCommaInitializer_<T>& ci) : Mat(ci.m) {}
template <typename T> inline Mat::Mat(const MatCommaInitializer_<T>& ci) { *this = ci.m; }

typedef Mat_<double> Matd_;

// ---- dense arithmetic on Mat_<double> (cv::gemm / MatExpr stand-ins; eager) ----
inline Mat_<double> operator*(const Mat_<double>& A, const Mat_<double>& B)
{
    assert(A.cols == B.rows);
    Mat_<double> C = Mat_<double>::zeros(A.rows, B.cols);
    // same cache-blocked product the oracle uses, so that timing either is comparable (operands may be ROI views)
    orc_gemm_acc((const double*)A.data, (int)(A.step / sizeof(double)), (const double*)B.data, (int)(B.step / sizeof(double)),
                 (double*)C.data, (int)(C.step / sizeof(double)), A.rows, A.cols, B.cols);
    return C;
}
inline Mat_<double> elementwise(const Mat_<double>& A, const Mat_<double>& B, double sb)
{
    assert(A.rows == B.rows && A.cols == B.cols);
    Mat_<double> C(A.rows, A.cols);
    for (int i = 0; i < A.rows; ++i)
        for (int j = 0; j < A.cols; ++j) C[i][j] = A[i][j] + sb * B[i][j];
    return C;
}
inline Mat_<double> operator+(const Mat_<double>& A, const Mat_<double>& B) { return elementwise(A, B, 1.0); }
inline Mat_<double> operator-(const Mat_<double>& A, const Mat_<double>& B) { return elementwise(A, B, -1.0); }
inline Mat_<double> scale(const Mat_<double>& A, double s)
{
    Mat_<double> C(A.rows, A.cols);
    for (int i = 0; i < A.rows; ++i)
        for (int j = 0; j < A.cols; ++j) C[i][j] = A[i][j] * s;
    return C;
}
inline Mat_<double> operator*(const Mat_<double>& A, double s) { return scale(A, s); }
inline Mat_<double> operator*(double s, const Mat_<double>& A) { return scale(A, s); }
inline Mat_<double> operator*(long double s, const Mat_<double>& A) { return scale(A, (double)s); }
inline Mat_<double> operator*(const Mat_<double>& A, long double s) { return scale(A, (double)s); }

inline Mat operator*(const Mat& m, int s)  // cv::Mat::ones(...) * 255 on 8-bit masks
{
    assert(m.depth() == CV_8U);
    Mat o = m.clone();
    for (int i = 0; i < o.rows; ++i)
        for (size_t j = 0; j < (size_t)o.cols * o.elemSize(); ++j) {
            const int v = o.data[(size_t)i * o.step + j] * s;
            o.data[(size_t)i * o.step + j] = (uchar)(v > 255 ? 255 : v < 0 ? 0 : v);
        }
    return o;
}

inline std::ostream& operator<<(std::ostream& os, const Mat& m)
{
    os << "[Mat " << m.rows << "x" << m.cols << "]";
    return os;
}

// cv::eigen for the symmetric 2x2 case used by the reference (Core/EKFMath.cpp:277)
inline bool eigen(const Mat& src, Mat& evals, Mat& evecs)
{
    assert(src.rows == 2 && src.cols == 2 && src.type() == CV_64F);
    double A[4] = {src.at<double>(0, 0), src.at<double>(0, 1), src.at<double>(1, 0), src.at<double>(1, 1)}, w[2], V[4];
    orc_eigen2x2(A, w, V);
    evals.create(2, 1, CV_64F);
    evecs.create(2, 2, CV_64F);
    evals.at<double>(0, 0) = w[0]; evals.at<double>(1, 0) = w[1];
    for (int i = 0; i < 4; ++i) evecs.at<double>(i / 2, i % 2) = V[i];
    return true;
}

// cv::ellipse: only the filled full ellipse on an 8-bit single-channel image has an effect (the matching mask and the
// new-feature mask);
// every other drawing call of the reference's GUI code is a no-op here.
inline void ellipse(Mat& img, Point center, Size axes, double angle, double startAngle, double endAngle, const Scalar& color,
                    int thickness = 1, int lineType = 8, int shift = 0)
{
    (void)startAngle; (void)endAngle; (void)lineType; (void)shift;
    if (thickness >= 0 || img.type() != CV_8UC1 || !img.isContinuous()) return;
    if (color.val[0] == 255) {
        orc_fill_ellipse(img.data, img.cols, img.rows, center.x, center.y, axes.width, axes.height, angle);
        return;
    }
    // any other colour (the black ellipses of the new-feature mask, DetectNewImageFeatures.cpp:115-121,286-291): rasterise
    // into a scratch image and copy the covered pixels in that colour
    std::vector<uchar> tmp((size_t)img.rows * img.cols, 0);
    orc_fill_ellipse(tmp.data(), img.cols, img.rows, center.x, center.y, axes.width, axes.height, angle);
    const uchar c = (uchar)color.val[0];
    for (size_t i = 0; i < tmp.size(); ++i)
        if (tmp[i]) img.data[i] = c;
}
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) {}
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
enum { FONT_HERSHEY_SCRIPT_SIMPLEX = 6, FONT_HERSHEY_SIMPLEX = 0, FONT_HERSHEY_PLAIN = 1 };
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1, int = 8, bool = false) {}
inline void namedWindow(const std::string&, int = 1) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
inline void destroyWindow(const std::string&) {}
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline Mat imread(const std::string&, int = 1) { return Mat(); }

class VideoWriter {
public:
    bool open(const std::string&, int, double, Size, bool = true) { return false; }
    bool isOpened() const { return false; }
    void release() {}
    void write(const Mat&) {}
};

// cv::FileStorage / FileNode: the reference only writes traces when an output path is given; nothing is written here
class FileNode {
public:
    bool empty() const { return true; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    operator std::string() const { return std::string(); }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    size_t size() const { return 0; }
    struct iterator {
        bool operator!=(const iterator&) const { return false; }
        iterator& operator++() { return *this; }
        FileNode operator*() const { return FileNode(); }
    };
    iterator begin() const { return iterator(); }
    iterator end() const { return iterator(); }
};
typedef FileNode::iterator FileNodeIterator;
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    void* fs = nullptr;
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool open(const std::string&, int) { return false; }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
template <typename T> inline void operator>>(const FileNode&, T&) {}

class FeatureDetector {
public:
    virtual ~FeatureDetector() {}
    virtual void detect(const Mat& image, std::vector<KeyPoint>& keypoints, const Mat& mask = Mat()) const = 0;
};
class DescriptorExtractor {
public:
    virtual ~DescriptorExtractor() {}
    virtual void compute(const Mat& image, std::vector<KeyPoint>& keypoints, Mat& descriptors) const = 0;
};

}  // namespace cv

inline void cvWriteComment(void*, const char*, int) {}

#endif
