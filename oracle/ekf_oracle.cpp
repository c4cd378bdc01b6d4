/*
 * ekf_oracle.cpp -- CPU ORACLE: a restatement of OpenEKFMonoSLAM's per-frame EKF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product path (openekfmonoslam_b200/).
 *
 * Every function cites the reference file:line it follows.  Paths are relative to
 * /root/reference/kalmanFilter/modules/ ; "E/" = 1PointRansacEKF/.
 * The restatement keeps the reference's literal (costly) forms: dense 2 x n Jacobians,
 * dense P*H^T with a >99% zero H, K = PHt * inv(S) by LU, P = (I - K H) P as an n x n x n
 * product, and a deep State copy plus a dense P*H_i^T per RANSAC hypothesis -- so that timing it
 * is an honest stand-in for the reference's CPU path.  x87 `L`-suffixed sub-expressions of the
 * reference are kept as long double.
 *
 * Third-party arithmetic (OpenCV 2.4.3.2, not under /root/reference) is hand-restated here from
 * the published algorithms: Mat::inv() DECOMP_LU (closed form for 2x2/3x3, partial-pivot LU
 * otherwise), cv::eigen (cyclic Jacobi, eigenvalues descending, eigenvectors as rows),
 * cv::ellipse filled (ellipse2Poly on the 1-degree float sine table + FillConvexPoly with its
 * fixed-point edge walk and Line2 outline).  These are pinned against cv2 4.13 fixtures in
 * tests/golden/ (tests/test_oracle_golden.py).
 *
 * PARITY STATUS: pinned.  tests/test_reference_pin.py compares every phase, the add-feature path and whole
 * EKF::step frames of this restatement with oracle/_ref/libref.so, the reference's own sources compiled
 * unmodified against oracle/cvshim (oracle/ref/Makefile): discrete results identical, FP64 to <= 1e-12.
 */
#include "ekf_oracle.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <vector>

// Core/EKFMath.h:37-45
#define EPSILON 2.22e-16L
#define DELTA 1.0e-12L
#define PI 3.14159265L
#define CHISQ_95_2 5.9915L
#define RAD_TO_DEG(a) (a) * 180.0L / PI

namespace {

// ---------------------------------------------------------------------------------------------
// minimal row-major double matrix (stands in for cv::Mat_<double>, Core/Base.h:66-67)
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Decision margins: every floating-point threshold decision of the frame records how far it was from
// flipping.  The sets the decisions produce are compared exactly against the GPU; the margins show how much numerical
// slack those comparisons had.  Test infrastructure: a process-wide recorder read through orc_margins_get().
// ---------------------------------------------------------------------------------------------
enum MarginClass { MG_FOV = 0, MG_FRAME, MG_GATE, MG_RATIO, MG_RANSAC, MG_CHI2, MG_DEADBAND, MG_COUNT };
struct MarginLog {
    double minAbs[MG_COUNT];
    int64_t count[MG_COUNT];
    int64_t zero[MG_COUNT];   // decisions whose operand was exactly on the threshold / exactly zero
};
MarginLog g_margins = {{1e300, 1e300, 1e300, 1e300, 1e300, 1e300, 1e300}, {0}, {0}};
inline void markMargin(int cls, double margin)
{
    g_margins.count[cls]++;
    const double a = fabs(margin);
    if (a == 0.0) g_margins.zero[cls]++;
    else if (a < g_margins.minAbs[cls]) g_margins.minAbs[cls] = a;
}

struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return d[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
    double* row(int i) { return &d[(size_t)i * c]; }
    const double* row(int i) const { return &d[(size_t)i * c]; }
};

Mat eye(int n)
{
    Mat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
}

// C += A * B (row-major).  Same dense product cv::gemm computes (summation order is not part of
// parity), cache-blocked over k and j with 4 rows of A per pass so that timing the oracle is a fair
// stand-in for OpenCV 2.4's blocked SSE2 gemm; target_clones picks AVX2+FMA at run time when present.
__attribute__((target_clones("avx2,fma", "default")))
void gemmAcc(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int K, int N)
{
    const int KB = 128, JB = 1024;
    for (int jb = 0; jb < N; jb += JB) {
        const int je = std::min(N, jb + JB);
        for (int kb = 0; kb < K; kb += KB) {
            const int ke = std::min(K, kb + KB);
            int i = 0;
            for (; i + 4 <= M; i += 4) {
                double* c0 = C + (size_t)i * ldc;
                double* c1 = c0 + ldc;
                double* c2 = c1 + ldc;
                double* c3 = c2 + ldc;
                for (int k = kb; k < ke; ++k) {
                    const double a0 = A[(size_t)i * lda + k], a1 = A[(size_t)(i + 1) * lda + k];
                    const double a2 = A[(size_t)(i + 2) * lda + k], a3 = A[(size_t)(i + 3) * lda + k];
                    const double* b = B + (size_t)k * ldb;
                    for (int j = jb; j < je; ++j) {
                        const double bv = b[j];
                        c0[j] += a0 * bv;
                        c1[j] += a1 * bv;
                        c2[j] += a2 * bv;
                        c3[j] += a3 * bv;
                    }
                }
            }
            for (; i < M; ++i) {
                double* c0 = C + (size_t)i * ldc;
                for (int k = kb; k < ke; ++k) {
                    const double a0 = A[(size_t)i * lda + k];
                    const double* b = B + (size_t)k * ldb;
                    for (int j = jb; j < je; ++j) c0[j] += a0 * b[j];
                }
            }
        }
    }
}

Mat mul(const Mat& A, const Mat& B)
{
    Mat C(A.r, B.c);
    gemmAcc(A.d.data(), A.c, B.d.data(), B.c, C.d.data(), C.c, A.r, A.c, B.c);
    return C;
}

Mat transpose(const Mat& A)
{
    Mat T(A.c, A.r);
    for (int i = 0; i < A.r; ++i)
        for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
    return T;
}

Mat block(const Mat& A, int r0, int r1, int c0, int c1)
{
    Mat B(r1 - r0, c1 - c0);
    for (int i = r0; i < r1; ++i)
        for (int j = c0; j < c1; ++j) B(i - r0, j - c0) = A(i, j);
    return B;
}

void setBlock(Mat& A, int r0, int c0, const Mat& B)
{
    for (int i = 0; i < B.r; ++i)
        for (int j = 0; j < B.c; ++j) A(r0 + i, c0 + j) = B(i, j);
}

// cv::Mat::inv() with DECOMP_LU (OpenCV 2.4 core/src/lapack.cpp, cv::invert): closed-form
// determinant/adjugate for 2x2 and 3x3, partial-pivot LU solving A X = I otherwise.
// Call sites: E/Update.cpp:108 (k x k), E/EKF.cpp:94 (2x2), E/MeasurementPrediction.cpp:216,353,672.
bool invertLU(const double* A, int n, double* out)
{
    if (n == 1) {
        if (A[0] == 0.) return false;
        out[0] = 1. / A[0];
        return true;
    }
    if (n == 2) {
        double d = A[0] * A[3] - A[1] * A[2];
        if (d == 0.) return false;
        d = 1. / d;
        double t0 = A[0] * d, t1 = A[3] * d;
        out[3] = t0;
        out[0] = t1;
        t0 = -A[1] * d;
        t1 = -A[2] * d;
        out[1] = t0;
        out[2] = t1;
        return true;
    }
    if (n == 3) {
#define S_(i, j) A[(i) * 3 + (j)]
        double d = S_(0, 0) * (S_(1, 1) * S_(2, 2) - S_(1, 2) * S_(2, 1)) -
                   S_(0, 1) * (S_(1, 0) * S_(2, 2) - S_(1, 2) * S_(2, 0)) +
                   S_(0, 2) * (S_(1, 0) * S_(2, 1) - S_(1, 1) * S_(2, 0));
        if (d == 0.) return false;
        d = 1. / d;
        double t[9];
        t[0] = (S_(1, 1) * S_(2, 2) - S_(1, 2) * S_(2, 1)) * d;
        t[1] = (S_(0, 2) * S_(2, 1) - S_(0, 1) * S_(2, 2)) * d;
        t[2] = (S_(0, 1) * S_(1, 2) - S_(0, 2) * S_(1, 1)) * d;
        t[3] = (S_(1, 2) * S_(2, 0) - S_(1, 0) * S_(2, 2)) * d;
        t[4] = (S_(0, 0) * S_(2, 2) - S_(0, 2) * S_(2, 0)) * d;
        t[5] = (S_(0, 2) * S_(1, 0) - S_(0, 0) * S_(1, 2)) * d;
        t[6] = (S_(1, 0) * S_(2, 1) - S_(1, 1) * S_(2, 0)) * d;
        t[7] = (S_(0, 1) * S_(2, 0) - S_(0, 0) * S_(2, 1)) * d;
        t[8] = (S_(0, 0) * S_(1, 1) - S_(0, 1) * S_(1, 0)) * d;
#undef S_
        for (int i = 0; i < 9; ++i) out[i] = t[i];
        return true;
    }
    // general n: LUImpl (partial pivoting) on a copy, right-hand side = identity
    std::vector<double> a(A, A + (size_t)n * n);
    std::vector<double> b((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) b[(size_t)i * n + i] = 1.0;
    const double eps = 2.220446049250313e-16 * 100;
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (std::fabs(a[(size_t)j * n + i]) > std::fabs(a[(size_t)k * n + i])) k = j;
        if (std::fabs(a[(size_t)k * n + i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < n; ++j) std::swap(a[(size_t)i * n + j], a[(size_t)k * n + j]);
            for (int j = 0; j < n; ++j) std::swap(b[(size_t)i * n + j], b[(size_t)k * n + j]);
        }
        double dd = -1 / a[(size_t)i * n + i];
        for (int j = i + 1; j < n; ++j) {
            double alpha = a[(size_t)j * n + i] * dd;
            if (alpha == 0.) continue;
            for (int kk = i + 1; kk < n; ++kk) a[(size_t)j * n + kk] += alpha * a[(size_t)i * n + kk];
            for (int kk = 0; kk < n; ++kk) b[(size_t)j * n + kk] += alpha * b[(size_t)i * n + kk];
        }
        a[(size_t)i * n + i] = -dd;
    }
    for (int i = n - 1; i >= 0; --i)
        for (int j = 0; j < n; ++j) {
            double s = b[(size_t)i * n + j];
            for (int k = i + 1; k < n; ++k) s -= a[(size_t)i * n + k] * b[(size_t)k * n + j];
            b[(size_t)i * n + j] = s * a[(size_t)i * n + i];
        }
    std::memcpy(out, b.data(), sizeof(double) * (size_t)n * n);
    return true;
}

Mat inv(const Mat& A)
{
    Mat R(A.r, A.c);
    if (!invertLU(A.d.data(), A.r, R.d.data())) std::fill(R.d.begin(), R.d.end(), 0.0);
    return R;
}

// cv::eigen on a symmetric 2x2 (OpenCV JacobiImpl_): one Jacobi rotation, eigenvalues sorted
// descending, eigenvectors are the ROWS of V.  Call site: Core/EKFMath.cpp:277.
void eigen2x2(const double* A, double* W, double* V)
{
    const double eps = 2.220446049250313e-16;
    V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 1;
    W[0] = A[0];
    W[1] = A[3];
    double p = A[1];
    if (std::fabs(p) > eps) {
        double y = (W[1] - W[0]) * 0.5;
        double t = std::fabs(y) + hypot(p, y);
        double s = hypot(p, t);
        double c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        W[0] -= t;
        W[1] += t;
        for (int i = 0; i < 2; ++i) {
            double a0 = V[i], b0 = V[2 + i];
            V[i] = a0 * c - b0 * s;
            V[2 + i] = a0 * s + b0 * c;
        }
    }
    if (W[0] < W[1]) {
        std::swap(W[0], W[1]);
        std::swap(V[0], V[2]);
        std::swap(V[1], V[3]);
    }
}

inline int cvRoundD(double v) { return (int)lrint(v); }  // round-half-to-even like cvRound
inline int cvRoundF(float v) { return (int)lrintf(v); }

// ---------------------------------------------------------------------------------------------
// cv::ellipse(..., thickness=-1): ellipse2Poly + FillConvexPoly (OpenCV core/src/drawing.cpp)
// Call site: Gui/Draw.cpp:58.
// ---------------------------------------------------------------------------------------------
const float kSinTable90[91] = {
    0.0000000f, 0.0174524f, 0.0348995f, 0.0523360f, 0.0697565f, 0.0871557f, 0.1045285f,
    0.1218693f, 0.1391731f, 0.1564345f, 0.1736482f, 0.1908090f, 0.2079117f, 0.2249511f,
    0.2419219f, 0.2588190f, 0.2756374f, 0.2923717f, 0.3090170f, 0.3255682f, 0.3420201f,
    0.3583679f, 0.3746066f, 0.3907311f, 0.4067366f, 0.4226183f, 0.4383711f, 0.4539905f,
    0.4694716f, 0.4848096f, 0.5000000f, 0.5150381f, 0.5299193f, 0.5446390f, 0.5591929f,
    0.5735764f, 0.5877853f, 0.6018150f, 0.6156615f, 0.6293204f, 0.6427876f, 0.6560590f,
    0.6691306f, 0.6819984f, 0.6946584f, 0.7071068f, 0.7193398f, 0.7313537f, 0.7431448f,
    0.7547096f, 0.7660444f, 0.7771460f, 0.7880108f, 0.7986355f, 0.8090170f, 0.8191520f,
    0.8290376f, 0.8386706f, 0.8480481f, 0.8571673f, 0.8660254f, 0.8746197f, 0.8829476f,
    0.8910065f, 0.8987940f, 0.9063078f, 0.9135455f, 0.9205049f, 0.9271839f, 0.9335804f,
    0.9396926f, 0.9455186f, 0.9510565f, 0.9563048f, 0.9612617f, 0.9659258f, 0.9702957f,
    0.9743701f, 0.9781476f, 0.9816272f, 0.9848078f, 0.9876883f, 0.9902681f, 0.9925462f,
    0.9945219f, 0.9961947f, 0.9975641f, 0.9986295f, 0.9993908f, 0.9998477f, 1.0000000f};

// SinTable[k], k in [0,450]: the table holds sin(k degrees) with the symmetries of the sine
inline float sinTable(int k)
{
    if (k >= 360) k -= 360;
    if (k <= 90) return kSinTable90[k];
    if (k <= 180) return kSinTable90[180 - k];
    if (k <= 270) return -kSinTable90[k - 180];
    return -kSinTable90[360 - k];
}

const int XY_SHIFT = 16;
const int64_t XY_ONE = 1 << XY_SHIFT;

struct Pt64 { int64_t x, y; };

inline void putPoint(uint8_t* img, int w, int h, int64_t x, int64_t y)
{
    if (0 <= x && x < w && 0 <= y && y < h) img[(size_t)y * w + x] = 255;
}

// cv::clipLine on 64-bit fixed-point coordinates
bool clipLine64(int64_t width, int64_t height, Pt64& pt1, Pt64& pt2)
{
    int c1, c2;
    int64_t right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
    int64_t &x1 = pt1.x, &y1 = pt1.y, &x2 = pt2.x, &y2 = pt2.y;
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (x2 - x1) / (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (x2 - x1) / (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Line2: fixed-point DDA outline used by FillConvexPoly for shift != 0
void line2(uint8_t* img, int w, int h, Pt64 pt1, Pt64 pt2)
{
    if (!clipLine64((int64_t)w << XY_SHIFT, (int64_t)h << XY_SHIFT, pt1, pt2)) return;
    int64_t dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
    int64_t j = dx < 0 ? -1 : 0;
    int64_t ax = (dx ^ j) - j;
    int64_t i = dy < 0 ? -1 : 0;
    int64_t ay = (dy ^ i) - i;
    int64_t x_step, y_step;
    int ecount;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        pt1.x ^= pt2.x & j; pt2.x ^= pt1.x & j; pt1.x ^= pt2.x & j;
        pt1.y ^= pt2.y & j; pt2.y ^= pt1.y & j; pt1.y ^= pt2.y & j;
        x_step = XY_ONE;
        y_step = dy * (1 << XY_SHIFT) / (ax | 1);
        ecount = (int)((pt2.x - pt1.x) >> XY_SHIFT);
    } else {
        dx = (dx ^ i) - i;
        pt1.x ^= pt2.x & i; pt2.x ^= pt1.x & i; pt1.x ^= pt2.x & i;
        pt1.y ^= pt2.y & i; pt2.y ^= pt1.y & i; pt1.y ^= pt2.y & i;
        x_step = dx * (1 << XY_SHIFT) / (ay | 1);
        y_step = XY_ONE;
        ecount = (int)((pt2.y - pt1.y) >> XY_SHIFT);
    }
    pt1.x += (XY_ONE >> 1);
    pt1.y += (XY_ONE >> 1);
    putPoint(img, w, h, (pt2.x + (XY_ONE >> 1)) >> XY_SHIFT, (pt2.y + (XY_ONE >> 1)) >> XY_SHIFT);
    if (ax > ay) {
        pt1.x >>= XY_SHIFT;
        while (ecount >= 0) {
            putPoint(img, w, h, pt1.x, pt1.y >> XY_SHIFT);
            pt1.x++;
            pt1.y += y_step;
            ecount--;
        }
    } else {
        pt1.y >>= XY_SHIFT;
        while (ecount >= 0) {
            putPoint(img, w, h, pt1.x >> XY_SHIFT, pt1.y);
            pt1.x += x_step;
            pt1.y++;
            ecount--;
        }
    }
}

void fillConvexPoly(uint8_t* img, int width, int height, const std::vector<Pt64>& v)
{
    const int shift = XY_SHIFT;
    const int npts = (int)v.size();
    struct { int idx, di; int64_t x, dx; int ye; } edge[2];
    const int delta = 1 << shift >> 1;
    int i, y, imin = 0;
    int edges = npts;
    int64_t xmin, xmax, ymin, ymax;
    const int delta1 = XY_ONE >> 1, delta2 = XY_ONE >> 1;

    Pt64 p0 = v[npts - 1];
    xmin = xmax = v[0].x;
    ymin = ymax = v[0].y;
    for (i = 0; i < npts; i++) {
        Pt64 p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        ymax = std::max(ymax, p.y);
        xmax = std::max(xmax, p.x);
        xmin = std::min(xmin, p.x);
        line2(img, width, height, p0, p);
        p0 = p;
    }
    xmin = (xmin + delta) >> shift;
    xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift;
    ymax = (ymax + delta) >> shift;
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= width || (int)ymin >= height) return;
    ymax = std::min(ymax, (int64_t)height - 1);
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y = (int)ymin;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    do {
        for (i = 0; i < 2; i++) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                int ty = 0;
                for (; edges-- > 0;) {
                    ty = (int)((v[idx].y + delta) >> shift);
                    if (ty > y) {
                        int64_t xs = v[idx0].x;
                        int64_t xe = v[idx].x;
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta1) >> XY_SHIFT);
            int xx2 = (int)((edge[right].x + delta2) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < width) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= width) xx2 = width - 1;
                for (int x = xx1; x <= xx2; ++x) img[(size_t)y * width + x] = 255;
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

// cv::ellipse(img, Point center, Size axes, angle, 0, 360, color, -1)
void fillEllipse(uint8_t* img, int width, int height, int cx, int cy, int aw, int ah, double angleDeg)
{
    int angle = cvRoundD(angleDeg);
    int64_t centerx = (int64_t)cx << XY_SHIFT, centery = (int64_t)cy << XY_SHIFT;
    int64_t axw = std::llabs((int64_t)aw << XY_SHIFT), axh = std::llabs((int64_t)ah << XY_SHIFT);
    int delta = (int)((std::max(axw, axh) + (XY_ONE >> 1)) >> XY_SHIFT);
    delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;

    // ellipse2Poly(center, axes, angle, 0, 360, delta)
    while (angle < 0) angle += 360;
    while (angle > 360) angle -= 360;
    const int arc_start = 0, arc_end = 360;
    float alpha = sinTable(450 - angle);  // cos
    float beta = sinTable(angle);         // sin
    std::vector<Pt64> v;
    Pt64 prev = {(int64_t)-1, (int64_t)-1};
    bool havePrev = false;
    for (int i = arc_start; i < arc_end + delta; i += delta) {
        int a = i;
        if (a > arc_end) a = arc_end;
        if (a < 0) a += 360;
        double x = (double)axw * sinTable(450 - a);
        double y = (double)axh * sinTable(a);
        double px = (double)centerx + x * alpha - y * beta;
        double py = (double)centery + x * beta + y * alpha;
        Pt64 pt;
        pt.x = (int64_t)cvRoundD(px / XY_ONE) << XY_SHIFT;
        pt.y = (int64_t)cvRoundD(py / XY_ONE) << XY_SHIFT;
        pt.x += cvRoundD(px - pt.x);
        pt.y += cvRoundD(py - pt.y);
        if (!havePrev || pt.x != prev.x || pt.y != prev.y) {
            v.push_back(pt);
            prev = pt;
            havePrev = true;
        }
    }
    if (v.size() <= 1) {
        Pt64 c = {centerx, centery};
        v.assign(2, c);
    }
    fillConvexPoly(img, width, height, v);
}

// Core/EKFMath.cpp:271-298 matrix2x2ToUncertaintyEllipse2D
void ellipseParams(const double* S, float& axW, float& axH, double& angle)
{
    double ev[2], V[4];
    eigen2x2(S, ev, V);
    axW = static_cast<float>(2.0L * sqrt(ev[0] * CHISQ_95_2));
    axH = static_cast<float>(2.0L * sqrt(ev[1] * CHISQ_95_2));
    double tanv = V[2] / V[0];
    angle = atan(tanv);
}

// Gui/Draw.cpp:42-64 drawUncertaintyEllipse2D(img, Point2f center, cov, maxAxesSize, white, fill=true)
void drawUncertaintyEllipse(uint8_t* img, int width, int height, double cxd, double cyd, const double* S,
                            int maxAxes)
{
    float cxf = (float)cxd, cyf = (float)cyd;  // cv::Point2d -> cv::Point2f at the call, Matching.cpp:197
    int icx = (int)cxf, icy = (int)cyf;         // cv::Point intCenter(center.x, center.y): truncation
    float aw, ah;
    double angle;
    ellipseParams(S, aw, ah, angle);
    float mw = std::min(aw, (float)maxAxes), mh = std::min(ah, (float)maxAxes);
    int iw = (int)mw, ih = (int)mh;  // cv::Size(int,int) from floats: truncation (Draw.cpp:55)
    double angDeg = (double)(RAD_TO_DEG(angle));
    fillEllipse(img, width, height, icx, icy, iw, ih, angDeg);
}

// Core/EKFMath.cpp:302-351 pointIsInsideEllipse(Point2f point, Point2f center, Size axes, double angle)
bool pointIsInsideEllipse(float px, float py, float cx, float cy, int aw, int ah, double angle)
{
    double majorAxis = std::max(aw, ah);
    double minorAxis = std::min(aw, ah);
    double f = sqrt(majorAxis * majorAxis - minorAxis * minorAxis);
    double f1x, f1y, f2x, f2y;
    if (ah < aw) {
        f1x = f * cos(angle) + cx;
        f1y = f * sin(angle) + cy;
        f2x = -f * cos(angle) + cx;
        f2y = -f * sin(angle) + cy;
    } else {
        f1x = f * (-sin(angle)) + cx;
        f1y = f * cos(angle) + cy;
        f2x = -f * (-sin(angle)) + cx;
        f2y = -f * cos(angle) + cy;
    }
    double a1x = px - f1x, a1y = py - f1y, a2x = px - f2x, a2y = py - f2y;
    double norm_sum = sqrt(a1x * a1x + a1y * a1y) + sqrt(a2x * a2x + a2y * a2y);
    markMargin(MG_GATE, norm_sum - 2 * majorAxis);
    return norm_sum <= 2 * majorAxis;
}

// ---------------------------------------------------------------------------------------------
// Core/EKFMath.cpp scalar helpers
// ---------------------------------------------------------------------------------------------
double euclideanNorm3(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

void anglesToQuaternion(const double* v, double* quat)  // EKFMath.cpp:62-81
{
    double norm = euclideanNorm3(v);
    if (norm < EPSILON) {
        quat[0] = 1; quat[1] = 0; quat[2] = 0; quat[3] = 0;
    } else {
        double normDiv2 = norm / 2;
        double sinNormDiv2 = sin(normDiv2);
        quat[0] = cos(normDiv2);
        quat[1] = sinNormDiv2 * v[0] / norm;
        quat[2] = sinNormDiv2 * v[1] / norm;
        quat[3] = sinNormDiv2 * v[2] / norm;
    }
}

void multiplyQuaternions(const double* q1, const double* q2, double* q)  // EKFMath.cpp:85-101
{
    q[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
    q[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
    q[2] = q1[0] * q2[2] - q1[1] * q2[3] + q1[2] * q2[0] + q1[3] * q2[1];
    q[3] = q1[0] * q2[3] + q1[1] * q2[2] - q1[2] * q2[1] + q1[3] * q2[0];
}

void quaternionToRotationMatrix(const double* q, double* R)  // EKFMath.cpp:121-141
{
    double r = q[0], x = q[1], y = q[2], z = q[3];
    double r2 = r * r, x2 = x * x, y2 = y * y, z2 = z * z;
    R[0] = r2 + x2 - y2 - z2;
    R[1] = 2 * (x * y - r * z);
    R[2] = 2 * (z * x + r * y);
    R[3] = 2 * (x * y + r * z);
    R[4] = r2 - x2 + y2 - z2;
    R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (z * x - r * y);
    R[7] = 2 * (y * z + r * x);
    R[8] = r2 - x2 - y2 + z2;
}

void makeDirectionalVector(double theta, double fi, double* m)  // EKFMath.cpp:145-152
{
    double cosfi = cos(fi);
    m[0] = cosfi * sin(theta);
    m[1] = -sin(fi);
    m[2] = cosfi * cos(theta);
}

void rotMulVec(const double* R, const double* v, double* out)  // EKFMath.cpp:156-165
{
    double x = v[0], y = v[1], z = v[2];
    out[0] = R[0] * x + R[1] * y + R[2] * z;
    out[1] = R[3] * x + R[4] * y + R[5] * z;
    out[2] = R[6] * x + R[7] * y + R[8] * z;
}

void matrixMult(const double* L, int lr, int lc, const double* Rm, int rc, double* out)  // EKFMath.cpp:201-216
{
    for (int i = 0; i < lr; ++i)
        for (int j = 0; j < rc; ++j) {
            out[i * rc + j] = 0.0L;
            for (int k = 0; k < lc; ++k) out[i * rc + j] += L[i * lc + k] * Rm[k * rc + j];
        }
}

// ---------------------------------------------------------------------------------------------
// data model: E/State.h:40-81, E/MapFeature.h:48-78
// ---------------------------------------------------------------------------------------------
enum { TYPE_DEPTH = 1, TYPE_INVERSE_DEPTH = 2 };

struct Feature {
    int type;
    double position[6];
    int dim;
    int covPos;
    uint8_t desc[32];
    unsigned timesPredicted, timesMatched;
};

struct State {
    double position[3];
    double orientation[4];
    double R[9];
    double linearVelocity[3];
    double angularVelocity[3];
    std::vector<Feature> features;
    void setOrientation(const double* q)  // State.cpp:131-139
    {
        for (int i = 0; i < 4; ++i) orientation[i] = q[i];
        quaternionToRotationMatrix(orientation, R);
    }
    int nDepth() const
    {
        int c = 0;
        for (const Feature& f : features) c += (f.type == TYPE_DEPTH);
        return c;
    }
    int stateDim() const
    {
        int n = 13;
        for (const Feature& f : features) n += f.dim;
        return n;
    }
};

struct Prediction {  // E/ImageFeaturePrediction.h:37-50 + its Jacobian row pair
    int featureIndex;
    double h[2];
    double S[4];
    double Hx[26];  // 2 x 13 (cols 7..12 stay zero)
    double Hf[12];  // 2 x dim
    int covPos, dim;
};

struct Match {  // E/Matching.h:38-48
    int featureIndex;
    double z[2];
    uint8_t desc[32];
    float distance;
    int kpIndex;
};

}  // namespace

struct orc_filter {
    orc_params p;
    State state;
    Mat P;
    // per-frame scratch, in the reference's vector order
    std::vector<Prediction> preds;     // predictedDistortedFeatures (+ Jacobians)
    std::vector<int> unseen;           // feature indices not predicted
    std::vector<Match> matches;
    std::vector<uint8_t> mask, kpOk;
    std::vector<int> matchPred;        // index into preds for each match (EKF.cpp:368-392)
    std::vector<int> inlierIdx, outlierIdx;  // indices into matches
    std::vector<int> supportCounts;
    int nHyp = 0, bestHyp = -1;
    std::vector<Prediction> outlierPreds;    // re-prediction of outliers after the LI update
    std::vector<int> outlierMatchesKept;     // indices into matches, after the in-frame filter
    std::vector<int> rescuedIdx;             // indices into matches
    std::vector<int> rescuedPred;            // indices into outlierPreds
    orc_frame_info info;
};

namespace {

// ---------------------------------------------------------------------------------------------
// E/StateAndCovariancePrediction.cpp
// ---------------------------------------------------------------------------------------------
void predictState(State& s, double dt)  // :43-65
{
    for (int i = 0; i < 3; ++i) s.position[i] += s.linearVelocity[i] * dt;
    double w[3];
    for (int i = 0; i < 3; ++i) w[i] = s.angularVelocity[i] * dt;
    double q2[4], q[4];
    anglesToQuaternion(w, q2);
    multiplyQuaternions(s.orientation, q2, q);
    s.setOrientation(q);
}

inline double derivQuatWByOmegaA(double omegaA, double omega, double dt)  // :100-103
{
    return (-dt / 2.0L) * (omegaA / omega) * sinl(omega * dt / 2.0L);
}
inline double derivQuatAByOmegaA(double omegaA, double omega, double dt)  // :107-111
{
    return (dt / 2.0L) * omegaA * omegaA / (omega * omega) * cosl(omega * dt / 2.0L) +
           (1.0L / omega) * (1.0L - omegaA * omegaA / (omega * omega)) * sinl(omega * dt / 2.0L);
}
inline double derivQuatAByOmegaB(double omegaA, double omegaB, double omega, double dt)  // :115-119
{
    return (omegaA * omegaB / (omega * omega)) *
           ((dt / 2.0L) * cosl(omega * dt / 2.0L) - (1.0L / omega) * sinl(omega * dt / 2.0L));
}

void predictCovariance(Mat& P, const State& s, double dt, const orc_params& prm)  // :154-240
{
    Mat F = eye(13);
    for (int i = 0; i < 3; ++i) F(i, i + 7) = dt;

    // jacobianDynmodelEq3to7Quat :71-92
    {
        double w[3], q[4];
        for (int i = 0; i < 3; ++i) w[i] = s.angularVelocity[i] * dt;
        anglesToQuaternion(w, q);
        double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
        double m[16] = {qw, -qx, -qy, -qz, qx, qw, qz, -qy, qy, -qz, qw, qx, qz, qy, -qx, qw};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) F(3 + i, 3 + j) = m[i * 4 + j];
    }

    Mat G(13, 6);
    bool omegaZero = fabs(s.angularVelocity[0]) < EPSILON && fabs(s.angularVelocity[1]) < EPSILON &&
                     fabs(s.angularVelocity[2]) < EPSILON;
    if (omegaZero) {
        for (int i = 0; i < 3; ++i) F(i + 10, i + 10) = 0;
        // :211-212 copies the 4x4 F[3:7,3:7] view into a 4x3 view of G: cv::Mat::copyTo re-allocates
        // the destination header, so G[3:7,3:6] stays zero in this branch.
    } else {
        // jacobianDynmodelEq3to7Omega :123-148
        double no = euclideanNorm3(s.angularVelocity);
        double ox = s.angularVelocity[0], oy = s.angularVelocity[1], oz = s.angularVelocity[2];
        const double* q = s.orientation;
        double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
        Mat QM(4, 4);  // quaternionToQuaternionMatrix, EKFMath.cpp:105-116
        double qm[16] = {qw, -qx, -qy, -qz, qx, qw, -qz, qy, qy, qz, qw, -qx, qz, -qy, qx, qw};
        QM.d.assign(qm, qm + 16);
        Mat D(4, 3);
        double dm[12] = {derivQuatWByOmegaA(ox, no, dt),     derivQuatWByOmegaA(oy, no, dt),
                         derivQuatWByOmegaA(oz, no, dt),     derivQuatAByOmegaA(ox, no, dt),
                         derivQuatAByOmegaB(ox, oy, no, dt), derivQuatAByOmegaB(ox, oz, no, dt),
                         derivQuatAByOmegaB(oy, ox, no, dt), derivQuatAByOmegaA(oy, no, dt),
                         derivQuatAByOmegaB(oy, oz, no, dt), derivQuatAByOmegaB(oz, ox, no, dt),
                         derivQuatAByOmegaB(oz, oy, no, dt), derivQuatAByOmegaA(oz, no, dt)};
        D.d.assign(dm, dm + 12);
        Mat QD = mul(QM, D);
        setBlock(F, 3, 10, QD);
        setBlock(G, 3, 3, QD);
    }
    for (int i = 0; i < 3; ++i) {
        G(i + 7, i) = 1.0L;
        G(i + 10, i + 3) = 1.0L;
        G(i, i) = 1.0L * dt;
    }
    Mat Q(6, 6);
    double lin = prm.linear_accel_sd * prm.linear_accel_sd * dt * dt;
    double ang = prm.angular_accel_sd * prm.angular_accel_sd * dt * dt;
    for (int i = 0; i < 3; ++i) {
        Q(i, i) = lin;
        Q(i + 3, i + 3) = ang;
    }
    const int n = P.r;
    Mat Ft = transpose(F);
    Mat P13 = block(P, 0, 13, 0, 13);
    Mat A = mul(mul(F, P13), Ft);
    Mat B = mul(mul(G, Q), transpose(G));
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) P(i, j) = A(i, j) + B(i, j);
    if (n > 13) {
        Mat Pxm = block(P, 0, 13, 13, n);
        setBlock(P, 0, 13, mul(F, Pxm));
        Mat Pmx = block(P, 13, n, 0, 13);
        setBlock(P, 13, 0, mul(Pmx, Ft));
    }
}

// ---------------------------------------------------------------------------------------------
// E/MeasurementPrediction.cpp
// ---------------------------------------------------------------------------------------------
void distortPoint_matlab(const orc_params& c, const double* in, double* out)  // :47-83
{
    double pixelDistX = in[0] - c.cx;
    double pixelDistY = in[1] - c.cy;
    double ddx = c.dx * pixelDistX;
    double ddy = c.dy * pixelDistY;
    double distPow2 = ddx * ddx + ddy * ddy;
    double ru = sqrt(distPow2);
    double rd = ru / (1.0L + c.k1 * distPow2 + c.k2 * distPow2 * distPow2);
    for (int k = 0; k < 10; ++k) {
        double rd2 = rd * rd;
        double rd3 = rd2 * rd;
        double rd4 = rd2 * rd2;
        double rd5 = rd4 * rd;
        double f = rd + c.k1 * rd3 + c.k2 * rd5 - ru;
        double fp = 1 + 3 * c.k1 * rd2 + 5 * c.k2 * rd4;
        rd = rd - f / fp;
    }
    double rd2 = rd * rd;
    double rd4 = rd2 * rd2;
    double d = (1.0L + c.k1 * rd2 + c.k2 * rd4);
    out[0] = c.cx + pixelDistX / d;
    out[1] = c.cy + pixelDistY / d;
}

void projectToCameraFrame(const orc_params& c, const double* pt, double* uv)  // :110-120
{
    uv[0] = c.cx + (c.fx * pt[0] / pt[2]);
    uv[1] = c.cy + (c.fy * pt[1] / pt[2]);
}

void toCameraAxisInverseDepth(const double* pt, const double* camPos, const double* Rm, double* out)  // :127-140
{
    double rho = pt[5];
    double m[3];
    makeDirectionalVector(pt[3], pt[4], m);
    double a[3];
    a[0] = rho * (pt[0] - camPos[0]) + m[0];
    a[1] = rho * (pt[1] - camPos[1]) + m[1];
    a[2] = rho * (pt[2] - camPos[2]) + m[2];
    rotMulVec(Rm, a, out);
}

void toCameraAxis(const double* pt, const double* camPos, const double* Rm, double* out)  // :147-156
{
    double a[3] = {pt[0] - camPos[0], pt[1] - camPos[1], pt[2] - camPos[2]};
    rotMulVec(Rm, a, out);
}

bool isInFrontOfCamera(const orc_params& c, const double* f)  // :162-171
{
    double atanxz = RAD_TO_DEG(atan2(f[0], f[2]));
    double atanyz = RAD_TO_DEG(atan2(f[1], f[2]));
    markMargin(MG_FOV, std::min(c.angular_vision_x - fabs(atanxz), c.angular_vision_y - fabs(atanyz)));
    return -c.angular_vision_x < atanxz && atanxz < c.angular_vision_x && -c.angular_vision_y < atanyz &&
           atanyz < c.angular_vision_y;
}

bool isVisibleInImageFrame(const orc_params& c, const double* p)  // :176-181
{
    markMargin(MG_FRAME, std::min(std::min(p[0], c.pixels_x - p[0]), std::min(p[1], c.pixels_y - p[1])));
    return (p[0] > 0 && p[0] < c.pixels_x && p[1] > 0 && p[1] < c.pixels_y);
}

// predictMeasurementState :203-265.  `subset` empty = all features in map order.
void predictMeasurementState(const orc_params& prm, const State& s, const std::vector<int>& subset,
                             std::vector<Prediction>& preds, std::vector<int>& notPredicted)
{
    const bool all = subset.empty();
    size_t count = all ? s.features.size() : subset.size();
    if (count == 0) return;
    double Rinv[9], Rt[9];
    invertLU(s.R, 3, Rinv);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = s.R[j * 3 + i];
    for (size_t i = 0; i < count; ++i) {
        int fi = all ? (int)i : subset[i];
        const Feature& f = s.features[fi];
        double c3[3], uv[2];
        if (f.type == TYPE_INVERSE_DEPTH)
            toCameraAxisInverseDepth(f.position, s.position, Rt, c3);
        else
            toCameraAxis(f.position, s.position, Rinv, c3);
        bool predicted = false;
        if (isInFrontOfCamera(prm, c3)) {
            projectToCameraFrame(prm, c3, uv);
            distortPoint_matlab(prm, uv, uv);
            if (isVisibleInImageFrame(prm, uv)) {
                Prediction p;
                std::memset(&p, 0, sizeof(p));
                p.featureIndex = fi;
                p.h[0] = uv[0];
                p.h[1] = uv[1];
                preds.push_back(p);
                predicted = true;
            }
        }
        if (!predicted) notPredicted.push_back(fi);
    }
}

void jacFrameProjection(const orc_params& c, const State& s, const double* Rinv, const double* pt, bool invDepth,
                        double* J)  // :273-297
{
    double pc[3];
    if (invDepth)
        toCameraAxisInverseDepth(pt, s.position, Rinv, pc);
    else
        toCameraAxis(pt, s.position, Rinv, pc);
    J[0] = c.fx / pc[2];
    J[1] = 0;
    J[2] = -pc[0] * c.fx / (pc[2] * pc[2]);
    J[3] = 0;
    J[4] = c.fy / pc[2];
    J[5] = -pc[1] * c.fy / (pc[2] * pc[2]);
}

void jacDistortion(const orc_params& c, const double* hd, double* J)  // :308-337
{
    double pdx = hd[0] - c.cx;
    double pdy = hd[1] - c.cy;
    double ddx = c.dx * pdx;
    double ddy = c.dy * pdy;
    double d2 = ddx * ddx + ddy * ddy;
    double rad = 1 + c.k1 * d2 + c.k2 * d2 * d2;
    double xx = rad + pdx * (c.k1 + 2 * c.k2 * d2) * (2 * pdx * c.dx * c.dx);
    double yy = rad + pdy * (c.k1 + 2 * c.k2 * d2) * (2 * pdy * c.dy * c.dy);
    double xy = pdx * (c.k1 + 2 * c.k2 * d2) * (2 * pdy * c.dy * c.dy);
    double yx = pdy * (c.k1 + 2 * c.k2 * d2) * (2 * pdx * c.dx * c.dx);
    J[0] = xx;
    J[1] = xy;
    J[2] = yx;
    J[3] = yy;
}

void jacProjection(const orc_params& c, const State& s, const double* Rinv, const double* hd, const double* pt,
                   bool invDepth, double* J)  // :343-362
{
    double dj[4], fpj[6], idj[4];
    jacDistortion(c, hd, dj);
    jacFrameProjection(c, s, Rinv, pt, invDepth, fpj);
    if (!invertLU(dj, 2, idj)) idj[0] = idj[1] = idj[2] = idj[3] = 0;
    J[0] = idj[0] * fpj[0] + idj[1] * fpj[3];
    J[1] = idj[0] * fpj[1] + idj[1] * fpj[4];
    J[2] = idj[0] * fpj[2] + idj[1] * fpj[5];
    J[3] = idj[2] * fpj[0] + idj[3] * fpj[3];
    J[4] = idj[2] * fpj[1] + idj[3] * fpj[4];
    J[5] = idj[2] * fpj[2] + idj[3] * fpj[5];
}

// :369-399.  Reproduces the reference's indexing slip: element [1] is never written and
// element [2] is written twice (and, in the rho overload, scaled by rho twice).
void jacCameraAxisRightPart(const double* Rinv, double* J)
{
    J[0] = -Rinv[0];
    J[2] = -Rinv[1];
    J[2] = -Rinv[2];
    J[3] = -Rinv[3];
    J[4] = -Rinv[4];
    J[5] = -Rinv[5];
    J[6] = -Rinv[6];
    J[7] = -Rinv[7];
    J[8] = -Rinv[8];
}
void jacCameraAxisRightPart(const double* Rinv, double rho, double* J)
{
    jacCameraAxisRightPart(Rinv, J);
    J[0] *= rho;
    J[2] *= rho;
    J[2] *= rho;
    J[3] *= rho;
    J[4] *= rho;
    J[5] *= rho;
    J[6] *= rho;
    J[7] *= rho;
    J[8] *= rho;
}

// E/CommonFunctions.cpp:87-145 makeJacobianOfQuaternionToRotationMatrix (3x4)
void jacQuatToRot(const double* q, const double* a, double* J)
{
    double q0 = q[0], qx = q[1], qy = q[2], qz = q[3];
    double T[9], t[3];
    T[0] = 2 * q0; T[1] = -2 * qz; T[2] = 2 * qy;
    T[3] = 2 * qz; T[4] = 2 * q0; T[5] = -2 * qx;
    T[6] = -2 * qy; T[7] = 2 * qx; T[8] = 2 * q0;
    rotMulVec(T, a, t);
    J[0] = t[0]; J[4] = t[1]; J[8] = t[2];
    T[0] = 2 * qx; T[1] = 2 * qy; T[2] = 2 * qz;
    T[3] = 2 * qy; T[4] = -2 * qx; T[5] = -2 * q0;
    T[6] = 2 * qz; T[7] = 2 * q0; T[8] = -2 * qx;
    rotMulVec(T, a, t);
    J[1] = t[0]; J[5] = t[1]; J[9] = t[2];
    T[0] = -2 * qy; T[1] = 2 * qx; T[2] = 2 * q0;
    T[3] = 2 * qx; T[4] = 2 * qy; T[5] = 2 * qz;
    T[6] = -2 * q0; T[7] = 2 * qz; T[8] = -2 * qy;
    rotMulVec(T, a, t);
    J[2] = t[0]; J[6] = t[1]; J[10] = t[2];
    T[0] = -2 * qz; T[1] = -2 * q0; T[2] = 2 * qx;
    T[3] = 2 * q0; T[4] = -2 * qz; T[5] = 2 * qy;
    T[6] = 2 * qx; T[7] = 2 * qy; T[8] = 2 * qz;
    rotMulVec(T, a, t);
    J[3] = t[0]; J[7] = t[1]; J[11] = t[2];
}

// makeJacobianOfMeasurementByState :491-503 -> Hx (2x13, only cols 0..6 written)
void jacMeasurementByState(const orc_params& c, const State& s, const double* Rinv, const double* hd,
                           const double* pt, bool invDepth, double* Hx /*2x13*/)
{
    // dh/dr :411-436
    double carp[9] = {0}, pj[6] = {0};
    if (invDepth)
        jacCameraAxisRightPart(Rinv, pt[5], carp);
    else
        jacCameraAxisRightPart(Rinv, carp);
    jacProjection(c, s, Rinv, hd, pt, invDepth, pj);
    double r23[6];
    matrixMult(pj, 2, 3, carp, 3, r23);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) Hx[i * 13 + j] = r23[i * 3 + j];
    // dh/dq :442-484
    double qc[4] = {s.orientation[0], -s.orientation[1], -s.orientation[2], -s.orientation[3]};
    double ca[3] = {pt[0] - s.position[0], pt[1] - s.position[1], pt[2] - s.position[2]};
    if (invDepth) {
        double m[3];
        double rho = pt[5];
        makeDirectionalVector(pt[3], pt[4], m);
        ca[0] = ca[0] * rho + m[0];
        ca[1] = ca[1] * rho + m[1];
        ca[2] = ca[2] * rho + m[2];
    }
    double qj[12] = {0};
    jacQuatToRot(qc, ca, qj);
    qj[1] = -qj[1]; qj[5] = -qj[5]; qj[9] = -qj[9];
    qj[2] = -qj[2]; qj[6] = -qj[6]; qj[10] = -qj[10];
    qj[3] = -qj[3]; qj[7] = -qj[7]; qj[11] = -qj[11];
    jacProjection(c, s, Rinv, hd, pt, invDepth, pj);
    double r24[8];
    matrixMult(pj, 2, 3, qj, 4, r24);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 4; ++j) Hx[i * 13 + 3 + j] = r24[i * 4 + j];
}

void jacMeasurementByFeatureDepth(const orc_params& c, const State& s, const double* Rinv, const double* hd,
                                  const double* pt, double* Hf /*2x3*/)  // :510-523
{
    double pj[6];
    jacProjection(c, s, Rinv, hd, pt, false, pj);
    matrixMult(pj, 2, 3, Rinv, 3, Hf);
}

void jacMeasurementByFeatureInvDepth(const orc_params& c, const State& s, const double* Rinv, const double* hd,
                                     const double* pt, double* Hf /*2x6*/)  // :530-589
{
    double theta = pt[3], phi = pt[4], rho = pt[5];
    double cosphi = cos(phi), costheta = cos(theta), sintheta = sin(theta), sinphi = sin(phi);
    double dTheta[3] = {cosphi * costheta, 0, -cosphi * sintheta};
    double dPhi[3] = {-sinphi * sintheta, -cosphi, -sinphi * costheta};
    double pc[3] = {pt[0] - s.position[0], pt[1] - s.position[1], pt[2] - s.position[2]};
    double rTheta[3], rPhi[3], rPc[3], rRho[9];
    rotMulVec(Rinv, dTheta, rTheta);
    rotMulVec(Rinv, dPhi, rPhi);
    rotMulVec(Rinv, pc, rPc);
    for (int i = 0; i < 9; ++i) rRho[i] = rho * Rinv[i];
    double D[18];
    for (int i = 0; i < 3; ++i) {
        D[i * 6 + 0] = rRho[0 + 3 * i];
        D[i * 6 + 1] = rRho[1 + 3 * i];
        D[i * 6 + 2] = rRho[2 + 3 * i];
        D[i * 6 + 3] = rTheta[i];
        D[i * 6 + 4] = rPhi[i];
        // :571 places the UN-rotated (y - r)[i] in the rho column (rotationByPointInCameraAxis is
        // computed at :553 but not used)
        D[i * 6 + 5] = pc[i];
    }
    (void)rPc;
    double pj[6];
    jacProjection(c, s, Rinv, hd, pt, true, pj);
    matrixMult(pj, 2, 3, D, 6, Hf);
}

// makeMeasurementCovariance :595-658 (keeps the literal dense 2 x n hiByP)
void makeMeasurementCovariance(const orc_params& c, const State& s, const double* Rinv, const Mat& P, Prediction& p)
{
    const Feature& f = s.features[p.featureIndex];
    bool invDepth = f.type == TYPE_INVERSE_DEPTH;
    const int dim = f.dim, cp = f.covPos, n = P.c;
    std::memset(p.Hx, 0, sizeof(p.Hx));
    std::memset(p.Hf, 0, sizeof(p.Hf));
    jacMeasurementByState(c, s, Rinv, p.h, f.position, invDepth, p.Hx);
    if (invDepth)
        jacMeasurementByFeatureInvDepth(c, s, Rinv, p.h, f.position, p.Hf);
    else
        jacMeasurementByFeatureDepth(c, s, Rinv, p.h, f.position, p.Hf);
    p.covPos = cp;
    p.dim = dim;
    std::vector<double> hiByP((size_t)2 * n, 0.0);
    for (int r = 0; r < 2; ++r) {
        double* o = &hiByP[(size_t)r * n];
        for (int k = 0; k < dim; ++k) {
            double a = p.Hf[r * dim + k];
            const double* pr = P.row(cp + k);
            for (int j = 0; j < n; ++j) o[j] += a * pr[j];
        }
        for (int k = 0; k < 13; ++k) {
            double a = p.Hx[r * 13 + k];
            const double* pr = P.row(k);
            for (int j = 0; j < n; ++j) o[j] += a * pr[j];
        }
    }
    for (int r = 0; r < 2; ++r)
        for (int q = 0; q < 2; ++q) {
            double a = 0, b = 0;
            for (int k = 0; k < 13; ++k) a += hiByP[(size_t)r * n + k] * p.Hx[q * 13 + k];
            for (int k = 0; k < dim; ++k) b += hiByP[(size_t)r * n + cp + k] * p.Hf[q * dim + k];
            p.S[r * 2 + q] = a + b + (r == q ? 1.0 : 0.0);
        }
}

void predictMeasurementCovariance(const orc_params& c, const State& s, const Mat& P, std::vector<Prediction>& preds)  // :666-700
{
    double Rinv[9];
    invertLU(s.R, 3, Rinv);
    for (Prediction& p : preds) makeMeasurementCovariance(c, s, Rinv, P, p);
}

// build the dense 2 x n Jacobian row pair (:656-657)
void denseRows(const Prediction& p, int n, double* row0, double* row1)
{
    std::fill(row0, row0 + n, 0.0);
    std::fill(row1, row1 + n, 0.0);
    for (int j = 0; j < 13; ++j) {
        row0[j] = p.Hx[j];
        row1[j] = p.Hx[13 + j];
    }
    for (int j = 0; j < p.dim; ++j) {
        row0[p.covPos + j] = p.Hf[j];
        row1[p.covPos + j] = p.Hf[p.dim + j];
    }
}

// ---------------------------------------------------------------------------------------------
// E/Matching.cpp
// ---------------------------------------------------------------------------------------------
extern const uint8_t kPopCount[256];

double computeDistance(const uint8_t* a, const uint8_t* b)  // :47-93 (CV_8U branch, 32 bytes)
{
    unsigned d = 0;
    for (int j = 0; j < 32; ++j) d += kPopCount[a[j] ^ b[j]];
    return (double)d;
}

struct DMatch { int trainIdx; float distance; };

void findBestNMatches(unsigned nBest, const uint8_t* desc, const uint8_t* cand, int nCand, const uint8_t* mask,
                      std::list<DMatch>& best)  // :116-144
{
    best.clear();
    double minDistance = -1.0L;
    for (int i = 0; i < nCand; ++i) {
        if (!mask || (mask && mask[i])) {
            double distance = computeDistance(desc, cand + (size_t)i * 32);
            if (distance < minDistance || best.size() < 2) {
                minDistance = minDistance < 0 ? distance : std::min(minDistance, distance);
                DMatch m = {i, static_cast<float>(distance)};
                best.push_front(m);
                if (best.size() > nBest) best.pop_back();
            }
        }
    }
}

void matchPredictedFeatures(orc_filter* f, const float* kpxy, const uint8_t* kpdesc, int nkp)  // :181-264
{
    const orc_params& c = f->p;
    const int W = c.pixels_x, H = c.pixels_y;
    f->mask.assign((size_t)W * H, 0);
    int maxAxes = (int)(2.0L * std::max(c.pixels_x, c.pixels_y));
    for (const Prediction& p : f->preds) drawUncertaintyEllipse(f->mask.data(), W, H, p.h[0], p.h[1], p.S, maxAxes);

    // detector->detect(image, keypoints, mask) (:206): the front end runs on the full frame and the mask
    // is applied as OpenCV's KeyPointsFilter::runByPixelsMask post-filter (order preserving);
    // extractor->compute (:210) is the external front end's job (descriptors arrive with the keypoints).
    f->kpOk.assign(nkp, 0);
    std::vector<int> kept;
    for (int j = 0; j < nkp; ++j) {
        int yy = (int)(kpxy[2 * j + 1] + 0.5f), xx = (int)(kpxy[2 * j] + 0.5f);
        bool ok = xx >= 0 && xx < W && yy >= 0 && yy < H && f->mask[(size_t)yy * W + xx] != 0;
        f->kpOk[j] = ok;
        if (ok) kept.push_back(j);
    }
    const int nk = (int)kept.size();
    std::vector<uint8_t> desc((size_t)nk * 32);
    for (int j = 0; j < nk; ++j) std::memcpy(&desc[(size_t)j * 32], kpdesc + (size_t)kept[j] * 32, 32);

    f->matches.clear();
    std::vector<uint8_t> m(nk);
    for (const Prediction& p : f->preds) {
        std::fill(m.begin(), m.end(), 0);
        const Feature& feat = f->state.features[p.featureIndex];
        float aw, ah;
        double angle;
        ellipseParams(p.S, aw, ah, angle);
        // cv::Size2f -> cv::Size at the call (:232-235): saturate_cast<int>(float) = round-half-even;
        // cv::Point2d -> cv::Point2f for the centre.
        int iw = cvRoundF(aw), ih = cvRoundF(ah);
        float cxf = (float)p.h[0], cyf = (float)p.h[1];
        for (int j = 0; j < nk; ++j) {
            int kj = kept[j];
            if (pointIsInsideEllipse(kpxy[2 * kj], kpxy[2 * kj + 1], cxf, cyf, iw, ih, angle)) m[j] = 1;
        }
        // matchICDescriptors :148-177
        std::list<DMatch> best;
        findBestNMatches(2, feat.desc, desc.data(), nk, m.data(), best);
        if (best.size() >= 2) markMargin(MG_RATIO, (double)best.front().distance - (double)best.back().distance * c.matching_coef);
        if (best.size() == 1 ||
            (best.size() >= 2 && best.front().distance <= best.back().distance * c.matching_coef)) {
            const DMatch& bm = best.front();
            Match mt;
            mt.featureIndex = p.featureIndex;
            mt.kpIndex = kept[bm.trainIdx];
            mt.z[0] = static_cast<double>(kpxy[2 * mt.kpIndex]);
            mt.z[1] = static_cast<double>(kpxy[2 * mt.kpIndex + 1]);
            mt.distance = bm.distance;
            std::memcpy(mt.desc, &desc[(size_t)bm.trainIdx * 32], 32);
            f->matches.push_back(mt);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// E/Update.cpp
// ---------------------------------------------------------------------------------------------
// determineKalmanGain :92-109 (dense H, dense products, LU inverse)
void determineKalmanGain(const Mat& H, const Mat& P, double pixelErr, Mat& K)
{
    const int k = H.r;
    Mat Ht = transpose(H);
    Mat PHt = mul(P, Ht);
    Mat S = mul(H, PHt);
    for (int i = 0; i < k; ++i) S(i, i) += pixelErr;
    K = mul(PHt, inv(S));
}

// stateUpdate :116-208
void stateUpdate(const Mat& K, const std::vector<const Match*>& ms, const std::vector<const Prediction*>& ps, State& s)
{
    const int np = (int)ps.size();
    Mat nu(2 * np, 1);
    for (int i = 0; i < np; ++i) {
        double dx = ms[i]->z[0] - ps[i]->h[0];
        double dy = ms[i]->z[1] - ps[i]->h[1];
        nu(2 * i, 0) = fabs(dx) > DELTA ? dx : 0.0L;
        nu(2 * i + 1, 0) = fabs(dy) > DELTA ? dy : 0.0L;
        markMargin(MG_DEADBAND, dx == 0.0 ? 0.0 : fabs(dx) - DELTA);
        markMargin(MG_DEADBAND, dy == 0.0 ? 0.0 : fabs(dy) - DELTA);
    }
    Mat kd = mul(K, nu);
    const double* v = kd.d.data();
    for (int i = 0; i < kd.r; ++i) markMargin(MG_DEADBAND, v[i] == 0.0 ? 0.0 : fabs(v[i]) - DELTA);
    for (int i = 0; i < 3; ++i)
        if (fabs(v[i]) > DELTA) s.position[i] += v[i];
    for (int i = 0; i < 4; ++i)
        if (fabs(v[i + 3]) > DELTA) s.orientation[i] += v[i + 3];
    s.setOrientation(s.orientation);
    for (int i = 0; i < 3; ++i)
        if (fabs(v[i + 7]) > DELTA) s.linearVelocity[i] += v[i + 7];
    for (int i = 0; i < 3; ++i)
        if (fabs(v[i + 10]) > DELTA) s.angularVelocity[i] += v[i + 10];
    int acc = 13;
    for (Feature& f : s.features)
        for (int j = 0; j < f.dim; ++j) {
            if (fabs(v[acc]) > DELTA) f.position[j] += v[acc];
            acc++;
        }
}

// updateStateAndCovariance :237-266
void updateStateAndCovariance(const orc_params& c, const std::vector<const Prediction*>& ps,
                              const std::vector<const Match*>& ms, bool updateCov, State& s, Mat& P)
{
    const int np = (int)ps.size();
    const int n = s.stateDim();
    Mat H(2 * np, n);  // joinJacobians :222-232
    for (int i = 0; i < np; ++i) denseRows(*ps[i], n, H.row(2 * i), H.row(2 * i + 1));
    Mat K;
    determineKalmanGain(H, P, c.pixel_error_x, K);
    stateUpdate(K, ms, ps, s);
    if (updateCov) {  // covarianceUpdate :214-218   P = (I - K H) P
        Mat KH = mul(K, H);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) KH(i, j) = (i == j ? 1.0 : 0.0) - KH(i, j);
        P = mul(KH, P);
    }
}

// update :282-319
void updateFull(const orc_params& c, State& s, Mat& P, const std::vector<const Match*>& ms,
                const std::vector<const Prediction*>& ps)
{
    if (ms.empty()) return;
    updateStateAndCovariance(c, ps, ms, true, s, P);
    const int n = P.r;
    for (int i = 0; i < n; ++i)  // 0.5 P + 0.5 P^T
        for (int j = i; j < n; ++j) {
            double a = 0.5L * P(i, j) + 0.5L * P(j, i);
            double b = 0.5L * P(j, i) + 0.5L * P(i, j);
            P(i, j) = a;
            P(j, i) = b;
        }
    // normalizeQuaternionJacobian :45-60
    double r = s.orientation[0], x = s.orientation[1], y = s.orientation[2], z = s.orientation[3];
    double norm = sqrt(r * r + x * x + y * y + z * z);
    double a = 1.0L / (pow(norm, 3));
    Mat J(4, 4);
    double jm[16] = {x * x + y * y + z * z, -r * x, -r * y, -r * z, -x * r, r * r + y * y + z * z, -x * y, -x * z,
                     -y * r, -y * x, r * r + x * x + z * z, -y * z, -z * r, -z * x, -z * y, r * r + x * x + y * y};
    for (int i = 0; i < 16; ++i) J.d[i] = jm[i] * a;
    double nq[4] = {s.orientation[0] / norm, s.orientation[1] / norm, s.orientation[2] / norm, s.orientation[3] / norm};
    s.setOrientation(nq);
    // normalizeCovariance :64-85
    Mat Jt = transpose(J);
    setBlock(P, 0, 3, mul(block(P, 0, 3, 3, 7), Jt));
    setBlock(P, 3, 0, mul(J, block(P, 3, 7, 0, 3)));
    setBlock(P, 3, 3, mul(mul(J, block(P, 3, 7, 3, 7)), Jt));
    if (n > 7) {
        setBlock(P, 3, 7, mul(J, block(P, 3, 7, 7, n)));
        setBlock(P, 7, 3, mul(block(P, 7, n, 3, 7), Jt));
    }
}

// ---------------------------------------------------------------------------------------------
// E/1PointRansac.cpp
// ---------------------------------------------------------------------------------------------
void matchesBelowAThreshold(const std::vector<Match>& matches, const std::vector<Prediction>& preds, double thr,
                            std::vector<int>& support)  // :48-84
{
    for (size_t i = 0; i < preds.size(); ++i) {
        const Prediction& p = preds[i];
        bool found = false;
        size_t mi = 0;
        while (!found && mi < matches.size()) {
            const Match& m = matches[mi];
            if (m.featureIndex == p.featureIndex) {
                double xd = m.z[0] - p.h[0];
                double yd = m.z[1] - p.h[1];
                double dist = sqrt(xd * xd + yd * yd);
                markMargin(MG_RANSAC, dist - thr);
                if (dist < thr) support.push_back((int)mi);
                found = true;
            }
            mi++;
        }
    }
}

void ransac(orc_filter* f)  // :101-234
{
    f->inlierIdx.clear();
    f->outlierIdx.clear();
    f->supportCounts.clear();
    f->nHyp = 0;
    f->bestHyp = -1;
    const size_t m = f->matches.size();
    if (m == 0) return;
    unsigned numberOfHipotesis = 1000;
    const double thr = f->p.ransac_threshold;
    std::vector<int> inliers;
    for (unsigned i = 0; i < numberOfHipotesis && i < m; ++i) {
        const Match* match = &f->matches[i];                     // selectRandomMatch :88-92 (deterministic)
        const Prediction* pred = &f->preds[f->matchPred[i]];
        std::vector<const Prediction*> ps(1, pred);
        std::vector<const Match*> ms(1, match);
        State tmp(f->state);                                     // deep copy, :147
        updateStateAndCovariance(f->p, ps, ms, false, tmp, f->P); // updateOnlyState :149
        std::vector<Prediction> hp;
        std::vector<int> unseen, none;
        predictMeasurementState(f->p, tmp, none, hp, unseen);    // :156
        std::vector<int> support;
        matchesBelowAThreshold(f->matches, hp, thr, support);
        f->supportCounts.push_back((int)support.size());
        f->nHyp = (int)i + 1;
        if (support.size() > inliers.size()) {
            inliers = support;
            f->bestHyp = (int)i;
            double e = 1.0L - (double)inliers.size() / (double)m;
            long double ratio = logl(1.0L - f->p.ransac_all_inliers_prob) / logl(1.0L - (1.0L - e));
            // static_cast<int>(long double) then assignment to uint (:177).  A non-finite or
            // out-of-range ratio (all matches are inliers => log(1) = 0 => -inf) is undefined
            // behaviour in C++; x86 (fistp / cvttsd2si) yields the "integer indefinite"
            // 0x80000000, i.e. 2147483648 hypotheses as uint.  The oracle adopts that value.
            int asInt;
            if (!(ratio > -2147483649.0L && ratio < 2147483648.0L))
                asInt = INT_MIN;
            else
                asInt = static_cast<int>(ratio);
            numberOfHipotesis = (unsigned)asInt;
        }
    }
    std::vector<bool> maskv(m, false);
    for (int idx : inliers) maskv[idx] = true;
    for (size_t i = 0; i < m; ++i) (maskv[i] ? f->inlierIdx : f->outlierIdx).push_back((int)i);
}

// ---------------------------------------------------------------------------------------------
// E/EKF.cpp glue
// ---------------------------------------------------------------------------------------------
void alignPredictionsWithMatches(orc_filter* f)  // EKF.cpp:368-392
{
    f->matchPred.clear();
    for (const Match& m : f->matches) {
        for (size_t j = 0; j < f->preds.size(); ++j)
            if (f->preds[j].featureIndex == m.featureIndex) {
                f->matchPred.push_back((int)j);
                break;
            }
    }
}

void rescue(orc_filter* f)  // EKF.cpp:448-506 + rescueOutliers :68-119
{
    f->outlierPreds.clear();
    f->rescuedIdx.clear();
    f->rescuedPred.clear();
    f->outlierMatchesKept = f->outlierIdx;
    const size_t no = f->outlierIdx.size();
    if (no == 0) return;
    std::vector<int> subset;
    for (int mi : f->outlierIdx) subset.push_back(f->matches[mi].featureIndex);
    std::vector<int> unseen;
    predictMeasurementState(f->p, f->state, subset, f->outlierPreds, unseen);
    predictMeasurementCovariance(f->p, f->state, f->P, f->outlierPreds);
    const size_t np = f->outlierPreds.size();
    if (0 < np && np < no) {
        std::vector<int> kept;
        size_t j = 0;
        for (size_t i = 0; i < no && j < np; ++i)
            if (f->matches[f->outlierIdx[i]].featureIndex == f->outlierPreds[j].featureIndex) {
                j++;
                kept.push_back(f->outlierIdx[i]);
            }
        f->outlierMatchesKept = kept;
    }
    // np == 0 with no > 0: the reference would index an empty prediction vector (undefined
    // behaviour, EKF.cpp:501-505 -> :82); the oracle rescues nothing in that case.
    if (np == 0) return;
    for (size_t i = 0; i < f->outlierMatchesKept.size(); ++i) {
        const Match& m = f->matches[f->outlierMatchesKept[i]];
        const Prediction& p = f->outlierPreds[i];
        double d0 = m.z[0] - p.h[0], d1 = m.z[1] - p.h[1];
        double Si[4];
        if (!invertLU(p.S, 2, Si)) Si[0] = Si[1] = Si[2] = Si[3] = 0;
        // dist^T * S^-1 * dist as two small products (EKF.cpp:94)
        double t0 = d0 * Si[0] + d1 * Si[2];
        double t1 = d0 * Si[1] + d1 * Si[3];
        double chi = t0 * d0 + t1 * d1;
        markMargin(MG_CHI2, chi - f->p.ransac_chi2);
        if (chi < f->p.ransac_chi2) {
            f->rescuedIdx.push_back(f->outlierMatchesKept[i]);
            f->rescuedPred.push_back((int)i);
        }
    }
}

double nowUs()
{
    using namespace std::chrono;
    return duration<double, std::micro>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

namespace {
// Core/EKFMath.h:48-58 popCountTable (bit count of a byte)
const uint8_t kPopCount[256] = {
#define B2(n) n, n + 1, n + 1, n + 2
#define B4(n) B2(n), B2(n + 1), B2(n + 1), B2(n + 2)
#define B6(n) B4(n), B4(n + 1), B4(n + 1), B4(n + 2)
    B6(0), B6(1), B6(1), B6(2)
#undef B2
#undef B4
#undef B6
};
}  // namespace

// =============================================================================================
// C interface
// =============================================================================================
extern "C" {

orc_filter* orc_create(const orc_params* p)
{
    orc_filter* f = new orc_filter();
    f->p = *p;
    std::memset(&f->info, 0, sizeof(f->info));
    orc_init(f);
    return f;
}

void orc_destroy(orc_filter* f) { delete f; }

void orc_init(orc_filter* f)  // CommonFunctions.cpp:39-80
{
    State& s = f->state;
    s.features.clear();
    for (int i = 0; i < 3; ++i) {
        s.position[i] = 0.0L;
        s.linearVelocity[i] = 0.0L;
        s.angularVelocity[i] = EPSILON;
    }
    double q[4] = {1.0L, 0.0L, 0.0L, 0.0L};
    s.setOrientation(q);
    f->P = Mat(13, 13);
    for (int i = 0; i < 7; ++i) f->P(i, i) = EPSILON;
    double l2 = f->p.init_linear_accel_sd * f->p.init_linear_accel_sd;
    double a2 = f->p.init_angular_accel_sd * f->p.init_angular_accel_sd;
    for (int i = 0; i < 3; ++i) {
        f->P(i + 7, i + 7) = l2;
        f->P(i + 10, i + 10) = a2;
    }
}

void orc_add_feature(orc_filter* f, const double* uv, const uint8_t* desc32)  // AddMapFeature.cpp:43-350
{
    const orc_params& c = f->p;
    State& s = f->state;
    // undistortPoint :43-59
    auto undistort = [&](const double* in, double* out) {
        double px = in[0] - c.cx, py = in[1] - c.cy;
        double dx = c.dx * px, dy = c.dy * py;
        double rd = dx * dx + dy * dy;
        double dist = 1 + c.k1 * rd + c.k2 * rd * rd;
        out[0] = c.cx + px * dist;
        out[1] = c.cy + py * dist;
    };
    Feature nf;
    std::memset(&nf, 0, sizeof(nf));
    for (int i = 0; i < 3; ++i) nf.position[i] = s.position[i];
    double und[2];
    undistort(uv, und);
    double rp[3] = {-(c.cx - und[0]) / c.fx, -(c.cy - und[1]) / c.fy, 1.0L};
    double rw[3];
    rotMulVec(s.R, rp, rw);
    nf.position[3] = atan2(rw[0], rw[2]);
    nf.position[4] = atan2(-rw[1], sqrt(rw[0] * rw[0] + rw[2] * rw[2]));
    nf.position[5] = c.init_inv_depth_rho;
    nf.type = TYPE_INVERSE_DEPTH;
    nf.dim = 6;
    nf.covPos = f->P.c;
    std::memcpy(nf.desc, desc32, 32);
    s.features.push_back(nf);

    // computeAddFeatureJacobian :116-216
    double Jpo[42] = {0}, Jhr[18] = {0};
    {
        double xyz_c[3] = {rp[0], rp[1], 1.0L};
        double xyz_w[3];
        rotMulVec(s.R, xyz_c, xyz_w);
        double xw = xyz_w[0], yw = xyz_w[1], zw = xyz_w[2];
        double xxzz = xw * xw + zw * zw;
        double dth[3] = {zw / xxzz, 0, -xw / xxzz};
        double sq = sqrt(xw * xw + zw * zw);
        double nsq = xxzz + yw * yw;
        double dph[3] = {xw * yw / (nsq * sq), -sq / nsq, zw * yw / (nsq * sq)};
        double dgw_dq[12] = {0};
        jacQuatToRot(s.orientation, xyz_c, dgw_dq);
        double dth_dq[4], dph_dq[4];
        matrixMult(dth, 1, 3, dgw_dq, 4, dth_dq);
        matrixMult(dph, 1, 3, dgw_dq, 4, dph_dq);
        for (int i = 0; i < 3; ++i) Jpo[i * 7 + i] = 1.0L;
        for (int i = 0; i < 4; ++i) {
            Jpo[3 * 7 + i + 3] = dth_dq[i];
            Jpo[4 * 7 + i + 3] = dph_dq[i];
        }
        double sub[6] = {0};
        matrixMult(dth, 1, 3, s.R, 3, &sub[0]);
        matrixMult(dph, 1, 3, s.R, 3, &sub[3]);
        double fku_inv = 1.0L / c.fx, fkv_inv = 1.0L / c.fy;
        double dgc_dhu[6] = {fku_inv, 0, 0, fkv_inv, 0, 0};
        double sub2[4];
        matrixMult(sub, 2, 3, dgc_dhu, 2, sub2);
        // computeUndistortPointJacobian :66-92
        double ud = uv[0], vd = uv[1];
        double xd = (ud - c.cx) * c.dx, yd = (vd - c.cy) * c.dy;
        double rd2 = xd * xd + yd * yd;
        double k12 = c.k1 + 2.0L * c.k2 * rd2;
        double k1p = 1.0L + c.k1 * rd2 + c.k2 * rd2 * rd2;
        double dx2 = 2.0L * c.dx * c.dx, dy2 = 2.0L * c.dy * c.dy;
        double dhu[4];
        dhu[0] = k1p + (ud - c.cx) * k12 * ((ud - c.cx) * dx2);
        dhu[1] = (ud - c.cx) * k12 * ((vd - c.cy) * dy2);
        dhu[2] = (vd - c.cy) * k12 * ((ud - c.cx) * dx2);
        dhu[3] = (vd - c.cy) * k12 * ((vd - c.cy) * dy2) + k1p;
        double sub3[4];
        matrixMult(sub2, 2, 2, dhu, 2, sub3);
        Jhr[9] = sub3[0];
        Jhr[10] = sub3[1];
        Jhr[12] = sub3[2];
        Jhr[13] = sub3[3];
        Jhr[17] = 1.0L;
    }
    // addFeatureToCovarianceMatrix :221-289
    const int n = f->P.r;
    Mat J(6, 7), JH(6, 3), N3(3, 3);
    J.d.assign(Jpo, Jpo + 42);
    JH.d.assign(Jhr, Jhr + 18);
    N3(0, 0) = c.pixel_error_x * c.pixel_error_x;
    N3(1, 1) = c.pixel_error_y * c.pixel_error_y;
    N3(2, 2) = c.inverse_depth_rho_sd * c.inverse_depth_rho_sd;
    Mat rows7 = block(f->P, 0, 7, 0, n);
    Mat cols7 = block(f->P, 0, n, 0, 7);
    Mat Jt = transpose(J);
    Mat newVsPrev = mul(J, rows7);       // 6 x n
    Mat prevVsNew = mul(cols7, Jt);      // n x 6
    Mat nv7 = block(newVsPrev, 0, 6, 0, 7);
    Mat a = mul(nv7, Jt);
    Mat b = mul(mul(JH, N3), transpose(JH));
    Mat NP(n + 6, n + 6);
    setBlock(NP, 0, 0, f->P);
    setBlock(NP, n, 0, newVsPrev);
    setBlock(NP, 0, n, prevVsNew);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) NP(n + i, n + j) = a(i, j) + b(i, j);
    f->P = NP;
}

void orc_remove_features(orc_filter* f, const uint8_t* flags)  // MapManagement.cpp:168-259
{
    State& s = f->state;
    const int n = f->P.r;
    std::vector<int> keep;
    for (int i = 0; i < 13; ++i) keep.push_back(i);
    std::vector<Feature> nf;
    for (size_t i = 0; i < s.features.size(); ++i) {
        if (flags[i]) continue;
        Feature ft = s.features[i];
        int newPos = (int)keep.size();
        for (int j = 0; j < ft.dim; ++j) keep.push_back(ft.covPos + j);
        ft.covPos = newPos;
        nf.push_back(ft);
    }
    (void)n;
    Mat NP((int)keep.size(), (int)keep.size());
    for (size_t i = 0; i < keep.size(); ++i)
        for (size_t j = 0; j < keep.size(); ++j) NP((int)i, (int)j) = f->P(keep[i], keep[j]);
    f->P = NP;
    s.features = nf;
}

// removeBadMapFeatures (MapManagement.cpp:279-307): float ratio timesMatched / timesPredicted below the threshold.
// 0/0 is NaN and x/0 is inf, both compare false: a feature never predicted is never "bad".
static void badFeatureFlags(const orc_filter* f, double goodPct, std::vector<uint8_t>& flags)
{
    const State& s = f->state;
    flags.assign(s.features.size(), 0);
    for (size_t i = 0; i < s.features.size(); ++i) {
        float pct = static_cast<float>(s.features[i].timesMatched) / static_cast<float>(s.features[i].timesPredicted);
        if (pct < goodPct) flags[i] = 1;
    }
}

// computeLinearityIndex (MapManagement.cpp:311-341)
static double linearityIndex(const orc_filter* f, const Feature& mf)
{
    const State& s = f->state;
    int idx = mf.covPos + mf.dim - 1;
    double invDepthError = sqrt(f->P(idx, idx));
    double invDepthValue = mf.position[mf.dim - 1];
    double sigma = invDepthError / (invDepthValue * invDepthValue);
    double xyz[3], m[3];
    makeDirectionalVector(mf.position[3], mf.position[4], m);       // changeInverseDepthToDepth, CommonFunctions.cpp
    for (int i = 0; i < 3; ++i) xyz[i] = mf.position[i] + m[i] / mf.position[5];
    double toCam[3], toFirst[3];
    for (int i = 0; i < 3; ++i) {
        toCam[i] = xyz[i] - s.position[i];
        toFirst[i] = xyz[i] - mf.position[i];
    }
    double dot = 0.0L;
    for (int i = 0; i < 3; ++i) dot += toCam[i] * toFirst[i];
    double dFirst = sqrt(toFirst[0] * toFirst[0] + toFirst[1] * toFirst[1] + toFirst[2] * toFirst[2]);
    double dCam = sqrt(toCam[0] * toCam[0] + toCam[1] * toCam[1] + toCam[2] * toCam[2]);
    double cosAlpha = dot / (dFirst * dCam);
    return 4.0L * sigma * cosAlpha / dCam;
}

// convertToDepth (MapManagement.cpp:345-497): 6 -> 3 rows/columns with the 3x6 Jacobian of (x,y,z) + m(theta,phi)/rho
static void convertToDepth(orc_filter* f, int fi)
{
    State& s = f->state;
    Feature& mf = s.features[fi];
    double theta = mf.position[3], phi = mf.position[4], rho = mf.position[5];
    double mi[3], xyz[3];
    makeDirectionalVector(theta, phi, mi);
    for (int i = 0; i < 3; ++i) xyz[i] = mf.position[i] + mi[i] / rho;
    Mat J(3, 6);
    J(0, 0) = J(1, 1) = J(2, 2) = 1.0L;
    J(0, 3) = cos(phi) * cos(theta) / rho;  J(1, 3) = 0.0;                J(2, 3) = -cos(phi) * sin(theta) / rho;
    J(0, 4) = -sin(phi) * sin(theta) / rho; J(1, 4) = -cos(phi) / rho;    J(2, 4) = -sin(phi) * cos(theta) / rho;
    for (int i = 0; i < 3; ++i) J(i, 5) = -mi[i] / (rho * rho);
    const int n = f->P.r, c0 = mf.covPos, c1 = c0 + 6;
    Mat P6n = block(f->P, c0, c1, 0, n);
    Mat sub3n = mul(J, P6n);                                   // 3 x n
    Mat sub36 = block(sub3n, 0, 3, c0, c1);
    Mat Jt = transpose(J);
    Mat d33 = mul(sub36, Jt);
    Mat colsK = mul(block(f->P, 0, c0, c0, c1), Jt);           // k x 3
    Mat colsL = (c1 < n) ? mul(block(f->P, c1, n, c0, c1), Jt) : Mat(0, 3);
    Mat NP(n - 3, n - 3);
    auto src = [&](int i) { return i < c0 + 3 ? i : i + 3; };  // new index -> old index (outside the feature block)
    for (int i = 0; i < n - 3; ++i)
        for (int j = 0; j < n - 3; ++j) {
            const bool fi_ = (i >= c0 && i < c0 + 3), fj_ = (j >= c0 && j < c0 + 3);
            double val;
            if (fi_ && fj_) val = d33(i - c0, j - c0);
            else if (fi_) val = sub3n(i - c0, src(j));
            else if (fj_) val = (i < c0) ? colsK(i, j - c0) : colsL(src(i) - c1, j - c0);
            else val = f->P(src(i), src(j));
            NP(i, j) = val;
        }
    f->P = NP;
    mf.dim = 3;
    mf.type = TYPE_DEPTH;
    for (int i = 0; i < 3; ++i) mf.position[i] = xyz[i];
    for (int i = 3; i < 6; ++i) mf.position[i] = 0.0;
    for (Feature& o : s.features)
        if (o.covPos > c0) o.covPos = o.covPos - 6 + 3;
}

void orc_set_hit_counters(orc_filter* f, const int32_t* tp, const int32_t* tm)
{
    for (size_t i = 0; i < f->state.features.size(); ++i) {
        f->state.features[i].timesPredicted = (unsigned)tp[i];
        f->state.features[i].timesMatched = (unsigned)tm[i];
    }
}

int32_t orc_remove_bad_features(orc_filter* f, double goodPct, uint8_t* removed)  // MapManagement.cpp:279-307
{
    std::vector<uint8_t> flags;
    badFeatureFlags(f, goodPct, flags);
    int32_t c = 0;
    for (size_t i = 0; i < flags.size(); ++i) {
        c += flags[i];
        if (removed) removed[i] = flags[i];
    }
    if (c) orc_remove_features(f, flags.data());
    return c;
}

int32_t orc_convert_features(orc_filter* f, double threshold)  // MapManagement.cpp:501-524: at most ONE per call
{
    State& s = f->state;
    for (size_t i = 0; i < s.features.size(); ++i) {
        if (s.features[i].type != TYPE_INVERSE_DEPTH) continue;
        if (linearityIndex(f, s.features[i]) < threshold) {
            convertToDepth(f, (int)i);
            return (int32_t)i;
        }
    }
    return -1;
}

// EKF.cpp:575-592 up to (not including) the detection of new features.  removed[i] for the N features present
// before the call: 0 kept, 1 removed as bad, 2 removed as unseen; *converted = index (in the NEW numbering) of the
// feature converted to XYZ or -1.  Returns newFeaturesNeededCount.
int32_t orc_map_management(orc_filter* f, const orc_map_policy* pol, uint8_t* removed, int32_t* converted)
{
    State& s = f->state;
    const size_t N0 = s.features.size();
    int needed = pol->min_matches_per_image - (int)(f->inlierIdx.size() + f->rescuedIdx.size());
    std::vector<uint8_t> rem(N0, 0), bad;
    badFeatureFlags(f, pol->good_feature_matching_percent, bad);
    std::vector<int> alive;                  // old index of every surviving feature
    for (size_t i = 0; i < N0; ++i)
        if (bad[i]) rem[i] = 1; else alive.push_back((int)i);
    bool anyBad = alive.size() != N0;
    if (anyBad) orc_remove_features(f, bad.data());
    if (needed > 0 && (pol->always_remove_unseen ||
                       (pol->max_map_features_count > 0 && (int)s.features.size() + needed > pol->max_map_features_count) ||
                       (pol->max_map_size > 0 && f->P.r + needed * 6 > pol->max_map_size))) {
        std::vector<uint8_t> un(s.features.size(), 0);
        bool any = false;
        for (int oldIdx : f->unseen)          // unseen features that survived removeBadMapFeatures
            for (size_t a = 0; a < alive.size(); ++a)
                if (alive[a] == oldIdx) { un[a] = 1; rem[oldIdx] = 2; any = true; }
        if (any) orc_remove_features(f, un.data());
    }
    int32_t conv = orc_convert_features(f, pol->linearity_index_threshold);
    if (removed) std::memcpy(removed, rem.data(), N0);
    if (converted) *converted = conv;
    return needed;
}

void orc_dims(const orc_filter* f, int32_t* n, int32_t* nf)
{
    *n = f->P.r;
    *nf = (int32_t)f->state.features.size();
}

void orc_get_state(const orc_filter* f, double* x, double* P)
{
    const State& s = f->state;
    if (x) {
        for (int i = 0; i < 3; ++i) {
            x[i] = s.position[i];
            x[7 + i] = s.linearVelocity[i];
            x[10 + i] = s.angularVelocity[i];
        }
        for (int i = 0; i < 4; ++i) x[3 + i] = s.orientation[i];
        for (const Feature& ft : s.features)
            for (int j = 0; j < ft.dim; ++j) x[ft.covPos + j] = ft.position[j];
    }
    if (P) std::memcpy(P, f->P.d.data(), sizeof(double) * f->P.d.size());
}

void orc_get_features(const orc_filter* f, int32_t* type, int32_t* off, uint8_t* desc, int32_t* tp, int32_t* tm)
{
    const State& s = f->state;
    for (size_t i = 0; i < s.features.size(); ++i) {
        if (type) type[i] = s.features[i].type;
        if (off) off[i] = s.features[i].covPos;
        if (desc) std::memcpy(desc + i * 32, s.features[i].desc, 32);
        if (tp) tp[i] = (int32_t)s.features[i].timesPredicted;
        if (tm) tm[i] = (int32_t)s.features[i].timesMatched;
    }
}

void orc_set_state(orc_filter* f, int32_t n, int32_t nf, const double* x, const int32_t* type, const int32_t* off,
                   const double* P, const uint8_t* desc)
{
    State& s = f->state;
    for (int i = 0; i < 3; ++i) {
        s.position[i] = x[i];
        s.linearVelocity[i] = x[7 + i];
        s.angularVelocity[i] = x[10 + i];
    }
    s.setOrientation(x + 3);
    s.features.clear();
    for (int i = 0; i < nf; ++i) {
        Feature ft;
        std::memset(&ft, 0, sizeof(ft));
        ft.type = type[i];
        ft.dim = type[i] == TYPE_INVERSE_DEPTH ? 6 : 3;
        ft.covPos = off[i];
        for (int j = 0; j < ft.dim; ++j) ft.position[j] = x[off[i] + j];
        if (desc) std::memcpy(ft.desc, desc + (size_t)i * 32, 32);
        s.features.push_back(ft);
    }
    f->P = Mat(n, n);
    std::memcpy(f->P.d.data(), P, sizeof(double) * (size_t)n * n);
}

void orc_predict(orc_filter* f)  // stateAndCovariancePrediction :244-253
{
    double dt = 1.0L;
    predictCovariance(f->P, f->state, dt, f->p);
    predictState(f->state, dt);
}

void orc_measure(orc_filter* f)  // predictCameraMeasurements :705-719
{
    f->preds.clear();
    f->unseen.clear();
    std::vector<int> none;
    predictMeasurementState(f->p, f->state, none, f->preds, f->unseen);
    predictMeasurementCovariance(f->p, f->state, f->P, f->preds);
}

void orc_match(orc_filter* f, const float* kp, const uint8_t* desc, int32_t nkp)
{
    matchPredictedFeatures(f, kp, desc, nkp);
    alignPredictionsWithMatches(f);
}

void orc_ransac(orc_filter* f) { ransac(f); }

void orc_update_li(orc_filter* f)  // EKF.cpp:430
{
    std::vector<const Match*> ms;
    std::vector<const Prediction*> ps;
    for (int mi : f->inlierIdx) {
        ms.push_back(&f->matches[mi]);
        ps.push_back(&f->preds[f->matchPred[mi]]);
    }
    updateFull(f->p, f->state, f->P, ms, ps);
}

void orc_rescue(orc_filter* f) { rescue(f); }

void orc_update_hi(orc_filter* f)  // EKF.cpp:527-532
{
    std::vector<const Match*> ms;
    std::vector<const Prediction*> ps;
    for (size_t i = 0; i < f->rescuedIdx.size(); ++i) {
        ms.push_back(&f->matches[f->rescuedIdx[i]]);
        ps.push_back(&f->outlierPreds[f->rescuedPred[i]]);
    }
    updateFull(f->p, f->state, f->P, ms, ps);
}

void orc_update_map_features(orc_filter* f)  // MapManagement.cpp:77-113 with EKF.cpp:552-572
{
    State& s = f->state;
    for (const Prediction& p : f->preds) s.features[p.featureIndex].timesPredicted++;
    std::vector<int> all = f->inlierIdx;
    all.insert(all.end(), f->rescuedIdx.begin(), f->rescuedIdx.end());
    for (int mi : all) s.features[f->matches[mi].featureIndex].timesMatched++;
    for (int mi : all) std::memcpy(s.features[f->matches[mi].featureIndex].desc, f->matches[mi].desc, 32);
}

void orc_step(orc_filter* f, const float* kp, const uint8_t* desc, int32_t nkp, orc_frame_info* info)
{
    orc_frame_info& I = f->info;
    std::memset(&I, 0, sizeof(I));
    double t0 = nowUs();
    orc_predict(f);
    orc_measure(f);
    double t1 = nowUs();
    I.us_prediction = t1 - t0;
    orc_match(f, kp, desc, nkp);
    double t2 = nowUs();
    I.us_matching = t2 - t1;
    orc_ransac(f);
    double t3 = nowUs();
    I.us_ransac = t3 - t2;
    orc_update_li(f);
    double t4 = nowUs();
    I.us_update_li = t4 - t3;
    orc_rescue(f);
    double t5 = nowUs();
    I.us_rescue = t5 - t4;
    orc_update_hi(f);
    double t6 = nowUs();
    I.us_update_hi = t6 - t5;
    orc_update_map_features(f);
    double t7 = nowUs();
    I.us_map = t7 - t6;
    I.n = f->P.r;
    I.n_features = (int32_t)f->state.features.size();
    I.n_predicted = (int32_t)f->preds.size();
    I.n_matches = (int32_t)f->matches.size();
    I.n_hypotheses = f->nHyp;
    I.n_inliers = (int32_t)f->inlierIdx.size();
    I.n_outliers = (int32_t)f->outlierIdx.size();
    I.n_rescued = (int32_t)f->rescuedIdx.size();
    if (info) *info = I;
}

void orc_last_info(const orc_filter* f, orc_frame_info* info) { *info = f->info; }

void orc_get_measure(const orc_filter* f, uint8_t* vis, double* h, double* S, double* Hx, double* Hf, double* ell)
{
    const size_t N = f->state.features.size();
    if (vis) std::memset(vis, 0, N);
    for (const Prediction& p : f->preds) {
        const int i = p.featureIndex;
        if (vis) vis[i] = 1;
        if (h) { h[2 * i] = p.h[0]; h[2 * i + 1] = p.h[1]; }
        if (S) std::memcpy(S + 4 * i, p.S, 4 * sizeof(double));
        if (Hx)
            for (int r = 0; r < 2; ++r)
                for (int j = 0; j < 7; ++j) Hx[14 * i + 7 * r + j] = p.Hx[13 * r + j];
        if (Hf) {
            std::memset(Hf + 12 * i, 0, 12 * sizeof(double));
            for (int r = 0; r < 2; ++r)
                for (int j = 0; j < p.dim; ++j) Hf[12 * i + 6 * r + j] = p.Hf[p.dim * r + j];
        }
        if (ell) {
            float aw, ah;
            double ang;
            ellipseParams(p.S, aw, ah, ang);
            ell[3 * i] = aw;
            ell[3 * i + 1] = ah;
            ell[3 * i + 2] = ang;
        }
    }
}

void orc_get_match(const orc_filter* f, uint8_t* matched, double* z, int32_t* kp, float* dist)
{
    const size_t N = f->state.features.size();
    if (matched) std::memset(matched, 0, N);
    if (kp) for (size_t i = 0; i < N; ++i) kp[i] = -1;
    for (const Match& m : f->matches) {
        const int i = m.featureIndex;
        if (matched) matched[i] = 1;
        if (z) { z[2 * i] = m.z[0]; z[2 * i + 1] = m.z[1]; }
        if (kp) kp[i] = m.kpIndex;
        if (dist) dist[i] = m.distance;
    }
}

void orc_get_ransac(const orc_filter* f, uint8_t* inl, uint8_t* outl, int32_t* nHyp, int32_t* best, int32_t* counts)
{
    const size_t N = f->state.features.size();
    if (inl) std::memset(inl, 0, N);
    if (outl) std::memset(outl, 0, N);
    if (inl) for (int mi : f->inlierIdx) inl[f->matches[mi].featureIndex] = 1;
    if (outl) for (int mi : f->outlierIdx) outl[f->matches[mi].featureIndex] = 1;
    if (nHyp) *nHyp = f->nHyp;
    if (best) *best = f->bestHyp;
    if (counts) for (size_t i = 0; i < f->supportCounts.size(); ++i) counts[i] = f->supportCounts[i];
}

void orc_get_rescue(const orc_filter* f, uint8_t* resc)
{
    const size_t N = f->state.features.size();
    std::memset(resc, 0, N);
    for (int mi : f->rescuedIdx) resc[f->matches[mi].featureIndex] = 1;
}

void orc_get_mask(const orc_filter* f, uint8_t* mask, uint8_t* kpok)
{
    if (mask && !f->mask.empty()) std::memcpy(mask, f->mask.data(), f->mask.size());
    if (kpok && !f->kpOk.empty()) std::memcpy(kpok, f->kpOk.data(), f->kpOk.size());
}

// decision margins since the last reset: 7 classes (field of view [deg], in-frame [px], foci gate [px], 2-best ratio
// [Hamming], RANSAC support distance [px], rescue chi-square, dead-band [state units]): smallest non-zero |margin|, number of
// decisions, number of decisions with an operand exactly on the threshold (dead-band: exactly zero)
void orc_margins_reset()
{
    for (int i = 0; i < MG_COUNT; ++i) { g_margins.minAbs[i] = 1e300; g_margins.count[i] = 0; g_margins.zero[i] = 0; }
}
void orc_margins_get(double* minAbs, int64_t* count, int64_t* zero)
{
    for (int i = 0; i < MG_COUNT; ++i) { minAbs[i] = g_margins.minAbs[i]; count[i] = g_margins.count[i]; zero[i] = g_margins.zero[i]; }
}

void orc_eigen2x2(const double* A, double* ev, double* V) { eigen2x2(A, ev, V); }
void orc_gemm_acc(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int K, int N)
{
    gemmAcc(A, lda, B, ldb, C, ldc, M, K, N);
}
int orc_invert(const double* A, int32_t n, double* out) { return invertLU(A, n, out) ? 1 : 0; }
void orc_ellipse_params(const double* S, double* o)
{
    float a, b;
    double ang;
    ellipseParams(S, a, b, ang);
    o[0] = a;
    o[1] = b;
    o[2] = ang;
}
void orc_fill_ellipse(uint8_t* img, int32_t w, int32_t h, int32_t cx, int32_t cy, int32_t aw, int32_t ah, double ang)
{
    fillEllipse(img, w, h, cx, cy, aw, ah, ang);
}
void orc_draw_uncertainty_ellipse(uint8_t* img, int32_t w, int32_t h, double cx, double cy, const double* S, int32_t mx)
{
    drawUncertaintyEllipse(img, w, h, cx, cy, S, mx);
}
int orc_point_in_ellipse(float px, float py, float cx, float cy, int32_t aw, int32_t ah, double ang)
{
    return pointIsInsideEllipse(px, py, cx, cy, aw, ah, ang) ? 1 : 0;
}

// The reference's literal covariance update on caller-provided dense H (k x n): K = P H^T inv(H P H^T + sigma I),
// P = (I - K H) P.  Used by bench.py to time the update alone at large n.
void orc_update_dense(double* Pd, double* /*x_unused*/, int32_t n, const double* Hd, int32_t k, double sigma, double* K_out)
{
    Mat P(n, n), H(k, n), K;
    std::memcpy(P.d.data(), Pd, sizeof(double) * (size_t)n * n);
    std::memcpy(H.d.data(), Hd, sizeof(double) * (size_t)k * n);
    determineKalmanGain(H, P, sigma, K);
    Mat KH = mul(K, H);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) KH(i, j) = (i == j ? 1.0 : 0.0) - KH(i, j);
    P = mul(KH, P);
    std::memcpy(Pd, P.d.data(), sizeof(double) * (size_t)n * n);
    if (K_out) std::memcpy(K_out, K.d.data(), sizeof(double) * (size_t)n * k);
}

}  // extern "C"
