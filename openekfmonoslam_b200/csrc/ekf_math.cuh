// ekf_math.cuh -- scalar geometry of the EKF hot path, usable from host and device.
//
// These are the per-feature / per-camera closed forms the kernels evaluate.  Each block cites the
// reference code whose result it reproduces ("E/" = kalmanFilter/modules/1PointRansacEKF/,
// "C/" = kalmanFilter/modules/Core/).  Everything is FP64 (the reference's x87 long-double
// sub-expressions differ at the 1e-19 level, far below the 1e-9 parity tolerance).
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define EKF_HD __host__ __device__ __forceinline__
#else
#define EKF_HD inline
#endif

namespace ekf {

constexpr double kEpsilon = 2.22e-16;       // C/EKFMath.h:37 EPSILON
constexpr double kDelta = 1.0e-12;          // C/EKFMath.h:38 DELTA (dead-band of E/Update.cpp:133-204)
constexpr double kPiTrunc = 3.14159265;     // C/EKFMath.h:39 PI (truncated on purpose)
constexpr double kChi2_95_2 = 5.9915;       // C/EKFMath.h:40
constexpr int kTypeXYZ = 1;                 // E/MapFeature.h:39-44
constexpr int kTypeInvDepth = 2;

struct CamParams {  // subset of ekfb_params used by the geometry
    double fx, fy, k1, k2, cx, cy, dx, dy;
    double fov_x, fov_y;
    double width, height;
};

// ---- small dense helpers -------------------------------------------------------------------
EKF_HD void quat_to_rot(const double* q, double* R)  // C/EKFMath.cpp:121-141
{
    const double r = q[0], x = q[1], y = q[2], z = q[3];
    const double r2 = r * r, x2 = x * x, y2 = y * y, z2 = z * z;
    R[0] = r2 + x2 - y2 - z2;  R[1] = 2 * (x * y - r * z);    R[2] = 2 * (z * x + r * y);
    R[3] = 2 * (x * y + r * z);  R[4] = r2 - x2 + y2 - z2;    R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (z * x - r * y);  R[7] = 2 * (y * z + r * x);  R[8] = r2 - x2 - y2 + z2;
}

// cv::Mat::inv() on 3x3 / 2x2: determinant + adjugate in OpenCV's operation order.
EKF_HD bool inv3(const double* m, double* o)
{
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
               m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0.) {
        for (int i = 0; i < 9; ++i) o[i] = 0.;
        return false;
    }
    d = 1. / d;
    double t[9];
    t[0] = (m[4] * m[8] - m[5] * m[7]) * d;  t[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    t[2] = (m[1] * m[5] - m[2] * m[4]) * d;  t[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    t[4] = (m[0] * m[8] - m[2] * m[6]) * d;  t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    t[6] = (m[3] * m[7] - m[4] * m[6]) * d;  t[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    for (int i = 0; i < 9; ++i) o[i] = t[i];
    return true;
}

EKF_HD bool inv2(const double* m, double* o)
{
    double d = m[0] * m[3] - m[1] * m[2];
    if (d == 0.) {
        o[0] = o[1] = o[2] = o[3] = 0.;
        return false;
    }
    d = 1. / d;
    const double a = m[3] * d, b = -m[1] * d, c = -m[2] * d, e = m[0] * d;
    o[0] = a; o[1] = b; o[2] = c; o[3] = e;
    return true;
}

EKF_HD void mat3_vec(const double* M, const double* v, double* o)
{
    const double x = v[0], y = v[1], z = v[2];
    o[0] = M[0] * x + M[1] * y + M[2] * z;
    o[1] = M[3] * x + M[4] * y + M[5] * z;
    o[2] = M[6] * x + M[7] * y + M[8] * z;
}

// rotation-vector -> quaternion, C/EKFMath.cpp:62-81
EKF_HD void rotvec_to_quat(const double* w, double* q)
{
    const double nrm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (nrm < kEpsilon) {
        q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0;
    } else {
        const double half = nrm / 2;
        const double s = sin(half);
        q[0] = cos(half);
        q[1] = s * w[0] / nrm;
        q[2] = s * w[1] / nrm;
        q[3] = s * w[2] / nrm;
    }
}

EKF_HD void quat_mul(const double* a, const double* b, double* q)  // C/EKFMath.cpp:85-101
{
    q[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    q[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    q[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    q[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

// ---- motion model --------------------------------------------------------------------------
// F (13x13, row-major) and G*Q*G^T (13x13) of the constant-velocity model evaluated at the
// PRE-prediction camera state xc[13], dt = 1 (E/StateAndCovariancePrediction.cpp:154-224,246).
// Includes the reference's omega ~ 0 branch: F[w,w] diagonal zeroed and G[q,alpha] left zero.
EKF_HD void motion_jacobians(const double* xc, double sd_lin, double sd_ang, double* F, double* GQG)
{
    const double dt = 1.0;
    for (int i = 0; i < 169; ++i) { F[i] = 0.; GQG[i] = 0.; }
    for (int i = 0; i < 13; ++i) F[i * 13 + i] = 1.;
    for (int i = 0; i < 3; ++i) F[i * 13 + 7 + i] = dt;
    const double* q = xc + 3;
    const double* om = xc + 10;
    double w[3] = {om[0] * dt, om[1] * dt, om[2] * dt};
    double qr[4];
    rotvec_to_quat(w, qr);
    {   // d(q x qr)/dq : right-multiplication matrix of qr (:71-92)
        const double a = qr[0], b = qr[1], c = qr[2], d = qr[3];
        const double m[16] = {a, -b, -c, -d, b, a, d, -c, c, -d, a, b, d, c, -b, a};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) F[(3 + i) * 13 + 3 + j] = m[i * 4 + j];
    }
    double Gq[12];  // the 4x3 block shared by F[q,w] and G[q,alpha]
    for (int i = 0; i < 12; ++i) Gq[i] = 0.;
    const bool omega_zero = fabs(om[0]) < kEpsilon && fabs(om[1]) < kEpsilon && fabs(om[2]) < kEpsilon;
    if (omega_zero) {
        for (int i = 0; i < 3; ++i) F[(10 + i) * 13 + 10 + i] = 0.;
    } else {
        const double n = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
        const double sh = sin(n * dt / 2.0), ch = cos(n * dt / 2.0);
        double D[12];  // d quat(w dt) / d w   (:100-146)
        for (int a = 0; a < 3; ++a) D[a] = (-dt / 2.0) * (om[a] / n) * sh;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double v;
                if (a == b)
                    v = (dt / 2.0) * om[a] * om[a] / (n * n) * ch + (1.0 / n) * (1.0 - om[a] * om[a] / (n * n)) * sh;
                else
                    v = (om[a] * om[b] / (n * n)) * ((dt / 2.0) * ch - (1.0 / n) * sh);
                D[(1 + a) * 3 + b] = v;
            }
        // left-multiplication matrix of q (C/EKFMath.cpp:105-116)
        const double a = q[0], b = q[1], c = q[2], d = q[3];
        const double L[16] = {a, -b, -c, -d, b, a, -d, c, c, d, a, -b, d, -c, b, a};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0.;
                for (int k = 0; k < 4; ++k) s += L[i * 4 + k] * D[k * 3 + j];
                Gq[i * 3 + j] = s;
                F[(3 + i) * 13 + 10 + j] = s;
            }
    }
    // G (13x6): G[v,a] = I, G[w,alpha] = I, G[r,a] = dt I, G[q,alpha] = Gq.  Q = diag(ql I3, qa I3).
    const double ql = sd_lin * sd_lin * dt * dt, qa = sd_ang * sd_ang * dt * dt;
    double G[78];
    for (int i = 0; i < 78; ++i) G[i] = 0.;
    for (int i = 0; i < 3; ++i) {
        G[(7 + i) * 6 + i] = 1.0;
        G[(10 + i) * 6 + 3 + i] = 1.0;
        G[i * 6 + i] = dt;
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) G[(3 + i) * 6 + 3 + j] = Gq[i * 3 + j];
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) {
            double s = 0.;
            for (int k = 0; k < 6; ++k) s += (G[i * 6 + k] * (k < 3 ? ql : qa)) * G[j * 6 + k];
            GQG[i * 13 + j] = s;
        }
}

// x <- f(x): r += v dt; q <- q x quat(w dt) (not renormalised), E/StateAndCovariancePrediction.cpp:43-65
EKF_HD void motion_predict(double* xc)
{
    const double dt = 1.0;
    for (int i = 0; i < 3; ++i) xc[i] += xc[7 + i] * dt;
    double w[3] = {xc[10] * dt, xc[11] * dt, xc[12] * dt};
    double qr[4], qn[4];
    rotvec_to_quat(w, qr);
    quat_mul(xc + 3, qr, qn);
    for (int i = 0; i < 4; ++i) xc[3 + i] = qn[i];
}

// ---- measurement model -----------------------------------------------------------------------
EKF_HD void direction(double theta, double phi, double* m)  // C/EKFMath.cpp:145-152
{
    const double cp = cos(phi);
    m[0] = cp * sin(theta);
    m[1] = -sin(phi);
    m[2] = cp * cos(theta);
}

// ideal pixel -> distorted pixel, 10 Newton steps (E/MeasurementPrediction.cpp:47-83)
EKF_HD void distort(const CamParams& c, const double* u, double* hd)
{
    const double px = u[0] - c.cx, py = u[1] - c.cy;
    const double mx = c.dx * px, my = c.dy * py;
    const double d2 = mx * mx + my * my;
    const double ru = sqrt(d2);
    double rd = ru / (1.0 + c.k1 * d2 + c.k2 * d2 * d2);
    for (int it = 0; it < 10; ++it) {
        const double r2 = rd * rd, r3 = r2 * rd, r4 = r2 * r2, r5 = r4 * rd;
        const double f = rd + c.k1 * r3 + c.k2 * r5 - ru;
        const double fp = 1 + 3 * c.k1 * r2 + 5 * c.k2 * r4;
        rd = rd - f / fp;
    }
    const double r2 = rd * rd, r4 = r2 * r2;
    const double d = 1.0 + c.k1 * r2 + c.k2 * r4;
    hd[0] = c.cx + px / d;
    hd[1] = c.cy + py / d;
}

// the un-rotated camera-frame ray of a feature: rho (y - r) + m(theta, phi)  or  y - r
EKF_HD void feature_ray(int type, const double* y, const double* r, double* a)
{
    if (type == kTypeInvDepth) {
        double m[3];
        direction(y[3], y[4], m);
        const double rho = y[5];
        a[0] = rho * (y[0] - r[0]) + m[0];
        a[1] = rho * (y[1] - r[1]) + m[1];
        a[2] = rho * (y[2] - r[2]) + m[2];
    } else {
        a[0] = y[0] - r[0];
        a[1] = y[1] - r[1];
        a[2] = y[2] - r[2];
    }
}

// h(x) for one feature; returns false if the feature is not predicted inside the frame.
// Inverse-depth features rotate with R^T, XYZ features with inv(R) -- they differ when q is not
// unit norm, as inside RANSAC (E/MeasurementPrediction.cpp:203-265, tests :162-181).
EKF_HD bool predict_pixel(const CamParams& c, const double* r, const double* Rt, const double* Rinv, int type,
                          const double* y, double* h)
{
    double a[3], p[3];
    feature_ray(type, y, r, a);
    mat3_vec(type == kTypeInvDepth ? Rt : Rinv, a, p);
    const double ax = atan2(p[0], p[2]) * 180.0 / kPiTrunc;
    const double ay = atan2(p[1], p[2]) * 180.0 / kPiTrunc;
    if (!(-c.fov_x < ax && ax < c.fov_x && -c.fov_y < ay && ay < c.fov_y)) return false;
    double u[2] = {c.cx + (c.fx * p[0] / p[2]), c.cy + (c.fy * p[1] / p[2])};
    distort(c, u, h);
    return (h[0] > 0 && h[0] < c.width && h[1] > 0 && h[1] < c.height);
}

// d(R(q) a)/dq, 3x4 row-major (E/CommonFunctions.cpp:87-145)
EKF_HD void drot_dq(const double* q, const double* a, double* J)
{
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double ax = a[0], ay = a[1], az = a[2];
    J[0] = 2 * w * ax - 2 * z * ay + 2 * y * az;   J[4] = 2 * z * ax + 2 * w * ay - 2 * x * az;
    J[8] = -2 * y * ax + 2 * x * ay + 2 * w * az;
    J[1] = 2 * x * ax + 2 * y * ay + 2 * z * az;   J[5] = 2 * y * ax - 2 * x * ay - 2 * w * az;
    J[9] = 2 * z * ax + 2 * w * ay - 2 * x * az;
    J[2] = -2 * y * ax + 2 * x * ay + 2 * w * az;  J[6] = 2 * x * ax + 2 * y * ay + 2 * z * az;
    J[10] = -2 * w * ax + 2 * z * ay - 2 * y * az;
    J[3] = -2 * z * ax - 2 * w * ay + 2 * x * az;  J[7] = 2 * w * ax - 2 * z * ay + 2 * y * az;
    J[11] = 2 * x * ax + 2 * y * ay + 2 * z * az;
}

// Measurement Jacobian of one predicted feature at distorted pixel h:
//   Hx (2x7: d h / d r, d h / d q; the v, w columns are zero) and Hf (2x6, first `dim` columns used).
// Reproduces E/MeasurementPrediction.cpp:273-589 including two quirks that change values:
//   * d c / d r uses a 3x3 whose entry (0,1) is 0 and (0,2) is -rho^2 Rinv[2] (inverse depth) or
//     -Rinv[2] (XYZ) because :371-373,:392-394 skip index 1 and touch index 2 twice;
//   * the rho column of d c / d y holds the UN-rotated (y_xyz - r) (:571 uses pointInCameraAxis).
EKF_HD void measurement_jacobian(const CamParams& c, const double* r, const double* q, const double* Rinv, int type,
                                 const double* y, const double* h, double* Hx, double* Hf)
{
    const bool idp = (type == kTypeInvDepth);
    double a[3], p[3];
    feature_ray(type, y, r, a);
    mat3_vec(Rinv, a, p);  // the Jacobians use inv(R) for both feature types
    // pinhole Jacobian (2x3), :273-297
    const double fpj[6] = {c.fx / p[2], 0., -p[0] * c.fx / (p[2] * p[2]), 0., c.fy / p[2], -p[1] * c.fy / (p[2] * p[2])};
    // d undistort / d h_d (2x2) and its inverse, :308-362
    const double px = h[0] - c.cx, py = h[1] - c.cy;
    const double mx = c.dx * px, my = c.dy * py;
    const double d2 = mx * mx + my * my;
    const double rad = 1 + c.k1 * d2 + c.k2 * d2 * d2;
    const double g = c.k1 + 2 * c.k2 * d2;
    const double dj[4] = {rad + px * g * (2 * px * c.dx * c.dx), px * g * (2 * py * c.dy * c.dy),
                          py * g * (2 * px * c.dx * c.dx), rad + py * g * (2 * py * c.dy * c.dy)};
    double idj[4];
    inv2(dj, idj);
    double pj[6];
    for (int j = 0; j < 3; ++j) {
        pj[j] = idj[0] * fpj[j] + idj[1] * fpj[3 + j];
        pj[3 + j] = idj[2] * fpj[j] + idj[3] * fpj[3 + j];
    }
    // d c / d r with the reference's index slip
    const double rho = idp ? y[5] : 1.0;
    double dr[9];
    dr[0] = -Rinv[0] * rho;
    dr[1] = 0.;
    dr[2] = idp ? (-Rinv[2] * rho) * rho : -Rinv[2];
    for (int i = 3; i < 9; ++i) dr[i] = -Rinv[i] * rho;
    // d c / d q = dR(conj q) a * diag(1,-1,-1,-1)
    const double qc[4] = {q[0], -q[1], -q[2], -q[3]};
    double dq[12];
    drot_dq(qc, a, dq);
    for (int i = 0; i < 3; ++i) {
        dq[i * 4 + 1] = -dq[i * 4 + 1];
        dq[i * 4 + 2] = -dq[i * 4 + 2];
        dq[i * 4 + 3] = -dq[i * 4 + 3];
    }
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 3; ++j) {
            double s = 0.;
            for (int k = 0; k < 3; ++k) s += pj[i * 3 + k] * dr[k * 3 + j];
            Hx[i * 7 + j] = s;
        }
        for (int j = 0; j < 4; ++j) {
            double s = 0.;
            for (int k = 0; k < 3; ++k) s += pj[i * 3 + k] * dq[k * 4 + j];
            Hx[i * 7 + 3 + j] = s;
        }
    }
    for (int i = 0; i < 12; ++i) Hf[i] = 0.;
    if (!idp) {  // :510-523
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0.;
                for (int k = 0; k < 3; ++k) s += pj[i * 3 + k] * Rinv[k * 3 + j];
                Hf[i * 6 + j] = s;
            }
    } else {  // :530-589
        const double th = y[3], ph = y[4];
        const double cph = cos(ph), cth = cos(th), sth = sin(th), sph = sin(ph);
        const double dmt[3] = {cph * cth, 0., -cph * sth};
        const double dmp[3] = {-sph * sth, -cph, -sph * cth};
        double rt[3], rp[3];
        mat3_vec(Rinv, dmt, rt);
        mat3_vec(Rinv, dmp, rp);
        double D[18];
        for (int i = 0; i < 3; ++i) {
            D[i * 6 + 0] = rho * Rinv[3 * i + 0];
            D[i * 6 + 1] = rho * Rinv[3 * i + 1];
            D[i * 6 + 2] = rho * Rinv[3 * i + 2];
            D[i * 6 + 3] = rt[i];
            D[i * 6 + 4] = rp[i];
            D[i * 6 + 5] = y[i] - r[i];
        }
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0.;
                for (int k = 0; k < 3; ++k) s += pj[i * 3 + k] * D[k * 6 + j];
                Hf[i * 6 + j] = s;
            }
    }
}

// ---- gate ellipse ------------------------------------------------------------------------------
// cv::eigen on a symmetric 2x2 (one Jacobi rotation, eigenvalues descending, eigenvectors as rows),
// then C/EKFMath.cpp:271-298: axes = float(2 sqrt(lambda * 5.9915)), angle = atan(V[1][0] / V[0][0]).
EKF_HD void gate_ellipse(const double* S, float* ax_w, float* ax_h, double* angle)
{
    double w0 = S[0], w1 = S[3];
    double v00 = 1, v01 = 0, v10 = 0, v11 = 1;
    const double p = S[1];
    if (fabs(p) > 2.220446049250313e-16) {
        const double yv = (w1 - w0) * 0.5;
        double t = fabs(yv) + hypot(p, yv);
        double s = hypot(p, t);
        const double cs = t / s;
        s = p / s;
        t = (p / t) * p;
        if (yv < 0) { s = -s; t = -t; }
        w0 -= t;
        w1 += t;
        v00 = cs; v01 = -s; v10 = s; v11 = cs;
    }
    if (w0 < w1) {
        double tmp = w0; w0 = w1; w1 = tmp;
        tmp = v00; v00 = v10; v10 = tmp;
        tmp = v01; v01 = v11; v11 = tmp;
    }
    (void)v01; (void)v11;
    *ax_w = (float)(2.0 * sqrt(w0 * kChi2_95_2));
    *ax_h = (float)(2.0 * sqrt(w1 * kChi2_95_2));
    *angle = atan(v10 / v00);
}

// foci test of C/EKFMath.cpp:302-351 (integer axes, float centre and point)
EKF_HD bool inside_gate(float px, float py, float cx, float cy, int aw, int ah, double angle)
{
    const double major = aw > ah ? aw : ah;
    const double minor = aw > ah ? ah : aw;
    const double f = sqrt(major * major - minor * minor);
    double f1x, f1y, f2x, f2y;
    const double ca = cos(angle), sa = sin(angle);
    if (ah < aw) {
        f1x = f * ca + cx;    f1y = f * sa + cy;
        f2x = -f * ca + cx;   f2y = -f * sa + cy;
    } else {
        f1x = f * (-sa) + cx;   f1y = f * ca + cy;
        f2x = -f * (-sa) + cx;  f2y = -f * ca + cy;
    }
    const double a1x = px - f1x, a1y = py - f1y, a2x = px - f2x, a2y = py - f2y;
    const double sum = sqrt(a1x * a1x + a1y * a1y) + sqrt(a2x * a2x + a2y * a2y);
    return sum <= 2 * major;
}

// dead-band of E/Update.cpp:133-134,153-199: keep v only if |v| > 1e-12
EKF_HD double deadband(double v) { return fabs(v) > kDelta ? v : 0.0; }

}  // namespace ekf
