// Test-only harness: runs the product's host+device geometry (openekfmonoslam_b200/csrc/ekf_math.cuh)
// on the HOST so the CPU test suite can compare it with the oracle without a GPU.
#include "../../openekfmonoslam_b200/csrc/ekf_math.cuh"
#include "../../include/ekf_b200.h"

using namespace ekf;

static CamParams cam_of(const ekfb_params* p)
{
    CamParams c;
    c.fx = p->fx; c.fy = p->fy; c.k1 = p->k1; c.k2 = p->k2; c.cx = p->cx; c.cy = p->cy; c.dx = p->dx; c.dy = p->dy;
    c.fov_x = p->angular_vision_x; c.fov_y = p->angular_vision_y; c.width = p->pixels_x; c.height = p->pixels_y;
    return c;
}

extern "C" void hd_motion(const ekfb_params* p, const double* xc, double* F, double* GQG, double* xc_next)
{
    motion_jacobians(xc, p->linear_accel_sd, p->angular_accel_sd, F, GQG);
    for (int i = 0; i < 13; ++i) xc_next[i] = xc[i];
    motion_predict(xc_next);
}

// per feature: visibility, h, Hx (2x7), Hf (2x6), gate ellipse from a caller-provided S
extern "C" void hd_measure(const ekfb_params* p, const double* x, int N, const int* ftype, const int* foff,
                           unsigned char* vis, double* h, double* Hx, double* Hf)
{
    CamParams c = cam_of(p);
    double R[9], Rinv[9], Rt[9];
    quat_to_rot(x + 3, R);
    inv3(R, Rinv);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Rt[a * 3 + b] = R[b * 3 + a];
    for (int j = 0; j < N; ++j) {
        double y[6] = {0, 0, 0, 0, 0, 0};
        const int d = ftype[j] == kTypeInvDepth ? 6 : 3;
        for (int a = 0; a < d; ++a) y[a] = x[foff[j] + a];
        vis[j] = predict_pixel(c, x, Rt, Rinv, ftype[j], y, h + 2 * j) ? 1 : 0;
        if (vis[j]) measurement_jacobian(c, x, x + 3, Rinv, ftype[j], y, h + 2 * j, Hx + 14 * j, Hf + 12 * j);
    }
}

extern "C" void hd_gate(const double* S, double* out3)
{
    float a, b;
    double ang;
    gate_ellipse(S, &a, &b, &ang);
    out3[0] = a; out3[1] = b; out3[2] = ang;
}

extern "C" int hd_inside(float px, float py, float cx, float cy, int aw, int ah, double ang)
{
    return inside_gate(px, py, cx, cy, aw, ah, ang) ? 1 : 0;
}
