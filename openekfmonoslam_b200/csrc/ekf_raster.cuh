// ekf_raster.cuh -- the matching mask: union of filled uncertainty ellipses, rasterised with the
// pixel semantics of the reference's drawUncertaintyEllipse2D (modules/Gui/Draw.cpp:42-64), i.e.
// OpenCV's cv::ellipse(img, Point, Size, angleDeg, 0, 360, 255, -1): integer centre and axes,
// integer-degree rotation, a polygon on OpenCV's 1-degree float sine table, outline by fixed-point
// DDA lines and a convex scanline fill in 16.16 fixed point.  One warp rasterises one ellipse:
// lanes compute vertices and draw outline edges in parallel, lane 0 walks the two polygon chains to
// produce one span per image row, then all lanes fill the spans.
//
// Integer paths must be bit-exact with the CPU oracle, so vertex arithmetic uses explicit
// round-to-nearest mul/add (no FMA contraction).
#pragma once

#include <stdint.h>

namespace ekf {

constexpr int kXYShift = 16;
constexpr long long kXYOne = 1LL << kXYShift;
constexpr int kMaxPolyVerts = 80;  // 360/5 + 1 = 73 vertices at the finest angular step

// sin(k degrees), k = 0..90, as the 7-digit float literals OpenCV tabulates (symmetry gives the rest)
__device__ __constant__ float c_sin90[91] = {
    0.0000000f, 0.0174524f, 0.0348995f, 0.0523360f, 0.0697565f, 0.0871557f, 0.1045285f, 0.1218693f, 0.1391731f,
    0.1564345f, 0.1736482f, 0.1908090f, 0.2079117f, 0.2249511f, 0.2419219f, 0.2588190f, 0.2756374f, 0.2923717f,
    0.3090170f, 0.3255682f, 0.3420201f, 0.3583679f, 0.3746066f, 0.3907311f, 0.4067366f, 0.4226183f, 0.4383711f,
    0.4539905f, 0.4694716f, 0.4848096f, 0.5000000f, 0.5150381f, 0.5299193f, 0.5446390f, 0.5591929f, 0.5735764f,
    0.5877853f, 0.6018150f, 0.6156615f, 0.6293204f, 0.6427876f, 0.6560590f, 0.6691306f, 0.6819984f, 0.6946584f,
    0.7071068f, 0.7193398f, 0.7313537f, 0.7431448f, 0.7547096f, 0.7660444f, 0.7771460f, 0.7880108f, 0.7986355f,
    0.8090170f, 0.8191520f, 0.8290376f, 0.8386706f, 0.8480481f, 0.8571673f, 0.8660254f, 0.8746197f, 0.8829476f,
    0.8910065f, 0.8987940f, 0.9063078f, 0.9135455f, 0.9205049f, 0.9271839f, 0.9335804f, 0.9396926f, 0.9455186f,
    0.9510565f, 0.9563048f, 0.9612617f, 0.9659258f, 0.9702957f, 0.9743701f, 0.9781476f, 0.9816272f, 0.9848078f,
    0.9876883f, 0.9902681f, 0.9925462f, 0.9945219f, 0.9961947f, 0.9975641f, 0.9986295f, 0.9993908f, 0.9998477f,
    1.0000000f};

__device__ __forceinline__ float sin_deg(int k)  // k in [0, 450]
{
    if (k >= 360) k -= 360;
    if (k <= 90) return c_sin90[k];
    if (k <= 180) return c_sin90[180 - k];
    if (k <= 270) return -c_sin90[k - 180];
    return -c_sin90[360 - k];
}

struct RasterPt { long long x, y; };

__device__ __forceinline__ void put_px(uint8_t* img, int W, int H, long long x, long long y, uint8_t val)
{
    if (0 <= x && x < W && 0 <= y && y < H) img[(size_t)y * W + x] = val;
}

// fixed-point clip of a segment to [0, W<<16) x [0, H<<16) (OpenCV clipLine on scaled coordinates)
__device__ inline bool clip_segment(long long width, long long height, RasterPt& p1, RasterPt& p2)
{
    const long long right = width - 1, bottom = height - 1;
    long long &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)(__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1)));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)(__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1)));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)(__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1)));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)(__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1)));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// outline edge: fixed-point DDA (OpenCV Line2 for 1-byte pixels)
__device__ inline void draw_edge(uint8_t* img, int W, int H, RasterPt p1, RasterPt p2, uint8_t val)
{
    if (!clip_segment((long long)W << kXYShift, (long long)H << kXYShift, p1, p2)) return;
    long long dx = p2.x - p1.x, dy = p2.y - p1.y;
    const long long j = dx < 0 ? -1 : 0;
    const long long ax = (dx ^ j) - j;
    const long long i = dy < 0 ? -1 : 0;
    const long long ay = (dy ^ i) - i;
    long long x_step, y_step;
    int ecount;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { RasterPt t = p1; p1 = p2; p2 = t; }
        x_step = kXYOne;
        y_step = dy * (1 << kXYShift) / (ax | 1);
        ecount = (int)((p2.x - p1.x) >> kXYShift);
    } else {
        dx = (dx ^ i) - i;
        if (i) { RasterPt t = p1; p1 = p2; p2 = t; }
        x_step = dx * (1 << kXYShift) / (ay | 1);
        y_step = kXYOne;
        ecount = (int)((p2.y - p1.y) >> kXYShift);
    }
    p1.x += (kXYOne >> 1);
    p1.y += (kXYOne >> 1);
    put_px(img, W, H, (p2.x + (kXYOne >> 1)) >> kXYShift, (p2.y + (kXYOne >> 1)) >> kXYShift, val);
    if (ax > ay) {
        p1.x >>= kXYShift;
        while (ecount >= 0) {
            put_px(img, W, H, p1.x, p1.y >> kXYShift, val);
            p1.x++;
            p1.y += y_step;
            ecount--;
        }
    } else {
        p1.y >>= kXYShift;
        while (ecount >= 0) {
            put_px(img, W, H, p1.x >> kXYShift, p1.y, val);
            p1.x += x_step;
            p1.y++;
            ecount--;
        }
    }
}

// Per-warp scratch in shared memory
struct RasterScratch {
    RasterPt raw[kMaxPolyVerts];
    RasterPt v[kMaxPolyVerts];
    int nv;
    int y_first, y_last;  // rows [y_first, y_last] have spans
};

// Rasterise one ellipse with the whole warp.  `spans` is per-warp shared scratch of 2*H ints.
// `val` is the colour: 255 for the matching mask, 0 for the new-feature mask (E/DetectNewImageFeatures.cpp:115-121).
__device__ inline void raster_ellipse_warp(uint8_t* img, int W, int H, int cx, int cy, int aw, int ah, double angle_deg,
                                           RasterScratch* sc, int* spans, int lane, uint8_t val = 255)
{
    int angle = __double2int_rn(angle_deg);
    const long long ctrx = (long long)cx << kXYShift, ctry = (long long)cy << kXYShift;
    long long axw = (long long)aw << kXYShift, axh = (long long)ah << kXYShift;
    if (axw < 0) axw = -axw;
    if (axh < 0) axh = -axh;
    int delta = (int)(((axw > axh ? axw : axh) + (kXYOne >> 1)) >> kXYShift);
    delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;
    while (angle < 0) angle += 360;
    while (angle > 360) angle -= 360;
    const float alpha = sin_deg(450 - angle), beta = sin_deg(angle);
    const int nraw = (360 + delta - 1) / delta + 1;  // i = 0, delta, ... while i < 360 + delta
    for (int t = lane; t < nraw; t += 32) {
        int a = t * delta;
        if (a > 360) a = 360;
        const double x = __dmul_rn((double)axw, (double)sin_deg(450 - a));
        const double y = __dmul_rn((double)axh, (double)sin_deg(a));
        const double px = __dadd_rn(__dadd_rn((double)ctrx, __dmul_rn(x, (double)alpha)), -__dmul_rn(y, (double)beta));
        const double py = __dadd_rn(__dadd_rn((double)ctry, __dmul_rn(x, (double)beta)), __dmul_rn(y, (double)alpha));
        RasterPt p;
        p.x = (long long)__double2int_rn(__ddiv_rn(px, (double)kXYOne)) << kXYShift;
        p.y = (long long)__double2int_rn(__ddiv_rn(py, (double)kXYOne)) << kXYShift;
        p.x += __double2int_rn(__dadd_rn(px, -(double)p.x));
        p.y += __double2int_rn(__dadd_rn(py, -(double)p.y));
        sc->raw[t] = p;
    }
    __syncwarp();
    {   // drop consecutive duplicates: a vertex is compared with the last one KEPT, which (by induction) always equals its raw
        // predecessor, so "keep" is a local test and the new positions are a ballot prefix sum -- all lanes, three rounds
        int nv = 0;
        for (int base = 0; base < nraw; base += 32) {
            const int t = base + lane;
            bool keep = false;
            RasterPt p = {0, 0};
            if (t < nraw) {
                p = sc->raw[t];
                keep = (t == 0) || p.x != sc->raw[t - 1].x || p.y != sc->raw[t - 1].y;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) sc->v[nv + __popc(bal & ((1u << lane) - 1))] = p;
            nv += __popc(bal);
        }
        if (nv <= 1) {
            if (lane == 0) {
                sc->v[0].x = ctrx; sc->v[0].y = ctry;
                sc->v[1] = sc->v[0];
            }
            nv = 2;
        }
        if (lane == 0) sc->nv = nv;
    }
    __syncwarp();
    const int npts = sc->nv;
    // outline: edge e joins v[e-1] (v[npts-1] for e = 0) and v[e]
    for (int e = lane; e < npts; e += 32) draw_edge(img, W, H, sc->v[e == 0 ? npts - 1 : e - 1], sc->v[e], val);

    // bounding box and the topmost vertex (the FIRST index of the smallest y, as a serial scan finds it): warp reduction
    long long xmin = sc->v[0].x, xmax = xmin, ymin = sc->v[0].y, ymax = ymin;
    int imin = 0;
    for (int t = lane; t < npts; t += 32) {
        const RasterPt p = sc->v[t];
        if (p.y < ymin) { ymin = p.y; imin = t; }   // (t ascends within a lane: the first index of the lane's minimum is kept)
        if (p.y > ymax) ymax = p.y;
        if (p.x > xmax) xmax = p.x;
        if (p.x < xmin) xmin = p.x;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const long long oy = __shfl_xor_sync(0xffffffffu, ymin, o), oY = __shfl_xor_sync(0xffffffffu, ymax, o);
        const long long ox = __shfl_xor_sync(0xffffffffu, xmin, o), oX = __shfl_xor_sync(0xffffffffu, xmax, o);
        const int oi = __shfl_xor_sync(0xffffffffu, imin, o);
        if (oy < ymin || (oy == ymin && oi < imin)) { ymin = oy; imin = oi; }
        if (oY > ymax) ymax = oY;
        if (ox < xmin) xmin = ox;
        if (oX > xmax) xmax = oX;
    }
    if (lane == 0) {  // convex scanline walk -> spans
        sc->y_first = 0;
        sc->y_last = -1;
        const int dlt = 1 << kXYShift >> 1;
        xmin = (xmin + dlt) >> kXYShift;
        xmax = (xmax + dlt) >> kXYShift;
        ymin = (ymin + dlt) >> kXYShift;
        ymax = (ymax + dlt) >> kXYShift;
        if (!(npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= W || (int)ymin >= H)) {
            if (ymax > H - 1) ymax = H - 1;
            struct { int idx, di; long long x, dx; int ye; } edge[2];
            int y = (int)ymin;
            int edges = npts;
            edge[0].idx = edge[1].idx = imin;
            edge[0].ye = edge[1].ye = y;
            edge[0].di = 1;
            edge[1].di = npts - 1;
            edge[0].x = edge[1].x = -kXYOne;
            edge[0].dx = edge[1].dx = 0;
            sc->y_first = y < 0 ? 0 : y;
            do {
                for (int s = 0; s < 2; ++s) {
                    if (y >= edge[s].ye) {
                        int idx0 = edge[s].idx;
                        const int di = edge[s].di;
                        int idx = idx0 + di;
                        if (idx >= npts) idx -= npts;
                        for (; edges-- > 0;) {
                            const int ty = (int)((sc->v[idx].y + dlt) >> kXYShift);
                            if (ty > y) {
                                const long long xs = sc->v[idx0].x, xe = sc->v[idx].x;
                                edge[s].ye = ty;
                                edge[s].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                                edge[s].x = xs;
                                edge[s].idx = idx;
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
                    int l = 0, r = 1;
                    if (edge[0].x > edge[1].x) { l = 1; r = 0; }
                    int xx1 = (int)((edge[l].x + (kXYOne >> 1)) >> kXYShift);
                    int xx2 = (int)((edge[r].x + (kXYOne >> 1)) >> kXYShift);
                    if (xx2 >= 0 && xx1 < W) {
                        if (xx1 < 0) xx1 = 0;
                        if (xx2 >= W) xx2 = W - 1;
                    } else {
                        xx1 = 1; xx2 = 0;  // empty
                    }
                    spans[2 * y] = xx1;
                    spans[2 * y + 1] = xx2;
                    sc->y_last = y;
                }
                edge[0].x += edge[0].dx;
                edge[1].x += edge[1].dx;
            } while (++y <= (int)ymax);
        }
    }
    __syncwarp();
    const int y0 = sc->y_first, y1 = sc->y_last;
    for (int y = y0; y <= y1; ++y) {
        const int xx1 = spans[2 * y], xx2 = spans[2 * y + 1];
        for (int x = xx1 + lane; x <= xx2; x += 32) img[(size_t)y * W + x] = val;
    }
    __syncwarp();
}

}  // namespace ekf
