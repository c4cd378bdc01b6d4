// ekf_frontend.cuh -- the detector / descriptor half of the front end on the device, so that a frame goes
// image -> keypoints + descriptors -> ekfb_step without a host hop.
//
//   k_fast_score   FAST-9/16 corner test and corner score per pixel: OpenCV's FAST_t<16> / cornerScore<16>
//                  (modules/features2d/src/fast.cpp, fast_score.cpp), the "FAST" detector of the reference's
//                  FeatureDetectorFactory (kalmanFilter/modules/Configuration/.../FeatureDetectorFactory.cpp:51-165).
//                  One thread per pixel, the 16 ring bytes come from a shared-memory tile with a 3-pixel apron.
//   k_fast_rows    3x3 non-maximum suppression (strictly greater than all eight neighbours) + border rule, counted per row
//   k_fast_scan    exclusive scan of the row counts: keypoints come out in raster order, exactly like the CPU scan
//   k_fast_emit    one warp per row: ballot-ordered compaction of the row's keypoints -> xy (float32)
//   k_brief        one warp per keypoint, 8 point pairs per lane: bit = [box5(p + a) < box5(p + b)] -> 32 bytes
// The detector is pinned against cv2.FastFeatureDetector through oracle/fast_oracle.py (identical keypoints, order and
// scores); the descriptor's point pairs are this repository's own (ekf_brief_pattern.h) because OpenCV 2.4's BRIEF table
// is not available here.  Bound: HBM/L2 streaming of the image (1 byte read, 1 byte written per pixel) -- a 640x480 frame
// is 0.3 MB, so launch latency dominates.
#pragma once

#include "ekf_brief_pattern.h"
#include "ekf_kernels.cuh"

namespace ekf {

constexpr int kFastBorder = 18;   // keypoints closer than this to the border are dropped (descriptor support 13 + 2, rounded up)

__device__ __constant__ signed char c_brief[256 * 4];

// interleaved 8-bit BGR / BGRA -> grey like cv::cvtColor(COLOR_BGR2GRAY): (B 1868 + G 9617 + R 4899 + 2^13) >> 14
// (what OpenCV's detectors do with a colour frame: E/Matching.cpp:206 passes the BGR image).  grid (ceil(W/32), ceil(H/8)), block (32, 8)
__global__ void k_bgr_to_gray(const uint8_t* src, int spitch, int channels, uint8_t* dst, int dpitch, int W, int H)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t* p = src + (size_t)y * spitch + (size_t)x * channels;
    dst[(size_t)y * dpitch + x] = (uint8_t)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + 8192) >> 14);
}

// grid (ceil(W/32), ceil(H/8)), block (32, 8)
__global__ void __launch_bounds__(256) k_fast_score(const uint8_t* img, int pitch, int W, int H, int threshold, uint8_t* score)
{
    __shared__ uint8_t tile[14][40];   // 8 + 6 rows, 32 + 6 columns (padded)
    const int x0 = blockIdx.x * 32 - 3, y0 = blockIdx.y * 8 - 3;
    for (int e = threadIdx.y * 32 + threadIdx.x; e < 14 * 38; e += 256) {
        const int ty = e / 38, tx = e % 38, gx = x0 + tx, gy = y0 + ty;
        tile[ty][tx] = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[(size_t)gy * pitch + gx] : 0;
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    int out = 0;
    if (x >= 3 && x < W - 3 && y >= 3 && y < H - 3) {
        const int cx = threadIdx.x + 3, cy = threadIdx.y + 3;
        const int v = tile[cy][cx];
        // OpenCV's ring order (dx, dy), patternSize 16
        const int rx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        const int ry[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
        int d[25];
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = v - tile[cy + ry[k]][cx + rx[k]];
#pragma unroll
        for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
        // corner: 9 contiguous ring pixels all brighter than v + t or all darker than v - t
        unsigned br = 0, dk = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            br |= (unsigned)(d[k] < -threshold) << k;
            dk |= (unsigned)(d[k] > threshold) << k;
        }
        br |= br << 16;
        dk |= dk << 16;
        unsigned rb = br, rd = dk;
#pragma unroll
        for (int s = 1; s < 9; ++s) { rb &= br >> s; rd &= dk >> s; }
        if (((rb | rd) & 0xffffu) != 0) {
            // cornerScore<16>: the largest threshold for which the pixel is still a corner
            int a0 = threshold;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                int a = min(min(d[k + 1], d[k + 2]), d[k + 3]);
                if (a <= a0) continue;
                a = min(min(min(a, d[k + 4]), min(d[k + 5], d[k + 6])), min(d[k + 7], d[k + 8]));
                a0 = max(a0, min(a, d[k]));
                a0 = max(a0, min(a, d[k + 9]));
            }
            int b0 = -a0;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                int b = max(max(d[k + 1], d[k + 2]), d[k + 3]);
                if (b >= b0) continue;
                b = max(max(max(b, d[k + 4]), max(d[k + 5], d[k + 6])), max(d[k + 7], d[k + 8]));
                b0 = min(b0, max(b, d[k]));
                b0 = min(b0, max(b, d[k + 9]));
            }
            out = -b0 - 1;   // <= 255 - 1
        }
    }
    score[(size_t)y * pitch + x] = (uint8_t)out;
}

__device__ __forceinline__ bool fast_is_keypoint(const uint8_t* score, int pitch, int W, int H, int x, int y)
{
    if (x < kFastBorder || x >= W - kFastBorder || y < kFastBorder || y >= H - kFastBorder) return false;
    const uint8_t* p = score + (size_t)y * pitch + x;
    const int s = p[0];
    if (s == 0) return false;
    return s > p[-1] && s > p[1] && s > p[-pitch - 1] && s > p[-pitch] && s > p[-pitch + 1] && s > p[pitch - 1] && s > p[pitch] &&
           s > p[pitch + 1];
}

// one warp per image row: count (emit = 0) or write (emit = 1) the row's keypoints in x order
__global__ void __launch_bounds__(256) k_fast_rows(const uint8_t* score, int pitch, int W, int H, int* rowCount, const int* rowOff,
                                                   float* xy, int cap, int emit)
{
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= H) return;
    int base = emit ? rowOff[y] : 0, cnt = 0;
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const bool k = x < W && fast_is_keypoint(score, pitch, W, H, x, y);
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (emit && k) {
            const int at = base + cnt + __popc(m & ((1u << lane) - 1u));
            if (at < cap) { xy[2 * at] = (float)x; xy[2 * at + 1] = (float)y; }
        }
        cnt += __popc(m);
    }
    if (!emit && lane == 0) rowCount[y] = cnt;
}

// exclusive scan of H row counts by one block; total (capped) -> *count
__global__ void __launch_bounds__(1024) k_fast_scan(const int* rowCount, int* rowOff, int H, int cap, int* count)
{
    __shared__ int part[1024];
    const int per = (H + 1023) / 1024, t = threadIdx.x, b = t * per;
    int s = 0;
    for (int i = b; i < min(b + per, H); ++i) s += rowCount[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int a = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += a;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int i = b; i < min(b + per, H); ++i) { rowOff[i] = run; run += rowCount[i]; }
    if (t == 1023) *count = min(part[1023], cap);
}

// one warp per keypoint; lane l produces byte l of the descriptor
__global__ void __launch_bounds__(256) k_brief(const uint8_t* img, int pitch, int W, int H, const float* xy, const int* count, uint8_t* desc)
{
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= *count) return;
    const int x = (int)xy[2 * n], y = (int)xy[2 * n + 1];
    unsigned byte = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const signed char* p = c_brief + (lane * 8 + bit) * 4;
        int sa = 0, sb = 0;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                sa += img[(size_t)(y + p[1] + dy) * pitch + x + p[0] + dx];
                sb += img[(size_t)(y + p[3] + dy) * pitch + x + p[2] + dx];
            }
        byte |= (unsigned)(sa < sb) << bit;
    }
    desc[(size_t)n * 32 + lane] = (uint8_t)byte;
}

}  // namespace ekf
