// ekf_ncc.cuh -- active search by normalised cross-correlation (the north star's matching path; the reference has no
// counterpart: its matcher is the descriptor path of ekf_kernels.cuh, see SURVEY.md 0.3 / 8b `ekfb_match_ncc`).
//
// Specification (restated on the CPU in oracle/ncc_oracle.py, which the GPU tests compare against bit for bit):
//   pyramid   L0 = the frame (8-bit grey), L(l+1)(y,x) = (a + b + c + d + 2) >> 2 over the 2x2 block, floor(W/2) x floor(H/2)
//   template  per feature and level an 11 x 11 8-bit patch T_l, supplied by the caller (ekfb_ncc_set_templates) or CAPTURED ON THE
//             DEVICE when the feature is created (ekfb_add_features with an image set: k_ncc_capture cuts the patch centred on
//             ((int)u >> l, (int)v >> l) of every level, zero where it leaves the level) together with the ANCHOR of the
//             feature: camera position r0, orientation q0 and the pixel (u0, v0) at that moment (10 doubles)
//   warp      (features with an anchor; EKFB_OPT_NCC_WARP) the patch is predicted for the current camera before it is compared:
//             the surface around the point X is taken as the plane through X facing the anchor camera, n = (X - r0)/|X - r0|;
//             pinhole model (fx, fy, cx, cy; lens distortion is ignored by the warp); A = d(u1, v1)/d(u0, v0) by unit
//             differences: anchor pixels (u0 + 1, v0), (u0, v0 + 1) -> rays R(q0) ((u - cx)/fx, (v - cy)/fy, 1) -> plane ->
//             current camera R(q1)^T (X' - r1) -> pixel, minus the same for (u0, v0).  Warped template
//             T'_l(tx, ty) = bilinear sample of T_l at (5, 5) + A^-1 (tx - 5, ty - 5), coordinates clamped to [0, 10], rounded
//             to the nearest byte.  A is replaced by the identity (raw template) when max|A - I| < 0.05 or when
//             det A is outside [0.25, 4] or not finite.  The same A serves all three levels.
//   search    for every predicted feature: centre c0 = (int)(float)h, integer gate axes (aw, ah) and angle as in the
//             descriptor matcher (E/Matching.cpp:227-239); start level l = smallest level with ceil(max(aw,ah) / 2^l) <= 12
//             (at most 2), window half-size R = min(12, that value); candidates = displacements in [-R, R]^2 around c0 >> l
//             whose patch lies inside the level and whose level-0 position passes the foci gate (C/EKFMath.cpp:302-351);
//             score = (121 S_tw - S_t S_w) / sqrt((121 S_tt - S_t^2)(121 S_ww - S_w^2)) from exact integer sums (flat patches
//             are skipped); best = highest score, ties to the smaller dy, then dx; then one 3 x 3 refinement around twice
//             the best position per finer level.  Match iff the level-0 score >= ncc_min: z = that pixel.
//
// One CTA per feature.  The search window -- (2R + 11)^2 <= 35 x 35 bytes -- is staged in shared memory by the TMA
// engine: one bulk asynchronous copy (cp.async.bulk.shared::cluster.global, 64 bytes from a 16-byte aligned source) per
// window row, all completing on one mbarrier (complete_tx); rows and columns outside the image are clamped away (their
// bytes are never read: a candidate's patch must lie inside the level).  Where the level is at least one box large the window
// comes instead by ONE tensor-map load (cp.async.bulk.tensor.2d, UTMALDG.2D; box 64 x 36 bytes at the 16-byte aligned origin,
// out-of-bounds zero fill) -- round 1 believed that form faulted on this pool; the fault was the probe's misaligned
// innermost coordinate (profiles/r02_tma_probe.txt).
// Warp w takes displacement rows w, w + 4, ..., lane = dx, so the 32 lanes of a warp read consecutive window bytes
// (conflict-free) and the same template byte (broadcast); template sums and the arg-max are warp-shuffle reductions.
// Bound: shared-memory bandwidth (2 x 121 byte reads per candidate) / latency; HBM traffic is the window bytes only.
#pragma once

#include <cuda.h>

#include "ekf_kernels.cuh"

namespace ekf {

constexpr int kNccP = 11, kNccPP = 121, kNccR = 12, kNccBoxW = 64, kNccBoxH = 36, kNccLevels = 3;

struct NccView {
    int W[kNccLevels], H[kNccLevels], pitch[kNccLevels];
    const uint8_t* img[kNccLevels];   // this filter's pyramid levels, 16-byte aligned rows
    const uint8_t* tmpl;   // [Nmax][3][128] (121 used)
    int tmaLevel[kNccLevels];   // 1: this level's window comes by ONE tensor-map TMA load (NccMaps), 0: by per-row bulk copies
    const double* anchor;  // [Nmax][10]: r0[3], q0[4], (u0, v0), valid (1.0) -- nullptr or valid == 0: no warp
    int warp;              // EKFB_OPT_NCC_WARP
    double* score;         // [Nmax] level-0 score of the last search (-2: none)
    int* level;            // [Nmax] start level
    double ncc_min;
};

// tensor maps of this filter's pyramid levels: UINT8, dims (W_l, H_l), row stride pitch_l, box kNccBoxW x kNccBoxH, no swizzle,
// out-of-bounds elements read as zero.  The innermost coordinate of a load must be a multiple of 16 BYTES (a misaligned one
// raises "illegal instruction" at the UTMALDG: round 1's probe used x = 100, tools/tma_probe.cu, profiles/r02_tma_probe.txt);
// the window origin is therefore rounded down to x0a = x0 & ~15, as for the bulk-copy form.
struct NccMaps { CUtensorMap m[kNccLevels]; };

// L(l+1) from L(l).  grid (ceil(Wo/32), ceil(Ho/8)), block (32, 8)
__global__ void k_pyr_down(const uint8_t* src, int sp, uint8_t* dst, int dp, int Wo, int Ho)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= Wo || y >= Ho) return;
    const uint8_t* s = src + (size_t)(2 * y) * sp + 2 * x;
    dst[(size_t)y * dp + x] = (uint8_t)((s[0] + s[1] + s[sp] + s[sp + 1] + 2) >> 2);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int phase)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// TMA bulk copy global -> shared (bytes a multiple of 16, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_load(void* smem, const void* gmem, int bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kNccAnchor = 10;

// Template + anchor capture for `count` new features first .. first + count - 1 of filter f at pixels uv (doubles, 2 each):
// one CTA (128 threads) per feature.  tmpl / anchor point at the filter's slices.
__global__ void __launch_bounds__(128) k_ncc_capture(DevView v, NccView nv, int f, int first, int count, const double* uv, uint8_t* tmpl,
                                                     double* anchor)
{
    const int i = blockIdx.x;
    if (i >= count) return;
    const int j = first + i;
    const int x = (int)uv[2 * i], y = (int)uv[2 * i + 1];
    for (int l = 0; l < kNccLevels; ++l) {
        const int cx = x >> l, cy = y >> l;
        uint8_t* T = tmpl + ((size_t)j * kNccLevels + l) * 128;
        for (int e = threadIdx.x; e < 128; e += blockDim.x) {
            uint8_t val = 0;
            if (e < kNccPP) {
                const int px = cx - 5 + e % kNccP, py = cy - 5 + e / kNccP;
                if (px >= 0 && px < nv.W[l] && py >= 0 && py < nv.H[l]) val = nv.img[l][(size_t)py * nv.pitch[l] + px];
            }
            T[e] = val;
        }
    }
    if (threadIdx.x < kNccAnchor) {
        const double* xs = v.x + (size_t)f * v.ld;
        const int a = threadIdx.x;
        anchor[(size_t)j * kNccAnchor + a] = a < 7 ? xs[a] : (a < 9 ? uv[2 * i + a - 7] : 1.0);
    }
}

// Map compaction (ekfb_map_management): templates and anchors follow their features.  keepList[r] = old index of the r-th
// surviving feature (ordered), built from the removal flags of k_map_plan.  One CTA per filter.
__global__ void __launch_bounds__(256) k_ncc_compact(DevView v, const uint8_t* tmpl, uint8_t* tmpl2, const double* anchor, double* anchor2,
                                                     int Nold_max)
{
    const int f = blockIdx.x;
    const size_t fo = (size_t)f * v.Nmax;
    extern __shared__ int keepList[];
    __shared__ int nKeep;
    if (threadIdx.x == 0) {   // (serial: a few hundred features, once per map change)
        int r = 0;
        const int Nold = Nold_max;
        for (int i = 0; i < Nold; ++i)
            if (v.mapflag[fo + i] == 0) keepList[r++] = i;
        nKeep = min(r, fdims(v, f)[D_MAP_NEW_NF]);
    }
    __syncthreads();
    const size_t tstride = (size_t)kNccLevels * 128;
    for (int r = 0; r < nKeep; ++r) {
        const int i = keepList[r];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tmpl + (fo + i) * tstride);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tmpl2 + (fo + r) * tstride);
        for (int e = threadIdx.x; e < (int)(tstride / 4); e += blockDim.x) dst[e] = src[e];
        if (threadIdx.x < kNccAnchor) anchor2[(fo + r) * kNccAnchor + threadIdx.x] = anchor[(fo + i) * kNccAnchor + threadIdx.x];
    }
}

// A = d(current pixel)/d(anchor pixel) of the plane-induced warp (see the header); false: use the raw template
__device__ inline bool ncc_warp_matrix(const CamParams& c, const double* an, const double* X, const double* r1, const double* q1, double* A)
{
    double R0[9], R1[9];
    quat_to_rot(an + 3, R0);
    quat_to_rot(q1, R1);
    double nrm[3] = {X[0] - an[0], X[1] - an[1], X[2] - an[2]};
    const double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    if (!(len > 0.0)) return false;
    for (int a = 0; a < 3; ++a) nrm[a] /= len;
    double uvw[3][2];
    for (int e = 0; e < 3; ++e) {
        const double u = an[7] + (e == 1 ? 1.0 : 0.0), w = an[8] + (e == 2 ? 1.0 : 0.0);
        const double dc[3] = {(u - c.cx) / c.fx, (w - c.cy) / c.fy, 1.0};
        double dw[3];
        for (int a = 0; a < 3; ++a) dw[a] = R0[a * 3] * dc[0] + R0[a * 3 + 1] * dc[1] + R0[a * 3 + 2] * dc[2];
        const double den = nrm[0] * dw[0] + nrm[1] * dw[1] + nrm[2] * dw[2];
        const double t = len / den;   // n . (X - r0) = len
        double pw[3], pc[3];
        for (int a = 0; a < 3; ++a) pw[a] = an[a] + t * dw[a] - r1[a];
        for (int a = 0; a < 3; ++a) pc[a] = R1[a] * pw[0] + R1[3 + a] * pw[1] + R1[6 + a] * pw[2];   // R1^T
        uvw[e][0] = c.cx + c.fx * pc[0] / pc[2];
        uvw[e][1] = c.cy + c.fy * pc[1] / pc[2];
    }
    A[0] = uvw[1][0] - uvw[0][0]; A[1] = uvw[2][0] - uvw[0][0];
    A[2] = uvw[1][1] - uvw[0][1]; A[3] = uvw[2][1] - uvw[0][1];
    const double det = A[0] * A[3] - A[1] * A[2];
    if (!(det >= 0.25 && det <= 4.0)) return false;   // (also false for NaN)
    const double dev = fmax(fmax(fabs(A[0] - 1.0), fabs(A[3] - 1.0)), fmax(fabs(A[1]), fabs(A[2])));
    return dev >= 0.05;
}

// T'(tx, ty) = bilinear sample of T at (5, 5) + Ainv (tx - 5, ty - 5), clamped to the patch, rounded to the nearest byte
__device__ __forceinline__ uint8_t ncc_warp_sample(const uint8_t* T, const double* Ainv, int tx, int ty)
{
    const double ox = tx - 5.0, oy = ty - 5.0;
    double sx = 5.0 + (Ainv[0] * ox + Ainv[1] * oy), sy = 5.0 + (Ainv[2] * ox + Ainv[3] * oy);
    sx = fmin(fmax(sx, 0.0), 10.0);
    sy = fmin(fmax(sy, 0.0), 10.0);
    const int x0 = min((int)sx, 9), y0 = min((int)sy, 9);
    const double fx = sx - x0, fy = sy - y0;
    const double t00 = T[y0 * kNccP + x0], t01 = T[y0 * kNccP + x0 + 1], t10 = T[(y0 + 1) * kNccP + x0], t11 = T[(y0 + 1) * kNccP + x0 + 1];
    const double top = t00 + fx * (t01 - t00), bot = t10 + fx * (t11 - t10);
    return (uint8_t)(int)(top + fy * (bot - top) + 0.5);
}

struct NccBest { double s; int dy, dx; };
__device__ __forceinline__ bool ncc_better(double s, int dy, int dx, const NccBest& b)
{
    return s > b.s || (s == b.s && (dy < b.dy || (dy == b.dy && dx < b.dx)));
}

// grid N (one CTA per feature of filter f), block 128
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

__global__ void __launch_bounds__(128) k_search_ncc(DevView v, NccView nv, const __grid_constant__ NccMaps maps, int f)
{
    __shared__ __align__(128) uint8_t win[kNccBoxW * kNccBoxH];
    __shared__ __align__(16) uint8_t tm[128];
    __shared__ __align__(8) uint64_t bar;
    __shared__ NccBest wbest[4];
    __shared__ int sBest[3];   // x, y (level coordinates), valid
    __shared__ int sSum[2];    // S_t, S_tt
    __shared__ __align__(16) uint8_t traw[128];
    __shared__ double sAinv[4];
    __shared__ int sWarp;
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = fdims(v, f)[D_N_FEAT];
    if (j >= N) return;
    const size_t fj = (size_t)f * v.Nmax + j;
    if (tid == 0) {   // unmatched unless the search succeeds
        v.mflag[fj] = 0;
        v.mkp[fj] = -1;
        nv.score[j] = -2.0;
        nv.level[j] = -1;
    }
    if (!v.vis[fj]) return;
    const int aw = __float2int_rn(v.ellax[fj * 2]), ah = __float2int_rn(v.ellax[fj * 2 + 1]);
    const float cxf = (float)v.h[fj * 2], cyf = (float)v.h[fj * 2 + 1];
    const double ang = v.ellang[fj];
    const int R0 = max(aw, ah);
    int lev = 0;
    while (lev < kNccLevels - 1 && ((R0 + (1 << lev) - 1) >> lev) > kNccR) ++lev;
    int R = min(kNccR, (R0 + (1 << lev) - 1) >> lev);
    int cx = ((int)cxf) >> lev, cy = ((int)cyf) >> lev;   // arithmetic shift: floor, also for negative centres
    if (tid == 0) {
        mbar_init(&bar, 1);
        nv.level[j] = lev;
        // the warp of the template for the current camera (see the header): one 2x2 matrix per feature, all levels
        sWarp = 0;
        if (nv.warp && nv.anchor != nullptr && nv.anchor[(size_t)j * kNccAnchor + 9] != 0.0) {
            const double* x = v.x + (size_t)f * v.ld;
            const double* y = x + v.foff[fj];
            double X[3], A[4];
            if (v.ftype[fj] == kTypeInvDepth) {
                double m[3];
                direction(y[3], y[4], m);
                for (int a = 0; a < 3; ++a) X[a] = y[a] + m[a] / y[5];
            } else
                for (int a = 0; a < 3; ++a) X[a] = y[a];
            if (ncc_warp_matrix(v.cam, nv.anchor + (size_t)j * kNccAnchor, X, x, x + 3, A)) {
                const double det = A[0] * A[3] - A[1] * A[2];
                sAinv[0] = A[3] / det; sAinv[1] = -A[1] / det; sAinv[2] = -A[2] / det; sAinv[3] = A[0] / det;
                sWarp = 1;
            }
        }
    }
    __syncthreads();
    int phase = 0;
    bool first = true;
    for (int l = lev; l >= 0; --l) {
        const int Wl = nv.W[l], Hl = nv.H[l];
        // window origin (x0, y0); columns are fetched from the 16-byte aligned x0a <= x0, clamped to the row
        const int x0 = cx - R - 5, y0 = cy - R - 5, x0a = x0 & ~15, xoff = x0 - x0a;
        if (tid == 0 && nv.tmaLevel[l]) {
            // the whole window in one tensor-map TMA load; rows / columns outside the level arrive as zeros
            mbar_expect_tx(&bar, kNccBoxW * kNccBoxH);
            tma_load_2d(win, &maps.m[l], x0a, y0, &bar);
        } else if (tid == 0) {
            const int xs = max(x0a, 0), xe = min(x0a + kNccBoxW, nv.pitch[l]);
            const int ys = max(y0, 0), ye = min(y0 + kNccBoxH, Hl);
            const int bytes = xe - xs, rows = ye - ys;
            const bool any = bytes > 0 && rows > 0;
            mbar_expect_tx(&bar, any ? bytes * rows : 0);
            if (any)
                for (int yy = ys; yy < ye; ++yy)
                    tma_bulk_load(win + (yy - y0) * kNccBoxW + (xs - x0a), nv.img[l] + (size_t)yy * nv.pitch[l] + xs, bytes, &bar);
        }
        if (tid < 32) reinterpret_cast<uint32_t*>(sWarp ? traw : tm)[tid] = reinterpret_cast<const uint32_t*>(nv.tmpl + ((size_t)j * kNccLevels + l) * 128)[tid];
        __syncthreads();
        if (sWarp) {
            if (tid < kNccPP) tm[tid] = ncc_warp_sample(traw, sAinv, tid % kNccP, tid / kNccP);
            __syncthreads();
        }
        if (warp == 0) {   // template sums by warp-shuffle reduction
            int st = 0, stt = 0;
            for (int e = lane; e < kNccPP; e += 32) { const int t = tm[e]; st += t; stt += t * t; }
            for (int o = 16; o > 0; o >>= 1) { st += __shfl_down_sync(0xffffffffu, st, o); stt += __shfl_down_sync(0xffffffffu, stt, o); }
            if (lane == 0) { sSum[0] = st; sSum[1] = stt; }
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        const long long St = sSum[0], Stt = sSum[1];
        const long long dt = kNccPP * Stt - St * St;
        NccBest best = {-3.0, 0, 0};
        for (int iy = warp; iy <= 2 * R; iy += 4) {
            const int dy = iy - R, dx = lane - R, px = cx + dx, py = cy + dy;
            bool ok = lane <= 2 * R && dt != 0 && px >= 5 && px < Wl - 5 && py >= 5 && py < Hl - 5;
            if (ok && first) ok = inside_gate((float)(px << l), (float)(py << l), cxf, cyf, aw, ah, ang);
            if (ok) {
                int sw = 0, sww = 0, stw = 0;
                const uint8_t* wp = win + iy * kNccBoxW + lane + xoff;
#pragma unroll
                for (int ty = 0; ty < kNccP; ++ty)
#pragma unroll
                    for (int tx = 0; tx < kNccP; ++tx) {
                        const int w = wp[ty * kNccBoxW + tx], t = tm[ty * kNccP + tx];
                        sw += w; sww += w * w; stw += t * w;
                    }
                const long long num = (long long)kNccPP * stw - St * sw;
                const long long dw = (long long)kNccPP * sww - (long long)sw * sw;
                if (dw != 0) {
                    const double s = __ddiv_rn((double)num, __dsqrt_rn(__dmul_rn((double)dt, (double)dw)));
                    if (ncc_better(s, dy, dx, best)) { best.s = s; best.dy = dy; best.dx = dx; }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {   // warp arg-max with the tie rule
            NccBest ot;
            ot.s = __shfl_down_sync(0xffffffffu, best.s, o);
            ot.dy = __shfl_down_sync(0xffffffffu, best.dy, o);
            ot.dx = __shfl_down_sync(0xffffffffu, best.dx, o);
            if (ncc_better(ot.s, ot.dy, ot.dx, best)) best = ot;
        }
        if (lane == 0) wbest[warp] = best;
        __syncthreads();
        if (tid == 0) {
            NccBest b = wbest[0];
            for (int w = 1; w < 4; ++w)
                if (ncc_better(wbest[w].s, wbest[w].dy, wbest[w].dx, b)) b = wbest[w];
            sBest[2] = b.s > -2.5;
            sBest[0] = cx + b.dx;
            sBest[1] = cy + b.dy;
            if (l == 0 && sBest[2]) nv.score[j] = b.s;
        }
        __syncthreads();
        if (!sBest[2]) return;   // no valid candidate at this level
        if (l == 0) {
            cx = sBest[0];
            cy = sBest[1];
        } else {
            cx = 2 * sBest[0];
            cy = 2 * sBest[1];
            R = 1;
            first = false;
        }
        __syncthreads();   // window and wbest are rewritten by the next level
    }
    if (tid == 0 && nv.score[j] >= nv.ncc_min) {
        v.mflag[fj] = 1;
        v.z[fj * 2] = (double)cx;
        v.z[fj * 2 + 1] = (double)cy;
        v.mdist[fj] = (float)(1.0 - nv.score[j]);
        v.mkp[fj] = -1;
    }
}

}  // namespace ekf
