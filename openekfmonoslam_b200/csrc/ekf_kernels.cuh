// ekf_kernels.cuh -- per-frame kernels other than the dense linear algebra: covariance/state
// prediction, measurement prediction + Jacobians, mask + descriptor matching, 1-point RANSAC
// hypothesis scoring, outlier rescue, map-feature bookkeeping.  sm_100a, FP64 maths, one grid
// dimension always indexes the filter of the batch.
//
// Data layout (device, per filter f, all arrays filter-major with fixed strides):
//   x[f][ld]            state vector: camera 13, then features at their covarianceMatrixPos
//   P[f][nmax][ld]      covariance, row-major, leading dimension ld (multiple of 16 doubles = 128 B)
//   ftype/foff[f][Nmax] feature type (1 XYZ / 2 inverse depth) and covarianceMatrixPos
//   per-frame, indexed by FEATURE (lists are ascending feature order everywhere in the reference,
//   so a list is a flag array plus a rank->feature table built by an ordered block scan):
//   vis,h,Si,Hx,Hf,ell  measurement prediction;  mflag,z,mkp,mdist  matches;  inl/outl/resc flags
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ekf_math.cuh"
#include "ekf_raster.cuh"

namespace ekf {

enum DimSlot {
    D_N_STATE = 0, D_N_FEAT = 1, D_N_KP = 2, D_N_PRED = 3, D_N_MATCH = 4, D_N_HYP = 5, D_BEST_HYP = 6, D_N_INL = 7,
    D_N_OUT = 8, D_N_RESC = 9, D_STATUS = 10, D_RANSAC_DONE = 11, D_RANSAC_NEXT = 12, D_RANSAC_CAP = 13,
    D_BEST_COUNT = 14, D_ULIST = 15, D_N_PRED2 = 16,
    // map management plan (ekf_map.cuh)
    D_MAP_CHANGED = 17, D_MAP_NEW_N = 18, D_MAP_NEW_NF = 19, D_MAP_CONVERT = 20, D_MAP_NEEDED = 21, D_MAP_NBAD = 22,
    D_MAP_NUNSEEN = 23, D_MAP_CONV_OLDOFF = 24, D_MAP_CONV_NEWOFF = 25,
    D_PRED_TICKET = 26 /* blocks of k_predict_cov that have finished (reset by the last one) */,
    D_STATUS_EVER = 27 /* OR of the per-frame status values since ekfb_set_state */,
    /* "last block done" tickets of kernels whose tail is run by the block that finishes last (each resets its own) */
    D_TICKET_MATCH = 28, D_TICKET_HYP = 29, D_TICKET_MEAS = 30, D_TICKET_UPD = 31, D_STRIDE = 32
};

struct DevView {
    int F, Nmax, nmax, ld, Kpmax, kmax, ldS, W, H, supWords, maxAxes;
    CamParams cam;
    double sd_lin, sd_ang, sigma_px, match_coef, ransac_thr, ransac_p, chi2;
    double* x; double* P;
    int* ftype; int* foff; uint8_t* desc; int* tpred; int* tmatch; int* dims;
    uint8_t* vis; double* h; double* Si; double* Hx; double* Hf; float* ellax; double* ellang;
    uint8_t* vis2; double* h2; double* Si2; double* Hx2; double* Hf2;
    uint8_t* mflag; double* z; int* mkp; float* mdist; int* mlist;
    uint8_t* inl; uint8_t* outl; uint8_t* resc; int* ulist;
    const float* const* kpxy; const uint8_t* const* kpdesc; uint8_t* kpok; uint8_t* mask;
    int* hypcount; uint32_t* hypsup;
    double* Bu; double* S; double* Sf; double* dx; double* Jq; double* Uinv; long long* dbg;
    // map management (ekf_map.cuh): second copies the compaction writes into (swapped with the live ones afterwards),
    // new row -> old row table, per-feature removal flags, conversion Jacobian, add-feature staging
    double* P2; double* x2; int* ftype2; int* foff2; uint8_t* desc2; int* tpred2; int* tmatch2;
    int* rowsrc; uint8_t* mapflag; double* convJ; double* addJ; double* adduv; uint8_t* adddesc;
    // zero-copy readback of the per-filter counters (mapped pinned host memory): hostDims[F][D_STRIDE], hostFlag[F]
    int* hostDims; volatile int* hostFlag;
    int faultInject;   // test hook (EKFB_OPT_FAULT_INJECT): factorisations report a non-positive pivot
};

__device__ __forceinline__ int* fdims(const DevView& v, int f) { return v.dims + (size_t)f * D_STRIDE; }

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be scheduled
// while its predecessor in the stream is still draining; it must not touch the predecessor's output before
// grid_dependency_wait(), which every kernel of the frame therefore executes first (a no-op for a plain launch).
// grid_launch_dependents() lets the successor be scheduled before this grid's blocks exit.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
#ifndef PDL_EARLY_TRIGGER
#define PDL_EARLY_TRIGGER 0   /* measured: triggering the successor at kernel start is slightly slower than letting it launch as blocks exit */
#endif

// The two kernel tails whose counters size the next launches (ransac_select_tail, rescue_gate_tail) end by writing the filter's counter
// block straight into mapped host memory and then a sequence number the host spins on: the host learns the counts a few
// microseconds after the kernel's last store instead of after a copy + stream synchronisation.  Call from all threads.
__device__ __forceinline__ void publish_dims(const DevView& v, int f, int seq)
{
    __syncthreads();
    if (threadIdx.x != 0 || v.hostDims == nullptr) return;
    const int* dm = fdims(v, f);
    int* hd = v.hostDims + (size_t)f * D_STRIDE;
#pragma unroll
    for (int i = 0; i < D_STRIDE; ++i) hd[i] = dm[i];
    __threadfence_system();
    v.hostFlag[f] = seq;
}

// "Last block done": called by ALL threads of a block once its global writes are issued; true in exactly one block per filter
// and launch -- the one that arrives last of `total` -- whose threads may then read what the other blocks wrote (through L2:
// use __ldcg / volatile for such reads) and run the launch's serial tail in place of a one-block follow-up kernel.
__device__ __forceinline__ bool last_block_done(int* ticket, int total)
{
    __shared__ int sLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        sLast = (t == total - 1);
        if (sLast) *ticket = 0;
        __threadfence();
    }
    __syncthreads();
    return sLast != 0;
}

// ---------------------------------------------------------------------------------------------
// P1: P <- F P F^T + G Q G^T on the 13 camera rows / columns (E/StateAndCovariancePrediction.cpp:154-240).
// Only rows 0..12 and their mirror columns change: a batched row-block update.  Block 0 does the
// 13x13 block, the others own 256 columns each: thread j reads P[0:13, j], forms F * col and writes
// the 13 new row entries (coalesced across j) plus the mirrored 13-double row segment P[j, 0:13].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_predict_cov(DevView v)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    const int f = blockIdx.y;
    const int n = fdims(v, f)[D_N_STATE];
    __shared__ double F[169], GQG[169], Pxx[169], T[169];
    double* P = v.P + (size_t)f * v.nmax * v.ld;
    const double* x = v.x + (size_t)f * v.ld;
    if (threadIdx.x == 0) motion_jacobians(x, v.sd_lin, v.sd_ang, F, GQG);
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < 169; e += blockDim.x) Pxx[e] = P[(size_t)(e / 13) * v.ld + (e % 13)];
        __syncthreads();
        for (int e = threadIdx.x; e < 169; e += blockDim.x) {
            const int i = e / 13, j = e % 13;
            double s = 0.;
            for (int k = 0; k < 13; ++k) s += F[i * 13 + k] * Pxx[k * 13 + j];
            T[e] = s;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 169; e += blockDim.x) {
            const int i = e / 13, j = e % 13;
            if (j > i) continue;  // lower triangle, mirrored: the block stays exactly symmetric
            double s = 0.;
            for (int k = 0; k < 13; ++k) s += T[i * 13 + k] * F[j * 13 + k];
            s += GQG[i * 13 + j];
            P[(size_t)i * v.ld + j] = s;
            P[(size_t)j * v.ld + i] = s;
        }
    } else {
        const int j = 13 + (blockIdx.x - 1) * blockDim.x + threadIdx.x;
        if (j < n) {
            double col[13], out[13];
#pragma unroll
            for (int k = 0; k < 13; ++k) col[k] = P[(size_t)k * v.ld + j];
#pragma unroll
            for (int i = 0; i < 13; ++i) {
                double s = 0.;
#pragma unroll
                for (int k = 0; k < 13; ++k) s += F[i * 13 + k] * col[k];
                out[i] = s;
            }
#pragma unroll
            for (int i = 0; i < 13; ++i) {
                P[(size_t)i * v.ld + j] = out[i];
                P[(size_t)j * v.ld + i] = out[i];
            }
        }
    }
    // P2, fused: x <- f(x) AFTER the covariance prediction (E/StateAndCovariancePrediction.cpp:43-65,252).  Every block read
    // the pre-prediction state before it took its ticket, so the block that draws the last ticket may overwrite it.
    __syncthreads();
    if (threadIdx.x == 0) {
        int* dmw = fdims(v, f);
        __threadfence();
        if (atomicAdd(&dmw[D_PRED_TICKET], 1) == (int)gridDim.x - 1) {
            dmw[D_PRED_TICKET] = 0;
            motion_predict(v.x + (size_t)f * v.ld);
        }
    }
}

// Establishes the device invariant "P is exactly symmetric" at upload: P <- 0.5 P + 0.5 P^T, which is
// what the reference applies at every update (E/Update.cpp:307).  grid (ceil(n/16), ceil(n/16), F), block (16,16)
__global__ void k_symmetrize(DevView v, int f)
{
    const int n = fdims(v, f)[D_N_STATE];
    const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
    if (i >= n || j > i) return;
    double* P = v.P + (size_t)f * v.nmax * v.ld;
    const double a = 0.5 * P[(size_t)i * v.ld + j] + 0.5 * P[(size_t)j * v.ld + i];
    P[(size_t)i * v.ld + j] = a;
    P[(size_t)j * v.ld + i] = a;
}

// ---------------------------------------------------------------------------------------------
// H1 + H2: one warp per feature.  Every lane evaluates the scalar chain h(x), H_x, H_f (no
// divergence); the 2x2 innovation covariance S_i = H_i P H_i^T + I is a gather of the (7+d)^2
// block of P split over the lanes and reduced with shuffles (E/MeasurementPrediction.cpp:203-265,
// 595-658).  mode 0: all features -> vis,h,Si,Hx,Hf,ell.  mode 1: the outlier subset after the
// low-innovation update -> vis2,h2,Si2,Hx2,Hf2 (E/EKF.cpp:464-468).
// ---------------------------------------------------------------------------------------------
__device__ void rescue_gate_tail(const DevView& v, int f, int seq);

__device__ __forceinline__ void measure_body(const DevView& v, int mode)
{
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    const int N = dm[D_N_FEAT];
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= N) return;
    const size_t fj = (size_t)f * v.Nmax + j;
    uint8_t* visOut = mode ? v.vis2 : v.vis;
    if (mode && !v.outl[fj]) {
        if (lane == 0) visOut[fj] = 0;
        return;
    }
    const double* x = v.x + (size_t)f * v.ld;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    const int type = v.ftype[fj], off = v.foff[fj];
    const int d = (type == kTypeInvDepth) ? 6 : 3;
    double R[9], Rinv[9], Rt[9], y[6], hh[2];
    quat_to_rot(x + 3, R);
    inv3(R, Rinv);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Rt[a * 3 + b] = R[b * 3 + a];
    for (int a = 0; a < 6; ++a) y[a] = (a < d) ? x[off + a] : 0.;
    const bool visible = predict_pixel(v.cam, x, Rt, Rinv, type, y, hh);
    if (!visible) {
        if (lane == 0) visOut[fj] = 0;
        return;
    }
    double Hx[14], Hf[12];
    measurement_jacobian(v.cam, x, x + 3, Rinv, type, y, hh, Hx, Hf);
    // S = sum_{a,b in A} H[.,a] P[ia,ib] H[.,b],  A = camera cols 0..6 and the feature's d cols
    const int na = 7 + d;
    double s00 = 0., s01 = 0., s10 = 0., s11 = 0.;
    for (int e = lane; e < na * na; e += 32) {
        const int a = e / na, b = e % na;
        const int ia = a < 7 ? a : off + a - 7, ib = b < 7 ? b : off + b - 7;
        const double p = P[(size_t)ia * v.ld + ib];
        const double h0a = a < 7 ? Hx[a] : Hf[a - 7], h1a = a < 7 ? Hx[7 + a] : Hf[6 + a - 7];
        const double h0b = b < 7 ? Hx[b] : Hf[b - 7], h1b = b < 7 ? Hx[7 + b] : Hf[6 + b - 7];
        s00 += h0a * p * h0b;
        s01 += h0a * p * h1b;
        s10 += h1a * p * h0b;
        s11 += h1a * p * h1b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s00 += __shfl_xor_sync(0xffffffffu, s00, o);
        s01 += __shfl_xor_sync(0xffffffffu, s01, o);
        s10 += __shfl_xor_sync(0xffffffffu, s10, o);
        s11 += __shfl_xor_sync(0xffffffffu, s11, o);
    }
    if (lane == 0) {
        double* hO = (mode ? v.h2 : v.h) + fj * 2;
        double* sO = (mode ? v.Si2 : v.Si) + fj * 4;
        double* hxO = (mode ? v.Hx2 : v.Hx) + fj * 14;
        double* hfO = (mode ? v.Hf2 : v.Hf) + fj * 12;
        visOut[fj] = 1;
        hO[0] = hh[0]; hO[1] = hh[1];
        const double S[4] = {s00 + 1.0, s01, s10, s11 + 1.0};
        for (int a = 0; a < 4; ++a) sO[a] = S[a];
        for (int a = 0; a < 14; ++a) hxO[a] = Hx[a];
        for (int a = 0; a < 12; ++a) hfO[a] = Hf[a];
        if (!mode) {
            float aw, ah;
            double ang;
            gate_ellipse(S, &aw, &ah, &ang);
            v.ellax[fj * 2] = aw;
            v.ellax[fj * 2 + 1] = ah;
            v.ellang[fj] = ang;
        }
    }
}

// seq > 0 (mode 1 only): the block of a filter that finishes last applies the chi-square gate to the re-predicted outliers
// (k_rescue_gate's body) and publishes the counters under that sequence number
// extra & 1 (mode 0): every block also clears its slice of the matching mask (the memset in front of the rasteriser);
// extra & 2 (mode 1): the tail also does the map-feature bookkeeping of the frame (k_update_map_features' body: it needs the
// inlier and rescued flags only, not the high-innovation update) -- ekfb_step uses both, the phase-by-phase API neither.
__device__ void map_features_tail(const DevView& v, int f);

__global__ void __launch_bounds__(256) k_measure(DevView v, int mode, int seq, int extra)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    if (extra & 1) {
        uint4* m4 = reinterpret_cast<uint4*>(v.mask + (size_t)blockIdx.y * v.W * v.H);
        const int n16 = (v.W * v.H) >> 4;   // (ekfb_measure only asks for this when W * H is a multiple of 16)
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n16; e += gridDim.x * blockDim.x) m4[e] = make_uint4(0, 0, 0, 0);
    }
    measure_body(v, mode);
    if (seq > 0 && last_block_done(fdims(v, blockIdx.y) + D_TICKET_MEAS, (int)gridDim.x)) {
        rescue_gate_tail(v, blockIdx.y, seq);
        if (extra & 2) map_features_tail(v, blockIdx.y);
    }
}

// ---------------------------------------------------------------------------------------------
// M1(1): mask = union of filled gate ellipses (E/Matching.cpp:193-202 -> Gui/Draw.cpp:42-64).
// One warp per predicted feature; dynamic smem: per warp a RasterScratch and 2*H span ints.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mask_raster(DevView v, uint8_t* maskBase, int maxAxes, int val)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f = blockIdx.y;
    const int N = fdims(v, f)[D_N_FEAT];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + wib;
    const size_t perWarp = sizeof(RasterScratch) + sizeof(int) * 2 * (size_t)v.H;
    RasterScratch* sc = reinterpret_cast<RasterScratch*>(smem_raw + perWarp * wib);
    int* spans = reinterpret_cast<int*>(smem_raw + perWarp * wib + sizeof(RasterScratch));
    if (j >= N) return;
    const size_t fj = (size_t)f * v.Nmax + j;
    if (!v.vis[fj]) return;
    // cv::Point2d -> Point2f -> Point (truncation); Size2f -> MIN(axis, maxAxes) -> int (truncation)
    const float cxf = (float)v.h[fj * 2], cyf = (float)v.h[fj * 2 + 1];
    const int icx = (int)cxf, icy = (int)cyf;
    const float mw = fminf(v.ellax[fj * 2], (float)maxAxes), mh = fminf(v.ellax[fj * 2 + 1], (float)maxAxes);
    const int iw = (int)mw, ih = (int)mh;
    const double angDeg = v.ellang[fj] * 180.0 / kPiTrunc;
    raster_ellipse_warp(maskBase + (size_t)f * v.W * v.H, v.W, v.H, icx, icy, iw, ih, angDeg, sc, spans, lane, (uint8_t)val);
}

// ---------------------------------------------------------------------------------------------
// M1(4) + M2 + M3: one warp per predicted feature.  Lanes stride over the keypoints, gate them
// with the foci test (C/EKFMath.cpp:302-351), ballot-compact the candidates IN KEYPOINT ORDER and
// replay the reference's order-dependent push-front 2-best rule (E/Matching.cpp:116-144) and its
// ratio test (:169-175).  Hamming distance of 32-byte descriptors with __popc.
// ---------------------------------------------------------------------------------------------
__device__ void after_match_tail(const DevView& v, int f);   // (below: counts, match list, RANSAC state reset)

// kpMask != 0: the detector-mask test of the keypoints is applied here as the keypoints are staged (block 0 also writes kpok);
// tail != 0: the block that finishes last runs k_after_match's body.  ekfb_match uses both: mask raster -> this kernel.
__global__ void __launch_bounds__(256) k_match(DevView v, int kpMask, int tail)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    // keypoint positions and mask flags are staged in shared memory, 1024 at a time, for the eight features of the CTA
    constexpr int CHUNK = 1024;
    __shared__ float2 sxy[CHUNK];
    __shared__ uint8_t sok[CHUNK];
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    const int N = dm[D_N_FEAT], Kp = dm[D_N_KP];
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t fj = (size_t)f * v.Nmax + (j < N ? j : 0);
    const bool active = j < N && v.vis[fj];
    const float2* xy2 = reinterpret_cast<const float2*>(v.kpxy[f]);
    const uint8_t* kd = v.kpdesc[f];
    const uint8_t* ok = v.kpok + (size_t)f * v.Kpmax;
    int iw = 0, ih = 0;
    float cxf = 0.f, cyf = 0.f;
    double ang = 0.0;
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) {
        iw = __float2int_rn(v.ellax[fj * 2]); ih = __float2int_rn(v.ellax[fj * 2 + 1]);  // Size2f -> Size rounds
        cxf = (float)v.h[fj * 2]; cyf = (float)v.h[fj * 2 + 1];
        ang = v.ellang[fj];
        const uint32_t* fd = reinterpret_cast<const uint32_t*>(v.desc + fj * 32);
#pragma unroll
        for (int a = 0; a < 8; ++a) q[a] = fd[a];
    }
    float minD = -1.f;
    int size = 0;
    float dFront = 0.f, dBack = 0.f;
    int iFront = -1, iBack = -1;
    float zx = 0.f, zy = 0.f;
    for (int c0 = 0; c0 < Kp; c0 += CHUNK) {
        const int cn = min(CHUNK, Kp - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < cn; e += blockDim.x) {
            const float2 pt = xy2[c0 + e];
            sxy[e] = pt;
            if (kpMask) {
                // keypoint survives the detector mask iff mask[(int)(y+0.5f)][(int)(x+0.5f)] != 0
                // (cv::KeyPointsFilter::runByPixelsMask, applied inside detector->detect, E/Matching.cpp:206)
                const int yy = (int)(pt.y + 0.5f), xx = (int)(pt.x + 0.5f);
                const bool o = xx >= 0 && xx < v.W && yy >= 0 && yy < v.H && v.mask[((size_t)f * v.H + yy) * v.W + xx] != 0;
                sok[e] = o;
                if (blockIdx.x == 0) v.kpok[(size_t)f * v.Kpmax + c0 + e] = o;
            } else
                sok[e] = ok[c0 + e];
        }
        __syncthreads();
        if (!active) continue;
        for (int base = 0; base < cn; base += 32) {
            const int kl = base + lane;
            bool cand = false;
            float dist = 0.f;
            float2 pt = make_float2(0.f, 0.f);
            if (kl < cn && sok[kl]) {
                pt = sxy[kl];
                cand = inside_gate(pt.x, pt.y, cxf, cyf, iw, ih, ang);
                if (cand) {
                    const uint32_t* cdp = reinterpret_cast<const uint32_t*>(kd + (size_t)(c0 + kl) * 32);
                    int dd = 0;
#pragma unroll
                    for (int a = 0; a < 8; ++a) dd += __popc(q[a] ^ cdp[a]);
                    dist = (float)dd;
                }
            }
            unsigned bal = __ballot_sync(0xffffffffu, cand);
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const float dcur = __shfl_sync(0xffffffffu, dist, src);
                if (dcur < minD || size < 2) {
                    minD = (minD < 0.f) ? dcur : fminf(minD, dcur);
                    dBack = dFront; iBack = iFront;  // push_front; a 3rd element falls off the back
                    dFront = dcur; iFront = c0 + base + src;
                    zx = __shfl_sync(0xffffffffu, pt.x, src); zy = __shfl_sync(0xffffffffu, pt.y, src);
                    if (size < 2) size++;
                }
            }
        }
    }
    (void)iBack;
    if (lane == 0 && j < N) {
        const bool accept = active && ((size == 1) || (size >= 2 && (double)dFront <= (double)dBack * v.match_coef));
        if (accept) {
            v.mflag[fj] = 1;
            v.mkp[fj] = iFront;
            v.mdist[fj] = dFront;
            v.z[fj * 2] = (double)zx;
            v.z[fj * 2 + 1] = (double)zy;
        } else {
            v.mflag[fj] = 0;
            if (active) v.mkp[fj] = -1;
        }
    }
    if (tail && last_block_done(fdims(v, f) + D_TICKET_MATCH, (int)gridDim.x)) after_match_tail(v, f);
}

// ---------------------------------------------------------------------------------------------
// ordered compaction of a per-feature flag array into a rank -> feature list (one CTA per filter)
// ---------------------------------------------------------------------------------------------
// (flags are read through L2: the callers include "last block done" tails that read what other blocks of the launch wrote)
__device__ inline int block_compact(const uint8_t* flags, int N, int* list)
{
    __shared__ int warpTot[32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < N; base += blockDim.x) {
        const int j = base + threadIdx.x;
        const bool fl = (j < N) && __ldcg(flags + j);
        const unsigned bal = __ballot_sync(0xffffffffu, fl);
        if (lane == 0) warpTot[wid] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int w = 0; w < wid; ++w) before += warpTot[w];
        if (fl) list[before + __popc(bal & ((1u << lane) - 1))] = j;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = running;
            for (int w = 0; w < nw; ++w) t += warpTot[w];
            running = t;
        }
        __syncthreads();
    }
    return running;
}

// after matching: counts, match list, RANSAC state reset (E/1PointRansac.cpp:101-125)
__device__ void after_match_tail(const DevView& v, int f)
{
    int* dm = fdims(v, f);
    const int N = dm[D_N_FEAT];
    const size_t fo = (size_t)f * v.Nmax;
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int c = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        c += v.vis[fo + j] ? 1 : 0;
        v.inl[fo + j] = 0;
        v.outl[fo + j] = 0;
        v.resc[fo + j] = 0;
    }
    atomicAdd(&cnt, c);
    const int m = block_compact(v.mflag + fo, N, v.mlist + fo);
    if (threadIdx.x == 0) {
        dm[D_N_PRED] = cnt;
        dm[D_N_MATCH] = m;
        dm[D_N_HYP] = 0;
        dm[D_BEST_HYP] = -1;
        dm[D_N_INL] = 0;
        dm[D_N_OUT] = 0;
        dm[D_N_RESC] = 0;
        dm[D_RANSAC_DONE] = (m == 0) ? 1 : 0;
        dm[D_RANSAC_NEXT] = 0;
        dm[D_RANSAC_CAP] = 1000;
        dm[D_BEST_COUNT] = 0;
        dm[D_ULIST] = 0;
        dm[D_N_PRED2] = 0;
        dm[D_STATUS_EVER] |= dm[D_STATUS];   // D_STATUS is per frame: a failed update only skips the rest of ITS frame
        dm[D_STATUS] = 0;
    }
}

// the same as its own launch (no features: nothing else of the matching phase runs)
__global__ void __launch_bounds__(256) k_after_match(DevView v)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    after_match_tail(v, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// R1: the hypotheses of 1-point RANSAC (E/1PointRansac.cpp:125-161).  Hypothesis i = i-th match.  The state-only
// EKF update uses K_i = P H_i^T (H_i P H_i^T + sigma I)^-1 with the symmetric-row gather P[:, c] = P[c, :]
// (13 coalesced row reads).  grid (chunk, F, ceil(Nmax / 64)): CTA (i, f, part) updates the 13 camera rows and the rows
// of its 64 features only, re-projects those features with the un-normalised q_i and ballot-packs the support of matched
// features within the pixel threshold into two words of the hypothesis' support set (the count is the popcount of the
// set, taken by k_ransac_select).  Hypotheses of a chunk are evaluated speculatively; the sequential acceptance rule is
// replayed by k_ransac_select.
// ---------------------------------------------------------------------------------------------
constexpr int kHypFeat = 64;  // features per CTA

__device__ void ransac_select_tail(const DevView& v, int f, int chunk0, int chunkLen, int seq);

__device__ __forceinline__ void ransac_hyp_body(const DevView& v, int chunk0)
{
    __shared__ double xc[13], xs[kHypFeat * 6];
    __shared__ double sK[4], sNu[2], sHx[14], sHf[12], sR[27];
    const int f = blockIdx.y, part = blockIdx.z;
    const int* dm = fdims(v, f);
    const int i = chunk0 + blockIdx.x;
    const int m = dm[D_N_MATCH];
    if (dm[D_RANSAC_DONE] || i >= m || (unsigned)i >= (unsigned)dm[D_RANSAC_CAP]) return;
    const int N = dm[D_N_FEAT];
    const int jBeg = part * kHypFeat;
    if (jBeg >= N && 2 * part >= v.supWords) return;
    const size_t fo = (size_t)f * v.Nmax;
    const int j0 = v.mlist[fo + i];
    const double* x = v.x + (size_t)f * v.ld;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    const int off0 = v.foff[fo + j0];
    const int d0 = v.ftype[fo + j0] == kTypeInvDepth ? 6 : 3;
    const int tid = threadIdx.x;
    if (tid == 0) {
        // S = H P H^T + sigma I: the measurement kernel already formed H P H^T + I
        const double* Si = v.Si + (fo + j0) * 4;
        const double S2[4] = {Si[0] - 1.0 + v.sigma_px, Si[1], Si[2], Si[3] - 1.0 + v.sigma_px};
        inv2(S2, sK);
        sNu[0] = deadband(v.z[(fo + j0) * 2] - v.h[(fo + j0) * 2]);
        sNu[1] = deadband(v.z[(fo + j0) * 2 + 1] - v.h[(fo + j0) * 2 + 1]);
    }
    if (tid < 14) sHx[tid] = v.Hx[(fo + j0) * 14 + tid];
    if (tid >= 32 && tid < 44) sHf[tid - 32] = v.Hf[(fo + j0) * 12 + tid - 32];
    __syncthreads();
    // rows of this CTA: slot < 13 -> camera row, else feature jBeg + (slot - 13) / 6, component (slot - 13) % 6
    for (int slot = tid; slot < 13 + kHypFeat * 6; slot += blockDim.x) {
        int row = slot;
        if (slot >= 13) {
            const int jj = (slot - 13) / 6, a = (slot - 13) % 6, j = jBeg + jj;
            row = -1;
            if (j < N && v.mflag[fo + j] && a < (v.ftype[fo + j] == kTypeInvDepth ? 6 : 3)) row = v.foff[fo + j] + a;
        }
        if (row < 0) continue;
        // all 13 row reads of P in flight at once (the loop over the feature's own columns has a fixed trip count of 6,
        // the XYZ case masks the last three: same sums in the same order)
        double pc[7], pf[6];
#pragma unroll
        for (int c = 0; c < 7; ++c) pc[c] = P[(size_t)c * v.ld + row];
#pragma unroll
        for (int c = 0; c < 6; ++c) pf[c] = c < d0 ? P[(size_t)(off0 + c) * v.ld + row] : 0.0;
        double p0 = 0., p1 = 0.;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            p0 += pc[c] * sHx[c];
            p1 += pc[c] * sHx[7 + c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c)
            if (c < d0) {
                p0 += pf[c] * sHf[c];
                p1 += pf[c] * sHf[6 + c];
            }
        const double k0 = p0 * sK[0] + p1 * sK[2];
        const double k1 = p0 * sK[1] + p1 * sK[3];
        const double dx = deadband(k0 * sNu[0] + k1 * sNu[1]);
        const double val = x[row] + dx;
        if (slot < 13) xc[slot] = val;
        else xs[slot - 13] = val;
    }
    __syncthreads();
    if (tid == 0) {
        quat_to_rot(xc + 3, sR);       // R(q_i), q_i not normalised (E/Update.cpp:168)
        inv3(sR, sR + 9);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) sR[18 + a * 3 + b] = sR[b * 3 + a];
    }
    __syncthreads();
    if (tid >= kHypFeat) return;
    uint32_t* sup = v.hypsup + ((size_t)f * v.Nmax + i) * v.supWords;
    const int j = jBeg + tid;
    bool s = false;
    if (j < N && v.mflag[fo + j]) {
        const int type = v.ftype[fo + j];
        double y[6], hh[2];
        const int d = type == kTypeInvDepth ? 6 : 3;
        for (int a = 0; a < 6; ++a) y[a] = a < d ? xs[tid * 6 + a] : 0.;
        if (predict_pixel(v.cam, xc, sR + 18, sR + 9, type, y, hh)) {
            const double ex = v.z[(fo + j) * 2] - hh[0], ey = v.z[(fo + j) * 2 + 1] - hh[1];
            s = sqrt(ex * ex + ey * ey) < v.ransac_thr;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, s);
    if ((tid & 31) == 0 && (j >> 5) < v.supWords) sup[j >> 5] = bal;
}

// seq > 0: the block of a filter that finishes last replays the acceptance rule (k_ransac_select's body) and publishes the
// counters under that sequence number: one launch per RANSAC round
__global__ void __launch_bounds__(448) k_ransac_hyp(DevView v, int chunk0, int seq)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    ransac_hyp_body(v, chunk0);
    if (seq > 0 && last_block_done(fdims(v, blockIdx.y) + D_TICKET_HYP, (int)(gridDim.x * gridDim.z)))
        ransac_select_tail(v, blockIdx.y, chunk0, (int)gridDim.x, seq);
}

// Sequential replay of the acceptance / adaptive-cap rule over one evaluated chunk
// (E/1PointRansac.cpp:125-186), then -- once the loop has ended -- the inlier/outlier split in
// match order (:201-227) and the inlier list for the low-innovation update.
__device__ void ransac_select_tail(const DevView& v, int f, int chunk0, int chunkLen, int seq)
{
    int* dm = fdims(v, f);
    const size_t fo = (size_t)f * v.Nmax;
    const int N = dm[D_N_FEAT], m = dm[D_N_MATCH];
    __shared__ int finished;
    __shared__ int cnts[64];   // support counts of the chunk's hypotheses = popcount of their support sets
    if ((int)threadIdx.x < chunkLen && threadIdx.x < 64) {
        const int i = chunk0 + threadIdx.x;
        int cnt = 0;
        if (i < m && !dm[D_RANSAC_DONE]) {
            const uint32_t* sw = v.hypsup + ((size_t)f * v.Nmax + i) * v.supWords;   // (written by other blocks: read through L2)
            for (int w = 0; w < v.supWords; ++w) cnt += __popc(__ldcg(sw + w));
        }
        cnts[threadIdx.x] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        finished = 0;
        if (!dm[D_RANSAC_DONE]) {
            int cap = dm[D_RANSAC_CAP], best = dm[D_BEST_COUNT], bestHyp = dm[D_BEST_HYP], i = dm[D_RANSAC_NEXT];
            const int end = chunk0 + chunkLen;
            bool done = false;
            for (; i < end; ++i) {
                if (!((unsigned)i < (unsigned)cap && i < m)) { done = true; break; }
                const int cnt = cnts[i - chunk0];
                if (cnt > best) {
                    best = cnt;
                    bestHyp = i;
                    // numberOfHipotesis = (uint) static_cast<int>( log(1-p) / log(1 - (1 - e)) ), e = 1 - best/m
                    const double e = 1.0 - (double)best / (double)m;
                    const double ratio = log(1.0 - v.ransac_p) / log(1.0 - (1.0 - e));
                    int asInt;
                    if (!(ratio > -2147483649.0 && ratio < 2147483648.0)) asInt = (int)0x80000000;  // x86 "indefinite"
                    else asInt = (int)ratio;
                    cap = asInt;  // compared as unsigned above, like the reference's uint
                }
            }
            if (!done && !((unsigned)i < (unsigned)cap && i < m)) done = true;
            dm[D_RANSAC_CAP] = cap;
            dm[D_BEST_COUNT] = best;
            dm[D_BEST_HYP] = bestHyp;
            dm[D_RANSAC_NEXT] = i;
            dm[D_N_HYP] = i;
            if (done) {
                dm[D_RANSAC_DONE] = 1;
                finished = 1;
            }
        }
    }
    __syncthreads();
    if (!finished) {
        publish_dims(v, f, seq);
        return;
    }
    const int bestHyp = dm[D_BEST_HYP];
    const uint32_t* sup = v.hypsup + ((size_t)f * v.Nmax + (bestHyp < 0 ? 0 : bestHyp)) * v.supWords;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const bool mt = v.mflag[fo + j];
        const bool in = mt && bestHyp >= 0 && ((__ldcg(sup + (j >> 5)) >> (j & 31)) & 1u);
        v.inl[fo + j] = in;
        v.outl[fo + j] = mt && !in;
    }
    __syncthreads();
    const int ni = block_compact(v.inl + fo, N, v.ulist + fo);
    if (threadIdx.x == 0) {
        dm[D_N_INL] = ni;
        dm[D_N_OUT] = m - ni;
        dm[D_ULIST] = ni;
    }
    publish_dims(v, f, seq);
}


// ---------------------------------------------------------------------------------------------
// X1: chi-square gate on the re-predicted outliers (E/EKF.cpp:477-506 + :68-119), then the
// rescued list for the high-innovation update.  One CTA per filter.
// ---------------------------------------------------------------------------------------------
__device__ void rescue_gate_tail(const DevView& v, int f, int seq)
{
    int* dm = fdims(v, f);
    const size_t fo = (size_t)f * v.Nmax;
    const int N = dm[D_N_FEAT];
    __shared__ int npred2;
    if (threadIdx.x == 0) npred2 = 0;
    __syncthreads();
    int c = 0;
    // (vis2, h2, Si2 were written by the other blocks of this launch: read through L2)
    for (int j = threadIdx.x; j < N; j += blockDim.x) c += (v.outl[fo + j] && __ldcg(v.vis2 + fo + j)) ? 1 : 0;
    atomicAdd(&npred2, c);
    __syncthreads();
    // npred2 == 0 with outliers present: the reference indexes an empty vector (undefined); rescue nothing.
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        bool r = false;
        if (npred2 > 0 && v.outl[fo + j] && __ldcg(v.vis2 + fo + j)) {
            const double d0 = v.z[(fo + j) * 2] - __ldcg(v.h2 + (fo + j) * 2), d1 = v.z[(fo + j) * 2 + 1] - __ldcg(v.h2 + (fo + j) * 2 + 1);
            double S2[4], Sinv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) S2[a] = __ldcg(v.Si2 + (fo + j) * 4 + a);
            inv2(S2, Sinv);
            const double t0 = d0 * Sinv[0] + d1 * Sinv[2];
            const double t1 = d0 * Sinv[1] + d1 * Sinv[3];
            r = (t0 * d0 + t1 * d1) < v.chi2;
        }
        v.resc[fo + j] = r;
    }
    __syncthreads();
    const int nr = block_compact(v.resc + fo, N, v.ulist + fo);
    if (threadIdx.x == 0) {
        dm[D_N_RESC] = nr;
        dm[D_ULIST] = nr;
        dm[D_N_PRED2] = npred2;
    }
    publish_dims(v, f, seq);
}

// updateMapFeatures (E/MapManagement.cpp:77-113): hit counters and descriptor refresh of inliers + rescued
__device__ __forceinline__ void map_feature_one(const DevView& v, int f, int j)
{
    const size_t fj = (size_t)f * v.Nmax + j;
    if (v.vis[fj]) v.tpred[fj] += 1;
    if (v.inl[fj] || v.resc[fj]) {
        v.tmatch[fj] += 1;
        if (v.mkp[fj] < 0) return;   // matched by the NCC search: there is no keypoint descriptor to refresh from
        const uint32_t* src = reinterpret_cast<const uint32_t*>(v.kpdesc[f] + (size_t)v.mkp[fj] * 32);
        uint32_t* dst = reinterpret_cast<uint32_t*>(v.desc + fj * 32);
#pragma unroll
        for (int a = 0; a < 8; ++a) dst[a] = src[a];
    }
}

__global__ void k_update_map_features(DevView v)
{
    grid_dependency_wait();
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid has started
    const int f = blockIdx.y;
    const int N = fdims(v, f)[D_N_FEAT];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) map_feature_one(v, f, j);
}

// the same by one block (tail of the rescue pass; resc was written by this block just before: a barrier lies in between)
__device__ void map_features_tail(const DevView& v, int f)
{
    __syncthreads();
    const int N = fdims(v, f)[D_N_FEAT];
    for (int j = threadIdx.x; j < N; j += blockDim.x) map_feature_one(v, f, j);
}

// keypoint counts of all filters in one launch (ekfb_select_frame / ekfb_set_keypoints_batch)
__global__ void k_set_kp_counts(DevView v, const int* counts)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < v.F) fdims(v, f)[D_N_KP] = counts[f];
}

// fixed-size result record per filter (13-state, 13x13 covariance block, counters)
struct RecordDev {
    double x_cam[13];
    double P_cam[169];
    int info[12];
};

__global__ void k_write_records(DevView v, RecordDev* out)
{
    const int f = blockIdx.x;
    const int* dm = fdims(v, f);
    const double* x = v.x + (size_t)f * v.ld;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    RecordDev* r = out + f;
    for (int e = threadIdx.x; e < 169; e += blockDim.x) r->P_cam[e] = P[(size_t)(e / 13) * v.ld + (e % 13)];
    if (threadIdx.x < 13) r->x_cam[threadIdx.x] = x[threadIdx.x];
    if (threadIdx.x == 32) {
        r->info[0] = dm[D_N_STATE];  r->info[1] = dm[D_N_FEAT];   r->info[2] = dm[D_N_KP];
        r->info[3] = dm[D_N_PRED];   r->info[4] = dm[D_N_MATCH];  r->info[5] = dm[D_N_HYP];
        r->info[6] = dm[D_BEST_HYP]; r->info[7] = dm[D_N_INL];    r->info[8] = dm[D_N_OUT];
        r->info[9] = dm[D_N_RESC];   r->info[10] = dm[D_STATUS];  r->info[11] = dm[D_STATUS_EVER] | dm[D_STATUS];
    }
}

}  // namespace ekf
