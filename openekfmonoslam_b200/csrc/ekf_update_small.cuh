// ekf_update_small.cuh -- U1 + U2 for a SMALL update (k <= 128 measurement rows: the high-innovation update of every frame,
// both updates of a small map): the factorisation of [S | nu], the slab TRSM  W^T = U^-T B  and  dx = W y  in ONE launch
// (reference: E/Update.cpp:92-141 -- S.inv(), K = P H^T S^-1, K nu).
//
// With one or two 64-row blocks the factorisation is 10-20 us of latency-bound work on one SM, while the per-block-step
// launches of ekf_schain.cuh cost a launch + a reload of every tile per step and leave 147 SMs idle, and the slab TRSM
// behind them is tiny (n k^2 = 13 MFLOP at n = 3013, k = 66).  So every CTA -- one per slab of SW columns of B -- factors
// S REDUNDANTLY in its own shared memory (no flags, no second launch, nothing published) and goes straight on to its slab:
//     T(0,0) = U00^T U00, Uinv0;   X01 = Uinv0^T S(0,1);   T(1,1) = S(1,1) - X01^T X01 = U11^T U11, Uinv1;   y = U^-T nu
//     X0 = Uinv0^T B0;   X1 = Uinv1^T (B1 - X01^T X0);   dx = X^T y;   W^T = X -> Bu
// All CTAs compute bit-identical factors (same code, same inputs), so W^T is consistent across slabs; a non-positive pivot is
// seen by every CTA, CTA 0 records EKFB_ERR_NUMERIC and nobody writes (x and P stay untouched).  The CTA that finishes last
// applies the state correction x += deadband(dx) and normalises the quaternion (U2, U4a).
// grid (ceil(n / SW), F), 256 threads, dynamic shared memory update_small_smem_bytes(SW).
#pragma once

#include "ekf_chain.cuh"

namespace ekf {

inline size_t update_small_smem_bytes(int SW)
{
    return sizeof(double) * ((size_t)5 * kNB * kSS + 4 * kNB + (size_t)3 * kNB * (SW + 4) + 8 * SW);
}

// identity outside the valid kb x kb part of a diagonal tile, W = 0 (the input of factor_tile64)
__device__ __forceinline__ void pad_diag_tile(double* T, double* W, int kb, int tid)
{
    for (int e = tid; e < kNB * 32; e += 256) {
        const int i = e >> 5, j = (e & 31) * 2;
        if (i >= kb || j + 1 >= kb) {
            if (i >= kb || j >= kb) T[i * kSS + j] = (i == j) ? 1.0 : 0.0;
            T[i * kSS + j + 1] = (i == j + 1) ? 1.0 : 0.0;
        }
        *reinterpret_cast<double2*>(W + i * kSS + j) = make_double2(0.0, 0.0);
    }
}

template <int SW>
__device__ __forceinline__ void update_small_body(const DevView& v, double* usm)
{
    constexpr int SWP = SW + 4, NT = SW / 8;
    double* W0 = usm;                     // Uinv_0
    double* W1 = W0 + kNB * kSS;          // Uinv_1
    double* Xa = W1 + kNB * kSS;          // S(0,1) -> X01 (K-major: row of block 0, column of block 1)
    double* T0 = Xa + kNB * kSS;          // S(0,0) -> U00
    double* T1 = T0 + kNB * kSS;          // S(1,1) -> U11
    double* nu = T1 + kNB * kSS;          // [128]
    double* y = nu + 2 * kNB;             // [128]
    double* Xs = y + 2 * kNB;             // [128][SWP] slab of B -> W^T
    double* Tt = Xs + 2 * kNB * SWP;      // [64][SWP]
    double* red = Tt + kNB * SWP;         // [8][SW]
    __shared__ int bad;
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3, p = w & 3, nh = w >> 2;
    int* dm = fdims(v, f);
    // the update list and the slab of B were written two or more launches ago (k_gain_rows precedes k_build_S): with
    // programmatic dependent launch this CTA may run while k_build_S drains, and loads its slab meanwhile
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (k == 0 || k > 2 * kNB) return;
    const int c0 = blockIdx.x * SW;
    if (c0 >= n) return;
    const int nbR = (k + kNB - 1) / kNB, kpad = nbR * kNB;
    double* Bg = v.Bu + (size_t)f * v.kmax * v.ld;
    for (int e = tid; e < kpad * SW; e += 256) {
        const int r = e / SW, c = e % SW;
        Xs[(size_t)r * SWP + c] = (r < k && c0 + c < n) ? Bg[(size_t)r * v.ld + c0 + c] : 0.0;
    }
    grid_dependency_wait();
    if (dm[D_STATUS] != 0) return;   // an earlier update of this frame failed: the rest of the frame is skipped
    if (tid == 0) bad = 0;
    const double* Sg = v.S + (size_t)f * v.kmax * v.ldS;
    load_tile64(T0, Sg, v.ldS, 0, k, 0, k, tid);
    if (nbR == 2) {
        load_tile64(Xa, Sg, v.ldS, 0, k, kNB, k, tid);
        load_tile64(T1, Sg, v.ldS, kNB, k, kNB, k, tid);
    }
    cp_async_commit();
    if (tid < kpad) nu[tid] = (tid < k) ? Sg[(size_t)tid * v.ldS + k] : 0.0;   // column k of S carries the innovation
    cp_async_wait<0>();
    __syncthreads();

    // ---- factorisation, redundantly in every CTA ----
    const int kb0 = min(kNB, k), kb1 = k - kNB;
    pad_diag_tile(T0, W0, kb0, tid);
    __syncthreads();
    factor_tile64(T0, W0, tid, &bad, nullptr, NoHook(), nullptr, (kb0 + 7) >> 3);
    __syncthreads();
    if (tid < kNB) {
        double s = 0.0;
        for (int pp = 0; pp <= tid; ++pp) s += W0[pp * kSS + tid] * nu[pp];
        y[tid] = s;
    }
    if (nbR == 2) {
        double x[2][4][2], xd[2][4][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) x[a][b][0] = x[a][b][1] = 0.0;
        x_gemm<false>(W0, Xa, Xa, p, nh, g, q, x, xd);
        __syncthreads();
        store_x_smem(x, Xa, p, nh, g, q);
        __syncthreads();
        double acc5[5][2];
#pragma unroll
        for (int t = 0; t < 5; ++t) acc5[t][0] = acc5[t][1] = 0.0;
        tile_mma_upper(Xa, w, g, q, acc5);
        tile_sub_upper(T1, w, g, q, acc5);
        if (tid < kNB) {   // nu_1 - X01^T y_0
            double s = nu[kNB + tid];
            for (int r = 0; r < kNB; ++r) s -= Xa[r * kSS + tid] * y[r];
            nu[kNB + tid] = s;
        }
        __syncthreads();
        pad_diag_tile(T1, W1, kb1, tid);
        __syncthreads();
        factor_tile64(T1, W1, tid, &bad, nullptr, NoHook(), nullptr, (kb1 + 7) >> 3);
        __syncthreads();
        if (tid < kNB) {
            double s = 0.0;
            for (int pp = 0; pp <= tid; ++pp) s += W1[pp * kSS + tid] * nu[kNB + pp];
            y[kNB + tid] = s;
        }
    }
    __syncthreads();
    if (bad || v.faultInject) {   // not positive definite: no W, x and P stay as they are
        if (blockIdx.x == 0 && tid == 0) dm[D_STATUS] = 4;   // EKFB_ERR_NUMERIC
        return;
    }

    // ---- slab TRSM: warp w owns rows 8w .. 8w+7 of a block; Uinv is upper triangular, so its k-range ends at 8w+8 ----
    const int m = 8 * w + g;
    double acc[NT][2];
#pragma unroll
    for (int b = 0; b < NT; ++b) acc[b][0] = acc[b][1] = 0.0;
    for (int k4 = 0; k4 < 8 * w + 8; k4 += 4) {
        const double a = W0[(k4 + q) * kSS + m];
#pragma unroll
        for (int b = 0; b < NT; ++b) dmma8x8x4(acc[b][0], acc[b][1], a, Xs[(size_t)(k4 + q) * SWP + 8 * b + g]);
    }
    __syncthreads();   // every warp has read rows 0..63 of the slab
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) Xs[(size_t)m * SWP + 8 * b + 2 * q + e] = (m < kb0) ? acc[b][e] : 0.0;
    __syncthreads();
    if (nbR == 2) {
#pragma unroll
        for (int b = 0; b < NT; ++b) acc[b][0] = acc[b][1] = 0.0;
#pragma unroll 4
        for (int k4 = 0; k4 < kNB; k4 += 4) {
            const double a = Xa[(k4 + q) * kSS + m];
#pragma unroll
            for (int b = 0; b < NT; ++b) dmma8x8x4(acc[b][0], acc[b][1], a, Xs[(size_t)(k4 + q) * SWP + 8 * b + g]);
        }
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * b + 2 * q + e;
                Tt[m * SWP + c] = (m < kb1) ? Xs[(size_t)(kNB + m) * SWP + c] - acc[b][e] : 0.0;
                acc[b][e] = 0.0;
            }
        __syncthreads();
        for (int k4 = 0; k4 < 8 * w + 8; k4 += 4) {
            const double a = W1[(k4 + q) * kSS + m];
#pragma unroll
            for (int b = 0; b < NT; ++b) dmma8x8x4(acc[b][0], acc[b][1], a, Tt[(size_t)(k4 + q) * SWP + 8 * b + g]);
        }
#pragma unroll
        for (int b = 0; b < NT; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) Xs[(size_t)(kNB + m) * SWP + 8 * b + 2 * q + e] = (m < kb1) ? acc[b][e] : 0.0;
        __syncthreads();
    }
    // ---- W^T back to global (rows k .. end of the last 16-row chunk as zeros: the TMA-fed downdate reads whole chunks), dx = W y ----
    const int kz = min(kpad, (k + 15) & ~15);
    for (int e = tid; e < kz * SW; e += 256) {
        const int r = e / SW, c = e % SW;
        if (c0 + c < n) Bg[(size_t)r * v.ld + c0 + c] = Xs[(size_t)r * SWP + c];
    }
    if (lane < SW) {
        double s = 0.;
        for (int r = w; r < k; r += 8) s += Xs[(size_t)r * SWP + lane] * y[r];
        red[w * SW + lane] = s;
    }
    __syncthreads();
    if (tid < SW && c0 + tid < n) {
        double s = 0.;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) s += red[ww * SW + tid];
        v.dx[(size_t)f * v.ld + c0 + tid] = s;
    }
}

// the filter's block that finishes last applies the state correction (state_apply_tail): no k_state_apply launch behind it
template <int SW>
__global__ void __launch_bounds__(256, 1) k_update_small(DevView v)
{
    extern __shared__ __align__(16) double usm[];
    update_small_body<SW>(v, usm);
    if (last_block_done(fdims(v, blockIdx.y) + D_TICKET_UPD, (int)gridDim.x)) state_apply_tail(v, blockIdx.y);
}

}  // namespace ekf
