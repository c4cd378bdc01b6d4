// ekf_linalg.cuh -- the dense part of the EKF update on sm_100a (E/Update.cpp:92-109,116-218,282-319).
//
// The reference forms K = P H^T (H P H^T + R)^-1 with a dense >99%-zero H and then P = (I - K H) P
// (2 n^3 flop).  Here the update is the partial Cholesky of the augmented matrix
//
//        [ S   B  nu ]         S = H P H^T + sigma I   (k x k)
//        [ B^T P     ]         B = (P H^T)^T           (k x n, rows gathered from 7+d rows of P)
//
// in upper/row form:  S = U^T U,  W^T = U^-T B,  y = U^-T nu,  then  x += W y  and  P -= W W^T.
// All operands are stored K-major (row r of B / W^T is one measurement row, contiguous over the
// state index), so every kernel streams rows and the two big contractions are the same "TN" GEMM
//        C[m][n] -= sum_r A[r][m] * B[r][n]
// executed on the FP64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA.8x8x4 on sm_100a) from a
// cp.async-fed shared-memory ring.
#pragma once

#include "ekf_kernels.cuh"

namespace ekf {

constexpr int kNB = 64;  // Cholesky block

struct UpdSrc {  // which prediction arrays feed the update (all-features pass or the rescue re-prediction)
    const double* h; const double* Hx; const double* Hf;
};

__device__ __forceinline__ UpdSrc upd_src(const DevView& v, int which)
{
    UpdSrc s;
    s.h = which ? v.h2 : v.h;
    s.Hx = which ? v.Hx2 : v.Hx;
    s.Hf = which ? v.Hf2 : v.Hf;
    return s;
}

// ---------------------------------------------------------------------------------------------
// U1 front: B = (P H^T)^T for the update list.  H_a touches camera columns 0..6 and the d columns
// of its own feature, so row pair a of B is a combination of 7 + d ROWS of P (P symmetric) --
// 13 coalesced row reads instead of the reference's dense n x n x k product (E/Update.cpp:105).
// Column n of B carries the dead-banded innovation nu (E/Update.cpp:133-134); padding columns are 0.
// grid (ceil(ld/256), ku_max, F)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gain_rows(DevView v, int which)
{
    const int f = blockIdx.z, a = blockIdx.y;
    const int* dm = fdims(v, f);
    if (a >= dm[D_ULIST]) return;
    const int n = dm[D_N_STATE];
    const size_t fo = (size_t)f * v.Nmax;
    const int j = v.ulist[fo + a];
    const UpdSrc s = upd_src(v, which);
    __shared__ double sHx[14], sHf[12];
    if (threadIdx.x < 14) sHx[threadIdx.x] = s.Hx[(fo + j) * 14 + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 44) sHf[threadIdx.x - 32] = s.Hf[(fo + j) * 12 + threadIdx.x - 32];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.ld) return;
    const int off = v.foff[fo + j];
    const int d = v.ftype[fo + j] == kTypeInvDepth ? 6 : 3;
    const double* P = v.P + (size_t)f * v.nmax * v.ld;
    double* B0 = v.Bu + ((size_t)f * v.kmax + 2 * a) * v.ld;
    double b0 = 0., b1 = 0.;
    if (i < n) {
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const double p = P[(size_t)c * v.ld + i];
            b0 += sHx[c] * p;
            b1 += sHx[7 + c] * p;
        }
        for (int c = 0; c < d; ++c) {
            const double p = P[(size_t)(off + c) * v.ld + i];
            b0 += sHf[c] * p;
            b1 += sHf[6 + c] * p;
        }
    } else if (i == n) {
        b0 = deadband(v.z[(fo + j) * 2] - s.h[(fo + j) * 2]);
        b1 = deadband(v.z[(fo + j) * 2 + 1] - s.h[(fo + j) * 2 + 1]);
    }
    B0[i] = b0;
    B0[v.ld + i] = b1;
    if (a == 0) v.dx[(size_t)f * v.ld + i] = 0.0;  // reset the state-correction accumulator
}

// S = H B^T + sigma I, sparse in H again (E/Update.cpp:95-107).  grid (ceil(k/16), ceil(k/16), F), block (16,16)
__global__ void k_build_S(DevView v, int which)
{
    const int f = blockIdx.z;
    const int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    const int ra = blockIdx.y * 16 + threadIdx.y, rb = blockIdx.x * 16 + threadIdx.x;
    if (ra >= k || rb >= k) return;
    const size_t fo = (size_t)f * v.Nmax;
    const int a = ra >> 1, r = ra & 1;
    const int j = v.ulist[fo + a];
    const UpdSrc s = upd_src(v, which);
    const double* Hx = s.Hx + (fo + j) * 14 + 7 * r;
    const double* Hf = s.Hf + (fo + j) * 12 + 6 * r;
    const int off = v.foff[fo + j];
    const int d = v.ftype[fo + j] == kTypeInvDepth ? 6 : 3;
    const double* Brow = v.Bu + ((size_t)f * v.kmax + rb) * v.ld;
    double acc = 0.;
#pragma unroll
    for (int c = 0; c < 7; ++c) acc += Hx[c] * Brow[c];
    for (int c = 0; c < d; ++c) acc += Hf[c] * Brow[off + c];
    if (ra == rb) acc += v.sigma_px;
    v.S[((size_t)f * v.kmax + ra) * v.ldS + rb] = acc;
    if (rb == 0)  // column k of S carries the dead-banded innovation nu (it becomes y = U^-T nu)
        v.S[((size_t)f * v.kmax + ra) * v.ldS + k] = deadband(v.z[(fo + j) * 2 + r] - s.h[(fo + j) * 2 + r]);
}

// ---------------------------------------------------------------------------------------------
// Cholesky step J, panel kernel: the 64-row block J of the augmented matrix [S | B | nu] becomes
// [U_JJ | X_J | y_J] with X_J = U_JJ^-T A_J.  Every CTA owns 128 columns (one per thread) and
//   (1) redundantly eliminates the 64x64 diagonal tile (plus the nu column) with one column per
//       thread in registers -- one barrier per pivot, reciprocal instead of division -- leaving the
//       multipliers M[c][i] = A(c)[c][i] / A(c)[c][c] and 1/sqrt(pivot) in shared memory;
//   (2) applies the same row operations to its own columns: 64 values per thread in registers,
//       2016 FMAs, no barrier (forward substitution without forming an inverse);
//   (3) accumulates its share of the state correction dx[col] += sum_r X[r][col] * y[r]
//       (E/Update.cpp:136-141: K nu = W y), own columns only, so no atomics and a fixed order.
// Column tiles run over the virtual concatenation of S's columns right of the block and B's n+1
// columns.  grid (tiles, F), 128 threads.
// ---------------------------------------------------------------------------------------------
// forward substitution of one column of the row block (64 values in registers, multipliers read as
// double2 broadcasts); a separate function so that its register allocation is its own
__device__ __noinline__ double panel_substitute(double* colp, int ldx, int kb, const double* Msm, const double* dinvs,
                                                const double* ysm)
{
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) x[i] = (i < kb) ? colp[(size_t)i * ldx] : 0.0;
    const double2* M2 = reinterpret_cast<const double2*>(Msm);
#pragma unroll
    for (int c = 0; c < kNB - 1; ++c) {
        const double xc = x[c];
#pragma unroll
        for (int p = c / 2; p < kNB / 2; ++p) {  // M[c][i] is zero for i <= c, so whole pairs are safe
            const double2 m = M2[c * (kNB / 2) + p];
            if (2 * p > c) x[2 * p] -= m.x * xc;
            x[2 * p + 1] -= m.y * xc;
        }
    }
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
        x[i] *= dinvs[i];
        part += x[i] * ysm[i];
        if (i < kb) colp[(size_t)i * ldx] = x[i];
    }
    return part;
}

constexpr int kPanelCols = 128;

// ---------------------------------------------------------------------------------------------
// Cholesky step J, panel kernel: the 64-row block J of the augmented matrix [S | B | nu] becomes
// [U_JJ | X_J | y_J] with X_J = U_JJ^-T A_J.  Every CTA owns 128 columns (one per thread) and
//   (1) redundantly eliminates the 64x64 diagonal tile plus the nu column: thread (i, h) keeps the
//       entries of row i in columns j = h (mod 2) in registers, the pivot row is broadcast through
//       a double-buffered shared row (one barrier per pivot, reciprocal instead of division),
//       leaving the multipliers M[c][i] = A(c)[c][i] / A(c)[c][c], 1/sqrt(pivot) and y_J in smem;
//   (2) applies the same row operations to its own columns: 64 values per thread in registers,
//       2016 FMAs, no barrier (forward substitution without forming an inverse);
//   (3) accumulates its share of the state correction dx[col] += sum_r X[r][col] * y[r]
//       (E/Update.cpp:136-141: K nu = W y), own columns only, so no atomics and a fixed order.
// Column tiles run over the virtual concatenation of S's columns right of the block and B's n+1
// columns.  grid (tiles, F), 128 threads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chol_panel(DevView v, int J)
{
    __shared__ __align__(16) double rowbuf[2][66];
    __shared__ __align__(16) double Msm[kNB * kNB];
    __shared__ double dinvs[kNB], ysm[kNB];
    const int f = blockIdx.y;
    int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    const int J0 = J * kNB, J1 = J0 + kNB;
    if (J0 >= k) return;
    const int kb = min(kNB, k - J0);
    const int nS = (k > J1) ? (k - J1 + kPanelCols - 1) / kPanelCols : 0;
    const int nB = (n + 1 + kPanelCols - 1) / kPanelCols;
    int t = blockIdx.x;
    double* base;
    int ldx, c0, cEnd;
    const bool isB = t >= nS;
    if (!isB) {
        base = v.S + ((size_t)f * v.kmax + J0) * v.ldS;
        ldx = v.ldS; c0 = J1 + t * kPanelCols; cEnd = k;
    } else {
        t -= nS;
        if (t >= nB) return;
        base = v.Bu + ((size_t)f * v.kmax + J0) * v.ld;
        ldx = v.ld; c0 = t * kPanelCols; cEnd = n + 1;
    }
    const int tid = threadIdx.x;
    double* Sd = v.S + ((size_t)f * v.kmax + J0) * v.ldS + J0;
    const double* nuCol = v.Bu + ((size_t)f * v.kmax + J0) * v.ld + n;
    for (int e = tid; e < kNB * kNB; e += blockDim.x) Msm[e] = 0.0;

    // ---- (1) diagonal tile elimination: thread (i, h) owns row i, columns j = 2 q + h, q = 0..32 ----
    {
        const int i = tid >> 1, h = tid & 1;
        double a[33];
#pragma unroll
        for (int q = 0; q < 33; ++q) {
            const int j = 2 * q + h;
            double val = 0.0;
            if (j < kNB) {
                if (i < kb && j < kb) { if (j >= i) val = Sd[(size_t)i * v.ldS + j]; }
                else if (i == j) val = 1.0;  // identity padding of a partial last block
            } else if (j == kNB) {
                if (i < kb) val = nuCol[(size_t)i * v.ld];
            }
            a[q] = val;
        }
        if (i == 0) {
#pragma unroll
            for (int q = 0; q < 33; ++q) rowbuf[0][2 * q + h] = a[q];
        }
        for (int c = 0; c < kNB; ++c) {
            __syncthreads();
            const double* rb = rowbuf[c & 1];
            const double piv = rb[c];
            const double pinv = 1.0 / piv;
            const double dinv = rsqrt(piv);
            if (tid == 0) {
                if (!(piv > 0.0)) dm[D_STATUS] = 4;  // EKFB_ERR_NUMERIC
                dinvs[c] = dinv;
                ysm[c] = rb[kNB] * dinv;
            }
            if (i > c) {
                const double m = rb[i] * pinv;
                if (h == 0) Msm[c * kNB + i] = m;
                const double2* rb2 = reinterpret_cast<const double2*>(rb);
#pragma unroll
                for (int q = 0; q < 33; ++q) {
                    const double2 r2 = rb2[q];
                    a[q] -= m * (h ? r2.y : r2.x);
                }
                if (i == c + 1) {
                    double* nb = rowbuf[(c + 1) & 1];
#pragma unroll
                    for (int q = 0; q < 33; ++q) nb[2 * q + h] = a[q];
                }
            }
        }
        __syncthreads();
    }

    // ---- (2) forward substitution on this CTA's columns ----
    const int col = c0 + tid;
    if (col >= cEnd) return;
    const double part = panel_substitute(base + col, ldx, kb, Msm, dinvs, ysm);
    // ---- (3) state correction share ----
    if (isB && col < n) v.dx[(size_t)f * v.ld + col] += part;
}

// ---------------------------------------------------------------------------------------------
// The TN contraction on the FP64 tensor pipe:   C[m][n] -= sum_{r<K} A[r][m] * B[r][n]
//   KIND 0: trailing update at Cholesky step J of S (A = B = X rows in S, upper tiles only) and of
//           B (A = X rows in S, B = X rows in B) in one launch: blockIdx.x runs over the virtual
//           concatenation of S's and B's column tiles
//   KIND 2: covariance downdate P -= W W^T            (A = B = W^T, K = k, lower tiles, mirrored
//           store so P stays exactly symmetric: replaces 0.5 P + 0.5 P^T of E/Update.cpp:307)
// CTA tile 128 x 128, 8 warps as 2 (m) x 4 (n), warp tile 64 x 32 = 8 x 4 DMMA m8n8k4 tiles.
// K is consumed in chunks of 16 rows through a 3-stage cp.async ring; the row pitch of 132 doubles
// makes both fragment loads bank-conflict free (lane -> (k = lane%4, m = lane/4) hits 16 distinct
// 8-byte banks per half warp).
// ---------------------------------------------------------------------------------------------
constexpr int kTM = 128, kTN = 128, kKC = 16, kLDS = 132, kStages = 3;
constexpr int kGemmSmemBytes = kStages * 2 * kKC * kLDS * (int)sizeof(double);

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// load a kKC x 128 slab (rows k0.., columns col0..) of a K-major operand into shared memory
__device__ __forceinline__ void load_slab(double* dst, const double* src, int ldsrc, int k0, int K, int col0,
                                          int colLimit, int tid)
{
#pragma unroll
    for (int it = 0; it < (kKC * 128 / 2) / 256; ++it) {
        const int c = tid + it * 256;
        const int r = c >> 6, c2 = (c & 63) * 2;
        double* d = dst + r * kLDS + c2;
        if (k0 + r < K && col0 + c2 < colLimit) {
            cp_async16(d, src + (size_t)(k0 + r) * ldsrc + col0 + c2);
        } else {
            d[0] = 0.0;
            d[1] = 0.0;
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256, 1) k_gemm_tn(DevView v, int J)
{
    extern __shared__ __align__(16) double gsm[];
    const int f = blockIdx.z;
    const int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    const double *A, *B;
    double* C;
    int lda, ldb, ldc, mBeg, mEnd, nBeg, nEnd, K, aLim, bLim;
    int tileN = blockIdx.x;
    bool sPart = false;
    if (KIND == 2) {
        A = B = v.Bu + (size_t)f * v.kmax * v.ld;
        C = v.P + (size_t)f * v.nmax * v.ld;
        lda = ldb = ldc = v.ld;
        mBeg = nBeg = 0; mEnd = nEnd = n; K = k; aLim = bLim = v.ld;
    } else {
        const int J0 = J * kNB, J1 = J0 + kNB;
        if (J1 >= k) return;
        A = v.S + ((size_t)f * v.kmax + J0) * v.ldS;
        lda = v.ldS; aLim = v.ldS;
        K = kNB;  // J1 < k, so the panel block is full
        mBeg = J1; mEnd = k;
        const int nS = (k - J1 + kTN - 1) / kTN;
        if ((int)blockIdx.x < nS) {
            B = A; ldb = lda; bLim = aLim;
            C = v.S + (size_t)f * v.kmax * v.ldS; ldc = v.ldS;
            nBeg = J1; nEnd = k; tileN = blockIdx.x; sPart = true;
        } else {
            B = v.Bu + ((size_t)f * v.kmax + J0) * v.ld; ldb = v.ld; bLim = v.ld;
            C = v.Bu + (size_t)f * v.kmax * v.ld; ldc = v.ld;
            nBeg = 0; nEnd = n + 1; tileN = blockIdx.x - nS;
        }
    }
    const int tm0 = mBeg + blockIdx.y * kTM, tn0 = nBeg + tileN * kTN;
    if (tm0 >= mEnd || tn0 >= nEnd) return;
    if (sPart && tn0 + kTN <= tm0) return;      // S tile strictly below the diagonal
    if (KIND == 2 && tn0 > tm0) return;         // lower tiles only

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int nk = (K + kKC - 1) / kKC;
    auto stageA = [&](int s) { return gsm + (size_t)s * 2 * kKC * kLDS; };
    auto stageB = [&](int s) { return gsm + (size_t)s * 2 * kKC * kLDS + kKC * kLDS; };
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nk) {
            load_slab(stageA(s), A, lda, s * kKC, K, tm0, aLim, tid);
            load_slab(stageB(s), B, ldb, s * kKC, K, tn0, bLim, tid);
        }
        cp_async_commit();
    }
    const int kq = lane & 3, mq = lane >> 2;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // prefetch chunk kt + kStages - 1 into the slot freed at iteration kt - 1
            const int nx = kt + kStages - 1;
            if (nx < nk) {
                load_slab(stageA(nx % kStages), A, lda, nx * kKC, K, tm0, aLim, tid);
                load_slab(stageB(nx % kStages), B, ldb, nx * kKC, K, tn0, bLim, tid);
            }
            cp_async_commit();
        }
        const double* As = stageA(kt % kStages) + wm * 64 + mq;
        const double* Bs = stageB(kt % kStages) + wn * 32 + mq;
#pragma unroll
        for (int k4 = 0; k4 < kKC; k4 += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int a = 0; a < 8; ++a) af[a] = As[(k4 + kq) * kLDS + a * 8];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[(k4 + kq) * kLDS + b * 8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    cp_async_wait<0>();
    // epilogue: C -= acc
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int gm = tm0 + wm * 64 + a * 8 + mq;
        if (gm >= mEnd) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gn = tn0 + wn * 32 + b * 8 + 2 * kq;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int g = gn + e;
                if (g >= nEnd) continue;
                if (KIND == 2) {
                    if (g > gm) continue;
                    const double val = C[(size_t)gm * ldc + g] - acc[a][b][e];
                    C[(size_t)gm * ldc + g] = val;
                    if (g != gm) C[(size_t)g * ldc + gm] = val;
                } else {
                    C[(size_t)gm * ldc + g] -= acc[a][b][e];
                }
            }
        }
    }
}

// =============================================================================================
// Fast path of the factorisation (k small enough for a shared-memory slab, the normal case):
//   S-chain : right-looking blocked Cholesky of [S | nu] ONLY (k x (k+1)); small latency-bound kernels
//   k_invert_diag : inverses of the 64x64 diagonal blocks of U, all blocks in parallel
//   k_trsm_slab   : W^T = U^-T B, one CTA per slab of SW columns of B held entirely in shared memory,
//                   left-looking over row blocks, all products on the FP64 tensor pipe; also dx = W y
// B is read once and W^T written once; no kernel of the chain touches the n-wide part.
// =============================================================================================

// S-chain panel, step J: 256 threads.  Phase 1: the diagonal tile is eliminated with thread (i, h) owning
// row i, columns j = h (mod 4); the pivot row goes through a double-buffered shared row, one barrier
// and one reciprocal per pivot.  Phase 2: one column of [S(J, >J) | nu] per thread, register-resident
// forward substitution.  grid (ceil((k + 1 - Jr) / 256), F), Jr = min(J1, k).
__global__ void __launch_bounds__(256) k_schain_panel(DevView v, int J)
{
    __shared__ __align__(16) double rowbuf[2][kNB];
    __shared__ __align__(16) double Msm[kNB * kNB];
    __shared__ double dinvs[kNB], ysm[kNB];
    const int f = blockIdx.y;
    int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    const int J0 = J * kNB;
    if (J0 >= k) return;
    const int kb = min(kNB, k - J0);
    const int Jr = J0 + kb;
    const int tid = threadIdx.x;
    double* Srow = v.S + ((size_t)f * v.kmax + J0) * v.ldS;
    const double* Sd = Srow + J0;
    double* Ublk = v.Uinv + ((size_t)f * (v.kmax / kNB) + J) * kNB * kNB;
    for (int e = tid; e < kNB * kNB; e += blockDim.x) Msm[e] = 0.0;
    if (tid < kNB) ysm[tid] = 0.0;
    {
        const int i = tid >> 2, h = tid & 3;
        double a[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int j = 4 * q + h;
            double val = 0.0;
            if (i < kb && j < kb) { if (j >= i) val = Sd[(size_t)i * v.ldS + j]; }
            else if (i == j) val = 1.0;
            a[q] = val;
        }
        if (i == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) rowbuf[0][4 * q + h] = a[q];
        }
        for (int c = 0; c < kNB; ++c) {
            __syncthreads();
            const double* rb = rowbuf[c & 1];
            const double piv = rb[c];
            const double pinv = __drcp_rn(piv);
            if (i > c) {
                const double m = rb[i] * pinv;
                if (h == 0) Msm[c * kNB + i] = m;
#pragma unroll
                for (int q = 0; q < 16; ++q) a[q] -= m * rb[4 * q + h];
                if (i == c + 1) {
                    double* nb = rowbuf[(c + 1) & 1];
#pragma unroll
                    for (int q = 0; q < 16; ++q) nb[4 * q + h] = a[q];
                }
            }
            // off the critical path: 1/sqrt(pivot) and the U row for the write-back
            if (tid < kNB) {
                const double dinv = rsqrt(piv);
                if (tid == c) {
                    if (!(piv > 0.0)) dm[D_STATUS] = 4;  // EKFB_ERR_NUMERIC
                    dinvs[c] = dinv;
                }
                // U_JJ row c goes to the Uinv slot of this block (NOT in place: the other CTAs of this launch
                // are still reading the original tile); k_invert_diag inverts it there
                if (blockIdx.x == 0) Ublk[c * kNB + tid] = (tid >= c) ? rb[tid] * dinv : 0.0;
            }
        }
        __syncthreads();
    }
    const int col = Jr + blockIdx.x * blockDim.x + tid;
    if (col > k) return;  // columns Jr .. k (column k = nu)
    panel_substitute(Srow + col, v.ldS, kb, Msm, dinvs, ysm);
}

// S-chain trailing update, step J (J1 < k):  S[I, c] -= X_J[:, I]^T X_J[:, c] for rows I >= J1 and columns
// c in [I, k] (column k = nu).  64x64 tiles, 4 warps of 32x32, K = 64 loaded in one shot.
// grid (ceil((k + 1 - J1) / 64), ceil((k - J1) / 64), F), dynamic smem 2 * 64 * 68 doubles.
constexpr int kSTrailSmem = 2 * kNB * 68 * (int)sizeof(double);

__global__ void __launch_bounds__(128) k_schain_trail(DevView v, int J)
{
    extern __shared__ __align__(16) double ssm[];
    double* As = ssm;
    double* Bs = ssm + kNB * 68;
    const int f = blockIdx.z;
    const int k = 2 * fdims(v, f)[D_ULIST];
    const int J0 = J * kNB, J1 = J0 + kNB;
    if (J1 >= k) return;
    const int tm0 = J1 + blockIdx.y * 64, tn0 = J1 + blockIdx.x * 64;
    if (tm0 >= k || tn0 > k) return;
    if (tn0 + 64 <= tm0) return;  // strictly below the diagonal
    const double* X = v.S + ((size_t)f * v.kmax + J0) * v.ldS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool same = (tm0 == tn0);
    for (int e = tid; e < kNB * 32; e += blockDim.x) {  // 64 rows x 32 double2 per operand
        const int r = e >> 5, c2 = (e & 31) * 2;
        if (tm0 + c2 < v.ldS) cp_async16(As + r * 68 + c2, X + (size_t)r * v.ldS + tm0 + c2);
        else { As[r * 68 + c2] = 0.0; As[r * 68 + c2 + 1] = 0.0; }
        if (!same) {
            if (tn0 + c2 < v.ldS) cp_async16(Bs + r * 68 + c2, X + (size_t)r * v.ldS + tn0 + c2);
            else { Bs[r * 68 + c2] = 0.0; Bs[r * 68 + c2 + 1] = 0.0; }
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const double* Bsrc = same ? As : Bs;
    const int g = lane >> 2, q = lane & 3;
    const int wm = w >> 1, wn = w & 1;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll 4
    for (int k4 = 0; k4 < kNB; k4 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = As[(k4 + q) * 68 + wm * 32 + a * 8 + g];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = Bsrc[(k4 + q) * 68 + wn * 32 + b * 8 + g];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    double* C = v.S + (size_t)f * v.kmax * v.ldS;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gm = tm0 + wm * 32 + a * 8 + g;
        if (gm >= k) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = tn0 + wn * 32 + b * 8 + 2 * q + e;
                if (gn <= k) C[(size_t)gm * v.ldS + gn] -= acc[a][b][e];
            }
    }
}

// Inverse of every 64x64 diagonal block U_JJ (upper triangular; identity padding of a partial last
// block) into Uinv[f][J][64][64].  4 lanes per column split the dot products.  grid (steps, F), 256 threads.
constexpr int kInvSmem = 2 * kNB * (kNB + 1) * (int)sizeof(double);

__global__ void __launch_bounds__(256) k_invert_diag(DevView v)
{
    extern __shared__ __align__(16) double ism[];
    double (*U)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(ism);
    double (*Ui)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(ism + kNB * (kNB + 1));
    const int f = blockIdx.y, J = blockIdx.x;
    const int k = 2 * fdims(v, f)[D_ULIST];
    const int J0 = J * kNB;
    if (J0 >= k) return;
    const int kb = min(kNB, k - J0);
    double* blk = v.Uinv + ((size_t)f * (v.kmax / kNB) + J) * kNB * kNB;  // holds U_JJ (identity padded)
    (void)kb;
    for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
        const int i = e / kNB, j = e % kNB;
        U[i][j] = (j >= i) ? blk[e] : 0.0;
        Ui[i][j] = 0.0;
    }
    __syncthreads();
    const int t = threadIdx.x >> 2, l = threadIdx.x & 3;  // column t, 4 lanes
    if (l == 0) Ui[t][t] = 1.0 / U[t][t];
    __syncwarp();
    for (int i = kNB - 2; i >= 0; --i) {  // warp-uniform trip count: the shuffles need every lane
        double s = 0.;
        if (i < t)
            for (int qq = i + 1 + l; qq <= t; qq += 4) s += U[i][qq] * Ui[qq][t];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (l == 0 && i < t) Ui[i][t] = -s / U[i][i];
        __syncwarp();
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) blk[e] = Ui[e / kNB][e % kNB];
}

// W^T = U^-T B on the tensor pipe.  CTA = slab of SW columns of B kept in shared memory (K-major,
// pitch SW + 4).  For row block J (left-looking):
//     T   = B_J - sum_{r < J0} U[r][J0 + m] * X[r][c]        (A chunks: 32 rows x 64 cols of U)
//     X_J = Uinv_J^T T                                       (A chunks: the two halves of Uinv_J)
// All A chunks of all row blocks form one stream that is double-buffered with cp.async one chunk ahead,
// so L2 latency is hidden across block boundaries.  8 warps: warp w owns rows 8w..8w+7 of the block.
// Finally dx[c] = sum_r X[r][c] y[r] (y = column k of the factored S) and W^T goes back to global.
// grid (ceil(n / SW), F), 256 threads, dynamic smem trsm_smem_bytes(k, SW).
template <int SW>
__host__ __device__ constexpr int trsm_pitch() { return SW + 4; }

inline size_t trsm_smem_bytes(int k, int SW)
{
    const int kpad = (k + kNB - 1) / kNB * kNB;
    return sizeof(double) * ((size_t)kpad * (SW + 4) + 2 * 32 * 68 + (size_t)kNB * (SW + 4) + 8 * SW);
}

template <int SW>
__global__ void __launch_bounds__(256, 1) k_trsm_slab(DevView v)
{
    constexpr int SWP = SW + 4, NT = SW / 8;
    extern __shared__ __align__(16) double tsm[];
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (k == 0) return;
    const int c0 = blockIdx.x * SW;
    if (c0 >= n) return;
    const int kpad = (k + kNB - 1) / kNB * kNB, steps = kpad / kNB;
    double* Xs = tsm;
    double* Us = Xs + (size_t)kpad * SWP;   // 2 stages x [32][68]
    double* Ts = Us + 2 * 32 * 68;          // [64][SWP]
    double* red = Ts + kNB * SWP;           // [8][SW]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* Bg = v.Bu + (size_t)f * v.kmax * v.ld;
    const double* Sg = v.S + (size_t)f * v.kmax * v.ldS;
    const double* Uinv = v.Uinv + (size_t)f * (v.kmax / kNB) * kNB * kNB;

    for (int e = tid; e < kpad * SW; e += blockDim.x) {
        const int r = e / SW, c = e % SW;
        Xs[(size_t)r * SWP + c] = (r < k && c0 + c < n) ? Bg[(size_t)r * v.ld + c0 + c] : 0.0;
    }
    // chunk stream: for block J: J0/32 chunks of U, then 2 chunks of Uinv_J
    auto issue = [&](int J, int ch, int stage) {
        const int J0 = J * kNB, nU = J0 / 32;
        double* dst = Us + stage * 32 * 68;
        const double* src;
        int lds;
        if (ch < nU) { src = Sg + (size_t)(ch * 32) * v.ldS + J0; lds = v.ldS; }
        else { src = Uinv + (size_t)J * kNB * kNB + (size_t)(ch - nU) * 32 * kNB; lds = kNB; }
        for (int e = tid; e < 32 * 32; e += blockDim.x) {
            const int r = e >> 5, c2 = (e & 31) * 2;
            cp_async16(dst + r * 68 + c2, src + (size_t)r * lds + c2);
        }
    };
    int J = 0, ch = 0, stage = 0;
    issue(0, 0, 0);
    cp_async_commit();
    double acc[NT][2];
    while (J < steps) {
        const int J0 = J * kNB, nU = J0 / 32, nCh = nU + 2;
        const int kb = min(kNB, k - J0);
        if (ch == 0) {
#pragma unroll
            for (int b = 0; b < NT; ++b) acc[b][0] = acc[b][1] = 0.0;
        }
        // prefetch the next chunk of the stream
        int Jn = J, chn = ch + 1;
        if (chn == nCh) { Jn = J + 1; chn = 0; }
        if (Jn < steps) issue(Jn, chn, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* Uc = Us + stage * 32 * 68;
        if (ch == nU) {
            // T = B_J - acc -> Ts (rows beyond the block are zero), then restart the accumulator
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = 8 * w + g, c = 8 * b + 2 * q + e;
                    Ts[m * SWP + c] = (m < kb) ? Xs[(size_t)(J0 + m) * SWP + c] - acc[b][e] : 0.0;
                    acc[b][e] = 0.0;
                }
            __syncthreads();
        }
        const double* Bsrc = (ch < nU) ? Xs + (size_t)(ch * 32) * SWP : Ts + (size_t)((ch - nU) * 32) * SWP;
#pragma unroll
        for (int k4 = 0; k4 < 32; k4 += 4) {
            const double a = Uc[(k4 + q) * 68 + 8 * w + g];
#pragma unroll
            for (int b = 0; b < NT; ++b) dmma8x8x4(acc[b][0], acc[b][1], a, Bsrc[(size_t)(k4 + q) * SWP + 8 * b + g]);
        }
        if (ch == nCh - 1) {
            // X_J complete: store into the slab
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = 8 * w + g, c = 8 * b + 2 * q + e;
                    Xs[(size_t)(J0 + m) * SWP + c] = (m < kb) ? acc[b][e] : 0.0;
                }
        }
        __syncthreads();  // stage buffer and Ts / Xs hazards before the next chunk
        stage ^= 1;
        J = Jn;
        ch = chn;
    }
    cp_async_wait<0>();
    // W^T back to global, dx = W y
    for (int e = tid; e < k * SW; e += blockDim.x) {
        const int r = e / SW, c = e % SW;
        if (c0 + c < n) Bg[(size_t)r * v.ld + c0 + c] = Xs[(size_t)r * SWP + c];
    }
    if (lane < SW) {
        double s = 0.;
        for (int r = w; r < k; r += 8) s += Xs[(size_t)r * SWP + lane] * Sg[(size_t)r * v.ldS + k];
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

// ---------------------------------------------------------------------------------------------
// U2: x += deadband(dx) (E/Update.cpp:143-204); dx = W y was accumulated by the panel kernels.
// grid (ceil(n/256), F)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_state_apply(DevView v)
{
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    if (dm[D_ULIST] == 0) return;
    const int n = dm[D_N_STATE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = v.dx[(size_t)f * v.ld + i];
    if (fabs(d) > kDelta) v.x[(size_t)f * v.ld + i] += d;
}

// U4 (a): J = d(q/|q|)/dq at the un-normalised q, then q <- q/|q| (E/Update.cpp:45-60,309-317)
__global__ void k_quat_norm(DevView v)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= v.F) return;
    if (fdims(v, f)[D_ULIST] == 0) return;
    double* x = v.x + (size_t)f * v.ld;
    const double r = x[3], a = x[4], b = x[5], c = x[6];
    const double nrm = sqrt(r * r + a * a + b * b + c * c);
    const double s = 1.0 / (nrm * nrm * nrm);
    const double M[16] = {a * a + b * b + c * c, -r * a, -r * b, -r * c, -a * r, r * r + b * b + c * c, -a * b, -a * c,
                          -b * r, -b * a, r * r + a * a + c * c, -b * c, -c * r, -c * a, -c * b, r * r + a * a + b * b};
    double* Jq = v.Jq + (size_t)f * 16;
    for (int e = 0; e < 16; ++e) Jq[e] = M[e] * s;
    x[3] = r / nrm; x[4] = a / nrm; x[5] = b / nrm; x[6] = c / nrm;
}

// U4 (b): P <- T P T^T with T = diag(I3, J, I) (normalizeCovariance, E/Update.cpp:64-85): only rows
// and columns 3..6 change.  grid (ceil(n/256), F)
__global__ void __launch_bounds__(256) k_quat_cov(DevView v)
{
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    if (dm[D_ULIST] == 0) return;
    const int n = dm[D_N_STATE];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* P = v.P + (size_t)f * v.nmax * v.ld;
    double Jm[16];
    for (int e = 0; e < 16; ++e) Jm[e] = v.Jq[(size_t)f * 16 + e];
    if (j >= 3 && j < 7) {
        if (j != 3) return;
        double B4[16], T4[16];
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) B4[a * 4 + b] = P[(size_t)(3 + a) * v.ld + 3 + b];
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) {
                double s = 0.;
                for (int c = 0; c < 4; ++c) s += Jm[a * 4 + c] * B4[c * 4 + b];
                T4[a * 4 + b] = s;
            }
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b <= a; ++b) {
                double s = 0.;
                for (int c = 0; c < 4; ++c) s += T4[a * 4 + c] * Jm[b * 4 + c];
                P[(size_t)(3 + a) * v.ld + 3 + b] = s;
                P[(size_t)(3 + b) * v.ld + 3 + a] = s;
            }
        return;
    }
    double col[4];
    for (int a = 0; a < 4; ++a) col[a] = P[(size_t)(3 + a) * v.ld + j];
    for (int a = 0; a < 4; ++a) {
        double s = 0.;
        for (int c = 0; c < 4; ++c) s += Jm[a * 4 + c] * col[c];
        P[(size_t)(3 + a) * v.ld + j] = s;
        P[(size_t)j * v.ld + 3 + a] = s;
    }
}

}  // namespace ekf
