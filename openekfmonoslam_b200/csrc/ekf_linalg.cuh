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
    grid_dependency_wait();   // launched with programmatic stream serialisation: nothing of the predecessor is read before this
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();
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
        double pc[7], pf[6];  // all row reads in flight at once; the XYZ case masks the last three feature rows
#pragma unroll
        for (int c = 0; c < 7; ++c) pc[c] = P[(size_t)c * v.ld + i];
#pragma unroll
        for (int c = 0; c < 6; ++c) pf[c] = c < d ? P[(size_t)(off + c) * v.ld + i] : 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            b0 += sHx[c] * pc[c];
            b1 += sHx[7 + c] * pc[c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c)
            if (c < d) {
                b0 += sHf[c] * pf[c];
                b1 += sHf[6 + c] * pf[c];
            }
    } else if (i == n) {
        b0 = deadband(v.z[(fo + j) * 2] - s.h[(fo + j) * 2]);
        b1 = deadband(v.z[(fo + j) * 2 + 1] - s.h[(fo + j) * 2 + 1]);
    }
    B0[i] = b0;
    B0[v.ld + i] = b1;
    if (a == 0) v.dx[(size_t)f * v.ld + i] = 0.0;  // reset the state-correction accumulator
}

// S = H B^T + sigma I, sparse in H again (E/Update.cpp:95-107); only the 32x32 tiles on or above the diagonal are formed
// (the factorisation reads the upper triangle).  CTA tile: 32 rows ra (16 features) x 32 columns rb.  The 7 camera columns
// and the 16 x 6 feature columns of the 32 rows rb of B are staged in shared memory with coalesced reads, the products run
// out of shared memory and S is written with the lanes along rb.  grid (ceil(k/32), ceil(k/32), F), block (32, 8).
__global__ void __launch_bounds__(256) k_build_S(DevView v, int which)
{
    grid_dependency_wait();   // launched with programmatic stream serialisation: nothing of the predecessor is read before this
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();
    constexpr int BP = 105;  // 7 + 16 * 6 columns, odd pitch
    __shared__ double Bs[32 * BP];
    __shared__ double Hs[32 * 13];
    __shared__ int sOff[16], sD[16], sJ[16];
    const int f = blockIdx.z;
    const int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    if (blockIdx.x < blockIdx.y) return;  // strictly below the diagonal
    const int ra0 = blockIdx.y * 32, rb0 = blockIdx.x * 32;
    if (ra0 >= k || rb0 >= k) return;
    const size_t fo = (size_t)f * v.Nmax;
    const UpdSrc s = upd_src(v, which);
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid < 16) {
        const int a = (ra0 >> 1) + tid;
        const int j = (2 * a < k) ? v.ulist[fo + a] : -1;
        sJ[tid] = j;
        sOff[tid] = j >= 0 ? v.foff[fo + j] : 0;
        sD[tid] = j >= 0 ? (v.ftype[fo + j] == kTypeInvDepth ? 6 : 3) : 0;
    }
    __syncthreads();
    for (int e = tid; e < 32 * 13; e += 256) {
        const int rl = e / 13, c = e % 13, j = sJ[rl >> 1], r = rl & 1;
        double h = 0.0;
        if (j >= 0) h = c < 7 ? s.Hx[(fo + j) * 14 + 7 * r + c] : s.Hf[(fo + j) * 12 + 6 * r + (c - 7)];
        Hs[e] = h;
    }
    const double* Bg = v.Bu + (size_t)f * v.kmax * v.ld;
#pragma unroll
    for (int it = 0; it < (32 * 103 + 255) / 256; ++it) {  // unrolled: all 13 reads of a thread in flight
        const int e = tid + 256 * it;
        if (e >= 32 * 103) break;
        const int rl = e / 103, cc = e % 103;
        double val = 0.0;
        if (rb0 + rl < k) {
            if (cc < 7) val = Bg[(size_t)(rb0 + rl) * v.ld + cc];
            else {
                const int fa = (cc - 7) / 6, c = (cc - 7) % 6;
                if (c < sD[fa]) val = Bg[(size_t)(rb0 + rl) * v.ld + sOff[fa] + c];
            }
        }
        Bs[rl * BP + cc] = val;
    }
    __syncthreads();
    const int rb = rb0 + threadIdx.x;
    const double* Brow = Bs + threadIdx.x * BP;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        const int rl = threadIdx.y * 4 + qq, ra = ra0 + rl;
        if (ra >= k) continue;
        const int fa = rl >> 1, d = sD[fa];
        const double* H = Hs + rl * 13;
        double acc = 0.;
#pragma unroll
        for (int c = 0; c < 7; ++c) acc += H[c] * Brow[c];
        for (int c = 0; c < d; ++c) acc += H[7 + c] * Brow[7 + 6 * fa + c];
        if (rb < k) {
            if (ra == rb) acc += v.sigma_px;
            v.S[((size_t)f * v.kmax + ra) * v.ldS + rb] = acc;
        }
        if (blockIdx.x == blockIdx.y && threadIdx.x == 0) {  // column k of S carries the dead-banded innovation nu (-> y = U^-T nu)
            const int j = sJ[fa], r = rl & 1;
            v.S[((size_t)f * v.kmax + ra) * v.ldS + k] = deadband(v.z[(fo + j) * 2 + r] - s.h[(fo + j) * 2 + r]);
        }
    }
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
//   KIND 3: same as KIND 2 without the register cap (batched filters: no co-resident remainder kernel)
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

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still running; it must not touch the predecessor's output before
// grid_dependency_wait().  grid_launch_dependents() lets the successor be scheduled early.
// (grid_dependency_wait / grid_launch_dependents are defined in ekf_kernels.cuh)

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
__global__ void __maxnreg__(KIND == 2 ? 200 : 255) k_gemm_tn(DevView v, int J)
{
    extern __shared__ __align__(16) double gsm[];
    const int f = blockIdx.z;
    const int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (KIND >= 2 && dm[D_STATUS] != 0) return;   // a failed factorisation leaves P untouched
    const double *A, *B;
    double* C;
    int lda, ldb, ldc, mBeg, mEnd, nBeg, nEnd, K, aLim, bLim;
    int tileN = blockIdx.x;
    bool sPart = false;
    if ((KIND >= 2)) {
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
    int tileM = blockIdx.y;
    if ((KIND >= 2)) {
        // 1-D grid over the lower-triangular 128x128 tiles, row block by row block; tiles with index >= J are
        // left to k_downdate64 (wave-quantisation remainder, see run_update)
        const int idx = blockIdx.x;
        if (idx >= J) return;
        int I = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
        while (I * (I + 1) / 2 > idx) --I;
        while ((I + 1) * (I + 2) / 2 <= idx) ++I;
        tileM = I;
        tileN = idx - I * (I + 1) / 2;
    }
    const int tm0 = mBeg + tileM * kTM, tn0 = nBeg + tileN * kTN;
    if (tm0 >= mEnd || tn0 >= nEnd) return;
    if (sPart && tn0 + kTN <= tm0) return;      // S tile strictly below the diagonal
    if ((KIND >= 2) && tn0 > tm0) return;         // lower tiles only

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    double acc[8][4][2];
    const int kq = lane & 3, mq = lane >> 2;

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
    // The accumulators start at -C: all 64 loads of the thread's part of the C tile are issued here, independent of
    // each other, and are in flight together with the operand ring's prologue; the epilogue is store-only
    // (C_new = -(-C + sum)).  A load -> subtract -> store epilogue serialises 64 memory round trips per thread because
    // the stores to C order the following loads from C.
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int gm = tm0 + wm * 64 + a * 8 + mq;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gn = tn0 + wn * 32 + b * 8 + 2 * kq;
            double c0 = 0.0, c1 = 0.0;
            if ((KIND >= 2) && gm < mEnd) {   // KIND 0 (generic factorisation path) keeps the in-place epilogue
                const double* cp = C + (size_t)gm * ldc + gn;
                if (gn + 1 < nEnd) {
                    const double2 t = *reinterpret_cast<const double2*>(cp);
                    c0 = t.x; c1 = t.y;
                } else if (gn < nEnd) {
                    c0 = *cp;
                }
            }
            acc[a][b][0] = -c0;
            acc[a][b][1] = -c1;
        }
    }
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
                if ((KIND >= 2)) {
                    if (g > gm) continue;
                    const double val = -acc[a][b][e];
                    C[(size_t)gm * ldc + g] = val;
                    if (g != gm) C[(size_t)g * ldc + gm] = val;
                } else {
                    C[(size_t)gm * ldc + g] -= acc[a][b][e];
                }
            }
        }
    }
}

// The wave-quantisation remainder of the covariance downdate: when the number T of 128x128 lower tiles is
// slightly above a multiple of the SM count (n = 3013: T = 300 = 2 * 148 + 4), the last T mod 148 tiles would
// cost a whole extra wave.  They are cut into 64x64 tiles and run by this small-footprint kernel (<= 112
// registers, 52 KB smem) on a second stream, co-resident with the big CTAs.  Big tile `idx` -> 4 small tiles.
constexpr int kSmallSmemBytes = kStages * kKC * (68 + 68) * (int)sizeof(double);

__global__ void __maxnreg__(112) k_downdate64(DevView v, int firstBig)
{
    grid_dependency_wait();   // launched with programmatic stream serialisation: nothing of the predecessor is read before this
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();
    extern __shared__ __align__(16) double ssm2[];
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    const int K = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (K == 0 || dm[D_STATUS] != 0) return;   // a failed factorisation (status != 0) leaves P untouched
    const int idx = firstBig + (blockIdx.x >> 2), sub = blockIdx.x & 3;
    int I = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while (I * (I + 1) / 2 > idx) --I;
    while ((I + 1) * (I + 2) / 2 <= idx) ++I;
    const int Jt = idx - I * (I + 1) / 2;
    const int tm0 = I * kTM + (sub >> 1) * 64, tn0 = Jt * kTN + (sub & 1) * 64;
    if (tm0 >= n || tn0 >= n || tn0 > tm0) return;
    const double* W = v.Bu + (size_t)f * v.kmax * v.ld;
    double* P = v.P + (size_t)f * v.nmax * v.ld;
    const int ld = v.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
    // warp tiles (32 x 32) that hold no element of the lower triangle inside the matrix do no arithmetic: the warp above the
    // diagonal of a diagonal tile, and the warps beyond the ragged last block row / column of P
    const bool warpIdle = (tm0 == tn0 && wn > wm) || (tm0 + wm * 32 >= n) || (tn0 + wn * 32 >= n);
    const int nk = (K + kKC - 1) / kKC;
    auto stA = [&](int st) { return ssm2 + (size_t)st * kKC * 136; };
    auto stB = [&](int st) { return ssm2 + (size_t)st * kKC * 136 + kKC * 68; };
    auto load = [&](int st, int kc) {
#pragma unroll
        for (int it = 0; it < (kKC * 32) / 128; ++it) {
            const int c = tid + it * 128;
            const int r = c >> 5, c2 = (c & 31) * 2;
            const bool rowOk = kc * kKC + r < K;
            double* da = stA(st) + r * 68 + c2;
            double* db = stB(st) + r * 68 + c2;
            if (rowOk && tm0 + c2 < ld) cp_async16(da, W + (size_t)(kc * kKC + r) * ld + tm0 + c2);
            else { da[0] = 0.0; da[1] = 0.0; }
            if (rowOk && tn0 + c2 < ld) cp_async16(db, W + (size_t)(kc * kKC + r) * ld + tn0 + c2);
            else { db[0] = 0.0; db[1] = 0.0; }
        }
    };
#pragma unroll
    for (int st = 0; st < kStages - 1; ++st) {
        if (st < nk) load(st, st);
        cp_async_commit();
    }
    // accumulators start at -P (see k_gemm_tn): the tile's loads overlap the ring prologue, the epilogue only stores
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gm = tm0 + wm * 32 + a * 8 + g;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gn = tn0 + wn * 32 + b * 8 + 2 * q;
            double c0 = 0.0, c1 = 0.0;
            if (gm < n) {
                const double* cp = P + (size_t)gm * ld + gn;
                if (gn + 1 < n) {
                    const double2 t = *reinterpret_cast<const double2*>(cp);
                    c0 = t.x; c1 = t.y;
                } else if (gn < n) {
                    c0 = *cp;
                }
            }
            acc[a][b][0] = -c0;
            acc[a][b][1] = -c1;
        }
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        if (kt + kStages - 1 < nk) load((kt + kStages - 1) % kStages, kt + kStages - 1);
        cp_async_commit();
        const double* As = stA(kt % kStages) + wm * 32 + g;
        const double* Bs = stB(kt % kStages) + wn * 32 + g;
        if (warpIdle) continue;
#pragma unroll
        for (int k4 = 0; k4 < kKC; k4 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[(k4 + q) * 68 + a * 8];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[(k4 + q) * 68 + b * 8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gm = tm0 + wm * 32 + a * 8 + g;
        if (gm >= n) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gn0 = tn0 + wn * 32 + b * 8 + 2 * q;
            if (gn0 + 1 < gm && gn0 + 1 < n) {
                // both elements strictly below the diagonal: one 16-byte store (whole 32-byte sectors per lane pair), two mirrors
                const double v0 = -acc[a][b][0], v1 = -acc[a][b][1];
                *reinterpret_cast<double2*>(P + (size_t)gm * ld + gn0) = make_double2(v0, v1);
                P[(size_t)gn0 * ld + gm] = v0;
                P[(size_t)(gn0 + 1) * ld + gm] = v1;
                continue;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = gn0 + e;
                if (gn > gm || gn >= n) continue;
                const double val = -acc[a][b][e];
                P[(size_t)gm * ld + gn] = val;
                if (gn != gm) P[(size_t)gn * ld + gm] = val;
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

// Register-resident forward substitution of one column against the pivot rows kept in shared memory:
//   x[i] -= U(c)[c][i] * (x[c] / piv_c)  for i > c,  then  x[c] *= 1/sqrt(piv_c).
__device__ __noinline__ void schain_substitute(double* colp, int ldx, int kb, const double* Urows, const double* pinvs,
                                               const double* dinvs)
{
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) x[i] = (i < kb) ? colp[(size_t)i * ldx] : 0.0;
    const double2* U2 = reinterpret_cast<const double2*>(Urows);
#pragma unroll
    for (int c = 0; c < kNB - 1; ++c) {
        const double xs = x[c] * pinvs[c];
#pragma unroll
        for (int p = c / 2; p < kNB / 2; ++p) {
            const double2 u = U2[c * (kNB / 2) + p];
            if (2 * p > c) x[2 * p] -= u.x * xs;
            if (2 * p + 1 > c) x[2 * p + 1] -= u.y * xs;
        }
    }
#pragma unroll
    for (int i = 0; i < kNB; ++i)
        if (i < kb) colp[(size_t)i * ldx] = x[i] * dinvs[i];
}

// S-chain panel, step J: 128 threads.  Phase 1 eliminates the 64x64 diagonal tile TWO pivots per barrier (the
// natural 2x2 blocks of the innovation covariance: one feature = two measurement rows).  Thread (ri, cj) owns
// rows 4ri..4ri+3 x columns 8cj..8cj+7 in registers (2-D blocking: 12 loaded doubles per 32 FMAs -- the
// one-row-per-thread variants were bound by shared-memory wavefronts).  Per step the two raw pivot rows
// (eliminated by all earlier pivots, not by each other) are published once into their own shared-memory rows;
// every thread redoes the 2x2 pivot algebra (two chained reciprocals) and applies a rank-2 update.  Measured on
// B200 the publish -> barrier -> load round trip costs ~450 cycles, a reciprocal chain ~100: pairing pivots
// halves the former.  1/sqrt and write-backs happen after the loop.  CTA 0 also carries an identity block through
// the same row operations (columns <= c only, complementary to the U part), which yields L^-1 and hence
// Uinv_J = U_JJ^-1 for the slab TRSM.  Phase 2: 32 threads own one column of [S(J, >J) | nu] each,
// register-resident forward substitution.  grid (ceil((k + 1 - Jr) / 32), F); dynamic smem kSPanelSmem.
constexpr int kSPanelSmem = (4 * kNB * kNB + 3 * kNB) * (int)sizeof(double);
constexpr int kSPanelCols = 32;

__global__ void __launch_bounds__(128) k_schain_panel(DevView v, int J, int colsPerCta)
{
    extern __shared__ __align__(16) double psm[];
    double* Rraw = psm;                    // [64][64] published raw pivot rows
    double* Ufin = psm + kNB * kNB;        // [64][64] final pivot rows (row c valid for j >= c)
    double* Eraw = Ufin + kNB * kNB;       // identity part, raw / final (CTA 0 only)
    double* Efin = Eraw + kNB * kNB;
    double* pinvs = Efin + kNB * kNB;
    double* pivs = pinvs + kNB;
    double* dinvs = pivs + kNB;
    const int f = blockIdx.y;
    int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    const int J0 = J * kNB;
    if (J0 >= k) return;
    const int kb = min(kNB, k - J0);
    const int Jr = J0 + kb;
    const int tid = threadIdx.x;
    const bool lead = (blockIdx.x == 0);
    double* Srow = v.S + ((size_t)f * v.kmax + J0) * v.ldS;
    const double* Sd = Srow + J0;
    const bool dbgT = (v.dbg != nullptr) && lead && tid == 0 && f == 0 && J == 0;
    if (dbgT) v.dbg[0] = clock64();
    {
        const int ri = tid >> 3, cj = tid & 7;
        const int r0 = 4 * ri, c0 = 8 * cj;
        double a[4][8], e[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = r0 + r, j = c0 + q;
                double val = 0.0;
                if (i < kb && j < kb) { if (j >= i) val = Sd[(size_t)i * v.ldS + j]; }
                else if (i == j) val = 1.0;
                a[r][q] = val;
                e[r][q] = (i == j) ? 1.0 : 0.0;
            }
        if (ri == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                Rraw[c0 + q] = a[0][q];
                Rraw[kNB + c0 + q] = a[1][q];
                if (lead) { Eraw[c0 + q] = e[0][q]; Eraw[kNB + c0 + q] = e[1][q]; }
            }
        }
        if (dbgT) v.dbg[1] = clock64();
        for (int c = 0; c < kNB; c += 2) {
            __syncthreads();
            const double* R0 = Rraw + c * kNB;
            const double* R1 = R0 + kNB;
            const double2 d0 = *reinterpret_cast<const double2*>(R0 + c);   // A[c][c], A[c][c+1]
            const double d11 = R1[c + 1];
            const double pinv0 = __drcp_rn(d0.x);
            const double l = d0.y * pinv0;
            const double piv1 = d11 - l * d0.y;
            const double pinv1 = __drcp_rn(piv1);
            if (tid == 0) { pinvs[c] = pinv0; pivs[c] = d0.x; pinvs[c + 1] = pinv1; pivs[c + 1] = piv1; }
            if (r0 + 3 > c + 1 || ri == 15) {
                // multipliers of this thread's rows (rows <= c+1 get zero)
                double m0[4], m1[4];
                {
                    const double2 g01 = *reinterpret_cast<const double2*>(R0 + r0);
                    const double2 g23 = *reinterpret_cast<const double2*>(R0 + r0 + 2);
                    const double2 h01 = *reinterpret_cast<const double2*>(R1 + r0);
                    const double2 h23 = *reinterpret_cast<const double2*>(R1 + r0 + 2);
                    m0[0] = g01.x * pinv0; m1[0] = (h01.x - m0[0] * d0.y) * pinv1;
                    m0[1] = g01.y * pinv0; m1[1] = (h01.y - m0[1] * d0.y) * pinv1;
                    m0[2] = g23.x * pinv0; m1[2] = (h23.x - m0[2] * d0.y) * pinv1;
                    m0[3] = g23.y * pinv0; m1[3] = (h23.y - m0[3] * d0.y) * pinv1;
                    if (!(r0 > c + 1)) { m0[0] = 0.0; m1[0] = 0.0; }
                    if (!(r0 + 1 > c + 1)) { m0[1] = 0.0; m1[1] = 0.0; }
                    if (!(r0 + 2 > c + 1)) { m0[2] = 0.0; m1[2] = 0.0; }
                    if (!(r0 + 3 > c + 1)) { m0[3] = 0.0; m1[3] = 0.0; }
                }
                if (c0 + 7 >= c) {  // U part
                    double u0[8], u1[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 t0 = *reinterpret_cast<const double2*>(R0 + c0 + 2 * q);
                        const double2 t1 = *reinterpret_cast<const double2*>(R1 + c0 + 2 * q);
                        u0[2 * q] = t0.x; u0[2 * q + 1] = t0.y;
                        u1[2 * q] = t1.x - l * t0.x; u1[2 * q + 1] = t1.y - l * t0.y;   // row c+1 eliminated by c
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int q = 0; q < 8; ++q) a[r][q] -= m0[r] * u0[q] + m1[r] * u1[q];
                    if (ri == 15) {  // rows 60..63 are active for every pivot: they record the final pivot rows
#pragma unroll
                        for (int q = 0; q < 8; ++q) { Ufin[c * kNB + c0 + q] = u0[q]; Ufin[(c + 1) * kNB + c0 + q] = u1[q]; }
                    }
                }
                if (lead && c0 <= c + 1) {  // identity part: rows c, c+1 are non-zero in columns <= c+1 only
                    const double* E0 = Eraw + c * kNB;
                    const double* E1 = E0 + kNB;
                    double v0[8], v1[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 t0 = *reinterpret_cast<const double2*>(E0 + c0 + 2 * q);
                        const double2 t1 = *reinterpret_cast<const double2*>(E1 + c0 + 2 * q);
                        v0[2 * q] = t0.x; v0[2 * q + 1] = t0.y;
                        v1[2 * q] = t1.x - l * t0.x; v1[2 * q + 1] = t1.y - l * t0.y;
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int q = 0; q < 8; ++q) e[r][q] -= m0[r] * v0[q] + m1[r] * v1[q];
                    if (ri == 15) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) { Efin[c * kNB + c0 + q] = v0[q]; Efin[(c + 1) * kNB + c0 + q] = v1[q]; }
                    }
                }
                if (c + 2 < kNB && ri == (c + 2) >> 2) {  // this row block holds rows c+2, c+3: publish them
                    const int rr = (c + 2) & 3;           // 0 or 2
                    double* n0 = Rraw + (c + 2) * kNB + c0;
                    double* q0 = Eraw + (c + 2) * kNB + c0;
#pragma unroll
                    for (int r = 0; r < 4; r += 2)
                        if (r == rr) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                n0[q] = a[r][q];
                                n0[kNB + q] = a[r + 1][q];
                                if (lead) { q0[q] = e[r][q]; q0[kNB + q] = e[r + 1][q]; }
                            }
                        }
                }
            }
        }
        __syncthreads();
        if (dbgT) v.dbg[2] = clock64();
        if (tid < kNB) {
            const double piv = pivs[tid];
            if (!(piv > 0.0)) dm[D_STATUS] = 4;  // EKFB_ERR_NUMERIC
            dinvs[tid] = rsqrt(piv);
        }
        __syncthreads();
        if (lead) {  // Uinv[s][m] = Linv[m][s] / sqrt(piv_m)  (upper triangular)
            double* Ublk = v.Uinv + ((size_t)f * (v.kmax / kNB) + J) * kNB * kNB;
            for (int idx = tid; idx < kNB * kNB; idx += blockDim.x) {
                const int sI = idx / kNB, m = idx % kNB;
                Ublk[idx] = (sI <= m) ? Efin[m * kNB + sI] * dinvs[m] : 0.0;
            }
        }
    }
    if (dbgT) v.dbg[3] = clock64();
    // colsPerCta = 32 for a single filter (latency: spread the columns over many SMs), 128 for batched filters
    // (throughput: fewer redundant eliminations of the diagonal tile)
    if (tid >= colsPerCta) return;
    const int col = Jr + blockIdx.x * colsPerCta + tid;
    if (col > k) return;  // columns Jr .. k (column k = nu)
    schain_substitute(Srow + col, v.ldS, kb, Ufin, pinvs, dinvs);
    if (dbgT) v.dbg[4] = clock64();
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

// U2 + U4 (a) as the tail of the kernel that produced dx (run by its last block, see last_block_done): x += deadband(dx)
// (E/Update.cpp:143-204), then J = d(q/|q|)/dq at the un-normalised q and q <- q/|q| (E/Update.cpp:45-60,309-317).  The same
// arithmetic as k_state_apply, one block for the whole state.
__device__ inline void state_apply_tail(const DevView& v, int f)
{
    const int* dm = fdims(v, f);
    if (dm[D_ULIST] == 0 || __ldcg(dm + D_STATUS) != 0) return;
    const int n = dm[D_N_STATE];
    double* x = v.x + (size_t)f * v.ld;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = __ldcg(v.dx + (size_t)f * v.ld + i);   // (written by the other blocks of this launch: read through L2)
        if (fabs(d) > kDelta) x[i] += d;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double r = x[3], a = x[4], b = x[5], c = x[6];
    const double nrm = sqrt(r * r + a * a + b * b + c * c);
    const double s = 1.0 / (nrm * nrm * nrm);
    const double M[16] = {a * a + b * b + c * c, -r * a, -r * b, -r * c, -a * r, r * r + b * b + c * c, -a * b, -a * c,
                          -b * r, -b * a, r * r + a * a + c * c, -b * c, -c * r, -c * a, -c * b, r * r + a * a + b * b};
    double* Jq = v.Jq + (size_t)f * 16;
    for (int e = 0; e < 16; ++e) Jq[e] = M[e] * s;
    x[3] = r / nrm; x[4] = a / nrm; x[5] = b / nrm; x[6] = c / nrm;
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

inline size_t trsm_smem_bytes(int k, int SW, int stages = 2, int pad = 4)
{
    const int kpad = (k + kNB - 1) / kNB * kNB;
    return sizeof(double) * ((size_t)kpad * (SW + pad) + (size_t)stages * 32 * 68 + (size_t)kNB * (SW + pad) + 8 * SW);
}

// NS = depth of the cp.async ring of 32-row operand chunks.  A chunk feeds only 8 x SW/8 DMMAs per warp, far less than
// the L2 round trip of its load, so the kernel's time is (number of chunks) x (load latency / chunks in flight): NS - 1
// chunks are kept in flight, as many as the shared memory left beside the slab allows (run_update picks NS).
template <int SW, int NS>
__device__ __forceinline__ void trsm_slab_body(const DevView& v, double* tsm)
{
    constexpr int SWP = SW + 4, NT = SW / 8;
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (k == 0) return;
    const int c0 = blockIdx.x * SW;
    if (c0 >= n) return;
    const int kpad = (k + kNB - 1) / kNB * kNB, steps = kpad / kNB;
    double* Xs = tsm;
    double* Us = Xs + (size_t)kpad * SWP;   // NS stages x [32][68]
    double* Ts = Us + NS * 32 * 68;         // [64][SWP]
    double* red = Ts + kNB * SWP;           // [8][SW]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* Bg = v.Bu + (size_t)f * v.kmax * v.ld;
    const double* Sg = v.Sf + (size_t)f * v.kmax * v.ldS;   // the factor buffer of the S-chain
    const double* Uinv = v.Uinv + (size_t)f * (v.kmax / kNB) * kNB * kNB;

    for (int e = tid; e < kpad * SW; e += blockDim.x) {
        const int r = e / SW, c = e % SW;
        Xs[(size_t)r * SWP + c] = (r < k && c0 + c < n) ? Bg[(size_t)r * v.ld + c0 + c] : 0.0;
    }
    // The slab of B does not depend on the S-chain: with programmatic dependent launch this CTA was scheduled while the last
    // chain step was still running; from here on the factor is read.
    grid_dependency_wait();
    if (dm[D_STATUS] != 0) return;   // the innovation covariance was not positive definite: no W, the update is skipped
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
    int Ji = 0, chi = 0, stagei = 0;   // issue position: NS - 1 chunks ahead of the consume position (J, ch, stage)
    auto issue_next = [&]() {
        if (Ji < steps) {
            issue(Ji, chi, stagei);
            if (++chi == (Ji * kNB) / 32 + 2) { ++Ji; chi = 0; }
            stagei = (stagei + 1 == NS) ? 0 : stagei + 1;
        }
        cp_async_commit();   // one group per slot, empty past the end of the stream: keeps wait_group's count uniform
    };
#pragma unroll
    for (int p = 0; p < NS - 1; ++p) issue_next();
    double acc[NT][2];
    while (J < steps) {
        const int J0 = J * kNB, nU = J0 / 32, nCh = nU + 2;
        const int kb = min(kNB, k - J0);
        if (ch == 0) {
#pragma unroll
            for (int b = 0; b < NT; ++b) acc[b][0] = acc[b][1] = 0.0;
        }
        int Jn = J, chn = ch + 1;
        if (chn == nCh) { Jn = J + 1; chn = 0; }
        // prefetch: the slot being refilled was consumed in the previous iteration (barrier at its end)
        issue_next();
        cp_async_wait<NS - 1>();
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
        stage = (stage + 1 == NS) ? 0 : stage + 1;
        J = Jn;
        ch = chn;
    }
    cp_async_wait<0>();
    // W^T back to global (rows k .. end of the last 16-row chunk as zeros: the TMA-fed downdate reads whole chunks), dx = W y
    const int kz = min(kpad, (k + 15) & ~15);
    for (int e = tid; e < kz * SW; e += blockDim.x) {
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

// the filter's block that finishes last applies the state correction (state_apply_tail): no k_state_apply launch behind it
template <int SW, int NS>
__global__ void __launch_bounds__(256, 1) k_trsm_slab(DevView v)
{
    extern __shared__ __align__(16) double tsm[];
    trsm_slab_body<SW, NS>(v, tsm);
    if (last_block_done(fdims(v, blockIdx.y) + D_TICKET_UPD, (int)gridDim.x)) state_apply_tail(v, blockIdx.y);
}

// ---------------------------------------------------------------------------------------------
// U3 (+ symmetrise of U4): the covariance downdate  P -= W W^T  on the FP64 tensor pipe, version 2.
//   * work items are the 128 x 64 tiles of the LOWER triangle only, in a 1-D grid (row block I has 2I+2
//     column tiles), 4 warps (2 x 2, warp tile 64 x 32 = 8 x 4 DMMA m8n8k4), two CTAs per SM so that the
//     hardware scheduler balances the 600 tiles of n = 3013 over 296 slots;
//   * K-major operands through a 3-stage cp.async ring (16 rows per stage), pitch 132 / 68 doubles;
//   * m- and n-tiles beyond the matrix edge are skipped (the ragged last block row);
//   * epilogue: lower elements are read-modified-written in place; the mirrored (upper) elements get the
//     SAME values through a shared-memory transpose so that they are stored as contiguous rows (the direct
//     scattered mirror store of version 1 cost 4x write amplification) -- P stays exactly symmetric.
// Algorithmic work per launch: n (n + 1) k flop; minimum traffic 16 n^2 bytes.
// ---------------------------------------------------------------------------------------------
constexpr int kDTM = 128, kDTN = 64, kDLA = 132, kDLB = 68, kDTP = 130;
constexpr int kDownSmemBytes = kStages * kKC * (kDLA + kDLB) * (int)sizeof(double);  // 76.8 KB >= 64 * 130 * 8

template <int PITCH, int WIDTH, int NTHREADS>
__device__ __forceinline__ void load_slab_w(double* dst, const double* src, int ldsrc, int k0, int K, int col0,
                                            int colLimit, int tid)
{
    constexpr int CPR = WIDTH / 2;  // 16-byte chunks per row
#pragma unroll
    for (int it = 0; it < (kKC * CPR) / NTHREADS; ++it) {
        const int c = tid + it * NTHREADS;
        const int r = c / CPR, c2 = (c % CPR) * 2;
        double* d = dst + r * PITCH + c2;
        if (k0 + r < K && col0 + c2 < colLimit) cp_async16(d, src + (size_t)(k0 + r) * ldsrc + col0 + c2);
        else { d[0] = 0.0; d[1] = 0.0; }
    }
}

__global__ void __launch_bounds__(128, 2) k_downdate(DevView v)
{
    extern __shared__ __align__(16) double dsm2[];
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    const int K = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (K == 0 || dm[D_STATUS] != 0) return;
    // linear tile index -> (row block I, column tile j), j in [0, 2I+1]
    const int idx = blockIdx.x;
    int I = (int)((sqrtf(4.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while (I * (I + 1) > idx) --I;
    while ((I + 1) * (I + 2) <= idx) ++I;
    const int j = idx - I * (I + 1);
    const int tm0 = I * kDTM, tn0 = j * kDTN;
    if (tm0 >= n || tn0 >= n) return;
    const bool dbgT = (v.dbg != nullptr) && idx == 40 && f == 0 && threadIdx.x == 0 && K > 128;
    if (dbgT) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        v.dbg[8] = clock64();
        v.dbg[9] = (long long)gt;
    }
    const double* W = v.Bu + (size_t)f * v.kmax * v.ld;
    double* P = v.P + (size_t)f * v.nmax * v.ld;
    const int ld = v.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, q = lane & 3;
    // number of valid 8-row / 8-column sub-tiles of this warp (ragged edge)
    const int aMax = min(8, max(0, (n - (tm0 + wm * 64) + 7) / 8));
    const int bMax = min(4, max(0, (n - (tn0 + wn * 32) + 7) / 8));
    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int nk = (K + kKC - 1) / kKC;
    auto stA = [&](int st) { return dsm2 + (size_t)st * kKC * (kDLA + kDLB); };
    auto stB = [&](int st) { return dsm2 + (size_t)st * kKC * (kDLA + kDLB) + kKC * kDLA; };
#pragma unroll
    for (int st = 0; st < kStages - 1; ++st) {
        if (st < nk) {
            load_slab_w<kDLA, kDTM, 128>(stA(st), W, ld, st * kKC, K, tm0, ld, tid);
            load_slab_w<kDLB, kDTN, 128>(stB(st), W, ld, st * kKC, K, tn0, ld, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {
            const int nx = kt + kStages - 1;
            if (nx < nk) {
                load_slab_w<kDLA, kDTM, 128>(stA(nx % kStages), W, ld, nx * kKC, K, tm0, ld, tid);
                load_slab_w<kDLB, kDTN, 128>(stB(nx % kStages), W, ld, nx * kKC, K, tn0, ld, tid);
            }
            cp_async_commit();
        }
        const double* As = stA(kt % kStages) + wm * 64 + g;
        const double* Bs = stB(kt % kStages) + wn * 32 + g;
#pragma unroll
        for (int k4 = 0; k4 < kKC; k4 += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int a = 0; a < 8; ++a) af[a] = As[(k4 + q) * kDLA + a * 8];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[(k4 + q) * kDLB + b * 8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
                if (a < aMax) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (b < bMax) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
                }
        }
    }
    cp_async_wait<0>();
    if (dbgT) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        v.dbg[10] = clock64();
        v.dbg[11] = (long long)gt;
        v.dbg[12] = K;
    }
    __syncthreads();  // pipeline buffers are free: reuse them as the transpose tile T[n][m], pitch 130
    double* T = dsm2;
    const bool needMirror = true;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int ml = wm * 64 + a * 8 + g, gm = tm0 + ml;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int nl = wn * 32 + b * 8 + 2 * q, gn = tn0 + nl;
            double v0 = 0.0, v1 = 0.0;
            if (gm < n && gn <= gm) {  // lower (incl. diagonal): read-modify-write in place
                if (gn + 1 <= gm && gn + 1 < n) {
                    double2 c2 = *reinterpret_cast<double2*>(P + (size_t)gm * ld + gn);
                    v0 = c2.x - acc[a][b][0];
                    v1 = c2.y - acc[a][b][1];
                    *reinterpret_cast<double2*>(P + (size_t)gm * ld + gn) = make_double2(v0, v1);
                } else {
                    v0 = P[(size_t)gm * ld + gn] - acc[a][b][0];
                    P[(size_t)gm * ld + gn] = v0;
                }
            }
            if (needMirror) {
                T[(size_t)nl * kDTP + ml] = v0;
                T[(size_t)(nl + 1) * kDTP + ml] = v1;
            }
        }
    }
    __syncthreads();
    // mirrored rows: P[gn][gm] = T[nl][ml] for gm > gn, contiguous in gm
    for (int nl = warp; nl < kDTN; nl += 4) {
        const int gn = tn0 + nl;
        if (gn >= n) break;
        double* Prow = P + (size_t)gn * ld + tm0;
        const double* Trow = T + (size_t)nl * kDTP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ml = h * 64 + 2 * lane, gm = tm0 + ml;
            const double2 t2 = *reinterpret_cast<const double2*>(Trow + ml);
            if (gm > gn && gm + 1 < n) {
                *reinterpret_cast<double2*>(Prow + ml) = t2;
            } else {
                if (gm > gn && gm < n) Prow[ml] = t2.x;
                if (gm + 1 > gn && gm + 1 < n) Prow[ml + 1] = t2.y;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// U2: x += deadband(dx) (E/Update.cpp:143-204); dx = W y was accumulated by the panel kernels.
// grid (ceil(n/256), F)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_state_apply(DevView v)
{
    grid_dependency_wait();   // launched with programmatic stream serialisation: nothing of the predecessor is read before this
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    if (dm[D_ULIST] == 0 || dm[D_STATUS] != 0) return;
    const int n = dm[D_N_STATE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double* x = v.x + (size_t)f * v.ld;
    if (i < n) {
        const double d = v.dx[(size_t)f * v.ld + i];
        if (fabs(d) > kDelta) x[i] += d;
    }
    if (blockIdx.x != 0) return;
    // U4 (a), fused: J = d(q/|q|)/dq at the un-normalised q, then q <- q/|q| (E/Update.cpp:45-60,309-317).  The quaternion
    // (elements 3..6) belongs to this block.
    __syncthreads();
    if (threadIdx.x != 0) return;
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
    grid_dependency_wait();   // launched with programmatic stream serialisation: nothing of the predecessor is read before this
    if (PDL_EARLY_TRIGGER) grid_launch_dependents();
    const int f = blockIdx.y;
    const int* dm = fdims(v, f);
    if (dm[D_ULIST] == 0 || dm[D_STATUS] != 0) return;
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
