// ekf_schain.cuh -- the S-chain: blocked Cholesky  S = U^T U  of the innovation covariance [S | nu] (k x (k+1), k <= ~900),
// the latency-critical part of update() (reference: Update.cpp:92-109 inverts S with cv::Mat::inv(); here the inverse is
// never formed: the factor and the inverses of its 64x64 diagonal blocks feed the slab TRSM of ekf_linalg.cuh).
//
// One launch per 64-row block step J = -1 .. nbR-2, one CTA per upper 64x64 tile (I, C), J < I <= C, of the trailing matrix:
//     X_I = Uinv_J^T S(J, I),  X_C = Uinv_J^T S(J, C)          (both redone by every tile CTA: no second launch, no grid sync)
//     S(I, C) -= X_I^T X_C                                      (FP64 tensor pipe, DMMA m8n8k4)
//     row I == J+1 : writes X_C = U(J, C) to the factor buffer Sf (Sf != S: other CTAs still read the raw S(J, C))
//     tile (J+1, J+1): factors its updated tile in shared memory, U_{J+1,J+1} and Uinv_{J+1} = U_{J+1,J+1}^-1, so the next
//                      launch can start from it.  This CTA is the critical path of the chain.
// The 64x64 factorisation is hierarchical (8x8 blocks): every thread redoes the 8x8 Cholesky of the current diagonal block in
// registers (no communication on the pivot chain, 1/sqrt by rsqrt.approx + one Halley step), block rows and the running
// inverse are finished by per-thread substitutions, and the rank-8 trailing updates are spread as 4x4 register blocks.
#pragma once

namespace ekf {

constexpr int kSS = 68;  // shared-memory pitch of a 64x64 tile (K-major DMMA operands, conflict-free)
constexpr int kStepSmem = (4 * kNB * kSS + 2 * kNB) * (int)sizeof(double);

__device__ __forceinline__ double rsqrt_fast(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double dy = d * y;
    const double e = fma(-dy, y, 1.0);            // 1 - d y^2
    return fma(y * e, fma(0.375, e, 0.5), y);     // y (1 + e/2 + 3 e^2 / 8): third-order, ~1 ulp from a 2^-20 seed
}

// acc[a][b][:] += sum_s A[s][wm*32 + a*8 + ..] * B[s][wn*32 + b*8 + ..]   (A, B: [64][kSS], K-major).  TRI: A is upper
// triangular (A[s][m] = 0 for s > m), the k-steps that only meet zeros are skipped.
template <bool TRI>
__device__ __forceinline__ void tile_gemm64(const double* __restrict__ A, const double* __restrict__ B, int wm, int wn, int g, int q,
                                            double (&acc)[4][4][2])
{
    const int kEnd = TRI ? wm * 32 + 32 : kNB;
#pragma unroll 2
    for (int k4 = 0; k4 < kEnd; k4 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = A[(k4 + q) * kSS + wm * 32 + a * 8 + g];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = B[(k4 + q) * kSS + wn * 32 + b * 8 + g];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (TRI && k4 >= wm * 32 + a * 8 + 8) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
}

// 64 x 64 tile of a row-major matrix into shared memory; rows >= rowLimit and columns >= colLimit become zero
__device__ __forceinline__ void load_tile64(double* dst, const double* src, int ld, int row0, int rowLimit, int col0, int colLimit, int tid)
{
    for (int e = tid; e < kNB * 32; e += 128) {
        const int r = e >> 5, c2 = (e & 31) * 2;
        double* d = dst + r * kSS + c2;
        const double* s = src + (size_t)(row0 + r) * ld + col0 + c2;
        if (row0 + r < rowLimit && col0 + c2 + 1 < colLimit) cp_async16(d, s);
        else {
            d[0] = (row0 + r < rowLimit && col0 + c2 < colLimit) ? s[0] : 0.0;
            d[1] = 0.0;
        }
    }
}

// In-place factorisation of the SPD tile T (upper triangle read, [64][kSS], valid kb x kb, the rest is replaced by identity):
// on return T holds U (upper, T = U^T U) and W holds U^-1 (upper, zeros below).  128 threads.  *bad is set if a pivot of the
// valid part is not positive.
__device__ void factor_tile64(double* T, double* W, int kb, int tid, int* bad)
{
    for (int e = tid; e < kNB * kNB; e += 128) {
        const int i = e >> 6, j = e & 63;
        if (i >= kb || j >= kb) T[i * kSS + j] = (i == j) ? 1.0 : 0.0;
        W[i * kSS + j] = 0.0;
    }
    __syncthreads();
    for (int b = 0; b < 8; ++b) {
        const int o = 8 * b;
        // ---- every thread: Cholesky of the 8x8 diagonal block, R upper, r = 1 / diag(R)
        double R[8][8], r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = i; j < 8; ++j) R[i][j] = T[(o + i) * kSS + o + j];
        bool neg = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            neg = neg || !(R[j][j] > 0.0);
            r[j] = rsqrt_fast(R[j][j]);
#pragma unroll
            for (int c = j; c < 8; ++c) R[j][c] *= r[j];
#pragma unroll
            for (int i = j + 1; i < 8; ++i)
#pragma unroll
                for (int c = i; c < 8; ++c) R[i][c] -= R[j][i] * R[j][c];
        }
        if (tid == 0 && neg && o < kb) *bad = 1;   // padded rows are identity, so any bad pivot is a real one
        // ---- per-thread substitutions against R
        const int ncol = kNB - o - 8;
        if (tid < ncol) {  // block row: column j of U(b, >b) = R^-T t
            double* col = T + o * kSS + o + 8 + tid;
            double u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double s = col[i * kSS];
#pragma unroll
                for (int p = 0; p < i; ++p) s -= R[p][i] * u[p];
                u[i] = s * r[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) col[i * kSS] = u[i];
        } else if (tid >= 64 && (tid < 72 || tid >= 128 - o)) {
            // column block b of W = U^-1: rows above the block solve x R = -G (G accumulated in place), the block's own
            // rows solve x R = e_i
            const bool own = tid < 72;
            const int row = own ? o + (tid - 64) : 127 - tid;
            double* wr = W + row * kSS + o;
            double x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double s = own ? ((tid - 64) == j ? 1.0 : 0.0) : -wr[j];
#pragma unroll
                for (int p = 0; p < j; ++p) s -= x[p] * R[p][j];
                x[j] = s * r[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[j] = x[j];
        }
        __syncthreads();
        if (tid == 56 + (b & 1)) {  // the diagonal block of U (after the barrier: the other warps read it during their Cholesky)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = i; j < 8; ++j) T[(o + i) * kSS + o + j] = R[i][j];
        }
        // ---- rank-8 updates as 4x4 register blocks: trailing tile (upper blocks only) and the running inverse
        const int nb4 = ncol / 4;
        const int nTrail = nb4 * (nb4 + 1) / 2;
        const int nGc = nb4, nG = 2 * (b + 1) * nGc;
        const double* Ub = T + o * kSS;  // the finished block row: Ub[s][col]
        for (int it = tid; it < nTrail + nG; it += 128) {
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
            if (it < nTrail) {
                int br = 0, t = it;
                while (t >= nb4 - br) { t -= nb4 - br; ++br; }
                const int r0 = o + 8 + 4 * br, c0 = o + 8 + 4 * (br + t);
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const double2 a01 = *reinterpret_cast<const double2*>(Ub + s * kSS + r0);
                    const double2 a23 = *reinterpret_cast<const double2*>(Ub + s * kSS + r0 + 2);
                    const double2 b01 = *reinterpret_cast<const double2*>(Ub + s * kSS + c0);
                    const double2 b23 = *reinterpret_cast<const double2*>(Ub + s * kSS + c0 + 2);
                    const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double2* p0 = reinterpret_cast<double2*>(T + (r0 + i) * kSS + c0);
                    double2 v0 = p0[0], v1 = p0[1];
                    v0.x -= acc[i][0]; v0.y -= acc[i][1]; v1.x -= acc[i][2]; v1.y -= acc[i][3];
                    p0[0] = v0; p0[1] = v1;
                }
            } else {
                const int t = it - nTrail;
                const int r0 = 4 * (t / nGc), c0 = o + 8 + 4 * (t % nGc);
                double av[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) {
                        const double2 w = *reinterpret_cast<const double2*>(W + (r0 + i) * kSS + o + 2 * s2);
                        av[i][2 * s2] = w.x; av[i][2 * s2 + 1] = w.y;
                    }
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const double2 b01 = *reinterpret_cast<const double2*>(Ub + s * kSS + c0);
                    const double2 b23 = *reinterpret_cast<const double2*>(Ub + s * kSS + c0 + 2);
                    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] += av[i][s] * bv[j];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double2* p0 = reinterpret_cast<double2*>(W + (r0 + i) * kSS + c0);
                    double2 v0 = p0[0], v1 = p0[1];
                    v0.x += acc[i][0]; v0.y += acc[i][1]; v1.x += acc[i][2]; v1.y += acc[i][3];
                    p0[0] = v0; p0[1] = v1;
                }
            }
        }
        __syncthreads();
    }
}

// grid (nbC - (J+1), nbR - J, F), 128 threads, dynamic smem kStepSmem.  J = -1 only factors tile (0, 0).
__global__ void __launch_bounds__(128, 1) k_schain_step(DevView v, int J)
{
    extern __shared__ __align__(16) double csm[];
    double* Ws = csm;                  // Uinv_J, later the inverse of the new diagonal block
    double* As = Ws + kNB * kSS;       // S(J, I) -> X_I
    double* Bs = As + kNB * kSS;       // S(J, C) -> X_C
    double* Ts = Bs + kNB * kSS;       // tile (I, C)
    double* nu = Ts + kNB * kSS;       // [64] innovation entries of the diagonal tile's rows
    const int f = blockIdx.z;
    int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    if (k == 0) return;
    const int nbR = (k + kNB - 1) / kNB, nbC = (k + kNB) / kNB;  // row blocks of S, column tiles of [S | nu]
    const int I = J + 1 + blockIdx.y, C = J + 1 + blockIdx.x;
    if (C < I || C >= nbC || I > nbR) return;
    if (J < 0 && C != 0) return;       // the first launch only factors tile (0, 0)
    const bool phantom = (I == nbR);   // no rows left: only the X_C of a column tile that holds just nu is produced
    if (phantom && blockIdx.y != 0) return;
    const bool diag = (I == C) && !phantom;     // X_C is X_I
    const bool fac = diag && blockIdx.y == 0;  // tile (J+1, J+1): factored here
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3, wm = w >> 1, wn = w & 1;
    const int J0 = J * kNB, I0 = I * kNB, C0 = C * kNB;
    double* Sg = v.S + (size_t)f * v.kmax * v.ldS;
    double* Sf = v.Sf + (size_t)f * v.kmax * v.ldS;
    double* UinvG = v.Uinv + (size_t)f * (v.kmax / kNB) * kNB * kNB;

    if (J >= 0) {
        for (int e = tid; e < kNB * 32; e += 128) {
            const int r = e >> 5, c2 = (e & 31) * 2;
            cp_async16(Ws + r * kSS + c2, UinvG + (size_t)J * kNB * kNB + r * kNB + c2);
        }
        if (!phantom) load_tile64(As, Sg, v.ldS, J0, k, I0, k + 1, tid);
        if (!diag) load_tile64(Bs, Sg, v.ldS, J0, k, C0, k + 1, tid);
    }
    if (!phantom) load_tile64(Ts, Sg, v.ldS, I0, k, C0, k + 1, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    if (J >= 0) {
        double xa[4][4][2], xb[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) xa[a][b][0] = xa[a][b][1] = xb[a][b][0] = xb[a][b][1] = 0.0;
        if (!phantom) tile_gemm64<true>(Ws, As, wm, wn, g, q, xa);
        if (!diag) tile_gemm64<true>(Ws, Bs, wm, wn, g, q, xb);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int m = wm * 32 + a * 8 + g, nn = wn * 32 + b * 8 + 2 * q;
                if (!phantom) *reinterpret_cast<double2*>(As + m * kSS + nn) = make_double2(xa[a][b][0], xa[a][b][1]);
                if (!diag) *reinterpret_cast<double2*>(Bs + m * kSS + nn) = make_double2(xb[a][b][0], xb[a][b][1]);
                if (blockIdx.y == 0) {  // row I == J+1 publishes U(J, C)
                    const double x0 = diag ? xa[a][b][0] : xb[a][b][0], x1 = diag ? xa[a][b][1] : xb[a][b][1];
                    double* dst = Sf + (size_t)(J0 + m) * v.ldS + C0 + nn;
                    if (C0 + nn + 1 <= k) *reinterpret_cast<double2*>(dst) = make_double2(x0, x1);
                    else if (C0 + nn <= k) dst[0] = x0;
                }
            }
        if (phantom) return;
        __syncthreads();
        double acc[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        tile_gemm64<false>(As, diag ? As : Bs, wm, wn, g, q, acc);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int m = wm * 32 + a * 8 + g, nn = wn * 32 + b * 8 + 2 * q;
                double2 t = *reinterpret_cast<double2*>(Ts + m * kSS + nn);
                t.x -= acc[a][b][0]; t.y -= acc[a][b][1];
                if (fac) *reinterpret_cast<double2*>(Ts + m * kSS + nn) = t;
                else if (I0 + m < k) {
                    double* dst = Sg + (size_t)(I0 + m) * v.ldS + C0 + nn;
                    if (C0 + nn + 1 <= k) *reinterpret_cast<double2*>(dst) = t;
                    else if (C0 + nn <= k) dst[0] = t.x;
                }
            }
        if (!fac) return;
        __syncthreads();
    }
    // ---- diagonal tile: factor, publish U_II, Uinv_I and (if nu lies in this tile) y_I
    const int kb = min(kNB, k - I0);
    const bool hasNu = (k - I0) < kNB;
    if (hasNu && tid < kNB) nu[tid] = (tid < kb) ? Ts[tid * kSS + kb] : 0.0;
    __syncthreads();
    __shared__ int bad;
    if (tid == 0) bad = 0;
    factor_tile64(Ts, Ws, kb, tid, &bad);
    if (tid == 0 && bad) dm[D_STATUS] = 4;  // EKFB_ERR_NUMERIC
    for (int e = tid; e < kNB * kNB; e += 128) {
        const int i = e >> 6, j = e & 63;
        UinvG[(size_t)I * kNB * kNB + e] = Ws[i * kSS + j];
        if (i < kb && j < kb) Sf[(size_t)(I0 + i) * v.ldS + I0 + j] = (j >= i) ? Ts[i * kSS + j] : 0.0;
    }
    if (hasNu && tid < kb) {
        double s = 0.0;
        for (int p = 0; p <= tid; ++p) s += Ws[p * kSS + tid] * nu[p];
        Sf[(size_t)(I0 + tid) * v.ldS + k] = s;
    }
}

}  // namespace ekf
