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

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void bar_factor() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 64 x 64 tile of a row-major matrix into shared memory (256 threads); rows >= rowLimit and columns >= colLimit become zero
__device__ __forceinline__ void load_tile64(double* dst, const double* src, int ld, int row0, int rowLimit, int col0, int colLimit, int tid)
{
#pragma unroll 4
    for (int e = tid; e < kNB * 32; e += 256) {
        const int r = e >> 5, c2 = (e & 31) * 2;
        double* d = dst + r * kSS + c2;
        const double* s = src + (size_t)(row0 + r) * ld + col0 + c2;
        const bool rowOk = row0 + r < rowLimit;
        if (rowOk && col0 + c2 + 1 < colLimit) cp_async16(d, s);
        else if (rowOk && col0 + c2 < colLimit) { cp_async8(d, s); d[1] = 0.0; }
        else { d[0] = 0.0; d[1] = 0.0; }
    }
}

// X = A^T Sx with A upper triangular (A[s][m] = 0 for s > m), for one or two right-hand tiles that share the A fragments.
// Warp (p, nh): m-tile rows {p, 7-p} (together 9 k-blocks: balanced), n-tiles 4nh .. 4nh+3.
template <bool TWO>
__device__ __forceinline__ void x_gemm(const double* __restrict__ A, const double* __restrict__ S0, const double* __restrict__ S1, int p,
                                       int nh, int g, int q, double (&x0)[2][4][2], double (&x1)[2][4][2])
{
    const int m0 = 8 * p, m1 = 8 * (7 - p), n0 = 32 * nh;
    int k4 = 0;
#pragma unroll 2
    for (; k4 < m0 + 8; k4 += 4) {
        const double* Ar = A + (k4 + q) * kSS;
        const double a0 = Ar[m0 + g], a1 = Ar[m1 + g];
        double b0[4], b1[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            b0[b] = S0[(k4 + q) * kSS + n0 + 8 * b + g];
            if (TWO) b1[b] = S1[(k4 + q) * kSS + n0 + 8 * b + g];
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            dmma8x8x4(x0[0][b][0], x0[0][b][1], a0, b0[b]);
            dmma8x8x4(x0[1][b][0], x0[1][b][1], a1, b0[b]);
            if (TWO) {
                dmma8x8x4(x1[0][b][0], x1[0][b][1], a0, b1[b]);
                dmma8x8x4(x1[1][b][0], x1[1][b][1], a1, b1[b]);
            }
        }
    }
#pragma unroll 2
    for (; k4 < m1 + 8; k4 += 4) {
        const double a1 = A[(k4 + q) * kSS + m1 + g];
        double b0[4], b1[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            b0[b] = S0[(k4 + q) * kSS + n0 + 8 * b + g];
            if (TWO) b1[b] = S1[(k4 + q) * kSS + n0 + 8 * b + g];
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            dmma8x8x4(x0[1][b][0], x0[1][b][1], a1, b0[b]);
            if (TWO) dmma8x8x4(x1[1][b][0], x1[1][b][1], a1, b1[b]);
        }
    }
}

// the 36 upper 8x8 tiles of a 64x64 tile, row-major
__constant__ unsigned char kUpperTile[36][2] = {
    {0, 0}, {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {0, 7}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {1, 7}, {2, 2}, {2, 3}, {2, 4},
    {2, 5}, {2, 6}, {2, 7}, {3, 3}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {4, 4}, {4, 5}, {4, 6}, {4, 7}, {5, 5}, {5, 6}, {5, 7}, {6, 6}, {6, 7}, {7, 7}};

// Left-looking tile updates of factor_tile64.  Before block step s >= 1 can start, eight 8x8 tiles must be brought up to date,
// one per warp (i = warp index):
//   i < 8-s : T(s, c) -= sum_{k < 8s} U[k][8s + .]^T U[k][8c + .],  c = s + i          (block row s of the tile)
//   else    : G(a, s)  = sum_{8a <= k < 8s} W[8a + .][k] U[k][8s + .], a = i - (8 - s)  (column block s of U^-1)
// (Accumulating the k < 8(s-1) part ahead of time on warps 4-7, while warps 0-3 run the in-register Cholesky, was measured:
// the DMMAs share the FP64 pipe with the Cholesky's dependent DFMA chain and slow it by as much as they save.)
struct LLTile {
    const double* ap;   // a-fragment A[k][m]: T[k][8s + m] (stride kSS in k) or W[8a + m][k] (stride 1)
    const double* bp;   // b-fragment B[k][n] = T[k][8c + n]
    double2* dst;
    int astep, k0;
    bool inv;
};

__device__ __forceinline__ LLTile ll_tile(double* T, double* W, int s, int i, int g, int q)
{
    LLTile t;
    t.inv = i >= 8 - s;
    const int a8 = t.inv ? i - (8 - s) : 0, c8 = t.inv ? s : s + i;
    t.k0 = t.inv ? 8 * a8 : 0;
    t.ap = t.inv ? W + (8 * a8 + g) * kSS + q : T + q * kSS + 8 * s + g;
    t.astep = t.inv ? 1 : kSS;
    t.bp = T + q * kSS + 8 * c8 + g;
    t.dst = reinterpret_cast<double2*>((t.inv ? W + (8 * a8 + g) * kSS : T + (8 * s + g) * kSS) + 8 * c8 + 2 * q);
    return t;
}

__device__ __forceinline__ void ll_accumulate(const LLTile& t, int kBegin, int kEnd, double (&c)[4])
{
    // four independent accumulator pairs: the dependent DMMA chain (25 cycles per link) is a quarter of the k-range long
    double d[4] = {0.0, 0.0, 0.0, 0.0};
    int k = kBegin;
    for (; k + 8 < kEnd; k += 16) {
        const double aa0 = t.ap[k * t.astep], aa1 = t.ap[(k + 4) * t.astep], aa2 = t.ap[(k + 8) * t.astep], aa3 = t.ap[(k + 12) * t.astep];
        const double bb0 = t.bp[k * kSS], bb1 = t.bp[(k + 4) * kSS], bb2 = t.bp[(k + 8) * kSS], bb3 = t.bp[(k + 12) * kSS];
        dmma8x8x4(c[0], c[1], aa0, bb0);
        dmma8x8x4(c[2], c[3], aa1, bb1);
        dmma8x8x4(d[0], d[1], aa2, bb2);
        dmma8x8x4(d[2], d[3], aa3, bb3);
    }
    if (k < kEnd) {
        const double aa0 = t.ap[k * t.astep], aa1 = t.ap[(k + 4) * t.astep];
        const double bb0 = t.bp[k * kSS], bb1 = t.bp[(k + 4) * kSS];
        dmma8x8x4(c[0], c[1], aa0, bb0);
        dmma8x8x4(c[2], c[3], aa1, bb1);
    }
    c[0] += d[0]; c[1] += d[1]; c[2] += d[2]; c[3] += d[3];
}

__device__ __forceinline__ void ll_store(const LLTile& t, const double (&c)[4])
{
    double2 v0 = *t.dst;
    if (t.inv) { v0.x += c[0] + c[2]; v0.y += c[1] + c[3]; }
    else { v0.x -= c[0] + c[2]; v0.y -= c[1] + c[3]; }
    *t.dst = v0;
}

// In-place factorisation of the SPD tile T (upper triangle read, [64][kSS]; the caller has replaced everything outside the
// valid part by identity and zeroed W): on return T holds U (upper, T = U^T U) and W holds U^-1 (upper, zeros below).
// Called by all 256 threads: warps 0-3 do the in-register work, warps 4-7 only join the rank-8 tile updates.  *bad is set if a
// pivot is not positive.
struct NoHook { __device__ __forceinline__ void operator()(int) const {} };

// hook(b): called by all threads at the start of every 8-row block step b: lets the caller overlap its own memory traffic
// (publishing finished tiles, prefetching its next tiles) with the latency-bound factorisation
// Xp != nullptr: a K-major 64x64 tile X whose product X^T X is still to be subtracted from T (the previous block row of the
// blocked factorisation): it is folded into the left-looking tile updates -- block row b of T gets its share right before it
// is needed -- instead of costing a separate 64^3 product in front of the factorisation.
// nblk < 8: only the leading 8 * nblk rows are factored -- the caller guarantees that the rest of T is the identity padding
// (then U and U^-1 are the identity there: W gets ones on that part of its diagonal and the remaining block steps are skipped).
// idle(b): called by warps 4-7 only, while warps 0-3 run the in-register Cholesky and the substitutions of block step b (they
// have nothing to do then): global-memory latency spent here (a flag poll) stays off the critical path.
template <typename Hook = NoHook, typename Idle = NoHook>
__device__ __forceinline__ void factor_tile64(double* T, double* W, int tid, int* bad, long long* dbg, Hook hook = Hook(),
                                              const double* Xp = nullptr, int nblk = 8, Idle idle = Idle())
{
    if (tid >= 8 * nblk && tid < kNB) W[tid * kSS + tid] = 1.0;   // (disjoint from everything the block steps below touch)
    long long tA = 0, tB = 0, tC = 0, t0 = 0;
    const int lane = tid & 31, w = tid >> 5, g = lane >> 2, q = lane & 3;
    for (int b = 0; b < nblk; ++b) {
        const int o = 8 * b;
        if (dbg) t0 = clock64();
        double R[8][8], r[8];
        hook(b);
        if (b > 0 || Xp != nullptr) {  // bring block row b of T and column block b of U^-1 up to date: one tile per warp
            const LLTile t = ll_tile(T, W, b, w, g, q);
            double ca[4] = {0.0, 0.0, 0.0, 0.0};
            if (b > 0) ll_accumulate(t, t.k0, o, ca);
            if (Xp != nullptr && !t.inv) {
                LLTile tx = t;
                tx.ap = Xp + q * kSS + o + g;
                tx.astep = kSS;
                tx.bp = Xp + q * kSS + 8 * (b + w) + g;
                ll_accumulate(tx, 0, kNB, ca);
            }
            ll_store(t, ca);
            __syncthreads();
        }
        if (dbg) { const long long t1 = clock64(); tC += t1 - t0; t0 = t1; }
        if (w < 4) {
        // ---- every thread of warps 0-3: Cholesky of the 8x8 diagonal block, R upper, r = 1 / diag(R)
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
        if (tid == 0 && neg) *bad = 1;
        if (dbg) { const long long t1 = clock64(); tA += t1 - t0; t0 = t1; }
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
        } else
            idle(b);
        __syncthreads();
        if (dbg) { const long long t1 = clock64(); tB += t1 - t0; t0 = t1; }
        if (tid == 56 + (b & 1)) {  // the diagonal block of U (after the barrier: the other warps read it during their Cholesky)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = i; j < 8; ++j) T[(o + i) * kSS + o + j] = R[i][j];
        }
    }
    if (dbg) { dbg[0] = tA; dbg[1] = tB; dbg[2] = tC; }
}

// grid (nbC - (J+1), max(nbR - (J+1), 1), F), 256 threads, dynamic smem kStepSmem.  J = -1 only factors tile (0, 0).
__global__ void __launch_bounds__(256, 1) k_schain_step(DevView v, int J)
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
    const bool diag = (I == C) && !phantom;     // X_C is X_I, only the upper 8x8 tiles of the update are needed
    const bool fac = diag && blockIdx.y == 0;  // tile (J+1, J+1): factored here
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3, p = w & 3, nh = w >> 2;
    const int J0 = J * kNB, I0 = I * kNB, C0 = C * kNB;
    double* Sg = v.S + (size_t)f * v.kmax * v.ldS;
    double* Sf = v.Sf + (size_t)f * v.kmax * v.ldS;
    double* UinvG = v.Uinv + (size_t)f * (v.kmax / kNB) * kNB * kNB;

    grid_launch_dependents();   // the next step (or the slab TRSM) may be scheduled now; it waits for this grid's completion
    grid_dependency_wait();     // block row J and Uinv_J come from the previous launch
    long long* dbg = (v.dbg != nullptr && fac && f == 0 && tid == 0 && J == 1) ? v.dbg + (k > 300 ? 16 : 40) : nullptr;
    if (dbg) dbg[0] = clock64();
    if (J >= 0) {
#pragma unroll 4
        for (int e = tid; e < kNB * 32; e += 256) {
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
    if (dbg) dbg[1] = clock64();

    if (J >= 0) {
        // ---- X_I (-> As) and X_C (-> Bs), both = Uinv_J^T (raw block row J)
        double xa[2][4][2], xb[2][4][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) xa[a][b][0] = xa[a][b][1] = xb[a][b][0] = xb[a][b][1] = 0.0;
        if (phantom) x_gemm<false>(Ws, Bs, Bs, p, nh, g, q, xb, xb);
        else if (diag) x_gemm<false>(Ws, As, As, p, nh, g, q, xa, xa);
        else x_gemm<true>(Ws, As, Bs, p, nh, g, q, xa, xb);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int m = (a == 0 ? 8 * p : 8 * (7 - p)) + g, nn = 32 * nh + 8 * b + 2 * q;
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
        if (dbg) dbg[2] = clock64();
        // ---- S(I, C) -= X_I^T X_C
        if (!diag) {
            double acc[2][4][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
            const int m0 = 16 * p, n0 = 32 * nh;
#pragma unroll 2
            for (int k4 = 0; k4 < kNB; k4 += 4) {
                double af[2], bf[4];
#pragma unroll
                for (int a = 0; a < 2; ++a) af[a] = As[(k4 + q) * kSS + m0 + 8 * a + g];
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = Bs[(k4 + q) * kSS + n0 + 8 * b + g];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int m = m0 + 8 * a + g;
                if (I0 + m >= k) continue;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int nn = n0 + 8 * b + 2 * q;
                    double2 t = *reinterpret_cast<double2*>(Ts + m * kSS + nn);
                    t.x -= acc[a][b][0]; t.y -= acc[a][b][1];
                    double* dst = Sg + (size_t)(I0 + m) * v.ldS + C0 + nn;
                    if (C0 + nn + 1 <= k) *reinterpret_cast<double2*>(dst) = t;
                    else if (C0 + nn <= k) dst[0] = t.x;
                }
            }
            return;
        }
        {
            double acc[5][2];
            int mi[5], ni[5];
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const int id = min(w + 8 * t, 35);
                mi[t] = 8 * kUpperTile[id][0]; ni[t] = 8 * kUpperTile[id][1];
                acc[t][0] = acc[t][1] = 0.0;
            }
            const bool five = (w + 32 < 36);
#pragma unroll 2
            for (int k4 = 0; k4 < kNB; k4 += 4) {
                const double* Ar = As + (k4 + q) * kSS + g;
                double af[5], bf[5];
#pragma unroll
                for (int t = 0; t < 5; ++t) { af[t] = Ar[mi[t]]; bf[t] = Ar[ni[t]]; }
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma8x8x4(acc[t][0], acc[t][1], af[t], bf[t]);
                if (five) dmma8x8x4(acc[4][0], acc[4][1], af[4], bf[4]);
            }
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                if (t == 4 && !five) break;
                const int m = mi[t] + g, nn = ni[t] + 2 * q;
                double2 tv = *reinterpret_cast<double2*>(Ts + m * kSS + nn);
                tv.x -= acc[t][0]; tv.y -= acc[t][1];
                if (fac) *reinterpret_cast<double2*>(Ts + m * kSS + nn) = tv;
                else if (I0 + m < k) {
                    double* dst = Sg + (size_t)(I0 + m) * v.ldS + C0 + nn;
                    if (C0 + nn + 1 <= k) *reinterpret_cast<double2*>(dst) = tv;
                    else if (C0 + nn <= k) dst[0] = tv.x;
                }
            }
        }
        if (!fac) return;
        __syncthreads();
        if (dbg) dbg[3] = clock64();
    }
    // ---- diagonal tile: factor, publish U_II, Uinv_I and (if nu lies in this tile) y_I
    const int kb = min(kNB, k - I0);
    const bool hasNu = (k - I0) < kNB;
    if (hasNu && tid < kNB) nu[tid] = (tid < kb) ? Ts[tid * kSS + kb] : 0.0;
    __shared__ int bad;
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int e = tid; e < kNB * 32; e += 256) {  // identity outside the valid part; W starts at zero
        const int i = e >> 5, j = (e & 31) * 2;
        if (i >= kb || j + 1 >= kb) {
            if (i >= kb || j >= kb) Ts[i * kSS + j] = (i == j) ? 1.0 : 0.0;
            Ts[i * kSS + j + 1] = (i == j + 1) ? 1.0 : 0.0;
        }
        *reinterpret_cast<double2*>(Ws + i * kSS + j) = make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (dbg) dbg[6] = clock64();
    factor_tile64(Ts, Ws, tid, &bad, dbg ? dbg + 8 : nullptr, NoHook(), nullptr, (kb + 7) >> 3);
    __syncthreads();
    if (dbg) dbg[4] = clock64();
    if (tid == 0 && (bad || v.faultInject)) dm[D_STATUS] = 4;  // EKFB_ERR_NUMERIC
    for (int e = tid; e < kNB * 32; e += 256) {
        const int i = e >> 5, j = (e & 31) * 2;
        *reinterpret_cast<double2*>(UinvG + (size_t)I * kNB * kNB + i * kNB + j) = *reinterpret_cast<const double2*>(Ws + i * kSS + j);
        if (i < kb && j + 1 >= i) {
            double* dst = Sf + (size_t)(I0 + i) * v.ldS + I0 + j;
            if (j >= i && j < kb) dst[0] = Ts[i * kSS + j];
            if (j + 1 < kb) dst[1] = Ts[i * kSS + j + 1];
        }
    }
    if (hasNu && tid < kb) {
        double s = 0.0;
        for (int pp = 0; pp <= tid; ++pp) s += Ws[pp * kSS + tid] * nu[pp];
        Sf[(size_t)(I0 + tid) * v.ldS + k] = s;
    }
    if (dbg) dbg[5] = clock64();
}

}  // namespace ekf
