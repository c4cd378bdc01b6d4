// ekf_chain.cuh -- the S-chain as ONE launch: blocked Cholesky  S = U^T U  of the innovation covariance [S | nu]
// (k x (k+1)) with the inverses of the 64x64 diagonal blocks, for the slab TRSM (reference: E/Update.cpp:105-108 inverts S by
// LU; here neither the inverse nor K is ever formed).  Replaces the per-block-step launches of ekf_schain.cuh (12 launches of
// ~17 us each at the C3 size, 27 % of the frame at 1.2 % of the FP64 peak): the steps were bound by launch + reload latency,
// not by work, so the whole factorisation now runs as a dataflow of 64x64 tile tasks inside one resident grid that
// hand tiles to each other through global memory (L2) and acquire / release flags.  k comes from device memory; the grid is
// sized for the largest k of the handle and surplus CTAs exit at once (nothing here is sized on the host per frame).
//
// Tile (I, C), I <= C, of the upper triangle of [S | nu]; X(I, C) = U(I, C) is a final factor tile (in the factor buffer Sf).
//   D(I)      critical chain, one CTA per filter:   X(I-1, I) = Uinv_{I-1}^T T(I-1, I)   (Uinv_{I-1} is still in its shared memory)
//                                                   T(I, I)   = T'(I, I) - X(I-1, I)^T X(I-1, I);  factor -> U_II, Uinv_I
//   A(I, C)   off-diagonal tile, left-looking:      T(I, C)   = S(I, C) - sum_{r < I} X(r, I)^T X(r, C)   as the rows arrive;
//             C == I+1: T handed to D(I+1);  else   X(I, C)   = Uinv_I^T T(I, C)
//   PD(C)     partial diagonal:                     T'(C, C)  = S(C, C) - sum_{r <= C-2} X(r, C)^T X(r, C)
// Per step the critical CTA does two 64^3 products and one 64x64 factorisation on tiles that are ready before it needs them;
// everything else runs beside it on other SMs.  Tasks are taken from a per-filter queue in row-major order (atomic counter),
// so a task only ever waits for tasks taken earlier by CTAs that are already running: no co-residency assumption, no
// dependence on block dispatch order.  All products on the FP64 tensor pipe (DMMA m8n8k4).
#pragma once

#include "ekf_schain.cuh"

namespace ekf {

#ifndef CHAIN_ABSORB_UPDATE
#define CHAIN_ABSORB_UPDATE 0   /* measured (profiles/r02_chain_phase_clocks.txt): the extra DMMAs land on the critical path of every block step of the factorisation: 28.9 k -> 33.0 k cycles per 64-row step */
#endif

// per-filter control block in global memory (ints): generation, ticket and queue counters, then the flags.  A flag is "set" when
// it holds the current generation, so nothing is cleared between launches; k_chain_finish advances the generation.
constexpr int CH_GEN = 0, CH_TICKET = 1, CH_QUEUE = 2, CH_FLAGS = 4;
__host__ __device__ inline int chain_ctl_ints(int nbMax) { return CH_FLAGS + 3 * (nbMax + 1) + (nbMax + 1) * (nbMax + 1); }

struct ChainCtl {
    int* base;
    int nbMax;
    __device__ __forceinline__ int* fdone(int I) const { return base + CH_FLAGS + I; }                     // U_II, Uinv_I published
    __device__ __forceinline__ int* tready(int I) const { return base + CH_FLAGS + (nbMax + 1) + I; }      // T(I, I+1) in S
    __device__ __forceinline__ int* pdready(int C) const { return base + CH_FLAGS + 2 * (nbMax + 1) + C; } // T'(C, C) in S
    __device__ __forceinline__ int* xready(int I, int C) const { return base + CH_FLAGS + 3 * (nbMax + 1) + I * (nbMax + 1) + C; }
};

__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr int kChainTimeoutStatus = 5;   // EKFB_ERR_INTERNAL: a flag did not arrive within ~1 s (never observed; bounds a hang)

// all threads: wait until *flag == gen (thread 0 polls).  Returns with the data published before the flag visible to the CTA.
// backoff: the 140-odd helper / TRSM CTAs that wait for a flag sleep ~100 ns between polls, so their polling does not load
// the L2 slice the critical CTA's flags and tiles live in; the critical CTA itself (backoff = false) polls back to back
__device__ __forceinline__ void chain_wait(const int* flag, int gen, int* status, bool backoff = true)
{
    if (threadIdx.x == 0) {
        if (ld_acquire(flag) != gen) {
            const long long t0 = clock64();
            while (ld_acquire(flag) != gen) {
                if (backoff) __nanosleep(100);
                if (clock64() - t0 > (1ll << 31)) { atomicExch(status, kChainTimeoutStatus); break; }
            }
        }
    }
    __syncthreads();
}
// all threads: the CTA's global stores so far become visible, then the flag is raised
__device__ __forceinline__ void chain_signal(int* flag, int gen)
{
    __syncthreads();                      // every thread's stores happen-before the signalling thread's release store, which is
    if (threadIdx.x == blockDim.x - 1)    // cumulative (the pattern of cutlass::Barrier::arrive_inc: barrier, then ONE fence +
        st_release(flag, gen);            // store; a __threadfence() in front of the st.release paid for the fence twice).  The
                                          // last thread signals: its warp idles through the in-register phases of the factorisation
}
constexpr int kChainSmem = (5 * kNB * kSS + 2 * kNB) * (int)sizeof(double);

struct ChainCtx {
    double *Ws, *As, *Bs, *Ts, *nu;     // shared-memory tiles (pitch kSS)
    double *A2;                         // fifth tile: the critical CTA prefetches T(I, I+1) into it while As still feeds the factorisation
    double *Sg, *Sf, *UinvG;            // this filter's S, factor buffer, diagonal-block inverses
    ChainCtl ctl;
    int* dm;
    int k, nbR, nbC, ldS, gen, faultInject;
    long long* dbg;
    int wsBlock;                         // which Uinv block Ws holds (-1: none)
    int pre;                             // D(pre)'s tiles were prefetched into As / Bs (-1: none)
};

// acc += A^T B for two K-major 64x64 tiles in shared memory; warp (p, nh) owns rows 16p.., columns 32nh..
__device__ __forceinline__ void tile_mma_full(const double* As, const double* Bs, int p, int nh, int g, int q, double (&acc)[2][4][2])
{
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
}

// acc += A^T A restricted to the 36 upper 8x8 tiles; warp w owns tiles w, w+8, ..., (five for w < 4)
__device__ __forceinline__ void tile_mma_upper(const double* As, int w, int g, int q, double (&acc)[5][2])
{
    int mi[5], ni[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        const int id = min(w + 8 * t, 35);
        mi[t] = 8 * kUpperTile[id][0]; ni[t] = 8 * kUpperTile[id][1];
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
}

// Ts(upper 8x8 tiles) -= acc
__device__ __forceinline__ void tile_sub_upper(double* Ts, int w, int g, int q, const double (&acc)[5][2])
{
#pragma unroll
    for (int t = 0; t < 5; ++t) {
        if (t == 4 && !(w + 32 < 36)) break;
        const int id = min(w + 8 * t, 35);
        const int m = 8 * kUpperTile[id][0] + g, nn = 8 * kUpperTile[id][1] + 2 * q;
        double2 tv = *reinterpret_cast<double2*>(Ts + m * kSS + nn);
        tv.x -= acc[t][0]; tv.y -= acc[t][1];
        *reinterpret_cast<double2*>(Ts + m * kSS + nn) = tv;
    }
}

// the 64x64 tile in shared memory -> rows row0.. / columns col0.. of a global matrix, clipped to rows < rowLimit, cols <= k
__device__ __forceinline__ void store_tile64(const double* src, double* dst, int ld, int row0, int rowLimit, int col0, int k, int tid)
{
    for (int e = tid; e < kNB * 32; e += 256) {
        const int r = e >> 5, c2 = (e & 31) * 2;
        if (row0 + r >= rowLimit) continue;
        double* d = dst + (size_t)(row0 + r) * ld + col0 + c2;
        const double2 t = *reinterpret_cast<const double2*>(src + r * kSS + c2);
        if (col0 + c2 + 1 <= k) *reinterpret_cast<double2*>(d) = t;
        else if (col0 + c2 <= k) d[0] = t.x;
    }
}

__device__ __forceinline__ void load_uinv(ChainCtx& cx, int I, int tid)
{
#pragma unroll 4
    for (int e = tid; e < kNB * 32; e += 256) {
        const int r = e >> 5, c2 = (e & 31) * 2;
        cp_async16(cx.Ws + r * kSS + c2, cx.UinvG + (size_t)I * kNB * kNB + r * kNB + c2);
    }
    cx.wsBlock = I;
}

// X (registers, x_gemm layout) -> Sf rows row0.., columns C0.. and optionally into a shared-memory tile
__device__ __forceinline__ void store_x(const ChainCtx& cx, const double (&x)[2][4][2], int row0, int C0, double* smemTile, int p, int nh, int g,
                                        int q)
{
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int m = (a == 0 ? 8 * p : 8 * (7 - p)) + g, nn = 32 * nh + 8 * b + 2 * q;
            if (smemTile) *reinterpret_cast<double2*>(smemTile + m * kSS + nn) = make_double2(x[a][b][0], x[a][b][1]);
            double* dst = cx.Sf + (size_t)(row0 + m) * cx.ldS + C0 + nn;
            if (C0 + nn + 1 <= cx.k) *reinterpret_cast<double2*>(dst) = make_double2(x[a][b][0], x[a][b][1]);
            else if (C0 + nn <= cx.k) dst[0] = x[a][b][0];
        }
}

// X (registers, x_gemm layout) into a shared-memory tile only
__device__ __forceinline__ void store_x_smem(const double (&x)[2][4][2], double* smemTile, int p, int nh, int g, int q)
{
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int m = (a == 0 ? 8 * p : 8 * (7 - p)) + g, nn = 32 * nh + 8 * b + 2 * q;
            *reinterpret_cast<double2*>(smemTile + m * kSS + nn) = make_double2(x[a][b][0], x[a][b][1]);
        }
}

// ---- A(I, C): off-diagonal tile -------------------------------------------------------------------------------------
__device__ void chain_task_offdiag(ChainCtx& cx, int I, int C)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3, p = w & 3, nh = w >> 2;
    const int I0 = I * kNB, C0 = C * kNB, k = cx.k;
    load_tile64(cx.Ts, cx.Sg, cx.ldS, I0, k, C0, k + 1, tid);
    cp_async_commit();
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int r = 0; r < I; ++r) {
        chain_wait(cx.ctl.xready(r, I), cx.gen, cx.dm + D_STATUS);
        chain_wait(cx.ctl.xready(r, C), cx.gen, cx.dm + D_STATUS);
        load_tile64(cx.As, cx.Sf, cx.ldS, r * kNB, k, I0, k + 1, tid);
        load_tile64(cx.Bs, cx.Sf, cx.ldS, r * kNB, k, C0, k + 1, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        tile_mma_full(cx.As, cx.Bs, p, nh, g, q, acc);
        __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
    // T = S(I, C) - acc, in place in Ts
    {
        const int m0 = 16 * p, n0 = 32 * nh;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int m = m0 + 8 * a + g, nn = n0 + 8 * b + 2 * q;
                double2 t = *reinterpret_cast<double2*>(cx.Ts + m * kSS + nn);
                t.x -= acc[a][b][0]; t.y -= acc[a][b][1];
                *reinterpret_cast<double2*>(cx.Ts + m * kSS + nn) = t;
            }
    }
    __syncthreads();
    if (C == I + 1 && C < cx.nbR) {   // handed to the critical chain: D(C) multiplies by Uinv_I^T itself
        store_tile64(cx.Ts, cx.Sg, cx.ldS, I0, k, C0, k, tid);
        chain_signal(cx.ctl.tready(I), cx.gen);
        return;
    }
    chain_wait(cx.ctl.fdone(I), cx.gen, cx.dm + D_STATUS);
    if (cx.wsBlock != I) {
        load_uinv(cx, I, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    double x[2][4][2], xd[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) x[a][b][0] = x[a][b][1] = 0.0;
    x_gemm<false>(cx.Ws, cx.Ts, cx.Ts, p, nh, g, q, x, xd);
    store_x(cx, x, I0, C0, nullptr, p, nh, g, q);
    chain_signal(cx.ctl.xready(I, C), cx.gen);
}

// ---- PD(C): diagonal tile minus the contributions of block rows 0 .. C-2 -----------------------------------------------
__device__ void chain_task_partial_diag(ChainCtx& cx, int C)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int C0 = C * kNB, k = cx.k;
    load_tile64(cx.Ts, cx.Sg, cx.ldS, C0, k, C0, k + 1, tid);
    cp_async_commit();
    double acc[5][2];
#pragma unroll
    for (int t = 0; t < 5; ++t) acc[t][0] = acc[t][1] = 0.0;
    for (int r = 0; r + 2 <= C; ++r) {
        chain_wait(cx.ctl.xready(r, C), cx.gen, cx.dm + D_STATUS);
        load_tile64(cx.As, cx.Sf, cx.ldS, r * kNB, k, C0, k + 1, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        tile_mma_upper(cx.As, w, g, q, acc);
        __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
    tile_sub_upper(cx.Ts, w, g, q, acc);
    __syncthreads();
    store_tile64(cx.Ts, cx.Sg, cx.ldS, C0, k, C0, k, tid);   // (entries below the diagonal are never read by anyone)
    chain_signal(cx.ctl.pdready(C), cx.gen);
}

// ---- D(I): the critical chain ---------------------------------------------------------------------------------------
// `prefetch`: the dedicated critical CTA keeps the tile it factors in Ts and, half way through the factorisation, starts the
// loads of its NEXT step -- T(I, I+1) into As, T'(I+1, I+1) into Bs -- if their flags are already up (they normally are: the
// helpers run ahead); the next call then finds them in shared memory (cx.pre == I + 1) and swaps Bs / Ts.
__device__ void chain_task_diag(ChainCtx& cx, int I, int* bad, int* sFlag, bool prefetch)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3, p = w & 3, nh = w >> 2;
    const int I0 = I * kNB, k = cx.k;
    const int kb = min(kNB, k - I0);
    const bool hasNu = (k - I0) < kNB;
    // T(I, I) -= X(I-1, I)^T X(I-1, I) is folded into the factorisation's left-looking tile updates (factor_tile64, Xp) unless the
    // innovation column lies in this tile (the last block: its nu entries are needed, fully updated, before the factorisation)
    const bool absorb = (I > 0) && !hasNu && CHAIN_ABSORB_UPDATE;
    const bool deferX = false;   // (publishing X(I-1, I) from inside the factorisation was measured: the helpers then deliver
                                 // T(I, I+1) too late for the prefetch and the step gets longer, profiles/r02_chain_phase_clocks.txt)
    long long* dbg = (cx.dbg != nullptr && I == 5 && tid == 0) ? cx.dbg + 48 : nullptr;   // phase clocks of one mid-chain step
    if (dbg) { dbg[0] = clock64(); dbg[9] = cx.pre == I; }
    if (cx.pre == I) {
        double* t = cx.Ts; cx.Ts = cx.Bs; cx.Bs = t;     // T'(I, I) arrived in Bs,
        t = cx.As; cx.As = cx.A2; cx.A2 = t;             // T(I-1, I) in A2
    } else {
        if (I > 0) {
            chain_wait(cx.ctl.tready(I - 1), cx.gen, cx.dm + D_STATUS, false);
            if (I >= 2) chain_wait(cx.ctl.pdready(I), cx.gen, cx.dm + D_STATUS, false);
            if (cx.wsBlock != I - 1) load_uinv(cx, I - 1, tid);
            load_tile64(cx.As, cx.Sg, cx.ldS, I0 - kNB, k, I0, k + 1, tid);
        }
        load_tile64(cx.Ts, cx.Sg, cx.ldS, I0, k, I0, k + 1, tid);
        cp_async_commit();
    }
    cx.pre = -1;
    cp_async_wait<0>();
    __syncthreads();
    if (dbg) dbg[1] = clock64();
    if (I > 0) {
        double x[2][4][2], xd[2][4][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) x[a][b][0] = x[a][b][1] = 0.0;
        x_gemm<false>(cx.Ws, cx.As, cx.As, p, nh, g, q, x, xd);
        __syncthreads();
        if (dbg) dbg[2] = clock64();
        if (deferX) {
            // X(I-1, I) K-major into As only; it goes to the factor buffer (and its flag up) from inside the factorisation
            store_x_smem(x, cx.As, p, nh, g, q);
            __syncthreads();
        } else {
            store_x(cx, x, I0 - kNB, I0, cx.As, p, nh, g, q);     // X(I-1, I): to the factor buffer and, K-major, into As
            chain_signal(cx.ctl.xready(I - 1, I), cx.gen);        // (its __syncthreads also orders the writes to As)
        }
        if (dbg) dbg[3] = clock64();
        if (!absorb) {
            double acc[5][2];
#pragma unroll
            for (int t = 0; t < 5; ++t) acc[t][0] = acc[t][1] = 0.0;
            tile_mma_upper(cx.As, w, g, q, acc);
            tile_sub_upper(cx.Ts, w, g, q, acc);
            __syncthreads();
        }
    }
    if (dbg) dbg[4] = clock64();
    // factor the diagonal tile: U_II, Uinv_I and (if nu lies in this tile) y_I
    if (hasNu && tid < kNB) cx.nu[tid] = (tid < kb) ? cx.Ts[tid * kSS + kb] : 0.0;
    __syncthreads();
    for (int e = tid; e < kNB * 32; e += 256) {   // identity outside the valid part; W starts at zero
        const int i = e >> 5, j = (e & 31) * 2;
        if (i >= kb || j + 1 >= kb) {
            if (i >= kb || j >= kb) cx.Ts[i * kSS + j] = (i == j) ? 1.0 : 0.0;
            cx.Ts[i * kSS + j + 1] = (i == j + 1) ? 1.0 : 0.0;
        }
        *reinterpret_cast<double2*>(cx.Ws + i * kSS + j) = make_double2(0.0, 0.0);
    }
    __syncthreads();
    const bool wantNext = prefetch && (I + 1 < cx.nbR);
    auto hook = [&](int b) {
        if (b == 1 && deferX) {   // publish X(I-1, I) while the in-register Cholesky of block 1 runs (As is not touched by the factor)
            store_tile64(cx.As, cx.Sf, cx.ldS, I0 - kNB, k, I0, k, tid);
            chain_signal(cx.ctl.xready(I - 1, I), cx.gen);
        }
        if (b != 5 || !wantNext) return;
        if (!*sFlag) return;   // (polled by an idle warp during block step 4, see below)
        load_tile64(cx.A2, cx.Sg, cx.ldS, I0, k, I0 + kNB, k + 1, tid);
        load_tile64(cx.Bs, cx.Sg, cx.ldS, I0 + kNB, k, I0 + kNB, k + 1, tid);
        cp_async_commit();
        cx.pre = I + 1;
    };
    // are the next step's tiles there?  Asked by a thread of the idle warps while block step 4's Cholesky runs: the two L2 round
    // trips of the acquire loads used to sit in front of block step 5 with all eight warps waiting for them
    auto idle = [&](int b) {
        if (b != 4 || tid != 255) return;
        *sFlag = wantNext && ld_acquire(cx.ctl.tready(I)) == cx.gen && (I + 1 < 2 || ld_acquire(cx.ctl.pdready(I + 1)) == cx.gen);
    };
    if (dbg) dbg[5] = clock64();
    factor_tile64(cx.Ts, cx.Ws, tid, bad, nullptr, hook, absorb ? cx.As : nullptr, (kb + 7) >> 3, idle);
    __syncthreads();
    if (dbg) dbg[6] = clock64();
    cx.wsBlock = I;
    if (tid == 0 && (*bad || cx.faultInject)) cx.dm[D_STATUS] = 4;  // EKFB_ERR_NUMERIC
    for (int e = tid; e < kNB * 32; e += 256) {
        const int i = e >> 5, j = (e & 31) * 2;
        *reinterpret_cast<double2*>(cx.UinvG + (size_t)I * kNB * kNB + i * kNB + j) = *reinterpret_cast<const double2*>(cx.Ws + i * kSS + j);
        if (i < kb && j + 1 >= i) {
            double* dst = cx.Sf + (size_t)(I0 + i) * cx.ldS + I0 + j;
            if (j >= i && j < kb) dst[0] = cx.Ts[i * kSS + j];
            if (j + 1 < kb) dst[1] = cx.Ts[i * kSS + j + 1];
        }
    }
    if (hasNu && tid < kb) {
        double s = 0.0;
        for (int pp = 0; pp <= tid; ++pp) s += cx.Ws[pp * kSS + tid] * cx.nu[pp];
        cx.Sf[(size_t)(I0 + tid) * cx.ldS + k] = s;
    }
    if (dbg) dbg[7] = clock64();
    chain_signal(cx.ctl.fdone(I), cx.gen);
    if (dbg) dbg[8] = clock64();
}

// Position p of the per-filter task queue (row-major; the D tasks are not in it): row I holds A(I, I+1) .. A(I, nbC-1), then
// PD(I+1) when I + 1 >= 2 and I + 1 < nbR.  Returns false past the end.
__device__ __forceinline__ bool chain_task_at(int pos, int nbR, int nbC, int& kind, int& I, int& C)
{
    for (int r = 0; r < nbR; ++r) {
        const int nA = nbC - r - 1, nP = (r + 1 >= 2 && r + 1 < nbR) ? 1 : 0;
        if (pos < nA) { kind = 0; I = r; C = r + 1 + pos; return true; }
        pos -= nA;
        if (pos < nP) { kind = 1; I = r; C = r + 1; return true; }
        pos -= nP;
    }
    return false;
}

// grid (G, F): G CTAs per filter.  The first CTA of a filter to arrive runs the critical chain, the others serve the task queue.
// With G == 1 the single CTA runs everything in dependency order.
__global__ void __launch_bounds__(256, 1) k_schain_fused(DevView v, int* ctlBase, int nbMax)
{
    extern __shared__ __align__(16) double csm[];
    __shared__ int sTicket, sPos, bad, sFlag;
    const int f = blockIdx.y, tid = threadIdx.x;
    grid_launch_dependents();   // the slab TRSM may be scheduled (it loads its slab of B first); it waits for this grid's completion
    grid_dependency_wait();
    int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    if (k == 0) return;
    ChainCtx cx;
    cx.Ws = csm; cx.As = cx.Ws + kNB * kSS; cx.Bs = cx.As + kNB * kSS; cx.Ts = cx.Bs + kNB * kSS; cx.nu = cx.Ts + kNB * kSS;
    cx.A2 = cx.nu + 2 * kNB;
    cx.Sg = v.S + (size_t)f * v.kmax * v.ldS;
    cx.Sf = v.Sf + (size_t)f * v.kmax * v.ldS;
    cx.UinvG = v.Uinv + (size_t)f * (v.kmax / kNB) * kNB * kNB;
    cx.ctl.base = ctlBase + (size_t)f * chain_ctl_ints(nbMax);
    cx.ctl.nbMax = nbMax;
    cx.dm = dm; cx.k = k; cx.ldS = v.ldS; cx.faultInject = v.faultInject; cx.dbg = (f == 0) ? v.dbg : nullptr;
    cx.nbR = (k + kNB - 1) / kNB; cx.nbC = (k + kNB) / kNB;
    cx.wsBlock = -1; cx.pre = -1;
    if (tid == 0) {
        sTicket = atomicAdd(cx.ctl.base + CH_TICKET, 1);
        bad = 0;
    }
    __syncthreads();
    cx.gen = ld_acquire(cx.ctl.base + CH_GEN) + 1;   // flags of this launch carry generation + 1 (k_chain_finish stores it back)
    const int ticket = sTicket;
    const int G = gridDim.x;
    if (ticket == 0) {
        if (G > 1) {
            for (int I = 0; I < cx.nbR; ++I) chain_task_diag(cx, I, &bad, &sFlag, true);
            return;
        }
        // single CTA per filter: row by row, D(I) then the row's queue tasks
        int pos = 0;
        for (int I = 0; I < cx.nbR; ++I) {
            chain_task_diag(cx, I, &bad, &sFlag, false);
            for (;;) {
                int kind, tI, tC;
                if (!chain_task_at(pos, cx.nbR, cx.nbC, kind, tI, tC) || tI != I) break;
                if (kind == 0) chain_task_offdiag(cx, tI, tC);
                else chain_task_partial_diag(cx, tC);
                ++pos;
            }
        }
        return;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) sPos = atomicAdd(cx.ctl.base + CH_QUEUE, 1);
        __syncthreads();
        int kind, tI, tC;
        if (!chain_task_at(sPos, cx.nbR, cx.nbC, kind, tI, tC)) return;
        if (kind == 0) chain_task_offdiag(cx, tI, tC);
        else chain_task_partial_diag(cx, tC);
    }
}

// after the chain (and everything that reads its flags) of one update: next generation, counters back to zero
__global__ void k_chain_finish(int* ctlBase, int nbMax, int F)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    int* c = ctlBase + (size_t)f * chain_ctl_ints(nbMax);
    if (c[CH_TICKET] == 0) return;   // this filter's chain did not run (k == 0)
    c[CH_GEN] += 1;
    c[CH_TICKET] = 0;
    c[CH_QUEUE] = 0;
}

// =====================================================================================================================
// The chain and the slab TRSM in ONE launch (single filter): the TRSM's row block J only needs block column J of U and
// Uinv_J, which the chain publishes when it raises fdone(J) -- so W^T = U^-T B is formed WHILE the chain runs, on the SMs the
// chain leaves idle, instead of after it.  grid (G): ticket 0 = critical chain, tickets 1 .. H = chain queue workers,
// tickets H+1 .. = one slab of SW columns of B each (the algorithm of k_trsm_slab, ekf_linalg.cuh, with a flag wait in
// front of every row block).  When the chain ends only the last row block of the TRSM is left.
// =====================================================================================================================
// PAD = 4: conflict-free pitch SW + 4 for the K-major slab; PAD = 0: dense pitch (two-way bank conflicts on the operand loads) for
// updates whose slab would not fit otherwise (k > 832 rows)
template <int SW, int NS, int PAD>
__device__ void trsm_slab_role(const DevView& v, int f, int slab, double* tsm, ChainCtx& cx)
{
    constexpr int SWP = SW + PAD, NT = SW / 8;
    const int* dm = cx.dm;
    const int k = cx.k, n = dm[D_N_STATE];
    const int c0 = slab * SW;
    if (c0 >= n) return;
    const int kpad = (k + kNB - 1) / kNB * kNB, steps = kpad / kNB;
    const int kz = min(kpad, (k + 15) & ~15);
    double* Xs = tsm;
    double* Us = Xs + (size_t)kpad * SWP;   // NS stages x [32][68]
    double* Ts = Us + NS * 32 * 68;         // [64][SWP]
    double* red = Ts + kNB * SWP;           // [8][SW]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    double* Bg = v.Bu + (size_t)f * v.kmax * v.ld;
    const double* Sg = cx.Sf;
    const double* Uinv = cx.UinvG;
    for (int e = tid; e < kpad * SW; e += blockDim.x) {
        const int r = e / SW, c = e % SW;
        Xs[(size_t)r * SWP + c] = (r < k && c0 + c < n) ? Bg[(size_t)r * v.ld + c0 + c] : 0.0;
    }
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
            // The accumulation over the earlier row blocks only needs the tiles X(r, Ji), which the chain publishes well before it
            // has factored tile (Ji, Ji): wait for each 64-row block of U as its first chunk comes up, and for the diagonal
            // factor (Uinv_Ji) only in front of the last two chunks -- so when the chain ends, the TRSM of the last row block
            // has four chunks left instead of its whole left-looking pass
            const int nUi = (Ji * kNB) / 32;
            if (chi < nUi) {
                if ((chi & 1) == 0) chain_wait(cx.ctl.xready(chi >> 1, Ji), cx.gen, cx.dm + D_STATUS);
            } else if (chi == nUi)
                chain_wait(cx.ctl.fdone(Ji), cx.gen, cx.dm + D_STATUS);
            issue(Ji, chi, stagei);
            if (++chi == (Ji * kNB) / 32 + 2) { ++Ji; chi = 0; }
            stagei = (stagei + 1 == NS) ? 0 : stagei + 1;
        }
        cp_async_commit();
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
        issue_next();
        cp_async_wait<NS - 1>();
        __syncthreads();
        const double* Uc = Us + stage * 32 * 68;
        if (ch == nU) {
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
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = 8 * w + g, c = 8 * b + 2 * q + e;
                    Xs[(size_t)(J0 + m) * SWP + c] = (m < kb) ? acc[b][e] : 0.0;
                }
        }
        __syncthreads();
        if (ch == nCh - 1) {
            // W^T rows of this block back to global right away (under the chain, not behind it); rows k .. end of the last 16-row
            // chunk go back as zeros (the TMA-fed downdate reads whole chunks).  Bu is scratch: if a later pivot fails, the rows
            // already written are never read (every consumer checks the status first).
            const int rEnd = min(J0 + kNB, kz);
            for (int e = tid; e < (rEnd - J0) * SW; e += blockDim.x) {
                const int r = J0 + e / SW, c = e % SW;
                if (c0 + c < n) Bg[(size_t)r * v.ld + c0 + c] = Xs[(size_t)r * SWP + c];
            }
        }
        stage = (stage + 1 == NS) ? 0 : stage + 1;
        J = Jn;
        ch = chn;
    }
    cp_async_wait<0>();
    // y = U^-T nu lives in column k of the factor buffer: rows of block I < the nu column tile come from A(I, Cnu)
    const int Cnu = cx.nbC - 1;
    for (int I = 0; I < cx.nbR; ++I)
        if (Cnu > I) chain_wait(cx.ctl.xready(I, Cnu), cx.gen, cx.dm + D_STATUS);
    if (__ldcg(dm + D_STATUS) != 0) return;   // not positive definite: the update is skipped
    if (lane < SW) {
        double s = 0.;
        for (int r = w; r < k; r += 8) s += Xs[(size_t)r * SWP + lane] * __ldcg(Sg + (size_t)r * v.ldS + k);
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

template <int SW, int NS, int PAD>
__device__ __forceinline__ void update_fused_body(const DevView& v, int* ctlBase, int nbMax, int nSlabs, double* csm)
{
    __shared__ int sTicket, sPos, bad, sFlag;
    const int f = 0, tid = threadIdx.x;
    grid_launch_dependents();
    grid_dependency_wait();
    int* dm = fdims(v, f);
    const int k = 2 * dm[D_ULIST];
    if (k == 0) return;
    ChainCtx cx;
    cx.Ws = csm; cx.As = cx.Ws + kNB * kSS; cx.Bs = cx.As + kNB * kSS; cx.Ts = cx.Bs + kNB * kSS; cx.nu = cx.Ts + kNB * kSS;
    cx.A2 = cx.nu + 2 * kNB;
    cx.Sg = v.S; cx.Sf = v.Sf; cx.UinvG = v.Uinv;
    cx.ctl.base = ctlBase;
    cx.ctl.nbMax = nbMax;
    cx.dm = dm; cx.k = k; cx.ldS = v.ldS; cx.faultInject = v.faultInject; cx.dbg = (f == 0) ? v.dbg : nullptr;
    cx.nbR = (k + kNB - 1) / kNB; cx.nbC = (k + kNB) / kNB;
    cx.wsBlock = -1; cx.pre = -1;
    if (tid == 0) {
        sTicket = atomicAdd(cx.ctl.base + CH_TICKET, 1);
        bad = 0;
    }
    __syncthreads();
    cx.gen = ld_acquire(cx.ctl.base + CH_GEN) + 1;
    const int ticket = sTicket;
    const int H = (int)gridDim.x - 1 - nSlabs;   // chain queue workers
    if (ticket == 0) {
        for (int I = 0; I < cx.nbR; ++I) chain_task_diag(cx, I, &bad, &sFlag, true);
        return;
    }
    if (ticket > H) {
        trsm_slab_role<SW, NS, PAD>(v, f, ticket - H - 1, csm, cx);
        return;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) sPos = atomicAdd(cx.ctl.base + CH_QUEUE, 1);
        __syncthreads();
        int kind, tI, tC;
        if (!chain_task_at(sPos, cx.nbR, cx.nbC, kind, tI, tC)) return;
        if (kind == 0) chain_task_offdiag(cx, tI, tC);
        else chain_task_partial_diag(cx, tC);
    }
}

// The block that finishes last (every flag of the dataflow has been consumed by then) advances the chain's generation and
// clears its counters -- what k_chain_finish does as a launch of its own -- and applies the state correction
// (state_apply_tail): the update is ONE launch from S to x.
template <int SW, int NS, int PAD = 4>
__global__ void __launch_bounds__(256, 1) k_update_fused(DevView v, int* ctlBase, int nbMax, int nSlabs)
{
    extern __shared__ __align__(16) double csm[];
    update_fused_body<SW, NS, PAD>(v, ctlBase, nbMax, nSlabs, csm);
    if (!last_block_done(fdims(v, 0) + D_TICKET_UPD, (int)gridDim.x)) return;
    if (threadIdx.x == 0 && ctlBase[CH_TICKET] != 0) {
        ctlBase[CH_GEN] += 1;
        ctlBase[CH_TICKET] = 0;
        ctlBase[CH_QUEUE] = 0;
    }
    state_apply_tail(v, 0);
}

}  // namespace ekf
