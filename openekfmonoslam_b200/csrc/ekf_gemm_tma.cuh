// ekf_gemm_tma.cuh -- W^T = U^-T B for updates too large for the shared-memory slab TRSM (k > ~1150 rows: the 1280x720
// stress configuration with 1000+ features), as a blocked left-looking solve on the GLOBAL-memory resident B
// (reference: E/Update.cpp:105-141 -- the n x k gain behind S.inv(), here never formed):
//     for every 64-row block J:   T   = B_J - U(0:J0, J)^T X(0:J0)        C -= A^T B, contraction over the J0 rows already solved
//                                 X_J = Uinv_J^T T                         C  = A^T B, contraction over the 64 rows of the block
// Both are the "TN" contraction of the covariance downdate with K-major operands, so the kernel below is the downdate's
// pipeline (ekf_downdate_tma.cuh) without the symmetry: one CTA per 64 x 64 tile of the block row, four consumer warps of
// 32 x 32 on the FP64 tensor pipe (DMMA m8n8k4), one producer lane feeding 16-row operand chunks through a 3-stage ring by
// tensor-map TMA (128-byte swizzled boxes, conflict-free fragment loads), the C tile by TMA in and out.  Both steps run in
// place on the rows of Bu: a CTA reads its own column tile completely before it stores it, and no other CTA touches it.
// Columns beyond n are zero on load (tensor-map bounds) and clipped on store.
// The long contraction (J0 rows) is what makes this efficient where the rank-64 right-looking update of the generic path
// (k_gemm_tn<0>, 8.5 TFLOP/s at n = 12013) is not: B is read once per block row instead of once per block step.
#pragma once

#include "ekf_downdate_tma.cuh"

namespace ekf {

struct GemmMaps {
    CUtensorMap A;     // operand A[r][m]: dims (cols, rows), box (16 | 64, 16)
    CUtensorMap B;     // operand B[r][n]: dims (n, rows),    box (16 | 64, 16)
    CUtensorMap C;     // C[m][n] (in and out): dims (n, rows), box (16 | 64, 64)
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<unsigned long long>(map)),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1)
                 : "memory");
}

constexpr int kGtSmemBytes = kTdTileBytes + kTdStages * 2 * kTdChunkBytes + 64;

// C(cRow0 .. +63, tile columns) = (sub ? C - A^T B : A^T B) with A = rows aRow0 .. aRow0 + K - 1, columns aCol0 .. +63 of map A
// and B = rows bRow0 .. bRow0 + K - 1 of map B.  K is a multiple of 16.  grid (ceil(n / 64)), 160 threads, smem kGtSmemBytes.
__global__ void __launch_bounds__(160, 2) k_gemm_tn_tma(const __grid_constant__ GemmMaps maps, int aCol0, int aRow0, int bRow0, int cRow0,
                                                        int K, int sub)
{
    extern __shared__ __align__(1024) unsigned char gts[];
    unsigned char* Cbuf = gts;
    unsigned char* ring = gts + kTdTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gts + kTdTileBytes + kTdStages * 2 * kTdChunkBytes);
    uint64_t* full = bars;                  // [kTdStages]
    uint64_t* empty = bars + kTdStages;     // [kTdStages]
    uint64_t* cfull = bars + 2 * kTdStages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kTdStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        mbar_init(cfull, 1);
    }
    __syncthreads();
    grid_dependency_wait();
    const int tn0 = blockIdx.x * 64, nk = K >> 4;
    if (warp == 4) {
        if (lane != 0) return;
        if (sub) {
            mbar_expect_tx(cfull, kTdTileBytes);
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(Cbuf + j * (kTdTileBytes / 4), &maps.C, tn0 + j * 16, cRow0, cfull);
        }
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % kTdStages;
            mbar_wait(empty + s, ((kt / kTdStages) & 1) ^ 1);
            mbar_expect_tx(full + s, 2 * kTdChunkBytes);
            unsigned char* A = ring + s * 2 * kTdChunkBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(A + j * (kTdChunkBytes / 4), &maps.A, aCol0 + j * 16, aRow0 + kt * 16, full + s);
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(A + kTdChunkBytes + j * (kTdChunkBytes / 4), &maps.B, tn0 + j * 16, bRow0 + kt * 16, full + s);
        }
        return;
    }
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % kTdStages;
        mbar_wait(full + s, (kt / kTdStages) & 1);
        const unsigned char* A = ring + s * 2 * kTdChunkBytes;
        const unsigned char* B = A + kTdChunkBytes;
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            const int r = ((st >> 1) << 3) + (st & 1) + (q << 1);   // (row order of the downdate: conflict-free in the swizzled box)
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = *reinterpret_cast<const double*>(A + td_off<true>(r, wm * 32 + a * 8 + g, 16));
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = *reinterpret_cast<const double*>(B + td_off<true>(r, wn * 32 + b * 8 + g, 16));
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    if (sub) mbar_wait(cfull, 0);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int m = wm * 32 + a * 8 + g, nn = wn * 32 + b * 8 + 2 * q;
            double2* pc = reinterpret_cast<double2*>(Cbuf + td_off<true>(m, nn, 64));
            double2 o = make_double2(acc[a][b][0], acc[a][b][1]);
            if (sub) {
                const double2 c2 = *pc;
                o.x = c2.x - o.x;
                o.y = c2.y - o.y;
            }
            *pc = o;
        }
    fence_async_smem();
    bar_consumers();
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_store_2d(&maps.C, Cbuf + j * (kTdTileBytes / 4), tn0 + j * 16, cRow0);
        tma_store_commit();
        tma_store_wait_all();
    }
}

// dx[c] = sum_{r < k} W^T[r][c] y[r]  (K nu = W y, E/Update.cpp:136-141) for the global-memory TRSM; rows k .. end of the
// last 16-row chunk of W^T are cleared for the TMA-fed downdate.  grid (ceil(n / 256)), 256 threads, single filter.
__global__ void __launch_bounds__(256) k_wy(DevView v)
{
    grid_dependency_wait();
    const int* dm = fdims(v, 0);
    const int k = 2 * dm[D_ULIST], n = dm[D_N_STATE];
    if (k == 0 || dm[D_STATUS] != 0) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s0 = 0., s1 = 0.;
    int r = 0;
    for (; r + 1 < k; r += 2) {
        s0 += v.Bu[(size_t)r * v.ld + c] * v.Sf[(size_t)r * v.ldS + k];
        s1 += v.Bu[(size_t)(r + 1) * v.ld + c] * v.Sf[(size_t)(r + 1) * v.ldS + k];
    }
    if (r < k) s0 += v.Bu[(size_t)r * v.ld + c] * v.Sf[(size_t)r * v.ldS + k];
    v.dx[c] = s0 + s1;
    const int kz = (k + 15) & ~15;
    for (int rr = k; rr < kz; ++rr) v.Bu[(size_t)rr * v.ld + c] = 0.0;
}

}  // namespace ekf
