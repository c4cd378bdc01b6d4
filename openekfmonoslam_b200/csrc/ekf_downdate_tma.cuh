// ekf_downdate_tma.cuh -- U3 (+ the symmetrise of U4): the covariance downdate  P -= W W^T  (E/Update.cpp:214-218, 307), fed by
// the TMA engine, for one filter or a batch.  Same contraction and tiling as k_downdate64 (ekf_linalg.cuh: lower 64x64 tiles,
// 4 warps of 32x32, FP64 DMMA m8n8k4), different plumbing:
//   * persistent CTAs (two per SM) walk the list of (filter, tile) items, so prologues / epilogues of one tile overlap the
//     arithmetic of the CTA that shares the SM, and the operand ring never drains inside a CTA;
//   * a PRODUCER WARP feeds everything with tensor-map TMA (cp.async.bulk.tensor.3d + mbarrier complete_tx; the third
//     coordinate is the filter): the P tile (64 x 64) into a staging buffer and the K-major operand chunks of W^T (16 rows x
//     64 columns for the tile's rows and for its columns) into a 3-stage ring -- full / empty mbarriers, no __syncthreads
//     and no cp.async address arithmetic in the four CONSUMER warps;
//   * 128-byte swizzled boxes (16 doubles wide) + a row permutation inside the 16-row chunk ({0,2,4,6}, {1,3,5,7}, ... per
//     k-step) make every DMMA fragment load bank-conflict free without padding (TMA writes dense boxes);
//   * the updated tile leaves through shared memory as TMA tensor stores, and its mirror image as a second set
//     (transposed in shared memory), so both triangles are written as full 128-byte rows: P stays exactly symmetric
//     (the two stores carry bit-identical values) and no thread issues a strided global store.  Tiles that cross the edge
//     of their filter's matrix (n is not a multiple of 64) are stored with predicated global stores instead, so nothing
//     beyond row / column n is ever written.
// Rows of W^T beyond a filter's K inside its last 16-row chunk must be zero: the slab TRSM writes them (k_trsm_slab).
// Algorithmic work per launch sum_f n_f (n_f + 1) K_f flop, minimum traffic 16 n^2 bytes per filter.
#pragma once

#include <cuda.h>

#include "ekf_linalg.cuh"
#include "ekf_ncc.cuh"   // mbarrier helpers

namespace ekf {

constexpr int kTdStages = 3;
constexpr int kTdTileBytes = 64 * 64 * 8;                 // 32 KB
constexpr int kTdChunkBytes = 16 * 64 * 8;                // one operand chunk: 16 rows x 64 columns
constexpr int kTdSmemBytes = 2 * kTdTileBytes + kTdStages * 2 * kTdChunkBytes + 64;

struct TdMaps {
    CUtensorMap P;   // dims (nmax cols, nmax rows, F filters), box (16 | 64, 64, 1)
    CUtensorMap W;   // dims (ld cols, kmax rows, F filters),  box (16 | 64, 16, 1)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(smem)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<unsigned long long>(map)),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// byte offset of element (row r, column c) of a tile / chunk whose columns come as 16-double boxes of `rows` rows each:
// SWZ: box (c / 16) at (c / 16) * rows * 128, row r at r * 128, 16-byte chunk (c % 16) / 2 XORed with r % 8 (128-byte swizzle);
// else dense rows of 64 doubles
template <bool SWZ>
__device__ __forceinline__ int td_off(int r, int c, int rows)
{
    if (SWZ) return (c >> 4) * rows * 128 + r * 128 + (((((c & 15) >> 1) ^ (r & 7)) << 4) | ((c & 1) << 3));
    return (r * 64 + c) * 8;
}

// work item -> filter, tile coordinates; false if the item has nothing to do (its filter has no update, or the tile lies
// outside that filter's matrix)
struct TdItem { int f, tm0, tn0, n, nk; bool diag, edge; };
__device__ __forceinline__ bool td_item(const DevView& v, int item, int tilesMax, TdItem& o)
{
    o.f = item / tilesMax;
    const int t = item - o.f * tilesMax;
    const int* dm = fdims(v, o.f);
    const int K = 2 * dm[D_ULIST];
    o.n = dm[D_N_STATE];
    if (K == 0 || dm[D_STATUS] != 0) return false;
    int I = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while (I * (I + 1) / 2 > t) --I;
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    const int J = t - I * (I + 1) / 2;
    o.tm0 = I * 64; o.tn0 = J * 64;
    if (o.tm0 >= o.n) return false;
    o.nk = (K + 15) >> 4;
    o.diag = (I == J);
    o.edge = (o.tm0 + 64 > o.n);      // (tn0 <= tm0: a tile that crosses the edge in its columns also crosses it in its rows)
    return true;
}

// grid (min(items, 2 * SMs)), 160 threads: warps 0-3 consume, warp 4 produces.  Dynamic shared memory kTdSmemBytes, 1024-byte aligned.
template <bool SWZ>
__global__ void __launch_bounds__(160, 2) k_downdate_tma(DevView v, const __grid_constant__ TdMaps maps, int tilesMax)
{
    extern __shared__ __align__(1024) unsigned char tds[];
    unsigned char* Pbuf = tds;
    unsigned char* Obuf = tds + kTdTileBytes;
    unsigned char* ring = tds + 2 * kTdTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tds + 2 * kTdTileBytes + kTdStages * 2 * kTdChunkBytes);
    uint64_t* full = bars;                  // [kTdStages]
    uint64_t* empty = bars + kTdStages;     // [kTdStages]
    uint64_t* pfull = bars + 2 * kTdStages;
    uint64_t* pempty = pfull + 1;
    grid_dependency_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kTdStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        mbar_init(pfull, 1);
        mbar_init(pempty, 4);
    }
    __syncthreads();
    const int items = v.F * tilesMax;
    constexpr int BOXES = SWZ ? 4 : 1, BOXC = SWZ ? 16 : 64;

    if (warp == 4) {
        // ---------------- producer: one elected lane issues every TMA load ----------------
        if (lane != 0) return;
        int c = 0, it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            TdItem w;
            if (!td_item(v, item, tilesMax, w)) continue;
            mbar_wait(pempty, (it & 1) ^ 1);
            mbar_expect_tx(pfull, kTdTileBytes);
#pragma unroll
            for (int j = 0; j < BOXES; ++j) tma_load_3d(Pbuf + j * (kTdTileBytes / BOXES), &maps.P, w.tn0 + j * BOXC, w.tm0, w.f, pfull);
            for (int kt = 0; kt < w.nk; ++kt, ++c) {
                const int s = c % kTdStages;
                mbar_wait(empty + s, ((c / kTdStages) & 1) ^ 1);
                mbar_expect_tx(full + s, w.diag ? kTdChunkBytes : 2 * kTdChunkBytes);
                unsigned char* A = ring + s * 2 * kTdChunkBytes;
#pragma unroll
                for (int j = 0; j < BOXES; ++j) tma_load_3d(A + j * (kTdChunkBytes / BOXES), &maps.W, w.tm0 + j * BOXC, kt * 16, w.f, full + s);
                if (!w.diag) {
#pragma unroll
                    for (int j = 0; j < BOXES; ++j)
                        tma_load_3d(A + kTdChunkBytes + j * (kTdChunkBytes / BOXES), &maps.W, w.tn0 + j * BOXC, kt * 16, w.f, full + s);
                }
            }
            ++it;
        }
        return;
    }

    // ---------------- consumers: 4 warps, warp tile 32 x 32 ----------------
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
    int c = 0, it = 0;
    bool storePending = false;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        TdItem w;
        if (!td_item(v, item, tilesMax, w)) continue;
        const int tm0 = w.tm0, tn0 = w.tn0, n = w.n;
        const bool diag = w.diag;
        // warp tiles with no element of the lower triangle inside the matrix do no arithmetic
        const bool idle = (diag && wn > wm) || (tm0 + wm * 32 >= n) || (tn0 + wn * 32 >= n);
        // accumulators start at -P (the tile arrived by TMA)
        mbar_wait(pfull, it & 1);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double2 p2 = *reinterpret_cast<const double2*>(Pbuf + td_off<SWZ>(wm * 32 + a * 8 + g, wn * 32 + b * 8 + 2 * q, 64));
                acc[a][b][0] = -p2.x;
                acc[a][b][1] = -p2.y;
            }
        __syncwarp();
        if (lane == 0) mbar_arrive(pempty);     // the producer may fetch the next tile's P
        for (int kt = 0; kt < w.nk; ++kt, ++c) {
            const int s = c % kTdStages;
            mbar_wait(full + s, (c / kTdStages) & 1);
            const unsigned char* A = ring + s * 2 * kTdChunkBytes;
            const unsigned char* B = diag ? A : A + kTdChunkBytes;
            if (!idle) {
#pragma unroll
                for (int st = 0; st < 4; ++st) {
                    // rows of this k-step: {0,2,4,6}, {1,3,5,7}, {8,10,12,14}, {9,11,13,15}.  An 8-byte fragment load is served per
                    // half-warp (g = 0..3 | 4..7: two 16-byte chunks of each of the four rows); with the swizzle XORing the
                    // chunk index with (row & 7), rows of equal parity put those chunk pairs on eight different chunks, i.e.
                    // all 32 banks exactly once ({0,1,4,5} collided pairwise: ncu counted one conflict per wavefront)
                    const int r = ((st >> 1) << 3) + (st & 1) + (q << 1);
                    double af[4], bf[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) af[a] = *reinterpret_cast<const double*>(A + td_off<SWZ>(r, wm * 32 + a * 8 + g, 16));
#pragma unroll
                    for (int b = 0; b < 4; ++b) bf[b] = *reinterpret_cast<const double*>(B + td_off<SWZ>(r, wn * 32 + b * 8 + g, 16));
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) dmma8x8x4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        ++it;
        if (w.edge) {
            // the tile crosses the edge of this filter's matrix: predicated stores from the registers (lower part + mirror)
            if (!idle) {
                double* P = v.P + (size_t)w.f * v.nmax * v.ld;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int gm = tm0 + wm * 32 + a * 8 + g;
                    if (gm >= n) continue;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int gn = tn0 + wn * 32 + b * 8 + 2 * q + e;
                            if (gn > gm || gn >= n) continue;
                            const double val = -acc[a][b][e];
                            P[(size_t)gm * v.ld + gn] = val;
                            if (gn != gm) P[(size_t)gn * v.ld + gm] = val;
                        }
                }
            }
            continue;
        }
        // ---- the tile (and its mirror image) through shared memory, TMA tensor stores ----
        if (storePending) {
            if (tid == 0) tma_store_wait_read();     // the previous tile's last store has read Obuf
            bar_consumers();
        }
        if (!idle) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int m = wm * 32 + a * 8 + g, nn = wn * 32 + b * 8 + 2 * q;
                    const double v0 = -acc[a][b][0], v1 = -acc[a][b][1];
                    *reinterpret_cast<double2*>(Obuf + td_off<SWZ>(m, nn, 64)) = make_double2(v0, v1);
                    if (diag) {   // the diagonal tile is stored once, whole: the idle warp's block is the mirror of warp (1, 0)'s
                        *reinterpret_cast<double*>(Obuf + td_off<SWZ>(nn, m, 64)) = v0;
                        *reinterpret_cast<double*>(Obuf + td_off<SWZ>(nn + 1, m, 64)) = v1;
                    }
                }
        }
        fence_async_smem();
        bar_consumers();
        if (tid == 0) {
#pragma unroll
            for (int j = 0; j < BOXES; ++j) tma_store_3d(&maps.P, Obuf + j * (kTdTileBytes / BOXES), tn0 + j * BOXC, tm0, w.f);
            tma_store_commit();
        }
        storePending = true;
        if (!diag) {
            if (tid == 0) tma_store_wait_read();
            bar_consumers();
            if (!idle) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int m = wm * 32 + a * 8 + g, nn = wn * 32 + b * 8 + 2 * q;
                        *reinterpret_cast<double*>(Obuf + td_off<SWZ>(nn, m, 64)) = -acc[a][b][0];
                        *reinterpret_cast<double*>(Obuf + td_off<SWZ>(nn + 1, m, 64)) = -acc[a][b][1];
                    }
            }
            fence_async_smem();
            bar_consumers();
            if (tid == 0) {
#pragma unroll
                for (int j = 0; j < BOXES; ++j) tma_store_3d(&maps.P, Obuf + j * (kTdTileBytes / BOXES), tm0 + j * BOXC, tn0, w.f);
                tma_store_commit();
            }
        }
    }
    if (tid == 0) tma_store_wait_all();
}

}  // namespace ekf
