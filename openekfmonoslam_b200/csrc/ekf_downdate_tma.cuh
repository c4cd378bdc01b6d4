// ekf_downdate_tma.cuh -- U3 (+ the symmetrise of U4): the covariance downdate  P -= W W^T  (E/Update.cpp:214-218, 307), fed by
// the TMA engine, for one filter or a batch.  Same contraction and tiling as k_downdate64 (ekf_linalg.cuh: lower 64x64 tiles,
// 4 warps of 32x32, FP64 DMMA m8n8k4), different plumbing:
//   * persistent CTAs (two per SM) walk the list of (filter, tile) items, so prologues / epilogues of one tile overlap the
//     arithmetic of the CTA that shares the SM, and the operand ring never drains inside a CTA;
//   * the work items come from a dynamic queue (one atomic per tile), heavy tiles first, so the last wave ends evenly;
//   * a PRODUCER WARP feeds everything with tensor-map TMA (cp.async.bulk.tensor.3d + mbarrier complete_tx; the third
//     coordinate is the filter): the K-major operand chunks of W^T (16 rows x 64 columns for the tile's rows and for its
//     columns) into a 3-stage ring -- full / empty mbarriers, no __syncthreads and no cp.async address arithmetic in the four
//     CONSUMER warps -- and the P tile (64 x 64) into the output buffer, where it is only needed when the products are done;
//   * 128-byte swizzled boxes (16 doubles wide) + a row permutation inside the 16-row chunk ({0,2,4,6}, {1,3,5,7}, ... per
//     k-step) make every DMMA fragment load bank-conflict free without padding (TMA writes dense boxes);
//   * the updated tile (P - W W^T, formed in place) leaves as TMA tensor stores, and its mirror image as a second set from a
//     second buffer (transposed in shared memory) issued back to back, so both triangles are written as full 128-byte rows: P stays exactly symmetric
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
constexpr int kTdSlots = 4;                               // work-item slots the producer hands to the consumers
constexpr int kTdTileBytes = 64 * 64 * 8;                 // 32 KB
constexpr int kTdChunkBytes = 16 * 64 * 8;                // one operand chunk: 16 rows x 64 columns
constexpr int kTdSmemBytes = 2 * kTdTileBytes + kTdStages * 2 * kTdChunkBytes + 128;

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

// queue position -> filter, tile coordinates; false if the item has nothing to do (its filter has no update, or the tile lies
// outside that filter's matrix).  Within a filter the positions run over the strictly-lower tiles first (row-major: neighbours
// in the queue share their row's operand chunks in L2), then the diagonal tiles (3/4 of the arithmetic), then the last tile
// row (mostly outside the matrix when n is not a multiple of 64): the cheap items come last, so the dynamic queue ends evenly.
struct TdItem { int f, tm0, tn0, n, nk; bool diag, edge; };
__device__ __forceinline__ bool td_item(const DevView& v, int pos, int tilesMax, int nT, TdItem& o)
{
    o.f = pos / tilesMax;
    const int t = pos - o.f * tilesMax;
    const int* dm = fdims(v, o.f);
    const int K = 2 * dm[D_ULIST];
    o.n = dm[D_N_STATE];
    if (K == 0 || dm[D_STATUS] != 0) return false;
    const int nA = (nT - 1) * (nT - 2) / 2, nB = nT - 1;
    int I, J;
    if (t < nA) {
        int Ip = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
        while (Ip * (Ip + 1) / 2 > t) --Ip;
        while ((Ip + 1) * (Ip + 2) / 2 <= t) ++Ip;
        J = t - Ip * (Ip + 1) / 2;
        I = Ip + 1;
    } else if (t < nA + nB) {
        I = J = t - nA;
    } else {
        I = nT - 1;
        J = t - nA - nB;
    }
    o.tm0 = I * 64; o.tn0 = J * 64;
    if (o.tm0 >= o.n) return false;
    o.nk = (K + 15) >> 4;
    o.diag = (I == J);
    o.edge = (o.tm0 + 64 > o.n);      // (tn0 <= tm0: a tile that crosses the edge in its columns also crosses it in its rows)
    return true;
}

// grid (min(items, 2 * SMs)), 160 threads: warps 0-3 consume, warp 4 produces.  Dynamic shared memory kTdSmemBytes, 1024-byte
// aligned.  queue[0] = next position of the work queue, queue[1] = CTAs that have drained it (the last one resets both).
// Per tile:  producer: item -> slot, operand chunks into the ring; P tile into Obuf once the previous tile's stores have read it
//            consumers: acc = sum_k W[k][m] W[k][n] over the ring (from zero); then Obuf (P) - acc in place, the mirror image
//            transposed into Mbuf, two TMA tensor stores back to back; thread 0 waits for the stores to have READ shared
//            memory and hands Obuf back to the producer, the other warps go straight on to the next tile's ring.
// probe (timing only, wrong results): 1 = no DMMA, 2 = no stores, 4 = no mirror store.
template <bool SWZ>
__global__ void __launch_bounds__(160, 2) k_downdate_tma(DevView v, const __grid_constant__ TdMaps maps, int tilesMax, int nT, int* queue,
                                                         int probe = 0)
{
    extern __shared__ __align__(1024) unsigned char tds[];
    unsigned char* Obuf = tds;
    unsigned char* Mbuf = tds + kTdTileBytes;
    unsigned char* ring = tds + 2 * kTdTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tds + 2 * kTdTileBytes + kTdStages * 2 * kTdChunkBytes);
    uint64_t* full = bars;                  // [kTdStages]
    uint64_t* empty = bars + kTdStages;     // [kTdStages]
    uint64_t* pfull = bars + 2 * kTdStages; // the P tile has landed in Obuf
    uint64_t* ofree = pfull + 1;            // the stores of the tile have read Obuf / Mbuf
    uint64_t* ifull = ofree + 1;            // [kTdSlots] a work item is in its slot
    volatile int* slots = reinterpret_cast<volatile int*>(ifull + kTdSlots);   // [kTdSlots]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kTdStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        mbar_init(pfull, 1);
        mbar_init(ofree, 1);
        for (int s = 0; s < kTdSlots; ++s) mbar_init(ifull + s, 1);
    }
    __syncthreads();
    grid_dependency_wait();
    const int items = v.F * tilesMax;
    constexpr int BOXES = SWZ ? 4 : 1, BOXC = SWZ ? 16 : 64;

    if (warp == 4) {
        // ---------------- producer: one elected lane takes the work items and issues every TMA load ----------------
        if (lane != 0) return;
        int c = 0;
        for (int it = 0;; ++it) {
            TdItem w;
            int pos;
            for (;;) {
                pos = atomicAdd(queue, 1);
                if (pos >= items) { pos = -1; break; }
                if (td_item(v, pos, tilesMax, nT, w)) break;
            }
            slots[it & (kTdSlots - 1)] = pos;
            mbar_arrive(ifull + (it & (kTdSlots - 1)));
            if (pos < 0) break;
            const int ktP = min(w.nk - 1, kTdStages - 1);   // the P tile is requested behind the first chunks: it is needed last
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
                if (kt == ktP) {
                    if (it > 0) mbar_wait(ofree, (it - 1) & 1);
                    mbar_expect_tx(pfull, kTdTileBytes);
#pragma unroll
                    for (int j = 0; j < BOXES; ++j) tma_load_3d(Obuf + j * (kTdTileBytes / BOXES), &maps.P, w.tn0 + j * BOXC, w.tm0, w.f, pfull);
                }
            }
        }
        __threadfence();
        if (atomicAdd(queue + 1, 1) == (int)gridDim.x - 1) {   // every CTA has drained the queue: ready for the next launch
            queue[0] = 0;
            queue[1] = 0;
        }
        return;
    }

    // ---------------- consumers: 4 warps, warp tile 32 x 32 ----------------
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
    int c = 0;
    for (int it = 0;; ++it) {
        mbar_wait(ifull + (it & (kTdSlots - 1)), (it / kTdSlots) & 1);
        const int pos = slots[it & (kTdSlots - 1)];
        if (pos < 0) break;
        TdItem w;
        td_item(v, pos, tilesMax, nT, w);
        const int tm0 = w.tm0, tn0 = w.tn0, n = w.n;
        const bool diag = w.diag;
        // warp tiles with no element of the lower triangle inside the matrix do no arithmetic
        const bool idle = (diag && wn > wm) || (tm0 + wm * 32 >= n) || (tn0 + wn * 32 >= n);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        // (software-pipelining the fragment loads across k-steps and chunk boundaries by hand was measured: 215 -> 230 us at
        // k = 640; the compiler's own schedule of the unrolled chunk is better)
        for (int kt = 0; kt < w.nk; ++kt, ++c) {
            const int s = c % kTdStages;
            mbar_wait(full + s, (c / kTdStages) & 1);
            const unsigned char* A = ring + s * 2 * kTdChunkBytes;
            const unsigned char* B = diag ? A : A + kTdChunkBytes;
            if (!idle && !(probe & 1)) {
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
        // ---- epilogue: P - acc ----
        mbar_wait(pfull, it & 1);
        if (w.edge || (probe & 2)) {
            // the tile crosses the edge of this filter's matrix: predicated stores from the registers (lower part + mirror)
            if (!idle && !(probe & 2)) {
                double* P = v.P + (size_t)w.f * v.nmax * v.ld;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int ml = wm * 32 + a * 8 + g, gm = tm0 + ml;
                    if (gm >= n) continue;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int nl = wn * 32 + b * 8 + 2 * q;
                        const double2 p2 = *reinterpret_cast<const double2*>(Obuf + td_off<SWZ>(ml, nl, 64));
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int gn = tn0 + nl + e;
                            if (gn > gm || gn >= n) continue;
                            const double val = (e ? p2.y : p2.x) - acc[a][b][e];
                            P[(size_t)gm * v.ld + gn] = val;
                            if (gn != gm) P[(size_t)gn * v.ld + gm] = val;
                        }
                    }
                }
            }
            bar_consumers();                         // every warp has read Obuf
            if (tid == 0) mbar_arrive(ofree);
            continue;
        }
        if (!idle) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int m = wm * 32 + a * 8 + g, nn = wn * 32 + b * 8 + 2 * q;
                    double2* po = reinterpret_cast<double2*>(Obuf + td_off<SWZ>(m, nn, 64));
                    const double2 p2 = *po;
                    const double v0 = p2.x - acc[a][b][0], v1 = p2.y - acc[a][b][1];
                    // the diagonal tile is stored once, whole, from Mbuf (the idle warp's block is the mirror of warp (1, 0)'s; mirrored
                    // elements carry bit-identical values); any other tile goes back in place and its transpose into Mbuf
                    if (diag) *reinterpret_cast<double2*>(Mbuf + td_off<SWZ>(m, nn, 64)) = make_double2(v0, v1);
                    else *po = make_double2(v0, v1);
                    *reinterpret_cast<double*>(Mbuf + td_off<SWZ>(nn, m, 64)) = v0;
                    *reinterpret_cast<double*>(Mbuf + td_off<SWZ>(nn + 1, m, 64)) = v1;
                }
        }
        fence_async_smem();
        bar_consumers();
        if (tid == 0) {
            if (diag) {
#pragma unroll
                for (int j = 0; j < BOXES; ++j) tma_store_3d(&maps.P, Mbuf + j * (kTdTileBytes / BOXES), tn0 + j * BOXC, tm0, w.f);
            } else {
#pragma unroll
                for (int j = 0; j < BOXES; ++j) tma_store_3d(&maps.P, Obuf + j * (kTdTileBytes / BOXES), tn0 + j * BOXC, tm0, w.f);
                if (!(probe & 4)) {
#pragma unroll
                    for (int j = 0; j < BOXES; ++j) tma_store_3d(&maps.P, Mbuf + j * (kTdTileBytes / BOXES), tm0 + j * BOXC, tn0, w.f);
                }
            }
            tma_store_commit();
            tma_store_wait_read();     // both buffers have been read: the producer may land the next P tile
            mbar_arrive(ofree);
        }
    }
    if (tid == 0) tma_store_wait_all();
}

}  // namespace ekf
