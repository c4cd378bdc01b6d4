// vendor_bar.cu -- the library bar for the covariance downdate: cuBLAS on the same box, same shapes.
//   P (n x n, FP64) -= W W^T with W^T stored K-major (k x n row-major, leading dimension ld), exactly the operand layout of
//   k_downdate64:  cublasDsyrk (lower triangle only -- it leaves the mirror to the caller), cublasDgemm (full n x n, what a
//   non-symmetric library call costs), and for the batched C4 shard shape cublasDgemmStridedBatched / a Dsyrk loop.
// Every timed launch is preceded by an L2 flush (256 MB memset) outside the timed interval, like bench.py.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/vendor_bar tools/vendor_bar.cu -lcublas
// usage: vendor_bar [n k [batch]] ...   (default: 3013 640 1   3013 72 1   3013 1000 1   1213 290 32   1213 64 32)
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
#define CB(x) do { cublasStatus_t s_ = (x); if (s_ != CUBLAS_STATUS_SUCCESS) { printf("cuBLAS error %d at %s:%d\n", (int)s_, __FILE__, __LINE__); return 1; } } while (0)

__global__ void fill(double* p, size_t n, double scale)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = scale * (double)((i * 2654435761u) % 1000) / 1000.0;
}

// what a caller of cublasDsyrk still has to do to keep P usable as a full symmetric matrix (every other kernel of the
// filter reads rows of P as columns): copy the lower triangle into the upper one.  32x32 tiles through shared memory.
__global__ void mirror_lower(double* P, int n, int ld)
{
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int i = bi * 32 + threadIdx.y, j = bj * 32 + threadIdx.x;
    for (int r = 0; r < 32; r += 8)
        if (i + r < n && j < n) t[threadIdx.y + r][threadIdx.x] = P[(size_t)j * ld + (i + r)];   // column-major lower (i >= j)
    __syncthreads();
    const int ii = bj * 32 + threadIdx.y, jj = bi * 32 + threadIdx.x;
    for (int r = 0; r < 32; r += 8)
        if (ii + r < n && jj < n && jj > ii + r) P[(size_t)jj * ld + (ii + r)] = t[threadIdx.x][threadIdx.y + r];
}

int main(int argc, char** argv)
{
    std::vector<int> cases;
    for (int i = 1; i + 2 < argc; i += 3)
        for (int j = 0; j < 3; ++j) cases.push_back(atoi(argv[i + j]));
    if (cases.empty()) cases = {3013, 640, 1, 3013, 72, 1, 3013, 1000, 1, 1213, 290, 32, 1213, 64, 32};
    cublasHandle_t h;
    CB(cublasCreate(&h));
    void* flush = nullptr;
    const size_t flushBytes = (size_t)256 << 20;
    CK(cudaMalloc(&flush, flushBytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (size_t c = 0; c + 3 <= cases.size(); c += 3) {
        const int n = cases[c], k = cases[c + 1], batch = cases[c + 2];
        const int ld = (n + 1 + 15) / 16 * 16;
        double *P, *W;
        CK(cudaMalloc(&P, sizeof(double) * (size_t)batch * n * ld));
        CK(cudaMalloc(&W, sizeof(double) * (size_t)batch * k * ld));
        fill<<<1024, 256>>>(P, (size_t)batch * n * ld, 1.0);
        fill<<<1024, 256>>>(W, (size_t)batch * k * ld, 1e-3);
        CK(cudaDeviceSynchronize());
        const double alpha = -1.0, beta = 1.0;
        // Row-major W^T (k x n, ld) is the column-major matrix A (n x k, lda = ld) = W.  C = C - A A^T.
        auto timeit = [&](auto&& launch, float* best) -> int {
            *best = 1e30f;
            for (int r = 0; r < 8; ++r) {
                CK(cudaMemsetAsync(flush, 0, flushBytes));
                CK(cudaEventRecord(e0));
                if (launch()) return 1;
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 2) *best = std::min(*best, ms);
            }
            return 0;
        };
        float tSyrk = 0.f, tGemm = 0.f, tBatched = 0.f, tSyrkMirror = 0.f;
        if (timeit([&]() -> int {
                for (int b = 0; b < batch; ++b) {
                    CB(cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, k, &alpha, W + (size_t)b * k * ld, ld, &beta,
                                   P + (size_t)b * n * ld, ld));
                    mirror_lower<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8)>>>(P + (size_t)b * n * ld, n, ld);
                }
                return 0; }, &tSyrkMirror)) return 1;
        if (timeit([&]() -> int {
                for (int b = 0; b < batch; ++b)
                    CB(cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, k, &alpha, W + (size_t)b * k * ld, ld, &beta,
                                   P + (size_t)b * n * ld, ld));
                return 0; }, &tSyrk)) return 1;
        if (timeit([&]() -> int {
                for (int b = 0; b < batch; ++b)
                    CB(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, k, &alpha, W + (size_t)b * k * ld, ld, W + (size_t)b * k * ld, ld,
                                   &beta, P + (size_t)b * n * ld, ld));
                return 0; }, &tGemm)) return 1;
        if (batch > 1) {
            if (timeit([&]() -> int {
                    CB(cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, k, &alpha, W, ld, (long long)k * ld, W, ld,
                                                 (long long)k * ld, &beta, P, ld, (long long)n * ld, batch));
                    return 0; }, &tBatched)) return 1;
        }
        const double fSym = (double)batch * n * (n + 1.0) * k, fFull = 2.0 * batch * (double)n * n * k;
        printf("{\"n\": %d, \"k\": %d, \"batch\": %d, \"cublasDsyrk_plus_mirror_us\": %.1f, \"cublasDsyrk_us\": %.1f, \"cublasDsyrk_tflops_symmetric_form\": %.2f, "
               "\"cublasDgemm_us\": %.1f, \"cublasDgemm_tflops_executed\": %.2f, \"cublasDgemm_tflops_symmetric_form\": %.2f",
               n, k, batch, tSyrkMirror * 1e3, tSyrk * 1e3, fSym / tSyrk / 1e9, tGemm * 1e3, fFull / tGemm / 1e9, fSym / tGemm / 1e9);
        if (batch > 1)
            printf(", \"cublasDgemmStridedBatched_us\": %.1f, \"cublasDgemmStridedBatched_tflops_symmetric_form\": %.2f", tBatched * 1e3,
                   fSym / tBatched / 1e9);
        printf("}\n");
        fflush(stdout);
        cudaFree(P); cudaFree(W);
    }
    return 0;
}
