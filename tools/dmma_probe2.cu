// probe: does the warp-tile shape limit the DMMA m8n8k4 rate?  Operands re-loaded from shared memory every k4-step, as in the
// downdate kernels.  AM x 4 fragments per warp (AM = 4: 32x32 warp tile, 8 LDS per 16 DMMA; AM = 8: 64x32, 12 LDS per 32 DMMA),
// W warps per SM (one CTA per SM).  Prints TFLOP/s over 148 SMs.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int AM, int W>
__global__ void __launch_bounds__(W * 32, 1) k(double* out, int iters)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    for (int e = tid; e < 16 * 68 * 2; e += W * 32) sm[e] = 1e-3 * (e % 97);
    __syncthreads();
    const double* As = sm + (warp & 1) * (AM == 4 ? 32 : 0) + g;
    const double* Bs = sm + 16 * 68 + ((warp >> 1) & 1) * 32 + g;
    double acc[AM][4][2];
    for (int a = 0; a < AM; ++a) for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            double af[AM], bf[4];
#pragma unroll
            for (int a = 0; a < AM; ++a) af[a] = As[(k4 + q) * 68 + a * 8];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[(k4 + q) * 68 + b * 8];
#pragma unroll
            for (int a = 0; a < AM; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    double s = 0;
    for (int a = 0; a < AM; ++a) for (int b = 0; b < 4; ++b) s += acc[a][b][0] + acc[a][b][1];
    out[blockIdx.x * W * 32 + tid] = s;
}
template <int AM, int W>
void run(double* out)
{
    cudaFuncSetAttribute(k<AM, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<AM, W><<<148, W * 32, 16 * 68 * 2 * 8>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fl = 512.0 * AM * 4 * 4 * iters * W * 148;
    printf("warp tile %dx32 (%d LDS per %d DMMA), %2d warps/SM: %.2f TFLOP/s\n", AM * 8, AM + 4, AM * 4, W, fl / best * 1e-9);
}
int main()
{
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    run<4, 4>(out); run<4, 8>(out); run<4, 16>(out); run<8, 4>(out); run<8, 8>(out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
