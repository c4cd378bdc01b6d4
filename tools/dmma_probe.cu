// probe: DMMA m8n8k4 throughput with the kernel's real operand pattern (8 A x 4 B fragments, 32 accumulators),
// (1) operands held in registers, (2) operands re-loaded from shared memory every k4-step (pitch 132/68).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    for (int e = tid; e < 16 * 132 * 2; e += 256) sm[e] = 1e-3 * (e % 97);
    __syncthreads();
    const double* As = sm + (warp >> 2) * 64 + g;
    const double* Bs = sm + 16 * 132 + (warp & 3) * 32 + g;
    double acc[8][4][2];
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    double af[8], bf[4];
    for (int a = 0; a < 8; ++a) af[a] = As[q * 132 + a * 8];
    for (int b = 0; b < 4; ++b) bf[b] = Bs[q * 132 + b * 8];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            if (MODE == 1) {
#pragma unroll
                for (int a = 0; a < 8; ++a) af[a] = As[(k4 + q) * 132 + a * 8];
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = Bs[(k4 + q) * 132 + b * 8];
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        if (MODE == 2) __syncthreads();
    }
    double s = 0;
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 4; ++b) s += acc[a][b][0] + acc[a][b][1];
    out[blockIdx.x * 256 + tid] = s;
}
template <int MODE>
void run(const char* name, double* out)
{
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<148, 256, 16 * 132 * 2 * 8>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fl = 512.0 * 32 * 4 * iters * 8 * 148;
    printf("%s: %.2f TFLOP/s\n", name, fl / best * 1e-9);
}
int main()
{
    double* out; cudaMalloc(&out, 148 * 256 * 8);
    run<0>("operands in registers, 8 warps/SM", out);
    run<1>("operands from smem every k4 (LDS.64 x12)", out);
    run<2>("registers + barrier every 16 rows", out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
