// fp64_peak.cu -- measures the FP64 issue-rate ceilings of this GPU: DFMA (CUDA-core path) and
// DMMA m8n8k4 (tensor path), register-resident, no memory traffic.  The roofline denominator for
// the covariance downdate is the better of the two and of a cuBLAS DGEMM (measured in bench.py).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters)
{
    double a[16], b = 1.0000001, c = 0.9999999;
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters)
{
    double c[16][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 20000, blocks = sms * 2, threads = warps * 32 / 2;
        float best1 = 1e9, best2 = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            float ms;
            cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); if (ms < best1) best1 = ms;
            cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); if (ms < best2) best2 = ms;
        }
        double fl1 = 2.0 * 16 * iters * (double)blocks * threads;
        double fl2 = 512.0 * 16 * iters * (double)blocks * (threads / 32);
        printf("{\"sms\": %d, \"warps_per_sm\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f}\n", sms, warps,
               fl1 / best1 * 1e-9, fl2 / best2 * 1e-9);
    }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
