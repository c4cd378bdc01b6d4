// fp64_latency.cu -- dependent-issue latencies on this GPU (single warp): DFMA, DMMA, 1/x, rsqrt, LDS.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double* sink, double seed)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    double c0 = 0, c1 = 0;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    long long t2 = clock64();
    double r = a + 2.0;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { r = __drcp_rn(r) + 1.5; }
    long long t3 = clock64();
    double s = a + 3.0;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { s = rsqrt(s) + 1.5; }
    long long t4 = clock64();
    int idx = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { idx = (int)sm[idx & 255] & 255; }
    long long t5 = clock64();
    double d = a + 4.0;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { d = 1.0 / d + 1.5; }
    long long t6 = clock64();
    if (threadIdx.x == 0) {
        out[0] = (t1 - t0) / 1024; out[1] = (t2 - t1) / 512; out[2] = (t3 - t2) / 256; out[3] = (t4 - t3) / 256;
        out[4] = (t5 - t4) / 256; out[5] = (t6 - t5) / 256;
    }
    sink[threadIdx.x] = a + c0 + c1 + r + s + idx + d;
}
int main()
{
    long long* o; double* s;
    cudaMalloc(&o, 64); cudaMalloc(&s, 8 * 32);
    k<<<1, 32>>>(o, s, 1.25); k<<<1, 32>>>(o, s, 1.25);
    long long h[8];
    cudaMemcpy(h, o, 48, cudaMemcpyDeviceToHost);
    printf("{\"dfma_dep_cycles\": %lld, \"dmma_dep_cycles\": %lld, \"drcp_plus_add_cycles\": %lld, \"rsqrt_plus_add_cycles\": %lld, "
           "\"lds_i2f_chain_cycles\": %lld, \"ddiv_plus_add_cycles\": %lld}\n", h[0], h[1], h[2], h[3], h[4], h[5]);
    return 0;
}
