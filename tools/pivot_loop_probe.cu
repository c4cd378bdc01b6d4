// probe: what limits one pivot iteration of the S-chain panel loop?  Variants toggle pieces of the body.
#include <cstdio>
#include <cuda_runtime.h>
template <int VAR>
__global__ void __launch_bounds__(128) k(double* out, long long* cyc, const double* in)
{
    extern __shared__ __align__(16) double sm[];
    double* U = sm;
    const int tid = threadIdx.x, ri = tid >> 3, cj = tid & 7, r0 = 4 * ri, c0 = 8 * cj;
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int q = 0; q < 8; ++q) a[r][q] = in[(r0 + r) * 64 + c0 + q];
    for (int e = tid; e < 64 * 64; e += 128) U[e] = in[e];
    __syncthreads();
    long long t0 = clock64();
    for (int c = 0; c < 64; ++c) {
        if (VAR != 5) __syncthreads();
        const double* rb = U + c * 64;
        const double piv = rb[c];
        double pinv;
        if (VAR == 1) pinv = piv * 0.5; else pinv = __drcp_rn(piv);
        if (r0 + 3 > c) {
            double m[4];
            const double2 m01 = *reinterpret_cast<const double2*>(rb + r0);
            const double2 m23 = *reinterpret_cast<const double2*>(rb + r0 + 2);
            m[0] = (r0 > c) ? m01.x * pinv : 0.0; m[1] = (r0 + 1 > c) ? m01.y * pinv : 0.0;
            m[2] = (r0 + 2 > c) ? m23.x * pinv : 0.0; m[3] = m23.y * pinv;
            if (VAR != 2 && c0 + 7 >= c) {
                double rv[8];
                for (int q = 0; q < 4; ++q) { const double2 t2 = *reinterpret_cast<const double2*>(rb + c0 + 2 * q); rv[2*q] = t2.x; rv[2*q+1] = t2.y; }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 8; ++q) a[r][q] -= m[r] * rv[q];
            } else if (VAR == 2) { a[0][0] += m[0] + m[1] + m[2] + m[3]; }
            if (VAR != 3 && ri == (c + 1) >> 2) {
                const int rr = (c + 1) & 3;
                double* nb = U + (c + 1) * 64 + c0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (r == rr)
#pragma unroll
                        for (int q = 0; q < 8; ++q) nb[q] = (VAR == 4) ? a[r][q] * 0.0 + 2.0 + q : a[r][q];
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int r = 0; r < 4; ++r) for (int q = 0; q < 8; ++q) s += a[r][q];
    out[tid] = s;
    if (tid == 0) cyc[VAR] = (t1 - t0) / 64;
}
int main()
{
    double h[4096];
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) h[i * 64 + j] = (i == j) ? 80.0 : 1.0 / (1 + i + j);
    double *in, *out; long long* cyc;
    cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 1024); cudaMalloc(&cyc, 64);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 128, 33000>>>(out, cyc, in); k<1><<<1, 128, 33000>>>(out, cyc, in); k<2><<<1, 128, 33000>>>(out, cyc, in);
        k<3><<<1, 128, 33000>>>(out, cyc, in); k<4><<<1, 128, 33000>>>(out, cyc, in); k<5><<<1, 128, 33000>>>(out, cyc, in);
    }
    long long c[8];
    cudaMemcpy(c, cyc, 48, cudaMemcpyDeviceToHost);
    printf("cycles/pivot: full %lld | no-rcp %lld | no-fma %lld | no-publish %lld | publish-indep %lld | no-barrier(racy) %lld  (%s)\n",
           c[0], c[1], c[2], c[3], c[4], c[5], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
