// Micro-benchmark: latency of a dependent DFMA chain and throughput of independent DFMAs on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *clk, int n)
{
    double a = out[0], b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) a = fma(a, b, 1e-9);
    long long t1 = clock64();
    out[2 + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void thr(double *out, long long *clk, int n)
{
    double a0 = out[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7, b = out[1];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        a0 = fma(a0, b, 1e-9); a1 = fma(a1, b, 1e-9); a2 = fma(a2, b, 1e-9); a3 = fma(a3, b, 1e-9);
        a4 = fma(a4, b, 1e-9); a5 = fma(a5, b, 1e-9); a6 = fma(a6, b, 1e-9); a7 = fma(a7, b, 1e-9);
    }
    long long t1 = clock64();
    out[2 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main()
{
    double *out; long long *clk, h;
    cudaMalloc(&out, 8 * 2048); cudaMalloc(&clk, 8);
    double init[2] = {1.0, 0.999999}; cudaMemcpy(out, init, 16, cudaMemcpyHostToDevice);
    const int n = 4096;
    lat<<<1, 32>>>(out, clk, n); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA chain, 1 warp: %.1f cycles per DFMA\n", (double)h / n);
    for (int warps = 1; warps <= 32; warps *= 2) {
        thr<<<1, 32 * warps>>>(out, clk, n); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("independent DFMAs, %2d warps on one SM: %.2f cycles per warp-DFMA (SM-wide: %.2f warp-DFMA/cycle)\n", warps,
               (double)h / (8.0 * n), 8.0 * n * warps / (double)h);
    }
    return 0;
}
