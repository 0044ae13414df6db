// Microbenchmark: FP64 dependent-op latency and per-SM throughput on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_dfma(double* out, long long* clk, double a, double b) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 1024; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void lat_rsqrt(double* out, long long* clk, double a) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) x = rsqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[1] = t1 - t0;
}
__global__ void lat_sqrt(double* out, long long* clk, double a) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) x = sqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[2] = t1 - t0;
}
__global__ void lat_div(double* out, long long* clk, double a) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) x = a / x + 1.0;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[3] = t1 - t0;
}
// throughput: each thread 8 independent chains
__global__ void thr_dfma(double* out, long long* clk, double a, double b) {
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = a + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 512; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, a);
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) clk[4] = t1 - t0;
}
__global__ void thr_dadd_min(double* out, long long* clk, double a, double b) {
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = a + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 512; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fmin(x[j] + b, a);
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) clk[5] = t1 - t0;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&clk, 64);
    cudaMemset(clk, 0, 64);
    lat_dfma<<<1, 32>>>(out, clk, 1.000001, 0.999999);
    lat_rsqrt<<<1, 32>>>(out, clk, 1.5);
    lat_sqrt<<<1, 32>>>(out, clk, 1.5);
    lat_div<<<1, 32>>>(out, clk, 1.5);
    long long h[8];
    cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.1f clk\n", h[0] / 1024.0);
    printf("rsqrt(double)+DADD chain: %.1f clk\n", h[1] / 256.0);
    printf("sqrt(double)+DADD chain: %.1f clk\n", h[2] / 256.0);
    printf("div(double)+DADD chain: %.1f clk\n", h[3] / 256.0);
    for (int threads = 128; threads <= 1024; threads *= 2) {
        thr_dfma<<<1, threads>>>(out, clk, 1.000001, 0.999999);
        thr_dadd_min<<<1, threads>>>(out, clk, 1.000001, 0.999999);
        cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d: DFMA %.1f lane-ops/clk/SM, DADD+DMNMX pair %.1f lane-ops/clk/SM\n", threads,
               (double)threads * 8 * 512 / h[4], (double)threads * 16 * 512 / h[5]);
    }
    return 0;
}
