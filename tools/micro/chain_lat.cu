// Microbenchmark: latencies of the dependent chains in the DTW sweep (single warp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_shfl(double* out, long long* clk, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) x = __shfl_up_sync(0xffffffffu, x, 1);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_dsetp(double* out, long long* clk, double a, double d1, double d2) {
    double best = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) {
        double c = best + d1;           // DADD
        double e = a + d2 * i;          // independent
        best = (e < c) ? e : c;         // DSETP + FSEL x2
    }
    long long t1 = clock64();
    out[threadIdx.x] = best;
    if (threadIdx.x == 0) clk[1] = t1 - t0;
}
__global__ void k_isetp(double* out, long long* clk, double a, double d1, double d2) {
    double best = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) {
        double c = best + d1;
        double e = a + d2 * i;
        best = (__double_as_longlong(e) < __double_as_longlong(c)) ? e : c;
    }
    long long t1 = clock64();
    out[threadIdx.x] = best;
    if (threadIdx.x == 0) clk[2] = t1 - t0;
}
__global__ void k_dmnmx(double* out, long long* clk, double a, double d1, double d2) {
    double best = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) {
        double c = best + d1;
        double e = a + d2 * i;
        best = fmin(e, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = best;
    if (threadIdx.x == 0) clk[3] = t1 - t0;
}
__global__ void k_dadd(double* out, long long* clk, double a, double d1) {
    double best = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) best = best + d1;
    long long t1 = clock64();
    out[threadIdx.x] = best;
    if (threadIdx.x == 0) clk[4] = t1 - t0;
}
// the sweep's step: shuffle + two cells
__global__ void k_step(double* out, long long* clk, double a, double d1, double d2) {
    double va = a, vb = a + 1, diag = a + 2, vap = a, vbp = a + 1;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 1024; ++i) {
        double up = __shfl_up_sync(0xffffffffu, vbp, 1);
        double best = up + d1, c = vap + d1;
        best = c < best ? c : best;
        c = diag + d1;
        va = c < best ? c : best;
        best = va + d2; c = vbp + d2;
        best = c < best ? c : best;
        c = vap + d2;
        vb = c < best ? c : best;
        diag = up; vap = va; vbp = vb;
    }
    long long t1 = clock64();
    out[threadIdx.x] = va + vb;
    if (threadIdx.x == 0) clk[5] = t1 - t0;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&clk, 64); cudaMemset(clk, 0, 64);
    k_shfl<<<1, 32>>>(out, clk, 1.5);
    k_dsetp<<<1, 32>>>(out, clk, 1.5, 0.25, 1e-3);
    k_isetp<<<1, 32>>>(out, clk, 1.5, 0.25, 1e-3);
    k_dmnmx<<<1, 32>>>(out, clk, 1.5, 0.25, 1e-3);
    k_dadd<<<1, 32>>>(out, clk, 1.5, 0.25);
    k_step<<<1, 32>>>(out, clk, 1.5, 0.25, 0.5);
    long long h[8];
    cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
    printf("shfl_up(double) chain: %.1f clk\n", h[0] / 1024.0);
    printf("DADD + DSETP + FSEL chain: %.1f clk\n", h[1] / 1024.0);
    printf("DADD + int64 compare + SEL chain: %.1f clk\n", h[2] / 1024.0);
    printf("DADD + DMNMX chain: %.1f clk\n", h[3] / 1024.0);
    printf("DADD chain: %.1f clk\n", h[4] / 1024.0);
    printf("sweep step (shfl + 2 cells): %.1f clk\n", h[5] / 1024.0);
    return 0;
}
