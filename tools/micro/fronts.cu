// Micro-benchmark: HBM copy bandwidth as a function of the number of concurrent sequential streams ("fronts").
// N doubles are split into F equal regions; CTA b works on region b % F and, inside it, the CTAs of that region sweep
// it together in 2 KB chunks (same as the lattice kernel: a front is a row being read by ~G/F CTAs at once).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fronts fronts.cu ; run: ./fronts
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_copy_fronts(const double* __restrict__ x, double* __restrict__ y, long n, int F) {
    const int f = blockIdx.x % F, w = blockIdx.x / F, W = (gridDim.x + F - 1 - f) / F == 0 ? 1 : (gridDim.x - f + F - 1) / F;
    const long per = n / F, base = (long)f * per;
    for (long i = (long)w * blockDim.x + threadIdx.x; i < per; i += (long)W * blockDim.x) y[base + i] = 2.0 * x[base + i];
}
int main() {
    const long n = 100000000;
    double *x, *y;
    cudaMalloc(&x, n * 8); cudaMalloc(&y, n * 8);
    cudaMemset(x, 0, n * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grids[] = {592, 1184};
    const int fronts[] = {1, 2, 4, 8, 15, 30, 60, 148, 592};
    for (int g : grids)
        for (int F : fronts) {
            if (F > g) continue;
            k_copy_fronts<<<g, 256>>>(x, y, n, F);
            cudaEventRecord(a);
            for (int r = 0; r < 10; ++r) k_copy_fronts<<<g, 256>>>(x, y, n, F);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("grid %4d fronts %3d: %.3f ms  %.0f GB/s\n", g, F, ms / 10, 16.0 * n / (ms / 10 * 1e-3) / 1e9);
        }
    return 0;
}
