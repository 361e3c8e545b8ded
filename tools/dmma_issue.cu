// How many warps per SM sub-partition does the FP64 tensor pipe need?  DMMA.8x8x4 rate for 1, 2, 3, 4 warps per SMSP
// (blocks of 128..512 threads, one block per SM) with 6 or 12 independent accumulator chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
    double d0[CH], d1[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { d0[i] = threadIdx.x * 1e-9; d1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma884(d0[i], d1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += d0[i] + d1[i];
    if (s == 12345.678) out[0] = s;
}
template <int CH>
void run(int threads, int sms, double* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<CH><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("warps/SMSP %d chains %2d: %.2f TFLOP/s\n", threads / 128, CH, (double)sms * (threads / 32) * CH * iters * 512.0 / best / 1e9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 1024);
    for (int threads = 128; threads <= 512; threads += 128) { run<3>(threads, p.multiProcessorCount, out); run<6>(threads, p.multiProcessorCount, out); run<12>(threads, p.multiProcessorCount, out); }
    return 0;
}
