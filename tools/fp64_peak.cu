// Microbenchmark: FP64 pipe peaks on B200 (sm_100a) -- DFMA, DMMA (m8n8k4 / m16n8k8 / m16n8k16), and mixed.
// Decides whether the 20x20 contraction of newview goes on the FP64 FMA pipe or the FP64 tensor path.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template<int CH>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
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
template<int CH>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double d[CH][4]; double av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < CH; i++) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma1688(d[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.678) out[0] = s;
}
template<int CH>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double d[CH][4]; double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < CH; i++) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i; d[i][2] = 1; d[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma16816(d[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.678) out[0] = s;
}
// mixed: CH dmma chains + CF dfma chains per thread
template<int CH, int CF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double d0[CH], d1[CH], acc[CF];
#pragma unroll
    for (int i = 0; i < CH; i++) { d0[i] = threadIdx.x * 1e-9; d1[i] = i; }
#pragma unroll
    for (int i = 0; i < CF; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma884(d0[i], d1[i], a, b);
#pragma unroll
        for (int i = 0; i < CF; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += d0[i] + d1[i];
#pragma unroll
    for (int i = 0; i < CF; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}
// FFMA for reference
template<int CH>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
    float acc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = threadIdx.x * 1e-9f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}
__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = in[i];
}

template<typename F> float timeit(F f, int rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < rep; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; printf("dev %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 1024));
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps; double thr = (double)grid * 256;
        float ms;
        ms = timeit([&]{ k_dfma<8><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d DFMA x8 chains: %.2f TFLOP/s\n", bps, thr * 8 * iters * 2 / ms / 1e9);
        ms = timeit([&]{ k_dfma<16><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d DFMA x16 chains: %.2f TFLOP/s\n", bps, thr * 16 * iters * 2 / ms / 1e9);
        ms = timeit([&]{ k_dmma884<4><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d DMMA884 x4: %.2f TFLOP/s\n", bps, (thr / 32) * 4 * iters * 512.0 / ms / 1e9);
        ms = timeit([&]{ k_dmma884<8><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d DMMA884 x8: %.2f TFLOP/s\n", bps, (thr / 32) * 8 * iters * 512.0 / ms / 1e9);
        ms = timeit([&]{ k_dmma1688<4><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d DMMA1688 x4: %.2f TFLOP/s\n", bps, (thr / 32) * 4 * iters * 2048.0 / ms / 1e9);
        ms = timeit([&]{ k_dmma16816<4><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d DMMA16816 x4: %.2f TFLOP/s\n", bps, (thr / 32) * 4 * iters * 4096.0 / ms / 1e9);
        ms = timeit([&]{ k_mixed<4,8><<<grid,256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d MIXED dmma4+dfma8: %.2f TFLOP/s total (dmma %.2f + dfma %.2f)\n", bps,
               ((thr / 32) * 4 * iters * 512.0 + thr * 8 * iters * 2) / ms / 1e9,
               (thr / 32) * 4 * iters * 512.0 / ms / 1e9, thr * 8 * iters * 2 / ms / 1e9);
        ms = timeit([&]{ k_ffma<16><<<grid,256>>>((float*)out, iters, 1.0000001f, 1e-9f); }, 5);
        printf("bps %d FFMA x16: %.2f TFLOP/s\n", bps, thr * 16 * iters * 2 / ms / 1e9);
    }
    // sustained DFMA for ~2s to see power-capped clocks
    {
        int grid = sms * 4; double thr = (double)grid * 256;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < 40; r++) k_dfma<16><<<grid,256>>>(out, iters * 4, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("sustained DFMA (%.0f ms): %.2f TFLOP/s\n", ms, thr * 16 * iters * 4.0 * 40 * 2 / ms / 1e9);
        cudaEventRecord(e0);
        for (int r = 0; r < 40; r++) k_dmma884<8><<<grid,256>>>(out, iters * 2, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("sustained DMMA884 (%.0f ms): %.2f TFLOP/s\n", ms, (thr / 32) * 8 * iters * 2.0 * 40 * 512.0 / ms / 1e9);
    }
    // HBM copy
    {
        size_t n = (size_t)1 << 27; // 128Mi double2 = 2 GiB each
        double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        cudaMemset(a, 1, n * 16); cudaMemset(b, 0, n * 16);
        for (int bps = 4; bps <= 32; bps *= 2) {
            float ms = timeit([&]{ k_copy<<<sms * bps, 256>>>(a, b, n); }, 5);
            printf("copy grid %d x256: %.1f GB/s\n", sms * bps, 2.0 * n * 16 / ms / 1e6);
        }
        float ms = timeit([&]{ cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice); }, 5);
        printf("cudaMemcpy D2D: %.1f GB/s\n", 2.0 * n * 16 / ms / 1e6);
    }
    return 0;
}
