// Does DMMA.8x8x4 issue rate depend on operand reuse between consecutive instructions?  1..3 warps per SM sub-partition.
//   mode 0: every DMMA reads the same A and B registers (the plain peak test)
//   mode 1: 12 chains, distinct A per (m, kt) and B per (nt, kt): nt innermost -> A shared by 3 consecutive DMMAs
//   mode 2: same operands, an order that never repeats A or B between consecutive DMMAs
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, int iters, const double* in) {
    double accL[2][3][2], accR[2][3][2], aL[2][5], aR[2][5], bL[3][5], bR[3][5];
    for (int m = 0; m < 2; m++) for (int kt = 0; kt < 5; kt++) { aL[m][kt] = in[threadIdx.x + m * 5 + kt]; aR[m][kt] = in[threadIdx.x + 64 + m * 5 + kt]; }
    for (int nt = 0; nt < 3; nt++) for (int kt = 0; kt < 5; kt++) { bL[nt][kt] = in[threadIdx.x + 128 + nt * 5 + kt]; bR[nt][kt] = in[threadIdx.x + 256 + nt * 5 + kt]; }
    for (int m = 0; m < 2; m++) for (int nt = 0; nt < 3; nt++) { accL[m][nt][0] = accL[m][nt][1] = accR[m][nt][0] = accR[m][nt][1] = 0.0; }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int kt = 0; kt < 5; kt++)
#pragma unroll
                for (int m = 0; m < 2; m++)
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) { dmma884(accL[m][nt][0], accL[m][nt][1], aL[0][0], bL[0][0]); dmma884(accR[m][nt][0], accR[m][nt][1], aL[0][0], bL[0][0]); }
        } else if (MODE == 1) {
#pragma unroll
            for (int kt = 0; kt < 5; kt++)
#pragma unroll
                for (int m = 0; m < 2; m++) {
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) dmma884(accL[m][nt][0], accL[m][nt][1], aL[m][kt], bL[nt][kt]);
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) dmma884(accR[m][nt][0], accR[m][nt][1], aR[m][kt], bR[nt][kt]);
                }
        } else {
#pragma unroll
            for (int kt = 0; kt < 5; kt++)
#pragma unroll
                for (int m = 0; m < 2; m++)
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) { dmma884(accL[m][nt][0], accL[m][nt][1], aL[m][kt], bL[nt][kt]); dmma884(accR[m][nt][0], accR[m][nt][1], aR[m][kt], bR[nt][kt]); }
        }
    }
    double s = 0;
    for (int m = 0; m < 2; m++) for (int nt = 0; nt < 3; nt++) s += accL[m][nt][0] + accL[m][nt][1] + accR[m][nt][0] + accR[m][nt][1];
    if (s == 12345.678) out[0] = s;
}
template <int MODE>
void run(int threads, int sms, double* out, double* in) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, threads>>>(out, iters, in); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<MODE><<<sms, threads>>>(out, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("mode %d warps/SMSP %d: %.2f TFLOP/s\n", MODE, threads / 128, (double)sms * (threads / 32) * 60 * iters * 512.0 / best / 1e9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out, *in; cudaMalloc(&out, 1024); cudaMalloc(&in, 8192 * 8); cudaMemset(in, 0, 8192 * 8);
    for (int threads = 128; threads <= 384; threads += 128) { run<0>(threads, p.multiProcessorCount, out, in); run<1>(threads, p.multiProcessorCount, out, in); run<2>(threads, p.multiProcessorCount, out, in); }
    return 0;
}
