// Small CUDA kernels of the likelihood path (sm_100a): the Newton-Raphson core on a stored
// product table, the fixed-order second stage of every reduction, replicate lnL.  The CLV-streaming kernels live in
// newview_mma.cu and branch_mma.cu.
#include "kernels.h"

#include <cstdio>

namespace pml {

namespace {

constexpr double kLogMinLik = -177.445678223345993274;  // ln 2^-256

constexpr int kPatternsPerBlock = 32;
constexpr int kThreads = kPatternsPerBlock * kCats;

__device__ __forceinline__ void load_row20(const double* __restrict__ src, double* v) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int j = 0; j < kStates / 2; ++j) {
        const double2 t = __ldg(s2 + j);
        v[2 * j] = t.x;
        v[2 * j + 1] = t.y;
    }
}

// ------------------------------------------------------------------------------------------------ reduce ---
// stage 2 of every reduction: fixed summation order -> bitwise reproducible results
__global__ void __launch_bounds__(1024) k_reduce(const double* __restrict__ partials, int nblocks, int nvals,
                                                 double* __restrict__ result) {
    __shared__ double s[1024];
    for (int v = 0; v < nvals; ++v) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += 1024) acc += partials[(int64_t)v * nblocks + i];
        s[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 512; w > 0; w >>= 1) {
            if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) result[v] = s[0];
        __syncthreads();
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ NR core --
__global__ void __launch_bounds__(kThreads) k_core(const DeviceModel* __restrict__ dm, const double* __restrict__ sumtable,
                                                   const int32_t* __restrict__ sum_scale, const int32_t* __restrict__ weights,
                                                   int64_t np, const double t, double* __restrict__ partials) {
    __shared__ double s_e[3][kCats][kStates];
    __shared__ double s_f[3][kCats][kPatternsPerBlock];
    if (threadIdx.x < kRow) {
        const int c = threadIdx.x / kStates, k = threadIdx.x % kStates;
        const double a = dm->lambda[k] * dm->rates[c], e = exp(a * t);
        s_e[0][c][k] = e;
        s_e[1][c][k] = a * e;
        s_e[2][c][k] = a * a * e;
    }
    __syncthreads();
    const int c = threadIdx.x / kPatternsPerBlock, lane = threadIdx.x % kPatternsPerBlock;
    const int64_t p = (int64_t)blockIdx.x * kPatternsPerBlock + lane;
    const bool live = p < np;
    double f = 0.0, f1 = 0.0, f2 = 0.0;
    if (live) {
        double s[kStates];
        load_row20(sumtable + p * kRow + c * kStates, s);
#pragma unroll
        for (int k = 0; k < kStates; ++k) {
            f = fma(s[k], s_e[0][c][k], f);
            f1 = fma(s[k], s_e[1][c][k], f1);
            f2 = fma(s[k], s_e[2][c][k], f2);
        }
    }
    s_f[0][c][lane] = f;
    s_f[1][c][lane] = f1;
    s_f[2][c][lane] = f2;
    __syncthreads();
    if (c != 0) return;
    double l = 0.0, d1 = 0.0, d2 = 0.0;
    if (live) {
        f = (s_f[0][0][lane] + s_f[0][1][lane]) + (s_f[0][2][lane] + s_f[0][3][lane]);
        f1 = (s_f[1][0][lane] + s_f[1][1][lane]) + (s_f[1][2][lane] + s_f[1][3][lane]);
        f2 = (s_f[2][0][lane] + s_f[2][1][lane]) + (s_f[2][2][lane] + s_f[2][3][lane]);
        const double w = (double)weights[p], inv = 1.0 / f, q = f1 * inv;
        l = w * (log(0.25 * f) + sum_scale[p] * kLogMinLik);
        d1 = w * q;
        d2 = w * (f2 * inv - q * q);
    }
    l = warp_sum(l);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (lane == 0) {
        partials[blockIdx.x] = l;
        partials[gridDim.x + blockIdx.x] = d1;
        partials[2 * (int64_t)gridDim.x + blockIdx.x] = d2;
    }
}

// ------------------------------------------------------------------------------------------------ replicates
__global__ void __launch_bounds__(256) k_replicate_lnl(const int32_t* __restrict__ W, int64_t np, int64_t ldw,
                                                       const double* __restrict__ site_lnl, double* __restrict__ lnl) {
    __shared__ double s[256];
    const int32_t* w = W + (int64_t)blockIdx.x * ldw;
    double acc = 0.0;
    for (int64_t p = threadIdx.x; p < np; p += 256) acc = fma((double)w[p], site_lnl[p], acc);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) lnl[blockIdx.x] = s[0];
}

inline int blocks_for(int64_t np) { return (int)((np + kPatternsPerBlock - 1) / kPatternsPerBlock); }

}  // namespace

void launch_core(const DeviceModel* dm, const double* sumtable, const int32_t* sum_scale, const int32_t* weights, int64_t np,
                 double t, double* partials, double* result, cudaStream_t stream) {
    const int grid = blocks_for(np);
    k_core<<<grid, kThreads, 0, stream>>>(dm, sumtable, sum_scale, weights, np, t, partials);
    k_reduce<<<1, 1024, 0, stream>>>(partials, grid, 3, result);
}

void launch_replicate_lnl(const int32_t* W, int nrep, int64_t np, int64_t ldw, const double* site_lnl, double* lnl,
                          cudaStream_t stream) {
    if (nrep > 0) k_replicate_lnl<<<nrep, 256, 0, stream>>>(W, np, ldw, site_lnl, lnl);
}

int64_t reduce_partials_capacity(int64_t np) { return 3 * (np / 16 + 1) + 8; }  // the branch kernel may run np/16 CTAs

}  // namespace pml
