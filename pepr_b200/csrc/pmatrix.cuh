// P(t) = V exp(Lambda r_c t) Vinv for the two branches of a CLV update, built inside the consuming kernel (raxmlHPC makeP,
// SURVEY a12).  Every CTA builds its own copy in shared memory: 2 x 4 x 400 entries x 20 FMA = 64 k FMA, about 1,000 cycles
// for eleven warps -- against a separate launch (10 us on the dependent chain of every branch visit) plus the strided
// fragment loads from global memory that followed it.  Branch lengths are read from the tree's device array, so a length a
// preceding kernel has just written (device-side Newton-Raphson) is picked up without the host in between.
//
// Usage (all threads of the CTA that take part, `nthreads` of them with ranks `tid` = 0 .. nthreads-1):
//   ModelRegs r = model_prefetch(...)      -- global loads issued; do this BEFORE the CTA queues its bulk loads
//   pdl_wait(); length_prefetch(r, ...)    -- the branch lengths may have been written by the preceding kernel
//   model_to_smem(...);  barrier           -- V, Vinv, exp tables in shared memory
//   build_p(...);        barrier           -- s_P[child][c][i][j]
//   build_tip_lookup(...) / load_p_fragments_smem(...)
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

namespace pml {
namespace pmat {

constexpr int kMat = kStates * kStates;      // 400
constexpr int kModelDoubles = 2 * kMat + 2 * kRow;  // V, Vinv, exp(lambda_k r_c t) of both branches
constexpr int kPDoubles = 2 * kCats * kMat;  // both children

struct ModelRegs {
    double v[2], vinv[2];
    double lambda, rate;
    double len[2];
};

// static part: nothing a preceding kernel may have written (safe before pdl_wait)
template <int kThreads>
__device__ __forceinline__ ModelRegs model_prefetch(const DeviceModel* dm, int tid) {
    static_assert(2 * kThreads >= kMat, "two rounds must cover a 20 x 20 matrix");
    ModelRegs r{};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kThreads;
        if (idx < kMat) {
            r.v[q] = (&dm->V[0][0])[idx];
            r.vinv[q] = (&dm->Vinv[0][0])[idx];
        }
    }
    if (tid < 2 * kRow) {
        r.lambda = dm->lambda[tid % kStates];
        r.rate = dm->rates[(tid % kRow) / kStates];
    }
    return r;
}
__device__ __forceinline__ void length_prefetch(ModelRegs& r, const double* len_l, const double* len_r) {
    r.len[0] = *len_l;
    r.len[1] = *len_r;
}

// s_model: [V 400][Vinv 400][exp child 0: 80][exp child 1: 80]
template <int kThreads>
__device__ __forceinline__ void model_to_smem(const ModelRegs& r, int tid, double* s_model) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kThreads;
        if (idx < kMat) {
            s_model[idx] = r.v[q];
            s_model[kMat + idx] = r.vinv[q];
        }
    }
    if (tid < 2 * kRow) s_model[2 * kMat + tid] = exp(r.lambda * r.rate * (tid < kRow ? r.len[0] : r.len[1]));
}

// s_P[child][c][i][j] = sum_k (V[i][k] e_c[k]) Vinv[k][j], k ascending (the order of the host model code)
template <int kThreads>
__device__ __forceinline__ void build_p(const double* s_model, int tid, double* s_P) {
    const double* s_V = s_model;
    const double* s_Vinv = s_model + kMat;
    const double* s_ex = s_model + 2 * kMat;
    for (int w = tid; w < 2 * kCats * kStates * 2; w += kThreads) {  // (child, c, i, half of the row)
        const int half = w & 1, i = (w >> 1) % kStates, cc = (w >> 1) / kStates;  // cc = child * 4 + c
        double acc[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) acc[j] = 0.0;
#pragma unroll 4
        for (int k = 0; k < kStates; ++k) {
            const double wk = s_V[i * kStates + k] * s_ex[cc * kStates + k];
            const double2* row = reinterpret_cast<const double2*>(s_Vinv + k * kStates + half * 10);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const double2 x = row[j];
                acc[2 * j] = fma(wk, x.x, acc[2 * j]);
                acc[2 * j + 1] = fma(wk, x.y, acc[2 * j + 1]);
            }
        }
        double2* out = reinterpret_cast<double2*>(s_P + (cc * kStates + i) * kStates + half * 10);
#pragma unroll
        for (int j = 0; j < 5; ++j) out[j] = make_double2(acc[2 * j], acc[2 * j + 1]);
    }
}

// tip[code][c*20+i] = sum_j P_c[i][j] * indicator(code)[j]; rows padded to `pad` doubles
template <int kThreads>
__device__ __forceinline__ void build_tip_lookup(const double* s_Pchild, int tid, double* s_tip, int pad) {
    for (int idx = tid; idx < kCodes * kRow; idx += kThreads) {
        const int code = idx / kRow, ci = idx % kRow;
        const double* p = s_Pchild + ci * kStates;  // P_c[i][*]
        double acc;
        if (code < 20) acc = p[code];
        else if (code == 20) acc = p[2] + p[3];
        else if (code == 21) acc = p[5] + p[6];
        else {
            acc = 0.0;
#pragma unroll
            for (int j = 0; j < kStates; ++j) acc += p[j];
        }
        s_tip[code * pad + ci] = acc;
    }
}

}  // namespace pmat
}  // namespace pml
