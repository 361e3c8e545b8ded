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

// V and Vinv only (static: may be staged before pdl_wait)
template <int kThreads>
__device__ __forceinline__ void matrices_to_smem(const ModelRegs& r, int tid, double* s_model) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kThreads;
        if (idx < kMat) {
            s_model[idx] = r.v[q];
            s_model[kMat + idx] = r.vinv[q];
        }
    }
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

// exp(x) for the arguments of this path, x = lambda_k r_c t <= 0 (an eigenvalue "0" may come out as a few ulp above it).
// The library exponential is a chain of ~35 dependent FP64 instructions, and a dependent FP64 instruction takes ~40 clk here:
// 1,440 clk measured between "length arrived" and "exponentials done" in every launch's prologue, a quarter of it.  This one
// has a chain of ten: Cody-Waite reduction by ln 2, then exp(r) = 1 + r + r^2 q(r) with q of degree 11 (Taylor; |r| <= 0.347:
// truncation 4e-18) evaluated by Estrin's scheme, then the power of two through the exponent bits.  Relative error
// <= 1.4 x 2^-53 over [-708, 0] (tests/test_host.py, against 60-digit arithmetic); below -708 -- and for a NaN -- the result is 0
// (the library gives < 3e-308 there).
__host__ __device__ __forceinline__ double exp_neg(double x) {
    if (!(x > -708.0)) return 0.0;
    const double n = rint(x * 1.4426950408889634);
    double r = fma(-n, 6.93147180369123816490e-01, x);
    r = fma(-n, 1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p0 = fma(1.0 / 6.0, r, 0.5), p1 = fma(1.0 / 120.0, r, 1.0 / 24.0), p2 = fma(1.0 / 5040.0, r, 1.0 / 720.0);
    const double p3 = fma(1.0 / 362880.0, r, 1.0 / 40320.0), p4 = fma(1.0 / 39916800.0, r, 1.0 / 3628800.0);
    const double p5 = fma(1.0 / 6227020800.0, r, 1.0 / 479001600.0);
    const double q0 = fma(p1, r2, p0), q1 = fma(p3, r2, p2), q2 = fma(p5, r2, p4);
    const double q = fma(q2, r8, fma(q1, r4, q0));
    const double e = fma(r2, q, r) + 1.0;
#ifdef __CUDA_ARCH__
    return e * __hiloint2double(((int)n + 1023) << 20, 0);
#else  // host build of the same arithmetic (pml_debug_exp_neg: the accuracy claim above is a CPU test)
    const unsigned long long bits = (unsigned long long)((int)n + 1023) << 52;
    double scale;
    __builtin_memcpy(&scale, &bits, sizeof scale);
    return e * scale;
#endif
}

// ---- the same product on the FP64 tensor path, one warp per (branch, category) ---------------------------------------
// P_c = W * Vinv with W[i][k] = V[i][k] e_c[k] as 3 x 3 tiles of m8n8k4 (45 DMMA, nine independent accumulator chains):
// the eight MMA warps of a CLV kernel build the eight matrices of a launch in ~1,500 cycles (FMA version above: ~3,200),
// and an inner child's matrix comes out of the accumulators directly in the B-fragment layout the CLV loop wants.
// Summation order over k differs from the host code in the last bit; the engine's results are compared with the oracle at
// 1e-10 relative.
__device__ __forceinline__ void dmma_p(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// acc[nt][jt] = P_c[nt*8 + g][jt*8 + 2t + {0,1}] for lane = 4g + t.  e_lane = exp(lambda_k r_c t) held by lane k (< 20):
// the warp computes its own twenty exponentials, so nothing but V and Vinv (static, staged before the dependency wait) has to
// be in shared memory.
// Columns 20, 21, 22 of the padded product are not wasted: with DeviceModel::vinv_codes as three more columns of Vinv they come
// out as P_c summed over the residues of the ambiguity codes B, Z and "undetermined" -- the three rows of a tip's look-up table
// that used to be added up with FP64 SIMT instructions and shuffles next to the other warps' DMMAs (0.68 us for a table against
// 0.12 us for a set of fragments).  extra_b() fetches a lane's share of those columns (static: before the dependency wait).
struct ExtraB {
    double v[5];  // for k = 4 ks + t: column 16 + g of the extended Vinv (g = 4, 5, 6; 0 elsewhere)
};
__device__ __forceinline__ ExtraB extra_b(const DeviceModel* dm, int lane) {
    const int g = lane >> 2, t = lane & 3;
    ExtraB x;
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) x.v[ks] = (g >= 4 && g < 7) ? dm->vinv_codes[4 * ks + t][g - 4] : 0.0;
    return x;
}
__device__ __forceinline__ void build_p_tiles(const double* s_model, double e_lane, int lane, const ExtraB& xb, double (&acc)[3][3][2]) {
    const double* s_V = s_model;
    const double* s_Vinv = s_model + kMat;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int jt = 0; jt < 3; ++jt) acc[nt][jt][0] = acc[nt][jt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
        const int k = 4 * ks + t;
        const double e = __shfl_sync(0xffffffffu, e_lane, k);
        double a[3], b[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int ij = q * 8 + g;
            a[q] = ij < kStates ? s_V[ij * kStates + k] * e : 0.0;
            b[q] = ij < kStates ? s_Vinv[k * kStates + ij] : (q == 2 ? xb.v[ks] : 0.0);
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int jt = 0; jt < 3; ++jt) dmma_p(acc[nt][jt][0], acc[nt][jt][1], a[nt], b[jt]);
    }
}
// accumulators -> B fragments of the CLV loop: frag[nt][kt] = P_c[nt*8 + g][kmap(kt, t)], kmap = 2t, 2t+1, 8+2t, 9+2t, 16+t
__device__ __forceinline__ void tiles_to_fragments(const double (&acc)[3][3][2], int lane, double (&frag)[3][5]) {
    const int t = lane & 3, src = (lane & ~3) | (t >> 1);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        frag[nt][0] = acc[nt][0][0];
        frag[nt][1] = acc[nt][0][1];
        frag[nt][2] = acc[nt][1][0];
        frag[nt][3] = acc[nt][1][1];
        const double v0 = __shfl_sync(0xffffffffu, acc[nt][2][0], src), v1 = __shfl_sync(0xffffffffu, acc[nt][2][1], src);
        frag[nt][4] = (t & 1) ? v1 : v0;
    }
}
// accumulators -> P_c[i][j] row-major in shared memory (a tip child's matrix feeds the lookup table)
__device__ __forceinline__ void tiles_to_smem(const double (&acc)[3][3][2], int lane, double* Pc) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int jt = 0; jt < 3; ++jt) {
            const int i = nt * 8 + g, j = jt * 8 + 2 * t;
            if (i < kStates && j < kStates) *reinterpret_cast<double2*>(Pc + i * kStates + j) = make_double2(acc[nt][jt][0], acc[nt][jt][1]);
        }
}
// accumulators -> the 23 x 80 look-up of a tip on that branch, tip[code][c*20 + i] = sum_j P_c[i][j] indicator(code)[j], written
// straight from the warp that built the matrix (category c): residue codes are single columns of P_c, and columns 20 - 22 of the
// extended product (build_p_tiles) are the rows of B = N|D, Z = Q|E and "undetermined".  Rows padded to `pad` doubles.
__device__ __forceinline__ void tiles_to_lookup(const double (&acc)[3][3][2], int lane, int c, double* table, int pad) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int i = nt * 8 + g;
        if (i < kStates) {
            double* col = table + c * kStates + i;
#pragma unroll
            for (int jt = 0; jt < 3; ++jt)
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    const int j = jt * 8 + 2 * t + x;
                    if (j < kCodes) col[j * pad] = acc[nt][jt][x];
                }
        }
    }
}
// fragment exchange between the two MMA groups: [warp slot][fragment][lane]
__device__ __forceinline__ void fragments_to_smem(const double (&frag)[3][5], int lane, double* slot) {
#pragma unroll
    for (int q = 0; q < 15; ++q) slot[q * 32 + lane] = frag[q / 5][q % 5];
}
__device__ __forceinline__ void fragments_from_smem(double (&frag)[3][5], int lane, const double* slot) {
#pragma unroll
    for (int q = 0; q < 15; ++q) frag[q / 5][q % 5] = slot[q * 32 + lane];
}
constexpr int kFragSlotDoubles = 15 * 32;

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
