// CUDA kernels of the likelihood path (sm_100a).  Baseline generation: one thread per (pattern, rate category),
// transition matrices staged in shared memory, deterministic two-stage reductions.
#include "kernels.h"

#include <cstdio>

namespace pml {

namespace {

constexpr double kTwo256 = 1.157920892373161954235709850086879078532699846656405640394575840079131296399e77;
constexpr double kMinLik = 8.636168555094444625386351862800399571116000364436281385023703470168591803162e-78;
constexpr double kLogMinLik = -177.445678223345993274;  // ln 2^-256

constexpr int kPatternsPerBlock = 32;
constexpr int kThreads = kPatternsPerBlock * kCats;

__device__ __forceinline__ void indicator(int code, double* v) {
#pragma unroll
    for (int j = 0; j < kStates; ++j) v[j] = code >= 22 ? 1.0 : 0.0;
    if (code < 20) {
#pragma unroll
        for (int j = 0; j < kStates; ++j)
            if (j == code) v[j] = 1.0;
    } else if (code == 20) {
        v[2] = v[3] = 1.0;
    } else if (code == 21) {
        v[5] = v[6] = 1.0;
    }
}

__device__ __forceinline__ void load_row20(const double* __restrict__ src, double* v) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int j = 0; j < kStates / 2; ++j) {
        const double2 t = __ldg(s2 + j);
        v[2 * j] = t.x;
        v[2 * j + 1] = t.y;
    }
}

// ------------------------------------------------------------------------------------------------ P(t) ----
__global__ void __launch_bounds__(256) k_make_p(const DeviceModel* __restrict__ dm, const double* __restrict__ lengths,
                                                const uint8_t* __restrict__ want_tip, PBlock* __restrict__ blocks) {
    __shared__ double s_exp[kCats][kStates];
    __shared__ double s_P[kCats][kStates][kStates];
    const int b = blockIdx.x;
    const double t = lengths[b];
    if (threadIdx.x < kRow) {
        const int c = threadIdx.x / kStates, k = threadIdx.x % kStates;
        s_exp[c][k] = exp(dm->lambda[k] * dm->rates[c] * t);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kCats * kStates * kStates; idx += blockDim.x) {
        const int c = idx / (kStates * kStates), i = (idx / kStates) % kStates, j = idx % kStates;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < kStates; ++k) acc = fma(dm->V[i][k] * s_exp[c][k], dm->Vinv[k][j], acc);
        s_P[c][i][j] = acc;
        blocks[b].P[c][i][j] = acc;
    }
    if (!want_tip[b]) return;
    __syncthreads();
    for (int idx = threadIdx.x; idx < kCodes * kRow; idx += blockDim.x) {
        const int code = idx / kRow, c = (idx % kRow) / kStates, i = idx % kStates;
        double acc;
        if (code < 20) acc = s_P[c][i][code];
        else if (code == 20) acc = s_P[c][i][2] + s_P[c][i][3];
        else if (code == 21) acc = s_P[c][i][5] + s_P[c][i][6];
        else {
            acc = 0.0;
#pragma unroll
            for (int j = 0; j < kStates; ++j) acc += s_P[c][i][j];
        }
        blocks[b].tip[code][c * kStates + i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ newview --
template <bool kTipL, bool kTipR>
__global__ void __launch_bounds__(kThreads) k_newview(NewviewOp op, int64_t np) {
    // per child either the 4x20x20 matrices (inner) or the 23x80 lookup (tip)
    __shared__ double s_l[kTipL ? kCodes * kRow : kCats * kStates * kStates];
    __shared__ double s_r[kTipR ? kCodes * kRow : kCats * kStates * kStates];
    __shared__ double s_max[kCats][kPatternsPerBlock];
    {
        const double* gl = kTipL ? &op.pleft->tip[0][0] : &op.pleft->P[0][0][0];
        const double* gr = kTipR ? &op.pright->tip[0][0] : &op.pright->P[0][0][0];
        for (int i = threadIdx.x; i < (kTipL ? kCodes * kRow : kCats * kStates * kStates); i += kThreads) s_l[i] = gl[i];
        for (int i = threadIdx.x; i < (kTipR ? kCodes * kRow : kCats * kStates * kStates); i += kThreads) s_r[i] = gr[i];
    }
    __syncthreads();
    const int c = threadIdx.x / kPatternsPerBlock, lane = threadIdx.x % kPatternsPerBlock;
    const int64_t p = (int64_t)blockIdx.x * kPatternsPerBlock + lane;
    const bool live = p < np;
    double res[kStates];
    double big = 0.0;
    if (live) {
        double x[kStates];
        if (kTipL) {
            const double* row = s_l + (int)op.left.codes[p] * kRow + c * kStates;
#pragma unroll
            for (int i = 0; i < kStates; ++i) res[i] = row[i];
        } else {
            load_row20(op.left.clv + p * kRow + c * kStates, x);
            const double* P = s_l + c * kStates * kStates;
#pragma unroll
            for (int i = 0; i < kStates; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < kStates; ++j) acc = fma(P[i * kStates + j], x[j], acc);
                res[i] = acc;
            }
        }
        if (kTipR) {
            const double* row = s_r + (int)op.right.codes[p] * kRow + c * kStates;
#pragma unroll
            for (int i = 0; i < kStates; ++i) res[i] *= row[i];
        } else {
            load_row20(op.right.clv + p * kRow + c * kStates, x);
            const double* P = s_r + c * kStates * kStates;
#pragma unroll
            for (int i = 0; i < kStates; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < kStates; ++j) acc = fma(P[i * kStates + j], x[j], acc);
                res[i] *= acc;
            }
        }
#pragma unroll
        for (int i = 0; i < kStates; ++i) big = fmax(big, fabs(res[i]));
    }
    s_max[c][lane] = big;
    __syncthreads();
    if (!live) return;
    const double m = fmax(fmax(s_max[0][lane], s_max[1][lane]), fmax(s_max[2][lane], s_max[3][lane]));
    const bool rescale = m < kMinLik;
    double2* dst = reinterpret_cast<double2*>(op.out + p * kRow + c * kStates);
#pragma unroll
    for (int i = 0; i < kStates / 2; ++i) {
        double2 v;
        v.x = rescale ? res[2 * i] * kTwo256 : res[2 * i];
        v.y = rescale ? res[2 * i + 1] * kTwo256 : res[2 * i + 1];
        dst[i] = v;
    }
    if (c == 0) {
        int32_t s = rescale ? 1 : 0;
        if (!kTipL) s += op.left.scale[p];
        if (!kTipR) s += op.right.scale[p];
        op.out_scale[p] = s;
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

// ------------------------------------------------------------------------------------------------ evaluate -
// b must be an inner node; a may be a tip
template <bool kTipA>
__global__ void __launch_bounds__(kThreads) k_evaluate(const DeviceModel* __restrict__ dm, Side a, Side b,
                                                       const PBlock* __restrict__ pb, const int32_t* __restrict__ weights,
                                                       int64_t np, double* __restrict__ site_lnl, double* __restrict__ partials) {
    __shared__ double s_P[kCats * kStates * kStates];
    __shared__ double s_pi[kStates];
    __shared__ double s_term[kCats][kPatternsPerBlock];
    for (int i = threadIdx.x; i < kCats * kStates * kStates; i += kThreads) s_P[i] = (&pb->P[0][0][0])[i];
    if (threadIdx.x < kStates) s_pi[threadIdx.x] = dm->pi[threadIdx.x];
    __syncthreads();
    const int c = threadIdx.x / kPatternsPerBlock, lane = threadIdx.x % kPatternsPerBlock;
    const int64_t p = (int64_t)blockIdx.x * kPatternsPerBlock + lane;
    const bool live = p < np;
    double term = 0.0;
    if (live) {
        double xa[kStates], xb[kStates];
        if (kTipA) indicator(a.codes[p], xa);
        else load_row20(a.clv + p * kRow + c * kStates, xa);
        load_row20(b.clv + p * kRow + c * kStates, xb);
        const double* P = s_P + c * kStates * kStates;
#pragma unroll
        for (int i = 0; i < kStates; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < kStates; ++j) acc = fma(P[i * kStates + j], xb[j], acc);
            term = fma(s_pi[i] * xa[i], acc, term);
        }
    }
    s_term[c][lane] = term;
    __syncthreads();
    if (c != 0) return;
    double contrib = 0.0;
    if (live) {
        const double f = (s_term[0][lane] + s_term[1][lane]) + (s_term[2][lane] + s_term[3][lane]);
        int32_t sc = b.scale[p];
        if (!kTipA) sc += a.scale[p];
        const double l = log(0.25 * f) + sc * kLogMinLik;
        site_lnl[p] = l;
        contrib = (double)weights[p] * l;
    }
    contrib = warp_sum(contrib);
    if (lane == 0) partials[blockIdx.x] = contrib;
}

// ------------------------------------------------------------------------------------------------ sumtable -
template <bool kTipA>
__global__ void __launch_bounds__(kThreads) k_sumtable(const DeviceModel* __restrict__ dm, Side a, Side b, int64_t np,
                                                       double* __restrict__ sumtable, int32_t* __restrict__ sum_scale) {
    __shared__ double s_piV[kStates][kStates];   // [i][k]
    __shared__ double s_Vinv[kStates][kStates];  // [k][i]
    for (int i = threadIdx.x; i < kStates * kStates; i += kThreads) {
        (&s_piV[0][0])[i] = (&dm->piV[0][0])[i];
        (&s_Vinv[0][0])[i] = (&dm->Vinv[0][0])[i];
    }
    __syncthreads();
    const int c = threadIdx.x / kPatternsPerBlock, lane = threadIdx.x % kPatternsPerBlock;
    const int64_t p = (int64_t)blockIdx.x * kPatternsPerBlock + lane;
    if (p >= np) return;
    double xa[kStates], xb[kStates];
    if (kTipA) indicator(a.codes[p], xa);
    else load_row20(a.clv + p * kRow + c * kStates, xa);
    load_row20(b.clv + p * kRow + c * kStates, xb);
    double2* dst = reinterpret_cast<double2*>(sumtable + p * kRow + c * kStates);
#pragma unroll
    for (int k = 0; k < kStates; k += 2) {
        double l0 = 0.0, r0 = 0.0, l1 = 0.0, r1 = 0.0;
#pragma unroll
        for (int i = 0; i < kStates; ++i) {
            l0 = fma(s_piV[i][k], xa[i], l0);
            r0 = fma(s_Vinv[k][i], xb[i], r0);
            l1 = fma(s_piV[i][k + 1], xa[i], l1);
            r1 = fma(s_Vinv[k + 1][i], xb[i], r1);
        }
        dst[k / 2] = make_double2(l0 * r0, l1 * r1);
    }
    if (c == 0) {
        int32_t sc = b.scale[p];
        if (!kTipA) sc += a.scale[p];
        sum_scale[p] = sc;
    }
}

// ------------------------------------------------------------------------------------------------ NR core --
__global__ void __launch_bounds__(kThreads) k_core(const DeviceModel* __restrict__ dm, const double* __restrict__ sumtable,
                                                   const int32_t* __restrict__ sum_scale, const int32_t* __restrict__ weights,
                                                   int64_t np, const double* __restrict__ d_t, double* __restrict__ partials) {
    __shared__ double s_e[3][kCats][kStates];
    __shared__ double s_f[3][kCats][kPatternsPerBlock];
    if (threadIdx.x < kRow) {
        const int c = threadIdx.x / kStates, k = threadIdx.x % kStates;
        const double a = dm->lambda[k] * dm->rates[c], e = exp(a * d_t[0]);
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

void launch_make_p(const DeviceModel* dm, const double* d_lengths, const uint8_t* d_want_tip, PBlock* d_blocks, int nblocks,
                   cudaStream_t stream) {
    if (nblocks > 0) k_make_p<<<nblocks, 256, 0, stream>>>(dm, d_lengths, d_want_tip, d_blocks);
}

void launch_newview(const NewviewOp& op, int64_t np, cudaStream_t stream) {
    const int grid = blocks_for(np);
    const bool tl = op.left.clv == nullptr, tr = op.right.clv == nullptr;
    if (tl && tr) k_newview<true, true><<<grid, kThreads, 0, stream>>>(op, np);
    else if (tl) k_newview<true, false><<<grid, kThreads, 0, stream>>>(op, np);
    else if (tr) k_newview<false, true><<<grid, kThreads, 0, stream>>>(op, np);
    else k_newview<false, false><<<grid, kThreads, 0, stream>>>(op, np);
}

void launch_evaluate(const DeviceModel* dm, const Side& a, const Side& b, const PBlock* p, const int32_t* weights, int64_t np,
                     double* site_lnl, double* partials, double* result, cudaStream_t stream) {
    const int grid = blocks_for(np);
    if (a.clv == nullptr) k_evaluate<true><<<grid, kThreads, 0, stream>>>(dm, a, b, p, weights, np, site_lnl, partials);
    else k_evaluate<false><<<grid, kThreads, 0, stream>>>(dm, a, b, p, weights, np, site_lnl, partials);
    k_reduce<<<1, 1024, 0, stream>>>(partials, grid, 1, result);
}

void launch_sumtable(const DeviceModel* dm, const Side& a, const Side& b, int64_t np, double* sumtable, int32_t* sum_scale,
                     cudaStream_t stream) {
    const int grid = blocks_for(np);
    if (a.clv == nullptr) k_sumtable<true><<<grid, kThreads, 0, stream>>>(dm, a, b, np, sumtable, sum_scale);
    else k_sumtable<false><<<grid, kThreads, 0, stream>>>(dm, a, b, np, sumtable, sum_scale);
}

void launch_core(const DeviceModel* dm, const double* sumtable, const int32_t* sum_scale, const int32_t* weights, int64_t np,
                 const double* d_t, double* partials, double* result, cudaStream_t stream) {
    const int grid = blocks_for(np);
    k_core<<<grid, kThreads, 0, stream>>>(dm, sumtable, sum_scale, weights, np, d_t, partials);
    k_reduce<<<1, 1024, 0, stream>>>(partials, grid, 3, result);
}

void launch_replicate_lnl(const int32_t* W, int nrep, int64_t np, int64_t ldw, const double* site_lnl, double* lnl,
                          cudaStream_t stream) {
    if (nrep > 0) k_replicate_lnl<<<nrep, 256, 0, stream>>>(W, np, ldw, site_lnl, lnl);
}

int64_t reduce_partials_capacity(int64_t np) { return 3 * (int64_t)blocks_for(np) + 8; }

}  // namespace pml
