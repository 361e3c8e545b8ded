// Branch kernel on the FP64 tensor path: for the two ends of a branch it forms, per pattern and rate category, the
// eigen-space product s[c][k] = (sum_i pi_i xa[c][i] V[i][k]) * (sum_i Vinv[k][i] xb[c][i])  (raxmlHPC sumGAMMAPROT) and
// contracts it at once with exp(lambda_k r_c t) * {1, lambda r, (lambda r)^2} (coreGTRGAMMAPROT), so that one pass over
// the two CLVs yields lnL, dlnL/dt and d2lnL/dt2 -- and, with t = the branch's own length, the root evaluate
// (evaluateGTRGAMMAPROT) including per-pattern lnL.  The product table is only written out when the caller wants to
// iterate Newton-Raphson on it (kStore).
//
// The streaming loop only leaves three sums (f, f', f'') per pattern; after its last tile every CTA turns the sums of its
// own tiles into per-pattern lnL and weighted totals with all 384 threads (logs, divisions and integer weights stay out of
// the loop whose FP64 pipe the MMAs need), and the CTA that draws the last ticket adds the CTA partials in a fixed order:
// one launch replaces sumtable + core + reduce of the first engine generation and its result is bit-reproducible.
//
// Same pipeline as newview_mma.cu: three groups of four warps (one per rate category), each with a private ring of
// shared-memory stages filled by TMA bulk copies, DMMA m8n8k4 with the fixed 20x20 matrices as B fragments in registers.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "mma_common.cuh"

namespace pml {

namespace {

using namespace mma;

constexpr double kLogMinLik = -177.445678223345993274;  // ln 2^-256
constexpr int kTipVecPad = 22;                          // doubles per row of the 23 x 20 tip table

template <bool kTipA>
struct BranchPlan {
    static constexpr int kInner = kTipA ? 1 : 2;
    static constexpr int kStages = kGroups * kDepth;
    static constexpr int kStageDoubles = kInner * kTileDoubles;
    static constexpr int kTipDoubles = kTipA ? kCodes * kTipVecPad : 0;
    static constexpr int kRedDoubles = 2 * kGroups * kCats * kTileRows * 3;  // [parity][group][cat][row][f,f1,f2]
    static constexpr int kFinalDoubles = 0;
    static constexpr size_t kBytes = 128 + sizeof(double) * (size_t)(kTipDoubles + kRedDoubles + kFinalDoubles + kStages * kStageDoubles);
};

template <bool kTipA, bool kStore>
__global__ void __launch_bounds__(kThreadsMma, 1) k_branch_mma(BranchArgs args, int ntiles) {
    using Plan = BranchPlan<kTipA>;
    constexpr int ST = Plan::kStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_tip = reinterpret_cast<double*>(smem_raw + 128);
    double* s_red = s_tip + Plan::kTipDoubles;
    double* s_final = s_red + Plan::kRedDoubles;
    double* s_stage = s_final + Plan::kFinalDoubles;
    const DeviceModel* dm = args.dm;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kTipA) {
        // tipvec[code][k] = sum over the residues the code allows of pi_i V[i][k]
        for (int idx = threadIdx.x; idx < kCodes * kStates; idx += kThreadsMma) {
            const int code = idx / kStates, k = idx % kStates;
            double acc = 0.0;
            if (code < 20) acc = dm->piV[code][k];
            else if (code == 20) acc = dm->piV[2][k] + dm->piV[3][k];
            else if (code == 21) acc = dm->piV[5][k] + dm->piV[6][k];
            else
                for (int i = 0; i < kStates; ++i) acc += dm->piV[i][k];
            s_tip[code * kTipVecPad + k] = acc;
        }
    }
    __syncthreads();

    const int c = warp & 3, grp = warp >> 2, g = lane >> 2, t = lane & 3;
    uint64_t* gfull = full + grp * kDepth;
    double* gstage = s_stage + (size_t)grp * kDepth * Plan::kStageDoubles;
    const int stride = kGroups * gridDim.x;
    auto refill = [&](int tile, int slot) {
        if (lane == 0) {
            constexpr uint32_t bytes = kTileDoubles * sizeof(double);
            mbar_expect_tx(gfull + slot, Plan::kInner * bytes);
            const size_t goff = (size_t)tile * kTileDoubles;
            double* dst = gstage + (size_t)slot * Plan::kStageDoubles;
            if (!kTipA) {
                bulk_g2s(dst, args.a.clv + goff, bytes, gfull + slot);
                dst += kTileDoubles;
            }
            bulk_g2s(dst, args.b.clv + goff, bytes, gfull + slot);
        }
    };
    const int first = blockIdx.x + grp * gridDim.x;
    if (c == 0)
        for (int d = 0; d < kDepth; ++d)
            if (first + d * stride < ntiles) refill(first + d * stride, d);
    // B fragments: a-side contracts x with pi_i V[i][k] (output k), b-side with Vinv[k][i]
    double fragA[3][5], fragB[3][5];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int k = nt * 8 + g;
#pragma unroll
        for (int kt = 0; kt < 5; ++kt) {
            fragA[nt][kt] = (!kTipA && k < kStates) ? dm->piV[kmap(kt, t)][k] : 0.0;
            fragB[nt][kt] = k < kStates ? dm->Vinv[k][kmap(kt, t)] : 0.0;
        }
    }
    // exp(lambda_k r_c t) and its first two t-derivatives at the D-fragment positions k = nt*8 + 2t + {0,1}
    double e0[3][2], e1[3][2], e2[3][2];
    {
        const double tt = args.t, rate = dm->rates[c];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = nt * 8 + 2 * t + j;
                const double a = k < kStates ? dm->lambda[k] * rate : 0.0;
                const double e = k < kStates ? exp(a * tt) : 0.0;
                e0[nt][j] = e;
                e1[nt][j] = a * e;
                e2[nt][j] = a * a * e;
            }
    }

    int next_code[2] = {0, 0};
    if (kTipA && first < ntiles) {
        next_code[0] = __ldg(args.a.codes + (int64_t)first * kTileRows + g);
        next_code[1] = __ldg(args.a.codes + (int64_t)first * kTileRows + 8 + g);
    }
    const int cta_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int rounds = (cta_tiles + kGroups - 1) / kGroups;
    mma_turn_init(grp);
    for (int it = 0; it < rounds; ++it) {
        const int tile = first + it * stride;
        if (tile >= ntiles) {  // no tile left for this group: keep the MMA token moving
            mma_turn_begin(grp);
            mma_turn_end(grp);
            continue;
        }
        const int slot = it % kDepth;
        const int64_t row0 = (int64_t)tile * kTileRows;
        int code[2] = {next_code[0], next_code[1]};
        if (kTipA && tile + stride < ntiles) {  // the codes of the following tile travel while this one is computed
            next_code[0] = __ldg(args.a.codes + row0 + (int64_t)stride * kTileRows + g);
            next_code[1] = __ldg(args.a.codes + row0 + (int64_t)stride * kTileRows + 8 + g);
        }
        // per-row integers are only needed after the MMAs: issue the loads now, consume them at the end
        mbar_wait(gfull + slot, (it / kDepth) & 1);
        const double* stage = gstage + (size_t)slot * Plan::kStageDoubles;
        AFrag fa[2], fb[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (!kTipA) fa[m] = load_a(stage + m * kBlockDoubles, c, lane);
            fb[m] = load_a(stage + (kTipA ? 0 : kTileDoubles) + m * kBlockDoubles, c, lane);
        }
        double accA[2][3][2], accB[2][3][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                accB[m][nt][0] = accB[m][nt][1] = 0.0;
                if (kTipA) {
                    const bool ok = nt < 2 || t < 2;
                    const double* row = s_tip + code[m] * kTipVecPad + nt * 8 + 2 * t;
                    accA[m][nt][0] = ok ? row[0] : 0.0;
                    accA[m][nt][1] = ok ? row[1] : 0.0;
                } else {
                    accA[m][nt][0] = accA[m][nt][1] = 0.0;
                }
            }
        mma_turn_begin(grp);  // see mma_common.cuh: the groups take turns on the FP64 tensor pipe
#pragma unroll
        for (int kt = 0; kt < 5; ++kt)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    if (!kTipA) dmma(accA[m][nt][0], accA[m][nt][1], fa[m].v[kt], fragA[nt][kt]);
                    dmma(accB[m][nt][0], accB[m][nt][1], fb[m].v[kt], fragB[nt][kt]);
                }
        mma_turn_end(grp);
        double* red = s_red + (((it & 1) * kGroups + grp) * kCats) * kTileRows * 3;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double f = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double sv = accA[m][nt][j] * accB[m][nt][j];
                    accA[m][nt][j] = sv;
                    f = fma(sv, e0[nt][j], f);
                    f1 = fma(sv, e1[nt][j], f1);
                    f2 = fma(sv, e2[nt][j], f2);
                }
            f += __shfl_xor_sync(0xffffffffu, f, 1);
            f1 += __shfl_xor_sync(0xffffffffu, f1, 1);
            f2 += __shfl_xor_sync(0xffffffffu, f2, 1);
            f += __shfl_xor_sync(0xffffffffu, f, 2);
            f1 += __shfl_xor_sync(0xffffffffu, f1, 2);
            f2 += __shfl_xor_sync(0xffffffffu, f2, 2);
            if (t == 0) {
                double* dst = red + (c * kTileRows + m * 8 + g) * 3;
                dst[0] = f;
                dst[1] = f1;
                dst[2] = f2;
            }
            if (kStore) {
                double* out = args.sumtable + (row0 + m * 8 + g) * kRow + c * kStates + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
                    if (nt < 2 || t < 2) *reinterpret_cast<double2*>(out + nt * 8) = make_double2(accA[m][nt][0], accA[m][nt][1]);
            }
        }
        named_barrier(1 + grp, 4 * 32);
        if (c == 0 && tile + kDepth * stride < ntiles) refill(tile + kDepth * stride, slot);
        // the four warps of a group share the row sums: warp c adds the categories of rows 4c .. 4c+3 (3 values each)
        if (lane < 12) {
            const int r = c * 4 + lane / 3, v = lane % 3;
            const double sum = (red[(0 * kTileRows + r) * 3 + v] + red[(1 * kTileRows + r) * 3 + v]) +
                               (red[(2 * kTileRows + r) * 3 + v] + red[(3 * kTileRows + r) * 3 + v]);
            args.rowsum[(row0 + r) * 3 + v] = sum;
        }
    }

    // ---- finish: the rows of this CTA's own tiles, all threads -------------------------------------------------
    __syncthreads();
    double sum_l = 0.0, sum_d1 = 0.0, sum_d2 = 0.0;
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (int idx = threadIdx.x; idx < my_tiles * kTileRows; idx += kThreadsMma) {
        const int64_t p = ((int64_t)blockIdx.x + (int64_t)(idx / kTileRows) * gridDim.x) * kTileRows + idx % kTileRows;
        const double f = args.rowsum[p * 3], f1 = args.rowsum[p * 3 + 1], f2 = args.rowsum[p * 3 + 2];
        int32_t sc = __ldg(args.b.scale + p);
        if (!kTipA) sc += __ldg(args.a.scale + p);
        const double w = (double)__ldg(args.weights + p), inv = 1.0 / f, q = f1 * inv;
        const double lnl = log(0.25 * f) + sc * kLogMinLik;
        if (args.site_lnl) args.site_lnl[p] = lnl;
        if (kStore) args.sum_scale[p] = sc;
        sum_l = fma(w, lnl, sum_l);
        sum_d1 = fma(w, q, sum_d1);
        sum_d2 = fma(w, f2 * inv - q * q, sum_d2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_l += __shfl_xor_sync(0xffffffffu, sum_l, o);
        sum_d1 += __shfl_xor_sync(0xffffffffu, sum_d1, o);
        sum_d2 += __shfl_xor_sync(0xffffffffu, sum_d2, o);
    }
    double* s_fin = s_red;           // the exchange buffer of the loop is free now
    __shared__ bool s_last;
    if (lane == 0) {
        s_fin[warp * 3 + 0] = sum_l;
        s_fin[warp * 3 + 1] = sum_d1;
        s_fin[warp * 3 + 2] = sum_d2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int k = 0; k < kComputeWarps; ++k) v += s_fin[k * 3 + threadIdx.x];
        args.partials[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(args.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp < 3) {  // warp v adds value v over the CTAs: lane-strided, then a shuffle tree -- the same order every run
        double acc = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) acc += args.partials[(int64_t)warp * gridDim.x + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            args.result[warp] = acc;
            if (args.host_result) args.host_result[warp] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *args.ticket = 0;
        if (args.host_result) {
            __threadfence_system();
            args.host_result[3] = args.sequence;
        }
    }
}

template <bool kTipA, bool kStore>
void launch_one(const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const int ntiles = (int)(np / kTileRows);
    const int grid = ntiles < sms ? ntiles : sms;
    k_branch_mma<kTipA, kStore><<<grid, kThreadsMma, BranchPlan<kTipA>::kBytes, stream>>>(args, ntiles);
}

}  // namespace

void configure_branch_kernels() {
    cudaFuncSetAttribute(k_branch_mma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BranchPlan<true>::kBytes);
    cudaFuncSetAttribute(k_branch_mma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BranchPlan<true>::kBytes);
    cudaFuncSetAttribute(k_branch_mma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BranchPlan<false>::kBytes);
    cudaFuncSetAttribute(k_branch_mma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BranchPlan<false>::kBytes);
}

// args.result[0..2] = lnL, dlnL/dt, d2lnL/dt2 of this rank's patterns.  np must be a multiple of 16.
void launch_branch_mma(const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const bool tip = args.a.clv == nullptr, store = args.sumtable != nullptr;
    if (tip) store ? launch_one<true, true>(args, np, sms, stream) : launch_one<true, false>(args, np, sms, stream);
    else store ? launch_one<false, true>(args, np, sms, stream) : launch_one<false, false>(args, np, sms, stream);
}

}  // namespace pml
