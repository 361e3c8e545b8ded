// newview on the FP64 tensor path (sm_100a): CLV tiles stream HBM -> shared memory through the TMA engine
// (cp.async.bulk + mbarrier ring), the 20x20 contractions run as DMMA m8n8k4 with P held in registers as B fragments,
// results go back with 128-bit stores.
//
// Why DMMA: tools/fp64_peak.cu measured 37.0 TFLOP/s for DMMA.8x8x4 against 33.5 for DFMA on B200 and showed that both
// share one pipe (profiles/r01_fp64_pipe_peaks.log).  An inner-inner site-update needs 6,480 flop per 1,920 B, i.e.
// 22 TFLOP/s at the measured 6.5 TB/s -- only the tensor path leaves issue slots and register bandwidth for the rest.
//
// Work split inside a CTA (288 threads): warps 0-7 compute, warp 8 produces.
//   compute warp w: rate category c = w & 3, rows (w >> 2) * 16 .. + 16 of the 32-pattern tile (two 8-row MMA tiles)
//   per 8 rows and child: D[8 x 24] = X[8 x 20] * P_c^T[20 x 24(20 used)] as 3 n-tiles x 5 k-tiles of m8n8k4
// Shared-memory rows are padded 640 -> 704 B so that the 8 rows of an A fragment fall into distinct banks.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace pml {

namespace {

constexpr int kTileRows = 32;                // patterns per pipeline stage
constexpr int kRowPad = 88;                  // doubles per padded smem row (704 B)
constexpr int kTipPad = 82;                  // doubles per padded tip-table row
constexpr int kComputeWarps = 8;
constexpr int kThreadsMma = (kComputeWarps + 1) * 32;
constexpr double kTwo256 = 1.157920892373161954235709850086879078532699846656405640394575840079131296399e77;
constexpr double kMinLik = 8.636168555094444625386351862800399571116000364436281385023703470168591803162e-78;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA engine bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// state index that lane t feeds into k-tile kt: pairs of k-tiles share one 128-bit shared-memory load
__device__ __forceinline__ int kmap(int kt, int t) { return kt < 4 ? (kt >> 1) * 8 + 2 * t + (kt & 1) : 16 + t; }

// B fragments of P_c^T for one child: frag[nt][kt] = P_c[nt*8 + g][kmap(kt, t)] (0 beyond state 19)
__device__ __forceinline__ void load_p_fragments(const PBlock* pb, int c, int g, int t, double (&frag)[3][5]) {
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int i = nt * 8 + g;
#pragma unroll
        for (int kt = 0; kt < 5; ++kt) frag[nt][kt] = i < kStates ? pb->P[c][i][kmap(kt, t)] : 0.0;
    }
}

// acc[nt][0..1] = sum_j P_c[nt*8 + 2t + {0,1}][j] * x[row g][c*20 + j] for the 8 rows starting at `rows`
__device__ __forceinline__ void contract_rows(const double* rows, int c, int g, int t, const double (&frag)[3][5], double (&acc)[3][2]) {
    const double* x = rows + g * kRowPad + c * kStates;
    const double2 a01 = *reinterpret_cast<const double2*>(x + 2 * t);
    const double2 a23 = *reinterpret_cast<const double2*>(x + 8 + 2 * t);
    const double a4 = x[16 + t];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        acc[nt][0] = 0.0;
        acc[nt][1] = 0.0;
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) dmma(acc[nt][0], acc[nt][1], a01.x, frag[nt][0]);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) dmma(acc[nt][0], acc[nt][1], a01.y, frag[nt][1]);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) dmma(acc[nt][0], acc[nt][1], a23.x, frag[nt][2]);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) dmma(acc[nt][0], acc[nt][1], a23.y, frag[nt][3]);
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) dmma(acc[nt][0], acc[nt][1], a4, frag[nt][4]);
}

// tip child: the same fragment positions read from the 23 x 80 lookup
__device__ __forceinline__ void lookup_rows(const double* table, int code, int c, int t, double (&acc)[3][2]) {
    const double* row = table + code * kTipPad + c * kStates + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        if (nt < 2 || t < 2) {
            const double2 v = *reinterpret_cast<const double2*>(row + nt * 8);
            acc[nt][0] = v.x;
            acc[nt][1] = v.y;
        } else {
            acc[nt][0] = 0.0;
            acc[nt][1] = 0.0;
        }
    }
}

template <bool kTipL, bool kTipR>
struct SmemPlan {
    static constexpr int kInner = (kTipL ? 0 : 1) + (kTipR ? 0 : 1);
    static constexpr int kStages = kInner == 2 ? 4 : (kInner == 1 ? 6 : 1);
    static constexpr int kStageDoubles = kInner * kTileRows * kRowPad;
    static constexpr int kTipDoubles = ((kTipL ? 1 : 0) + (kTipR ? 1 : 0)) * kCodes * kTipPad;
    static constexpr int kMaxDoubles = 2 * 2 * kCats * 16;  // [parity][half][cat][16 rows]
    static constexpr size_t kBytes = 128 /* barriers */ + sizeof(double) * (size_t)(kTipDoubles + kMaxDoubles + kStages * kStageDoubles);
};

template <bool kTipL, bool kTipR>
__global__ void __launch_bounds__(kThreadsMma, 1) k_newview_mma(NewviewOp op, int ntiles) {
    using Plan = SmemPlan<kTipL, kTipR>;
    constexpr int ST = Plan::kStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + ST;
    double* s_tip = reinterpret_cast<double*>(smem_raw + 128);
    double* s_max = s_tip + Plan::kTipDoubles;
    double* s_stage = s_max + Plan::kMaxDoubles;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kComputeWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kTipL || kTipR) {
        double* dst = s_tip;
        if (kTipL) {
            for (int i = threadIdx.x; i < kCodes * kRow; i += kThreadsMma) dst[(i / kRow) * kTipPad + i % kRow] = (&op.pleft->tip[0][0])[i];
            dst += kCodes * kTipPad;
        }
        if (kTipR)
            for (int i = threadIdx.x; i < kCodes * kRow; i += kThreadsMma) dst[(i / kRow) * kTipPad + i % kRow] = (&op.pright->tip[0][0])[i];
    }
    __syncthreads();
    const double* s_tipL = s_tip;
    const double* s_tipR = s_tip + (kTipL ? kCodes * kTipPad : 0);

    if (warp == kComputeWarps) {
        // ---------------------------------------------------------------- producer: one padded row per lane and child
        if (Plan::kInner == 0) return;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % ST;
            mbar_wait(empty + s, ((it / ST) & 1) ^ 1);
            if (lane == 0) mbar_expect_tx(full + s, Plan::kInner * kTileRows * kRow * (uint32_t)sizeof(double));
            __syncwarp();
            double* dst = s_stage + (size_t)s * Plan::kStageDoubles + lane * kRowPad;
            const size_t goff = ((size_t)tile * kTileRows + lane) * kRow;
            if (!kTipL) {
                bulk_g2s(dst, op.left.clv + goff, kRow * sizeof(double), full + s);
                dst += kTileRows * kRowPad;
            }
            if (!kTipR) bulk_g2s(dst, op.right.clv + goff, kRow * sizeof(double), full + s);
        }
        return;
    }

    // -------------------------------------------------------------------- consumers
    const int c = warp & 3, half = warp >> 2, g = lane >> 2, t = lane & 3;
    double fragL[3][5], fragR[3][5];
    if (!kTipL) load_p_fragments(op.pleft, c, g, t, fragL);
    if (!kTipR) load_p_fragments(op.pright, c, g, t, fragR);

    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % ST;
        const int64_t row0 = (int64_t)tile * kTileRows + half * 16;  // first of this warp's 16 rows
        int codeL[2] = {0, 0}, codeR[2] = {0, 0};
        if (kTipL) {
            codeL[0] = op.left.codes[row0 + g];
            codeL[1] = op.left.codes[row0 + 8 + g];
        }
        if (kTipR) {
            codeR[0] = op.right.codes[row0 + g];
            codeR[1] = op.right.codes[row0 + 8 + g];
        }
        double out[2][3][2];
        if (Plan::kInner > 0) mbar_wait(full + s, (it / ST) & 1);
        const double* stage = s_stage + (size_t)s * Plan::kStageDoubles + half * 16 * kRowPad;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double accL[3][2], accR[3][2];
            if (kTipL) lookup_rows(s_tipL, codeL[m], c, t, accL);
            else contract_rows(stage + m * 8 * kRowPad, c, g, t, fragL, accL);
            if (kTipR) lookup_rows(s_tipR, codeR[m], c, t, accR);
            else contract_rows(stage + (kTipL ? 0 : kTileRows * kRowPad) + m * 8 * kRowPad, c, g, t, fragR, accR);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                out[m][nt][0] = accL[nt][0] * accR[nt][0];
                out[m][nt][1] = accL[nt][1] * accR[nt][1];
            }
        }
        if (Plan::kInner > 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);  // this warp no longer reads the stage
        }
        // per-row magnitude over this category's 20 states, then over the four categories through shared memory
        double* mx = s_max + (((it & 1) * 2 + half) * kCats) * 16;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double big = 0.0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt)
                if (nt < 2 || t < 2) big = fmax(big, fmax(fabs(out[m][nt][0]), fabs(out[m][nt][1])));
            big = fmax(big, __shfl_xor_sync(0xffffffffu, big, 1));
            big = fmax(big, __shfl_xor_sync(0xffffffffu, big, 2));
            if (t == 0) mx[c * 16 + m * 8 + g] = big;
        }
        named_barrier(1 + half, 4 * 32);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int r = m * 8 + g;
            const double big = fmax(fmax(mx[r], mx[16 + r]), fmax(mx[32 + r], mx[48 + r]));
            const bool rescale = big < kMinLik;
            double* dst = op.out + (row0 + r) * kRow + c * kStates + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                if (nt < 2 || t < 2) {
                    double2 v = make_double2(out[m][nt][0], out[m][nt][1]);
                    if (rescale) {
                        v.x *= kTwo256;
                        v.y *= kTwo256;
                    }
                    *reinterpret_cast<double2*>(dst + nt * 8) = v;
                }
            }
            if (c == 0 && t == 0) {
                int32_t sc = rescale ? 1 : 0;
                if (!kTipL) sc += op.left.scale[row0 + r];
                if (!kTipR) sc += op.right.scale[row0 + r];
                op.out_scale[row0 + r] = sc;
            }
        }
    }
}

template <bool kTipL, bool kTipR>
void launch_one(const NewviewOp& op, int64_t np, int sms, cudaStream_t stream) {
    using Plan = SmemPlan<kTipL, kTipR>;
    const int ntiles = (int)(np / kTileRows);
    const int grid = ntiles < sms ? ntiles : sms;
    k_newview_mma<kTipL, kTipR><<<grid, kThreadsMma, Plan::kBytes, stream>>>(op, ntiles);
}

}  // namespace

void configure_mma_kernels() {
    cudaFuncSetAttribute(k_newview_mma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<true, true>::kBytes);
    cudaFuncSetAttribute(k_newview_mma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<true, false>::kBytes);
    cudaFuncSetAttribute(k_newview_mma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<false, true>::kBytes);
    cudaFuncSetAttribute(k_newview_mma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<false, false>::kBytes);
}

// np must be a multiple of 32 (the engine pads pattern rows to 128)
void launch_newview_mma(const NewviewOp& op, int64_t np, int sms, cudaStream_t stream) {
    const bool tl = op.left.clv == nullptr, tr = op.right.clv == nullptr;
    if (tl && tr) launch_one<true, true>(op, np, sms, stream);
    else if (tl) launch_one<true, false>(op, np, sms, stream);
    else if (tr) launch_one<false, true>(op, np, sms, stream);
    else launch_one<false, false>(op, np, sms, stream);
}

}  // namespace pml
