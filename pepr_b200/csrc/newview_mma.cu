// newview on the FP64 tensor path (sm_100a): CLV tiles stream HBM -> shared memory through the TMA engine
// (cp.async.bulk + mbarrier ring), the 20x20 contractions run as DMMA m8n8k4 with P held in registers as B fragments,
// results go back with 128-bit stores.
//
// Why DMMA: tools/fp64_peak.cu measured 37.0 TFLOP/s for DMMA.8x8x4 against 33.5 for DFMA on B200 and showed that both
// share one pipe (profiles/r01_fp64_pipe_peaks.log).  An inner-inner site-update needs 6,480 flop per 1,920 B, i.e.
// 22 TFLOP/s at the measured 6.5 TB/s -- only the tensor path leaves issue slots and register bandwidth for the rest.
//
// Work split inside a persistent CTA (384 threads = 3 groups x 4 warps):
//   warp w: rate category c = w & 3, group w >> 2; group k takes every third 16-pattern tile of the CTA and owns a private
//   ring of kDepth stages which its category-0 warp refills (one bulk row copy per lane) right after the group barrier,
//   so the three groups drift out of phase and one group's epilogue (scaling test, stores) overlaps the others' MMAs
//   per 8 rows and child: D[8 x 24] = X[8 x 20] * P_c^T[20 x 24(20 used)] as 3 n-tiles x 5 k-tiles of m8n8k4
// CLVs live in HBM in the blocked layout described in mma_common.cuh, so a stage is filled by one bulk copy per child.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "mma_common.cuh"

namespace pml {

namespace {

using namespace mma;

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
    static constexpr int kStages = kGroups * kDepth;
    static constexpr int kStageDoubles = kInner * kTileDoubles;
    static constexpr int kTipDoubles = ((kTipL ? 1 : 0) + (kTipR ? 1 : 0)) * kCodes * kTipPad;
    static constexpr int kMaxDoubles = 2 * kGroups * kCats * kTileRows;  // [parity][group][cat][16 rows]
    static constexpr size_t kBytes = 128 /* barriers */ + sizeof(double) * (size_t)(kTipDoubles + kMaxDoubles + kStages * kStageDoubles);
};

// at least one child is an inner node (the tip-tip case has its own kernel below)
template <bool kTipL, bool kTipR>
__global__ void __launch_bounds__(kThreadsMma, 1) k_newview_mma(NewviewOp op, int ntiles) {
    using Plan = SmemPlan<kTipL, kTipR>;
    constexpr int ST = Plan::kStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_tip = reinterpret_cast<double*>(smem_raw + 128);
    int* s_max = reinterpret_cast<int*>(s_tip + Plan::kTipDoubles);
    double* s_stage = s_tip + Plan::kTipDoubles + Plan::kMaxDoubles;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kTipL || kTipR) {
        const double* src = kTipL ? &op.pleft->tip[0][0] : &op.pright->tip[0][0];
        for (int i = threadIdx.x; i < kCodes * kRow; i += kThreadsMma) s_tip[(i / kRow) * kTipPad + i % kRow] = src[i];
    }
    __syncthreads();

    const int c = warp & 3, grp = warp >> 2, g = lane >> 2, t = lane & 3;
    uint64_t* gfull = full + grp * kDepth;
    double* gstage = s_stage + (size_t)grp * kDepth * Plan::kStageDoubles;
    const int stride = kGroups * gridDim.x;  // tile distance between two iterations of this group
    // refill of one ring slot: a 16-row tile is contiguous in the blocked layout -> one bulk copy per inner child
    auto refill = [&](int tile, int slot) {
        if (lane == 0) {
            constexpr uint32_t bytes = kTileDoubles * sizeof(double);
            mbar_expect_tx(gfull + slot, Plan::kInner * bytes);
            const size_t goff = (size_t)tile * kTileDoubles;
            double* dst = gstage + (size_t)slot * Plan::kStageDoubles;
            if (!kTipL) {
                bulk_g2s(dst, op.left.clv + goff, bytes, gfull + slot);
                dst += kTileDoubles;
            }
            if (!kTipR) bulk_g2s(dst, op.right.clv + goff, bytes, gfull + slot);
        }
    };
    const int first = blockIdx.x + grp * gridDim.x;
    if (c == 0)
        for (int d = 0; d < kDepth; ++d)
            if (first + d * stride < ntiles) refill(first + d * stride, d);

    double fragL[3][5], fragR[3][5];
    if (!kTipL) load_p_fragments(op.pleft, c, g, t, fragL);
    if (!kTipR) load_p_fragments(op.pright, c, g, t, fragR);

    int next_code[2] = {0, 0};
    if ((kTipL || kTipR) && first < ntiles) {
        const uint8_t* codes = kTipL ? op.left.codes : op.right.codes;
        next_code[0] = __ldg(codes + (int64_t)first * kTileRows + g);
        next_code[1] = __ldg(codes + (int64_t)first * kTileRows + 8 + g);
    }
    int32_t next_sc[4] = {0, 0, 0, 0};
    if (c == 0 && t == 0 && first < ntiles) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (!kTipL) next_sc[m] = __ldg(op.left.scale + (int64_t)first * kTileRows + m * 8 + g);
            if (!kTipR) next_sc[2 + m] = __ldg(op.right.scale + (int64_t)first * kTileRows + m * 8 + g);
        }
    }
    // every group runs the same number of rounds so that the MMA token keeps circulating; a group without a tile in the
    // last round just passes it on
    const int cta_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int rounds = (cta_tiles + kGroups - 1) / kGroups;
    mma_turn_init(grp);
    for (int it = 0; it < rounds; ++it) {
        const int tile = first + it * stride;
        if (tile >= ntiles) {
            mma_turn_begin(grp);
            mma_turn_end(grp);
            continue;
        }
        const int slot = it % kDepth;
        const int64_t row0 = (int64_t)tile * kTileRows;
        // small global reads first so that their latency hides behind the wait for the stage
        int code[2] = {next_code[0], next_code[1]};
        if ((kTipL || kTipR) && tile + stride < ntiles) {  // the codes of the following tile travel while this one is computed
            const uint8_t* codes = kTipL ? op.left.codes : op.right.codes;
            next_code[0] = __ldg(codes + row0 + (int64_t)stride * kTileRows + g);
            next_code[1] = __ldg(codes + row0 + (int64_t)stride * kTileRows + 8 + g);
        }
        // the children's scaling counts travel one iteration ahead (like the tip codes): their DRAM latency never stalls
        // the category-0 warp, which would otherwise hold up the other three at the group barrier
        const int32_t scl[2] = {next_sc[0], next_sc[1]}, scr[2] = {next_sc[2], next_sc[3]};
        if (c == 0 && t == 0 && tile + stride < ntiles) {
            const int64_t nrow = row0 + (int64_t)stride * kTileRows;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (!kTipL) next_sc[m] = __ldg(op.left.scale + nrow + m * 8 + g);
                if (!kTipR) next_sc[2 + m] = __ldg(op.right.scale + nrow + m * 8 + g);
            }
        }
        mbar_wait(gfull + slot, (it / kDepth) & 1);
        const double* stage = gstage + (size_t)slot * Plan::kStageDoubles;
        double accL[2][3][2], accR[2][3][2];
        AFrag aL[2], aR[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (kTipL) lookup_rows(s_tip, code[m], c, t, accL[m]);
            else aL[m] = load_a(stage + m * kBlockDoubles, c, lane);
            if (kTipR) lookup_rows(s_tip, code[m], c, t, accR[m]);
            else aR[m] = load_a(stage + (kTipL ? 0 : kTileDoubles) + m * kBlockDoubles, c, lane);
        }
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                if (!kTipL) accL[m][nt][0] = accL[m][nt][1] = 0.0;
                if (!kTipR) accR[m][nt][0] = accR[m][nt][1] = 0.0;
            }
        // this group's turn on the FP64 tensor pipe; up to 12 independent accumulator chains lie between two dependent MMAs
        mma_turn_begin(grp);
#pragma unroll
        for (int kt = 0; kt < 5; ++kt)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    if (!kTipL) dmma(accL[m][nt][0], accL[m][nt][1], aL[m].v[kt], fragL[nt][kt]);
                    if (!kTipR) dmma(accR[m][nt][0], accR[m][nt][1], aR[m].v[kt], fragR[nt][kt]);
                }
        mma_turn_end(grp);
        // per-row magnitude over this category's 20 states, then over the four categories through shared memory
        // magnitudes are compared through the high word of |x| on the integer pipe (2^-256 has a zero low word, so
        // "|x| < 2^-256" is exactly "hi(|x|) < hi(2^-256)"): the FP64 pipe is left to the MMAs
        int* mx = s_max + (((it & 1) * kGroups + grp) * kCats) * kTileRows;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            int big = 0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                accL[m][nt][0] *= accR[m][nt][0];
                accL[m][nt][1] *= accR[m][nt][1];
                if (nt < 2 || t < 2) big = max(big, max(__double2hiint(accL[m][nt][0]) & 0x7fffffff, __double2hiint(accL[m][nt][1]) & 0x7fffffff));
            }
            big = max(big, __shfl_xor_sync(0xffffffffu, big, 1));
            big = max(big, __shfl_xor_sync(0xffffffffu, big, 2));
            if (t == 0) mx[c * kTileRows + m * 8 + g] = big;
        }
        named_barrier(1 + grp, 4 * 32);
        // every warp of the group has its fragments in registers: the slot can take the tile kDepth iterations ahead
        if (c == 0 && tile + kDepth * stride < ntiles) refill(tile + kDepth * stride, slot);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int r = m * 8 + g;
            const int big = max(max(mx[r], mx[kTileRows + r]), max(mx[2 * kTileRows + r], mx[3 * kTileRows + r]));
            const bool rescale = big < kMinLikHi;
            if (rescale) {
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    accL[m][nt][0] *= kTwo256;
                    accL[m][nt][1] *= kTwo256;
                }
            }
            store_d(op.out + (size_t)tile * kTileDoubles + m * kBlockDoubles, c, lane, accL[m]);
            if (c == 0 && t == 0) op.out_scale[row0 + r] = scl[m] + scr[m] + (rescale ? 1 : 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------- tip-tip -----
// Both children are tips: the row is the product of two lookup rows, so the whole update is a gather + 640 B write.
// Whether a (code, code) pair needs the x2^256 rescale is decided once per CTA for all 529 pairs; the streaming loop
// then has no cross-thread traffic and its stores are fully coalesced (consecutive threads, consecutive 16 B).
constexpr int kTipTipThreads = 256;
constexpr int kTipTipRows = 128;  // one CTA per 128-row tile; many small CTAs per SM hide the code-load latency
__global__ void __launch_bounds__(kTipTipThreads) k_newview_tiptip(NewviewOp op) {
    __shared__ __align__(16) double s_l[kCodes * kTipPad];  // rows padded: the 8 rows a warp gathers from hit different banks
    __shared__ __align__(16) double s_r[kCodes * kTipPad];
    __shared__ double s_maxl[kCodes], s_maxr[kCodes];
    __shared__ int s_argl[kCodes];
    __shared__ uint8_t s_flag[kCodes * kCodes];
    __shared__ uint8_t s_pair[kTipTipRows][2];
    const int64_t row0 = (int64_t)blockIdx.x * kTipTipRows;
    // the residue codes of the tile are requested first; the table set-up below hides their latency
    uint8_t my_l = 0, my_r = 0;
    if (threadIdx.x < kTipTipRows) {
        my_l = __ldg(op.left.codes + row0 + threadIdx.x);
        my_r = __ldg(op.right.codes + row0 + threadIdx.x);
    }
    for (int i = threadIdx.x; i < kCodes * kRow; i += kTipTipThreads) {
        s_l[(i / kRow) * kTipPad + i % kRow] = (&op.pleft->tip[0][0])[i];
        s_r[(i / kRow) * kTipPad + i % kRow] = (&op.pright->tip[0][0])[i];
    }
    __syncthreads();
    // row maxima of both lookups bound every product: max_l * max_r from above, l[arg] * r[arg] from below
    for (int rowid = threadIdx.x >> 5; rowid < 2 * kCodes; rowid += kTipTipThreads / 32) {  // one warp per lookup row
        const bool right = rowid >= kCodes;
        const int code = rowid - (right ? kCodes : 0), lane = threadIdx.x & 31;
        const double* row = (right ? s_r : s_l) + code * kTipPad;
        double big = -1.0;
        int arg = 0;
        for (int k = lane; k < kRow; k += 32)
            if (fabs(row[k]) > big) {
                big = fabs(row[k]);
                arg = k;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, big, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > big || (ob == big && oa < arg)) {
                big = ob;
                arg = oa;
            }
        }
        if (lane == 0) {
            if (right) s_maxr[code] = big;
            else {
                s_maxl[code] = big;
                s_argl[code] = arg;
            }
        }
    }
    __syncthreads();
    for (int pair = threadIdx.x; pair < kCodes * kCodes; pair += kTipTipThreads) {
        const int cl = pair / kCodes, cr = pair % kCodes;
        const double* a = s_l + cl * kTipPad;
        const double* b = s_r + cr * kTipPad;
        uint8_t flag;
        if (s_maxl[cl] * s_maxr[cr] < kMinLik) flag = 1;                               // every product is below 2^-256
        else if (fabs(a[s_argl[cl]] * b[s_argl[cl]]) >= kMinLik) flag = 0;            // one product is certainly not
        else {                                                                         // undecided: test all 80
            double big = 0.0;
            for (int i = 0; i < kRow; ++i) big = fmax(big, fabs(a[i] * b[i]));
            flag = big < kMinLik ? 1 : 0;
        }
        s_flag[pair] = flag;
    }
    if (threadIdx.x < kTipTipRows) {
        s_pair[threadIdx.x][0] = my_l;
        s_pair[threadIdx.x][1] = my_r;
    }
    __syncthreads();
    double2* out = reinterpret_cast<double2*>(op.out + row0 * kRow);
#pragma unroll 4
    for (int q = threadIdx.x; q < kTipTipRows * (kRow / 2); q += kTipTipThreads) {
        // q walks the blocked layout in 16-byte steps: block, category, chunk (see mma_common.cuh)
        const int blk = q / (kBlockDoubles / 2), rem = q % (kBlockDoubles / 2);
        const int cat = rem / (kCatDoubles / 2), u = rem % (kCatDoubles / 2);
        int g, st;
        if (u < 64) {
            g = (u & 31) >> 2;
            st = (u >> 5) * 8 + (u & 3) * 2;
        } else {
            g = (u - 64) >> 1;
            st = 16 + ((u - 64) & 1) * 2;
        }
        const int r = blk * kBlockRows + g, k = (cat * kStates + st) >> 1;
        const int cl = s_pair[r][0], cr = s_pair[r][1];
        const double2 a = reinterpret_cast<const double2*>(s_l + cl * kTipPad)[k];
        const double2 b = reinterpret_cast<const double2*>(s_r + cr * kTipPad)[k];
        double2 v = make_double2(a.x * b.x, a.y * b.y);
        if (s_flag[cl * kCodes + cr]) {
            v.x *= kTwo256;
            v.y *= kTwo256;
        }
        out[q] = v;
    }
    if (threadIdx.x < kTipTipRows) op.out_scale[row0 + threadIdx.x] = s_flag[my_l * kCodes + my_r];
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
    cudaFuncSetAttribute(k_newview_mma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<true, false>::kBytes);
    cudaFuncSetAttribute(k_newview_mma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<false, true>::kBytes);
    cudaFuncSetAttribute(k_newview_mma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<false, false>::kBytes);
}

// np must be a multiple of 128 (the engine pads pattern rows to 128)
void launch_newview_mma(const NewviewOp& op, int64_t np, int sms, cudaStream_t stream) {
    const bool tl = op.left.clv == nullptr, tr = op.right.clv == nullptr;
    if (tl && tr) {
        k_newview_tiptip<<<(int)(np / kTipTipRows), kTipTipThreads, 0, stream>>>(op);
    } else if (tl) launch_one<true, false>(op, np, sms, stream);
    else if (tr) launch_one<false, true>(op, np, sms, stream);
    else launch_one<false, false>(op, np, sms, stream);
}

}  // namespace pml
