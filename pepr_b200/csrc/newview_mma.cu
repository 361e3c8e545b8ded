// newview on the FP64 tensor path (sm_100a): CLV tiles stream HBM -> shared memory through the TMA engine
// (cp.async.bulk + mbarrier ring), the 20x20 contractions run as DMMA m8n8k4 with P held in registers as B fragments,
// results go back with 128-bit stores.
//
// Why DMMA: tools/fp64_peak.cu measured 37.0 TFLOP/s for DMMA.8x8x4 against 33.5 for DFMA on B200 and showed that both
// share one pipe (profiles/r01_fp64_pipe_peaks.log).  An inner-inner site-update needs 6,480 flop per 1,920 B, i.e.
// 22 TFLOP/s at the measured 6.5 TB/s -- only the tensor path leaves issue slots and register bandwidth for the rest.
//
// Persistent CTAs of 12 warps, warp-specialised (MMA / epilogue / producer, see below);
//   per 8 rows and child: D[8 x 24] = X[8 x 20] * P_c^T[20 x 24(20 used)] as 3 n-tiles x 5 k-tiles of m8n8k4
// CLVs live in HBM in the blocked layout described in mma_common.cuh, so a stage is filled by one bulk copy per child.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "kernels.h"
#include "mma_common.cuh"
#include "pmatrix.cuh"

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

// ---- warp-specialised pipeline ---------------------------------------------------------------------------------
//   warps 0-7   MMA warps: two groups of four (one warp per rate category).  A group takes every second 16-pattern tile
//               of the CTA: wait for its stage, pull the A fragments, take the MMA turn (the two groups alternate on the
//               FP64 tensor pipe, see mma_common.cuh), multiply the two children, leave the products (blocked layout)
//               and the per-category row magnitudes in a shared-memory slot.  They never touch global memory.
//   warps 8-9   epilogue warps (alternate tiles): combine the four category magnitudes of every row, apply the rare
//               x2^256 rescale in place, hand the 10 KB tile to the TMA engine (cp.async.bulk shared -> global) and
//               write the scaling counts.
//   warp 10     producer: refills the stages with one bulk copy per inner child as soon as a stage has been consumed.
// Everything an MMA warp does besides its 60 DMMAs is ~70 instructions, so the pipe stays busy; stores cost no LSU work.
//
// A child is an inner node (its CLV tile streams in), a tip (23 x 80 look-up by residue code) or a cherry (product of two
// tip look-ups formed in registers, then multiplied by its branch matrix like an inner child -- see Side in kernels.h).
// The host orders the children tip < cherry < inner, which leaves five instances.
constexpr int kThreadsNewview = 384;   // warp 11 only helps with the prologue: 12 warps keep the register budget at 168
constexpr int kProdSlots = 4;
constexpr int kStagers = kThreadsNewview - 32;  // every warp but the producer takes part in the prologue
constexpr int kStageBarrier = 2;
constexpr int kTableDoubles = kCodes * kTipPad;

template <int KL, int KR>
struct SmemPlan {
    static constexpr int kInner = (KL == kSideInner ? 1 : 0) + (KR == kSideInner ? 1 : 0);
    static constexpr int kTablesL = KL == kSideTip ? 1 : (KL == kSideCherry ? 2 : 0);
    static constexpr int kTablesR = KR == kSideTip ? 1 : (KR == kSideCherry ? 2 : 0);
    static constexpr int kTables = kTablesL + kTablesR;
    // branches whose P matrices the prologue builds: the two of the update, then the tips of a cherry child
    static constexpr int kBranches = 2 + (KL == kSideCherry ? 2 : 0) + (KR == kSideCherry ? 2 : 0);
    // a stage = the CLV tiles of the inner children + 256 B of per-row side data that travels with them:
    // [0,64) / [64,128) scaling counts of the inner children, [128,160) residue codes on the left (tip: 16, cherry: 2 x 16),
    // [160,192) residue codes on the right
    static constexpr int kAuxDoubles = 32;
    static constexpr int kStageDoubles = kInner * kTileDoubles + kAuxDoubles;
    static constexpr int kMaxInts = kProdSlots * kCats * kTileRows;  // [slot][cat][row]
    static constexpr size_t kBarBytes = 256;
    static_assert(kProdSlots * kTileDoubles >= 8 * pmat::kFragSlotDoubles, "the product slots double as the fragment exchange area");
    static constexpr size_t kBytes = kBarBytes + sizeof(double) * (size_t)(2 * pmat::kMat + kTables * kTableDoubles) +
                                     sizeof(int) * (kMaxInts + kProdSlots * kTileRows) +
                                     sizeof(double) * (size_t)(kProdSlots * kTileDoubles + kMmaGroups * kDepth * kStageDoubles);
};

// at least one child is not a tip (the tip-tip case has its own kernel below, used only where a cherry must be stored)
// PML_TL_PROBE_A / _B (profiling builds only): two more %globaltimer stamps of thread 0 in the prologue, written to the timeline
// slots a CLV launch does not use ([4], [5]).  Points: 1 after the CTA barrier, 2 exponentials done, 3 matrix tiles done,
// 4 table / fragments stored, 5 after the first stage barrier, 6 after the second.  Stores only: nothing waits for them.
#ifndef PML_TL_PROBE_A
#define PML_TL_PROBE_A 0
#endif
#ifndef PML_TL_PROBE_B
#define PML_TL_PROBE_B 0
#endif
#define PML_TL_POINT(k)                                                   \
    do {                                                                  \
        if (PML_TL_PROBE_A == (k) && tl) op.timeline[4] = global_timer_ns(); \
        if (PML_TL_PROBE_B == (k) && tl) op.timeline[5] = global_timer_ns(); \
    } while (0)

template <int KL, int KR>
__global__ void __launch_bounds__(kThreadsNewview, 1) k_newview_mma(NewviewOp op, int ntiles) {
    using Plan = SmemPlan<KL, KR>;
    constexpr bool kInnerL = KL == kSideInner, kInnerR = KR == kSideInner;
    constexpr bool kTipL = KL == kSideTip, kTipR = KR == kSideTip;
    constexpr bool kChL = KL == kSideCherry, kChR = KR == kSideCherry;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* in_full = reinterpret_cast<uint64_t*>(smem_raw);    // [group][kDepth]
    uint64_t* in_empty = in_full + kMmaGroups * kDepth;            // [group][kDepth]
    uint64_t* prod_full = in_empty + kMmaGroups * kDepth;          // [kProdSlots]
    uint64_t* prod_empty = prod_full + kProdSlots;                 // [kProdSlots]
    double* s_model = reinterpret_cast<double*>(smem_raw + Plan::kBarBytes);   // V, Vinv
    double* s_tabL = s_model + 2 * pmat::kMat;                     // look-ups of the left child (tip: 1, cherry: 2)
    double* s_tabR = s_tabL + Plan::kTablesL * kTableDoubles;
    int* s_max = reinterpret_cast<int*>(s_tabR + Plan::kTablesR * kTableDoubles);
    int* s_sc = s_max + Plan::kMaxInts;                            // [kProdSlots][16] summed scaling counts of the children
    double* s_prod = reinterpret_cast<double*>(s_sc + kProdSlots * kTileRows);
    double* s_stage = s_prod + kProdSlots * kTileDoubles;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t_entry = 0;
    // model constants are requested before the producer queues the first tiles (see branch_mma.cu)
    const int stid = warp < kProducerWarp ? threadIdx.x : threadIdx.x - 32;  // rank among the staging threads
    pmat::ModelRegs regs{};
    const bool tl = op.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (tl) op.timeline[0] = global_timer_ns();
    pdl_launch_dependents();
    if (warp != kProducerWarp) regs = pmat::model_prefetch<kStagers>(op.dm, stid);
    // P matrices: MMA warp w builds category w & 3 of the branches at positions (w >> 2), (w >> 2) + 2, ... of the list
    // {left, right, tips of a left cherry, tips of a right cherry}: lane k < 20 holds lambda_k * r_c
    const int c_p = warp & 3;
    double lr = 0.0;
    if (warp < kMmaWarps && lane < kStates) lr = op.dm->lambda[lane] * op.dm->rates[c_p];
    pmat::ExtraB xb{};
    if (warp < kMmaWarps && Plan::kTablesL + Plan::kTablesR > 0) xb = pmat::extra_b(op.dm, lane);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kMmaGroups * kDepth; ++i) {
            mbar_init(in_full + i, 1);
            mbar_init(in_empty + i, 4);
        }
        for (int i = 0; i < kProdSlots; ++i) {
            mbar_init(prod_full + i, 4);
            mbar_init(prod_empty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // where this warp's branch lengths live and what they are multiplied by: kernel parameters, brought into registers AHEAD of
    // the dependency wait (left to the compiler their constant-bank reads sit behind it, in front of the loads they address)
    constexpr int kRounds = Plan::kBranches / 2;
    const double* len_src[kRounds];
    double len_mul[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int pos = 2 * r + (warp >> 2), id = (pos < 2 || kChL) ? pos : pos + 2;  // 0, 1: the update's branches; 2, 3 / 4, 5: cherry tips
        len_src[r] = id == 0 ? op.len_left : id == 1 ? op.len_right : id == 2 ? op.left.len1 : id == 3 ? op.left.len2
                     : id == 4 ? op.right.len1 : op.right.len2;
        len_mul[r] = id < 2 ? op.len_scale : 1.0;
        asm volatile("" : "+l"(len_src[r]), "+d"(len_mul[r]));
    }
    pdl_wait();  // from here on the kernel touches what its predecessors wrote: branch lengths, CLVs, scaling counts
    if (tl) op.timeline[1] = global_timer_ns();
    if (op.trace) t_entry = clock64();  // the cycle trace starts once the predecessor has drained
    // the lengths are requested before the model constants (in flight since the kernel started) are consumed, and nothing ahead
    // of the CTA barrier waits for them: one latency, not two
    double my_len[kRounds];
    if (warp < kMmaWarps) {
#pragma unroll
        for (int r = 0; r < kRounds; ++r) my_len[r] = *len_src[r];
    }
    if (warp != kProducerWarp) pmat::matrices_to_smem<kStagers>(regs, stid, s_model);  // V and Vinv
    __syncthreads();
    PML_TL_POINT(1);
    if (warp < kMmaWarps) {
#pragma unroll
        for (int r = 0; r < kRounds; ++r) my_len[r] *= len_mul[r];
    }

    // tiles of this CTA: n = 0 .. cta_tiles-1  <->  global tile blockIdx.x + n * gridDim.x ; MMA group n % 2, product slot n % 4
    const int cta_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == kProducerWarp) {
        // ---------------------------------------------------------------------------------------------- producer
        // lane 0 waits for the stage and announces the bytes, then lanes 0-7 hand one bulk copy each to the TMA engine
        constexpr uint32_t bytes = kTileDoubles * sizeof(double), sc_bytes = kTileRows * sizeof(int32_t);
        constexpr uint32_t tx = Plan::kInner * (bytes + sc_bytes) + (kTipL ? kTileRows : 0) + (kChL ? 2 * kTileRows : 0) +
                                (kTipR ? kTileRows : 0) + (kChR ? 2 * kTileRows : 0);
        for (int n = 0; n < cta_tiles; ++n) {
            const int grp = n % kMmaGroups, j = n / kMmaGroups, slot = j % kDepth;
            uint64_t* full = in_full + grp * kDepth + slot;
            if (lane == 0) {
                mbar_wait(in_empty + grp * kDepth + slot, ((j / kDepth) & 1) ^ 1);
                mbar_expect_tx(full, tx);
            }
            __syncwarp();
            const size_t tile = (size_t)blockIdx.x + (size_t)n * gridDim.x;
            const size_t goff = tile * kTileDoubles;
            double* dst = s_stage + (size_t)(grp * kDepth + slot) * Plan::kStageDoubles;
            unsigned char* aux = reinterpret_cast<unsigned char*>(dst + Plan::kInner * kTileDoubles);
            if (lane == 0) {
                if (kInnerL) bulk_g2s(dst, op.left.clv + goff, bytes, full);
                else bulk_g2s(aux + 128, op.left.codes + tile * kTileRows, kTileRows, full);
            } else if (lane == 1) {
                if (kInnerR) bulk_g2s(dst + (kInnerL ? kTileDoubles : 0), op.right.clv + goff, bytes, full);
                else bulk_g2s(aux + 160, op.right.codes + tile * kTileRows, kTileRows, full);
            } else if (lane == 2) {
                if (kInnerL) bulk_g2s(aux, op.left.scale + tile * kTileRows, sc_bytes, full);
                else if (kChL) bulk_g2s(aux + 144, op.left.codes2 + tile * kTileRows, kTileRows, full);
            } else if (lane == 3) {
                if (kInnerR) bulk_g2s(aux + sc_bytes, op.right.scale + tile * kTileRows, sc_bytes, full);
                else if (kChR) bulk_g2s(aux + 176, op.right.codes2 + tile * kTileRows, kTileRows, full);
            }
        }
        return;
    }

    // every other warp: the P matrices.  An inner or cherry child's matrix leaves the accumulators as B fragments (shared with
    // the other group's warp of the same category through the product slots, free until the first tile is done); a tip's
    // matrix -- a tip child's or a cherry's -- goes straight into its look-up table.
    const bool trp = op.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (trp) op.trace[90] += clock64() - t_entry;
    double fragL[3][5], fragR[3][5];
    double* s_x = s_prod;  // fragment exchange: [branch 0 / 1][category][15][32]
    if (warp < kMmaWarps) {
        // the exponentials of all rounds first: their dependency chains run side by side instead of one per round
        double e_round[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) e_round[r] = pmat::exp_neg(lr * my_len[r]);
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int pos = 2 * r + (warp >> 2), id = (pos < 2 || kChL) ? pos : pos + 2;
            double acc[3][3][2];
            const double e_lane = e_round[r];
            if (r == 0 && e_lane != 12345.678) PML_TL_POINT(2);  // (the comparisons tie a stamp to the value it follows; no code without probes)
            pmat::build_p_tiles(s_model, e_lane, lane, xb, acc);
            if (r == 0 && acc[2][2][1] != 12345.678) PML_TL_POINT(3);
            double* table = nullptr;
            if (id == 0 && kTipL) table = s_tabL;
            else if (id == 1 && kTipR) table = s_tabR;
            else if (id == 2 || id == 3) table = s_tabL + (id - 2) * kTableDoubles;
            else if (id >= 4) table = s_tabR + (id - 4) * kTableDoubles;
            if (table) pmat::tiles_to_lookup(acc, lane, c_p, table, kTipPad);
            else {
                double frag[3][5];
                pmat::tiles_to_fragments(acc, lane, frag);
                pmat::fragments_to_smem(frag, lane, s_x + (id * kCats + c_p) * pmat::kFragSlotDoubles);
            }
            if (PML_TL_PROBE_A == 4 || PML_TL_PROBE_B == 4) {
                __syncwarp();
                if (r == 0) PML_TL_POINT(4);
            }
        }
    }
    named_barrier(kStageBarrier, kStagers);
    if (trp) op.trace[92] += clock64() - t_entry;
    PML_TL_POINT(5);
    if (warp < kMmaWarps) {
        if (!kTipL) pmat::fragments_from_smem(fragL, lane, s_x + c_p * pmat::kFragSlotDoubles);
        if (!kTipR) pmat::fragments_from_smem(fragR, lane, s_x + (kCats + c_p) * pmat::kFragSlotDoubles);
    }
    named_barrier(kStageBarrier, kStagers);  // the product slots are free for their real purpose
    PML_TL_POINT(6);
    if (warp > kProducerWarp) return;

    if (warp >= kMmaWarps) {
        // ---------------------------------------------------------------------------------------------- epilogue
        const int e = warp - kMmaWarps;
        const int r = lane & 15;   // lanes 0-15 own one row each
        int prev_slot = -1;
        if (op.trace && blockIdx.x == 0 && lane == 0) op.trace[warp * 8 + 7] += clock64() - t_entry;  // entry -> loop
        for (int n = e; n < cta_tiles; n += kEpiWarps) {
            const int slot = n % kProdSlots;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)n * gridDim.x;
            const bool tr = op.trace != nullptr && blockIdx.x == 0 && lane == 0;
            const long long e0 = tr ? clock64() : 0;
            mbar_wait(prod_full + slot, (n / kProdSlots) & 1);
            const long long e1 = tr ? clock64() : 0;
            const int* mx = s_max + slot * kCats * kTileRows;
            const int big = max(max(mx[r], mx[kTileRows + r]), max(mx[2 * kTileRows + r], mx[3 * kTileRows + r]));
            const bool rescale = big < kMinLikHi;
            const unsigned flagged = __ballot_sync(0xffffffffu, rescale && lane < kTileRows);
            double* prod = s_prod + (size_t)slot * kTileDoubles;
            if (flagged) {  // rare: multiply the flagged rows in place before the tile leaves
                for (unsigned rest = flagged; rest; rest &= rest - 1) {
                    const int row = __ffs(rest) - 1;
                    double* blk = prod + (row >> 3) * kBlockDoubles;
                    const int gr = row & 7;
                    for (int k = lane; k < kRow; k += 32) {
                        const int cat = k / kStates, st = k % kStates;
                        const int off = cat * kCatDoubles + (st < 8 ? gr * 8 + st : (st < 16 ? 64 + gr * 8 + (st - 8) : 128 + gr * 4 + (st - 16)));
                        blk[off] *= kTwo256;
                    }
                }
                fence_async_smem();
                __syncwarp();
            }
            if (lane == 0) {
                bulk_s2g(op.out + (size_t)tile * kTileDoubles, prod, kTileDoubles * sizeof(double));
                bulk_commit();
            }
            if (lane < kTileRows) op.out_scale[tile * kTileRows + r] = s_sc[slot * kTileRows + r] + (rescale ? 1 : 0);
            // the previous tile of this warp has certainly been read out of shared memory once at most one store is pending
            const long long e2 = tr ? clock64() : 0;
            if (lane == 0) {
                if (prev_slot >= 0) {
                    bulk_wait_read<1>();
                    mbar_arrive(prod_empty + prev_slot);
                }
            }
            if (tr) {
                const long long e3 = clock64();
                op.trace[warp * 8 + 1] += e1 - e0;   // wait for products
                op.trace[warp * 8 + 2] += e2 - e1;   // magnitude test, store issue, scaling counts
                op.trace[warp * 8 + 3] += e3 - e2;   // wait until the previous store has left shared memory
                op.trace[warp * 8 + 4] += 1;
            }
            prev_slot = slot;
        }
        const long long e4 = op.trace ? clock64() : 0;
        if (lane == 0) bulk_wait_read<0>();
        if (op.trace && blockIdx.x == 0 && lane == 0) {
            op.trace[warp * 8 + 5] += clock64() - e4;  // final drain
            op.trace[warp * 8 + 0] += clock64() - t_entry;  // whole kernel as seen by this epilogue warp
            op.trace[warp * 8 + 6] += 1;
        }
        return;
    }

    // -------------------------------------------------------------------------------------------------- MMA warps
    const int c = warp & 3, grp = warp >> 2, g = lane >> 2, t = lane & 3;
    const int rounds = (cta_tiles + kMmaGroups - 1) / kMmaGroups;
    if (op.trace && blockIdx.x == 0 && lane == 0) op.trace[warp * 8 + 7] += clock64() - t_entry;  // prologue
    mma_turn_init(grp);
    if (tl) op.timeline[2] = global_timer_ns();
    for (int j = 0; j < rounds; ++j) {
        const int n = j * kMmaGroups + grp;
        if (n >= cta_tiles) {  // no tile left for this group: keep the MMA token moving
            mma_turn_begin(grp);
            mma_turn_end(grp);
            continue;
        }
        const int slot = j % kDepth;
        const bool tr = op.trace != nullptr && blockIdx.x == 0 && lane == 0;
        long long tk0 = tr ? clock64() : 0;
        mbar_wait(in_full + grp * kDepth + slot, (j / kDepth) & 1);
        long long tk1 = tr ? clock64() : 0;
        const double* stage = s_stage + (size_t)(grp * kDepth + slot) * Plan::kStageDoubles;
        const unsigned char* aux = reinterpret_cast<const unsigned char*>(stage + Plan::kInner * kTileDoubles);
        int32_t sc_sum = 0;  // category-0 warp, lanes 0-15: scaling counts of the inner children of row `lane`
        if (Plan::kInner > 0 && c == 0 && lane < kTileRows) {
            const int32_t* sci = reinterpret_cast<const int32_t*>(aux);
            sc_sum = (kInnerL ? sci[lane] : 0) + (kInnerR ? sci[kTileRows + lane] : 0);
        }
        double accL[2][3][2], accR[2][3][2];
        AFrag aL[2], aR[2];
        CherryIn cL{}, cR{};
        if (kChL) cL = cherry_begin(s_tabL, s_tabL + kTableDoubles, aux + 128, aux + 144, g, c, t);
        if (kChR) cR = cherry_begin(s_tabR, s_tabR + kTableDoubles, aux + 160, aux + 176, g, c, t);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int row = m * 8 + g;
            if (kTipL) lookup_rows(s_tabL, aux[128 + row], c, t, accL[m]);
            else if (kInnerL) aL[m] = load_a(stage + m * kBlockDoubles, c, lane);
            if (kTipR) lookup_rows(s_tabR, aux[160 + row], c, t, accR[m]);
            else if (kInnerR) aR[m] = load_a(stage + (kInnerL ? kTileDoubles : 0) + m * kBlockDoubles, c, lane);
        }
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                if (!kTipL) accL[m][nt][0] = accL[m][nt][1] = 0.0;
                if (!kTipR) accR[m][nt][0] = accR[m][nt][1] = 0.0;
            }
        // this group's turn on the FP64 tensor pipe; 12 (6) independent accumulator chains lie between two dependent MMAs
        mma_turn_begin(grp);
        long long tk2 = tr ? clock64() : 0;
        const int multiplied = children_mma<KL, KR>(aL, aR, cL, cR, fragL, fragR, accL, accR, t);
        long long tk3 = tr ? clock64() : 0;
        mma_turn_end(grp);
        __syncwarp();
        if (lane == 0) mbar_arrive(in_empty + grp * kDepth + slot);  // the stage may be refilled
        // products and, through the high word of |x| on the integer pipe, this category's magnitude of every row
        // (2^-256 has a zero low word, so "|x| < 2^-256" is exactly "hi(|x|) < hi(2^-256)")
        const int pslot = n % kProdSlots;
        int big[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            big[m] = 0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                if (!(multiplied >> m & 1)) {
                    accL[m][nt][0] *= accR[m][nt][0];
                    accL[m][nt][1] *= accR[m][nt][1];
                }
                if (nt < 2 || t < 2) big[m] = max(big[m], max(__double2hiint(accL[m][nt][0]) & 0x7fffffff, __double2hiint(accL[m][nt][1]) & 0x7fffffff));
            }
            big[m] = max(big[m], __shfl_xor_sync(0xffffffffu, big[m], 1));
            big[m] = max(big[m], __shfl_xor_sync(0xffffffffu, big[m], 2));
        }
        long long tk4 = tr ? clock64() : 0;
        mbar_wait(prod_empty + pslot, ((n / kProdSlots) & 1) ^ 1);
        long long tk5 = tr ? clock64() : 0;
        double* prod = s_prod + (size_t)pslot * kTileDoubles;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            store_d(prod + m * kBlockDoubles, c, lane, accL[m]);
            if (t == 0) s_max[(pslot * kCats + c) * kTileRows + m * 8 + g] = big[m];
        }
        if (c == 0 && lane < kTileRows) s_sc[pslot * kTileRows + lane] = sc_sum;
        fence_async_smem();  // the tile leaves through the async proxy (bulk store)
        __syncwarp();
        if (lane == 0) mbar_arrive(prod_full + pslot);
        if (tr) {  // phases: wait data | fragments + wait turn | MMAs | products + magnitudes | wait slot | store + fence | tiles
            long long* row = op.trace + warp * 8;
            const long long tk6 = clock64();
            row[0] += tk1 - tk0;
            row[1] += tk2 - tk1;
            row[2] += tk3 - tk2;
            row[3] += tk4 - tk3;
            row[4] += tk5 - tk4;
            row[5] += tk6 - tk5;
            row[6] += 1;
        }
    }
    if (tl) op.timeline[3] = global_timer_ns();
}

// ---------------------------------------------------------------------------------------------- tip-tip -----
// Both children are tips: the row is the product of two lookup rows, so the whole update is a gather + 640 B write.
// Persistent CTAs (one per SM): each builds the two P matrices and lookups once, decides once for all 529 (code, code)
// pairs whether the x2^256 rescale applies, and then streams 128-row chunks with fully coalesced 16-byte stores.
constexpr int kTipTipThreads = 512;
constexpr int kTipTipRows = 128;
struct TipTipSmem {
    double model[pmat::kModelDoubles];
    double P[pmat::kPDoubles];
    double l[kCodes * kTipPad];  // rows padded: the 8 rows a warp gathers from hit different banks
    double r[kCodes * kTipPad];
    double maxl[kCodes], maxr[kCodes];
    int argl[kCodes];
    uint8_t flag[kCodes * kCodes];
    uint8_t pair[2][kTipTipRows][2];
};
__global__ void __launch_bounds__(kTipTipThreads, 1) k_newview_tiptip(NewviewOp op, int64_t np) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TipTipSmem& sm = *reinterpret_cast<TipTipSmem*>(smem_raw);
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    pmat::ModelRegs regs = pmat::model_prefetch<kTipTipThreads>(op.dm, tid);
    pdl_wait();
    pmat::length_prefetch(regs, op.len_left, op.len_right);
    regs.len[0] *= op.len_scale;
    regs.len[1] *= op.len_scale;
    // every CTA streams one contiguous range of 8-row blocks (equal shares up to one block), in chunks of 128 rows
    const int64_t nblk = np / kBlockRows, per_cta = (nblk + gridDim.x - 1) / gridDim.x;
    const int64_t row_lo = min((long long)np, (long long)blockIdx.x * per_cta * kBlockRows);
    const int64_t row_hi = min((long long)np, (long long)(row_lo + per_cta * kBlockRows));
    // the residue codes of the first chunk are requested right away; the table set-up below hides their latency
    uint8_t my_l = 0, my_r = 0;
    if (tid < kTipTipRows && row_lo + tid < row_hi) {
        my_l = __ldg(op.left.codes + row_lo + tid);
        my_r = __ldg(op.right.codes + row_lo + tid);
    }
    pmat::model_to_smem<kTipTipThreads>(regs, tid, sm.model);
    __syncthreads();
    pmat::build_p<kTipTipThreads>(sm.model, tid, sm.P);
    __syncthreads();
    pmat::build_tip_lookup<kTipTipThreads>(sm.P, tid, sm.l, kTipPad);
    pmat::build_tip_lookup<kTipTipThreads>(sm.P + kCats * pmat::kMat, tid, sm.r, kTipPad);
    __syncthreads();
    // row maxima of both lookups bound every product: max_l * max_r from above, l[arg] * r[arg] from below
    for (int rowid = tid >> 5; rowid < 2 * kCodes; rowid += kTipTipThreads / 32) {  // one warp per lookup row
        const bool right = rowid >= kCodes;
        const int code = rowid - (right ? kCodes : 0), lane = tid & 31;
        const double* row = (right ? sm.r : sm.l) + code * kTipPad;
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
            if (right) sm.maxr[code] = big;
            else {
                sm.maxl[code] = big;
                sm.argl[code] = arg;
            }
        }
    }
    __syncthreads();
    for (int pair = tid; pair < kCodes * kCodes; pair += kTipTipThreads) {
        const int cl = pair / kCodes, cr = pair % kCodes;
        const double* a = sm.l + cl * kTipPad;
        const double* b = sm.r + cr * kTipPad;
        uint8_t flag;
        if (sm.maxl[cl] * sm.maxr[cr] < kMinLik) flag = 1;                               // every product is below 2^-256
        else if (fabs(a[sm.argl[cl]] * b[sm.argl[cl]]) >= kMinLik) flag = 0;            // one product is certainly not
        else {                                                                           // undecided: test all 80
            double big = 0.0;
            for (int i = 0; i < kRow; ++i) big = fmax(big, fabs(a[i] * b[i]));
            flag = big < kMinLik ? 1 : 0;
        }
        sm.flag[pair] = flag;
    }
    int buf = 0;
    for (int64_t row0 = row_lo; row0 < row_hi; row0 += kTipTipRows, buf ^= 1) {
        const int nrows = (int)min((long long)kTipTipRows, (long long)(row_hi - row0));
        if (tid < nrows) {
            sm.pair[buf][tid][0] = my_l;
            sm.pair[buf][tid][1] = my_r;
        }
        __syncthreads();  // also covers sm.flag on the first trip; the other pair buffer is free for the next trip
        if (tid < nrows) {
            op.out_scale[row0 + tid] = sm.flag[my_l * kCodes + my_r];
            const int64_t next = row0 + kTipTipRows + tid;
            if (next < row_hi) {  // the codes of the following chunk travel while this one is written
                my_l = __ldg(op.left.codes + next);
                my_r = __ldg(op.right.codes + next);
            }
        }
        // A warp takes four rows at a time and walks them in row order, 16 bytes per lane: the 32 lanes of one step read
        // CONSECUTIVE entries of (at most two) lookup rows -- conflict-free shared-memory reads; gathering in the order of
        // the blocked layout instead made 8 lookup rows collide on the same banks and held the kernel at 16 B/clk per SM.
        // The stores land in 64-byte runs of the blocked layout (mma_common.cuh).
        double2* out = reinterpret_cast<double2*>(op.out + row0 * kRow);
        const int lane = tid & 31, wid = tid >> 5;
        for (int grp = wid; grp < nrows / 4; grp += kTipTipThreads / 32) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int idx = lane + 32 * j;               // 0 .. 159 = 4 rows x 40 double2
                const int r = grp * 4 + idx / (kRow / 2), k = idx % (kRow / 2);
                const int cl = sm.pair[buf][r][0], cr = sm.pair[buf][r][1];
                const double2 a = reinterpret_cast<const double2*>(sm.l + cl * kTipPad)[k];
                const double2 b = reinterpret_cast<const double2*>(sm.r + cr * kTipPad)[k];
                double2 v = make_double2(a.x * b.x, a.y * b.y);
                if (sm.flag[cl * kCodes + cr]) {
                    v.x *= kTwo256;
                    v.y *= kTwo256;
                }
                // (row, category, state pair) -> 16-byte slot of the blocked layout
                const int cat = k / (kStates / 2), sp = k % (kStates / 2), g = r & 7;
                const int slot = sp < 4 ? g * 4 + sp : (sp < 8 ? 32 + g * 4 + (sp - 4) : 64 + g * 2 + (sp - 8));
                out[(r >> 3) * (kBlockDoubles / 2) + cat * (kCatDoubles / 2) + slot] = v;
            }
        }
    }
}

template <int KL, int KR>
void launch_one(const NewviewOp& op, int64_t np, int sms, cudaStream_t stream) {
    using Plan = SmemPlan<KL, KR>;
    const int ntiles = (int)(np / kTileRows);
    const int grid = ntiles < sms ? ntiles : sms;
    launch_pdl(k_newview_mma<KL, KR>, grid, kThreadsNewview, Plan::kBytes, stream, op, ntiles);
}

template <int KL, int KR>
void configure_one() {
    cudaFuncSetAttribute(k_newview_mma<KL, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPlan<KL, KR>::kBytes);
}

}  // namespace

void configure_mma_kernels() {
    configure_one<kSideTip, kSideCherry>();
    configure_one<kSideTip, kSideInner>();
    configure_one<kSideCherry, kSideCherry>();
    configure_one<kSideCherry, kSideInner>();
    configure_one<kSideInner, kSideInner>();
    cudaFuncSetAttribute(k_newview_tiptip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TipTipSmem));
}

NewviewOp canonical_children(const NewviewOp& in) {
    auto rank = [](const Side& s) { const int k = side_kind(s); return k == kSideTip ? 0 : (k == kSideCherry ? 1 : 2); };
    NewviewOp op = in;
    if (rank(op.left) > rank(op.right)) {  // the product of the two children commutes exactly
        std::swap(op.left, op.right);
        std::swap(op.len_left, op.len_right);
    }
    return op;
}

// np must be a multiple of 128 (the engine pads pattern rows to 128)
void launch_newview_mma(const NewviewOp& in, int64_t np, int sms, cudaStream_t stream) {
    const NewviewOp op = canonical_children(in);
    const int kl = side_kind(op.left), kr = side_kind(op.right);
    if (kl == kSideTip && kr == kSideTip) {
        const int64_t nblk = np / kBlockRows;
        launch_pdl(k_newview_tiptip, (int)(nblk < sms ? nblk : sms), kTipTipThreads, sizeof(TipTipSmem), stream, op, np);
    } else if (kl == kSideTip) kr == kSideCherry ? launch_one<kSideTip, kSideCherry>(op, np, sms, stream) : launch_one<kSideTip, kSideInner>(op, np, sms, stream);
    else if (kl == kSideCherry) kr == kSideCherry ? launch_one<kSideCherry, kSideCherry>(op, np, sms, stream) : launch_one<kSideCherry, kSideInner>(op, np, sms, stream);
    else launch_one<kSideInner, kSideInner>(op, np, sms, stream);
}

}  // namespace pml
