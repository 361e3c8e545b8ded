// Branch kernel on the FP64 tensor path: for the two ends of a branch it forms, per pattern and rate category, the
// eigen-space product s[c][k] = (sum_i pi_i xa[c][i] V[i][k]) * (sum_i Vinv[k][i] xb[c][i])  (raxmlHPC sumGAMMAPROT) and
// contracts it at once with exp(lambda_k r_c t) * {1, lambda r, (lambda r)^2} (coreGTRGAMMAPROT), so that one pass over
// the two CLVs yields lnL, dlnL/dt and d2lnL/dt2 -- and, with t = the branch's own length, the root evaluate
// (evaluateGTRGAMMAPROT) including per-pattern lnL.  The product table is only written out when the caller wants to
// iterate Newton-Raphson on it (kStore).
//
// Persistent CTAs of 11 warps, warp-specialised like newview_mma.cu:
//   warps 0-7   MMA warps, two groups of four (one warp per rate category) alternating on the FP64 tensor pipe; a group
//               takes every second 16-pattern tile, contracts its products with the exponentials and leaves f, f', f''
//               per row and category in a shared-memory slot.  They never touch global memory (except kStore).
//   warps 8-9   finishing warps: 32 rows (the tiles of both groups) at a time they add the four categories, take the log,
//               apply scaling counts and integer weights and keep running weighted sums; the logs and divisions stay off
//               the MMA warps.  Nothing per-pattern goes back to HBM unless per-pattern lnL was asked for.
//   warp 10     producer: TMA bulk copies of the CLV tiles and of the per-row side data (scaling counts, weights, codes).
// The CTA that draws the last ticket adds the CTA partials in a fixed order (bit-reproducible) and, when asked to, does
// the guarded Newton-Raphson step of raxmlHPC's topLevelMakenewz on the device: the branch length never visits the host.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "mma_common.cuh"
#include "pmatrix.cuh"

namespace pml {

namespace {

using namespace mma;

constexpr double kLogMinLik = -177.445678223345993274;  // ln 2^-256
constexpr int kTipVecPad = 22;                          // doubles per row of the 23 x 20 tip table
constexpr int kThreadsBranch = (kProducerWarp + 1) * 32;  // 352: eleven warps leave 184 registers per thread
constexpr int kBranchDepth = 4;                         // input stages per MMA group
constexpr int kRedSlots = 4;                            // row-sum slots between MMA and finishing warps (tile n -> n % 4)
constexpr int kFinishBarrier = 3;                       // named barrier of the two finishing warps
constexpr int kStageBarrier = 2;                        // named barrier of all warps but the producer (prologue)

// KA: what the a end is -- an inner node, a tip, or a folded cherry (Side in kernels.h; the b end is always an inner node)
template <int KA>
struct BranchPlan {
    static constexpr bool kTipA = KA == kSideTip, kChA = KA == kSideCherry;
    static constexpr int kInner = KA == kSideInner ? 2 : 1;
    // a stage = the CLV tiles of the inner ends + 256 B of per-row side data that travels with them:
    // [0,64) scaling counts of b, [64,128) scaling counts of a, [128,192) pattern weights, [192,208) residue codes of a
    // (a cherry's first tip), [208,224) residue codes of a cherry's second tip
    static constexpr int kAuxDoubles = 32;
    static constexpr int kStageDoubles = kInner * kTileDoubles + kAuxDoubles;
    // tip: 23 x 20 table of pi V sums; cherry: the two 23 x 80 look-ups of its tips
    static constexpr int kTipDoubles = kTipA ? kCodes * kTipVecPad : (kChA ? 2 * kCodes * kTipPad : 0);
    static constexpr int kRedDoubles = kRedSlots * kCats * kTileRows * 3;  // [slot][cat][row][f, f', f'']
    static constexpr int kSideInts = kRedSlots * kTileRows * 2;            // [slot][row][scaling count, weight]
    static constexpr int kExpDoubles = 3 * kCats * 24;                     // exp(lambda r t) * {1, lambda r, (lambda r)^2}, padded to 24 states
    static constexpr int kMatDoubles = 3 * kStates * kStates;              // V (cherry end only), Vinv and pi V staged once per CTA
    static constexpr size_t kBarBytes = 256;
    static constexpr size_t kBytes = kBarBytes + sizeof(double) * (size_t)(kTipDoubles + kRedDoubles + 8 + kExpDoubles + kMatDoubles) +
                                     sizeof(int) * kSideInts + sizeof(double) * (size_t)(kMmaGroups * kBranchDepth * kStageDoubles);
};

template <int KA, bool kStore>
__global__ void __launch_bounds__(kThreadsBranch, 1) k_branch_mma(BranchArgs args, int ntiles) {
    using Plan = BranchPlan<KA>;
    constexpr bool kTipA = Plan::kTipA, kChA = Plan::kChA, kInnerA = KA == kSideInner;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* in_full = reinterpret_cast<uint64_t*>(smem_raw);     // [group][kBranchDepth]
    uint64_t* in_empty = in_full + kMmaGroups * kBranchDepth;       // [group][kBranchDepth]
    uint64_t* red_full = in_empty + kMmaGroups * kBranchDepth;      // [kRedSlots]
    uint64_t* red_empty = red_full + kRedSlots;                     // [kRedSlots]
    double* s_tip = reinterpret_cast<double*>(smem_raw + Plan::kBarBytes);
    double* s_red = s_tip + Plan::kTipDoubles;
    double* s_fin = s_red + Plan::kRedDoubles;                      // [2 finishing warps][3] (+2 spare)
    double* s_exp = s_fin + 8;                                      // [3][kCats][24]
    double* s_v = s_exp + Plan::kExpDoubles;                        // [i][k]  (V, Vinv: the layout pmatrix.cuh builds P from)
    double* s_vinv = s_v + kStates * kStates;                       // [k][i]
    double* s_piv = s_vinv + kStates * kStates;                     // [i][k]
    int2* s_side = reinterpret_cast<int2*>(s_v + Plan::kMatDoubles);
    double* s_stage = reinterpret_cast<double*>(s_side + kRedSlots * kTileRows);
    const DeviceModel* dm = args.dm;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr = args.trace != nullptr && blockIdx.x == 0 && lane == 0;
    const long long t_entry = tr ? clock64() : 0;
    // The model constants are requested FIRST: once the producer has queued its first 160 KB of tiles every further load
    // of this SM waits behind them (measured: 10,000 cycles of prologue when the order was the other way round).
    constexpr int kStagers = kProducerWarp * 32;  // every warp but the producer
    const int tid = threadIdx.x;
    double pre_vinv[2] = {0.0, 0.0}, pre_piv[2] = {0.0, 0.0}, pre_v[2] = {0.0, 0.0}, pre_lambda = 0.0, pre_rate = 0.0, lr = 0.0;
    pmat::ExtraB xb{};
    pdl_launch_dependents();
    if (warp != kProducerWarp) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = tid + q * kStagers;
            if (idx < kStates * kStates) {
                pre_vinv[q] = (&dm->Vinv[0][0])[idx];
                pre_piv[q] = (&dm->piV[0][0])[idx];
                if (kChA) pre_v[q] = (&dm->V[0][0])[idx];
            }
        }
        // cherry end: MMA warp w builds the look-up of tip (w >> 2) for category w & 3; lane k < 20 holds lambda_k * r_c
        if (kChA && warp < kMmaWarps && lane < kStates) lr = dm->lambda[lane] * dm->rates[warp & 3];
        if (kChA && warp < kMmaWarps) xb = pmat::extra_b(dm, lane);
        if (tid < kCats * 24 && tid % 24 < kStates) {
            pre_lambda = dm->lambda[tid % 24];
            pre_rate = dm->rates[tid / 24];
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kMmaGroups * kBranchDepth; ++i) {
            mbar_init(in_full + i, 1);
            mbar_init(in_empty + i, 4);
        }
        for (int i = 0; i < kRedSlots; ++i) {
            mbar_init(red_full + i, 4);
            mbar_init(red_empty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // from here on the kernel touches what its predecessors wrote (see mma_common.cuh)
    // the length the sums are taken at: the device copy of the branch length is brought into the NR range first
    const double tt = args.t_ptr ? nr_clamp_length(*args.t_ptr) : args.t;
    double tip_len = 0.0;
    if (kChA && warp < kMmaWarps) tip_len = (warp >> 2) == 0 ? *args.a.len1 : *args.a.len2;
    __syncthreads();

    // tiles of this CTA: n = 0 .. cta_tiles-1  <->  global tile blockIdx.x + n * gridDim.x ; MMA group n % 2, row-sum slot n % 4
    const int cta_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == kProducerWarp) {
        // ---------------------------------------------------------------------------------------------- producer
        // lane 0 waits for the stage and announces the bytes, then lanes 0-4 hand one bulk copy each to the TMA engine
        constexpr uint32_t bytes = kTileDoubles * sizeof(double), int_bytes = kTileRows * sizeof(int32_t);
        for (int n = 0; n < cta_tiles; ++n) {
            const int grp = n % kMmaGroups, j = n / kMmaGroups, slot = j % kBranchDepth;
            uint64_t* full = in_full + grp * kBranchDepth + slot;
            if (lane == 0) {
                mbar_wait(in_empty + grp * kBranchDepth + slot, ((j / kBranchDepth) & 1) ^ 1);
                mbar_expect_tx(full, Plan::kInner * (bytes + int_bytes) + int_bytes + (kTipA ? kTileRows : 0) + (kChA ? 2 * kTileRows : 0));
            }
            __syncwarp();
            const size_t tile = (size_t)blockIdx.x + (size_t)n * gridDim.x;
            const size_t goff = tile * kTileDoubles;
            double* dst = s_stage + (size_t)(grp * kBranchDepth + slot) * Plan::kStageDoubles;
            unsigned char* aux = reinterpret_cast<unsigned char*>(dst + Plan::kInner * kTileDoubles);
            if (lane == 0) bulk_g2s(dst + (kInnerA ? kTileDoubles : 0), args.b.clv + goff, bytes, full);
            else if (lane == 1) {
                if (kInnerA) bulk_g2s(dst, args.a.clv + goff, bytes, full);
                else bulk_g2s(aux + 192, args.a.codes + tile * kTileRows, kTileRows, full);
            } else if (lane == 2) bulk_g2s(aux, args.b.scale + tile * kTileRows, int_bytes, full);
            else if (lane == 3) bulk_g2s(aux + 128, args.weights + tile * kTileRows, int_bytes, full);
            else if (lane == 4) {
                if (kInnerA) bulk_g2s(aux + 64, args.a.scale + tile * kTileRows, int_bytes, full);
                else if (kChA) bulk_g2s(aux + 208, args.a.codes2 + tile * kTileRows, kTileRows, full);
            }
        }
        return;
    }

    // every other warp: model constants into shared memory, one exponential per thread
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kStagers;
        if (idx < kStates * kStates) {
            s_vinv[idx] = pre_vinv[q];
            s_piv[idx] = pre_piv[q];
            if (kChA) s_v[idx] = pre_v[q];
        }
    }
    if (tid < kCats * 24) {
        const double a = pre_lambda * pre_rate;
        const double e = tid % 24 < kStates ? pmat::exp_neg(a * tt) : 0.0;
        s_exp[tid] = e;
        s_exp[kCats * 24 + tid] = a * e;
        s_exp[2 * kCats * 24 + tid] = a * a * e;
    }
    if (kTipA) {
        named_barrier(kStageBarrier, kStagers);
        // tipvec[code][k] = sum over the residues the code allows of pi_i V[i][k]
        for (int idx = tid; idx < kCodes * kStates; idx += kStagers) {
            const int code = idx / kStates, k = idx % kStates;
            double acc = 0.0;
            if (code < 20) acc = s_piv[code * kStates + k];
            else if (code == 20) acc = s_piv[2 * kStates + k] + s_piv[3 * kStates + k];
            else if (code == 21) acc = s_piv[5 * kStates + k] + s_piv[6 * kStates + k];
            else
                for (int i = 0; i < kStates; ++i) acc += s_piv[i * kStates + k];
            s_tip[code * kTipVecPad + k] = acc;
        }
    }
    if (kChA) {
        named_barrier(kStageBarrier, kStagers);  // V and Vinv are in place
        if (warp < kMmaWarps) {
            double acc[3][3][2];
            pmat::build_p_tiles(s_v, pmat::exp_neg(lr * tip_len), lane, xb, acc);
            pmat::tiles_to_lookup(acc, lane, warp & 3, s_tip + (warp >> 2) * kCodes * kTipPad, kTipPad);
        }
    }
    named_barrier(kStageBarrier, kStagers);

    if (warp >= kMmaWarps) {
        // ---------------------------------------------------------------------------------------------- finishing
        // a pair of tiles (one per MMA group) = 32 rows = one row per lane
        const int e = warp - kMmaWarps, half = lane >> 4, r = lane & 15;
        double sum_l = 0.0, sum_d1 = 0.0, sum_d2 = 0.0;
        const PublishEarly early = publish_prefetch(args.pub, tt);  // while the tiles are still streaming
        const int pairs = (cta_tiles + 1) / 2;
        if (tr) args.trace[warp * 8 + 7] += clock64() - t_entry;  // entry -> loop
        for (int q = e; q < pairs; q += kEpiWarps) {
            const int n0 = 2 * q, n1 = n0 + 1, n = n0 + half;
            const long long f0 = tr ? clock64() : 0;
            mbar_wait(red_full + n0 % kRedSlots, (n0 / kRedSlots) & 1);
            if (n1 < cta_tiles) mbar_wait(red_full + n1 % kRedSlots, (n1 / kRedSlots) & 1);
            const long long f1c = tr ? clock64() : 0;
            if (n < cta_tiles) {
                const int slot = n % kRedSlots;
                const double* red = s_red + (size_t)slot * kCats * kTileRows * 3 + r * 3;
                constexpr int cs = kTileRows * 3;
                const double f = (red[0] + red[cs]) + (red[2 * cs] + red[3 * cs]);
                const double f1 = (red[1] + red[cs + 1]) + (red[2 * cs + 1] + red[3 * cs + 1]);
                const double f2 = (red[2] + red[cs + 2]) + (red[2 * cs + 2] + red[3 * cs + 2]);
                const int2 side = s_side[slot * kTileRows + r];
                const double w = (double)side.y;
                const int64_t p = ((int64_t)blockIdx.x + (int64_t)n * gridDim.x) * kTileRows + r;
                if (kStore) args.sum_scale[p] = side.x;
                // every FP64 instruction of these two warps waits in line behind the MMA warps' DMMAs (~30 clk each):
                // only what the caller asked for is computed
                if (args.want_lnl) {
                    const double lnl = log(0.25 * f) + side.x * kLogMinLik;
                    if (args.site_lnl) args.site_lnl[p] = lnl;
                    sum_l = fma(w, lnl, sum_l);
                }
                if (args.want_derivs) {
                    const double inv = 1.0 / f, qd = f1 * inv;
                    sum_d1 = fma(w, qd, sum_d1);
                    sum_d2 = fma(w, f2 * inv - qd * qd, sum_d2);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(red_empty + n0 % kRedSlots);
                if (n1 < cta_tiles) mbar_arrive(red_empty + n1 % kRedSlots);
            }
            if (tr) {
                args.trace[warp * 8 + 1] += f1c - f0;          // wait for the row sums
                args.trace[warp * 8 + 2] += clock64() - f1c;   // logs, divisions, weights
                args.trace[warp * 8 + 4] += 1;
            }
        }
        const long long f2c = tr ? clock64() : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum_l += __shfl_xor_sync(0xffffffffu, sum_l, o);
            sum_d1 += __shfl_xor_sync(0xffffffffu, sum_d1, o);
            sum_d2 += __shfl_xor_sync(0xffffffffu, sum_d2, o);
        }
        if (lane == 0) {
            s_fin[e * 3 + 0] = sum_l;
            s_fin[e * 3 + 1] = sum_d1;
            s_fin[e * 3 + 2] = sum_d2;
        }
        named_barrier(kFinishBarrier, kEpiWarps * 32);
        if (e != 0) return;
        // CTA partials, then the CTA that draws the last ticket adds all of them in a fixed order
        if (lane < 3) {
            args.partials[(int64_t)lane * gridDim.x + blockIdx.x] = s_fin[lane] + s_fin[3 + lane];
            fence_acq_rel_gpu();
        }
        __syncwarp();
        unsigned int ticket = 0;
        if (lane == 0) ticket = atomicAdd(args.ticket, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (tr) {
            args.trace[warp * 8 + 3] += clock64() - f2c;     // CTA partial + ticket
            args.trace[warp * 8 + 0] += clock64() - t_entry;  // whole kernel as seen by CTA 0
            args.trace[warp * 8 + 6] += 1;
        }
        if (ticket != gridDim.x - 1) return;
        const long long f3c = args.trace ? clock64() : 0;
        fence_acq_rel_gpu();
        double r3[3] = {0.0, 0.0, 0.0};  // lane-strided, then a shuffle tree -- the same order every run
        for (int i = lane; i < (int)gridDim.x; i += 32) {
#pragma unroll
            for (int v = 0; v < 3; ++v) r3[v] += __ldcg(args.partials + (int64_t)v * gridDim.x + i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int v = 0; v < 3; ++v) r3[v] += __shfl_xor_sync(0xffffffffu, r3[v], o);
        }
        const bool lost = args.peer.mail != nullptr && !peer_allreduce3(args.peer, args.pub.seq, r3);
        if (lane == 0) {
            args.result[0] = r3[0];
            args.result[1] = r3[1];
            args.result[2] = r3[2];
            args.result[3] = tt;
            *args.ticket = 0;
            publish_result(args.pub, r3, tt, lost, early);
            if (args.trace) {
                args.trace[88] += clock64() - f3c;  // last CTA: final sum, NR step, publication
                args.trace[89] += 1;
            }
        }
        return;
    }

    // -------------------------------------------------------------------------------------------------- MMA warps
    const int c = warp & 3, grp = warp >> 2, g = lane >> 2, t = lane & 3;
    // B fragments: a-side contracts x with pi_i V[i][k] (output k), b-side with Vinv[k][i]
    double fragA[3][5], fragB[3][5];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int k = nt * 8 + g;
#pragma unroll
        for (int kt = 0; kt < 5; ++kt) {
            fragA[nt][kt] = (!kTipA && k < kStates) ? s_piv[kmap(kt, t) * kStates + k] : 0.0;   // inner or cherry end
            fragB[nt][kt] = k < kStates ? s_vinv[k * kStates + kmap(kt, t)] : 0.0;
        }
    }
    // The contraction of the products with exp(lambda_k r_c t) * {1, lambda r, (lambda r)^2} is one more small matrix
    // product, [8 rows x 24 states] x [24 states x 3], and runs on the tensor pipe as well: a product held at D-fragment
    // position (row g, state nt*8 + 2t + j) is the A element (row g, k = t) of k-tile (nt, j), so the matching B fragment is
    // efrag[nt][j] = e_n[nt*8 + 2t + j] for output column n = g < 3.  Lane t = 0 ends up with (f, f'), lane t = 1 with f''.
    // (As plain DFMAs the same contraction took 60 FP64 instructions per tile which fought the other group's DMMAs.)
    double efrag[3][2];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) efrag[nt][jj] = g < 3 ? s_exp[(g * kCats + c) * 24 + nt * 8 + 2 * t + jj] : 0.0;
    const int rounds = (cta_tiles + kMmaGroups - 1) / kMmaGroups;
    if (tr) args.trace[warp * 8 + 7] += clock64() - t_entry;  // prologue
    mma_turn_init(grp);
    for (int j = 0; j < rounds; ++j) {
        const int n = j * kMmaGroups + grp;
        if (n >= cta_tiles) {  // no tile left for this group: keep the MMA token moving
            mma_turn_begin(grp);
            mma_turn_end(grp);
            continue;
        }
        const int slot = j % kBranchDepth;
        const long long tk0 = tr ? clock64() : 0;
        mbar_wait(in_full + grp * kBranchDepth + slot, (j / kBranchDepth) & 1);
        const long long tk1 = tr ? clock64() : 0;
        const double* stage = s_stage + (size_t)(grp * kBranchDepth + slot) * Plan::kStageDoubles;
        const unsigned char* aux = reinterpret_cast<const unsigned char*>(stage + Plan::kInner * kTileDoubles);
        int2 side = make_int2(0, 0);  // category-0 warp, lanes 0-15: scaling counts and weight of row `lane`
        if (c == 0 && lane < kTileRows) {
            const int32_t* ai = reinterpret_cast<const int32_t*>(aux);
            side.x = ai[lane] + (kInnerA ? ai[kTileRows + lane] : 0);
            side.y = ai[2 * kTileRows + lane];
        }
        AFrag fa[2], fb[2];
        CherryIn cA{};
        if (kChA) cA = cherry_begin(s_tip, s_tip + kCodes * kTipPad, aux + 192, aux + 208, g, c, t);  // formed inside the turn (mma_common.cuh)
        int multiplied = 0;
        double accA[2][3][2], accB[2][3][2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (kInnerA) fa[m] = load_a(stage + m * kBlockDoubles, c, lane);
            fb[m] = load_a(stage + (kInnerA ? kTileDoubles : 0) + m * kBlockDoubles, c, lane);
            const int code = kTipA ? aux[192 + m * 8 + g] : 0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                accB[m][nt][0] = accB[m][nt][1] = 0.0;
                if (kTipA) {
                    const bool ok = nt < 2 || t < 2;
                    const double* row = s_tip + code * kTipVecPad + nt * 8 + 2 * t;
                    accA[m][nt][0] = ok ? row[0] : 0.0;
                    accA[m][nt][1] = ok ? row[1] : 0.0;
                } else {
                    accA[m][nt][0] = accA[m][nt][1] = 0.0;
                }
            }
        }
        mma_turn_begin(grp);  // see mma_common.cuh: the groups take turns on the FP64 tensor pipe
        const long long tk2 = tr ? clock64() : 0;
        // block by block; the whole contraction stays inside this group's turn (issued after the token is passed, its
        // FP64 instructions and the other group's DMMAs slow each other down: measured 1,500 instead of 1,215 clk per tile)
        if (kTipA) {
            // one side only: both blocks together keep six accumulator chains between two dependent DMMAs
#pragma unroll
            for (int kt = 0; kt < 5; ++kt)
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt) dmma(accB[m][nt][0], accB[m][nt][1], fb[m].v[kt], fragB[nt][kt]);
        } else if (kChA) {
            multiplied = children_mma<kSideCherry, kSideInner>(fa, fb, cA, cA, fragA, fragB, accA, accB, t);
        } else {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int kt = 0; kt < 5; ++kt)
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt) {
                        dmma(accA[m][nt][0], accA[m][nt][1], fa[m].v[kt], fragA[nt][kt]);
                        dmma(accB[m][nt][0], accB[m][nt][1], fb[m].v[kt], fragB[nt][kt]);
                    }
        }
        // products of the two sides, then the contraction: four accumulator chains (2 per block) of three DMMAs each
        double fs[2][2][2];
#pragma unroll
        for (int m = 0; m < 2; ++m) fs[m][0][0] = fs[m][0][1] = fs[m][1][0] = fs[m][1][1] = 0.0;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            if (multiplied & 1) break;
            accA[0][nt][0] *= accB[0][nt][0];
            accA[0][nt][1] *= accB[0][nt][1];
        }
        dmma(fs[0][0][0], fs[0][0][1], accA[0][0][0], efrag[0][0]);
        dmma(fs[0][1][0], fs[0][1][1], accA[0][0][1], efrag[0][1]);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            if (multiplied & 2) break;
            accA[1][nt][0] *= accB[1][nt][0];
            accA[1][nt][1] *= accB[1][nt][1];
        }
        dmma(fs[1][0][0], fs[1][0][1], accA[1][0][0], efrag[0][0]);
        dmma(fs[1][1][0], fs[1][1][1], accA[1][0][1], efrag[0][1]);
#pragma unroll
        for (int nt = 1; nt < 3; ++nt)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                dmma(fs[m][0][0], fs[m][0][1], accA[m][nt][0], efrag[nt][0]);
                dmma(fs[m][1][0], fs[m][1][1], accA[m][nt][1], efrag[nt][1]);
            }
        const long long tk3 = tr ? clock64() : 0;
        mma_turn_end(grp);
        __syncwarp();
        if (lane == 0) mbar_arrive(in_empty + grp * kBranchDepth + slot);  // the stage may be refilled
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            fs[m][0][0] += fs[m][1][0];
            fs[m][0][1] += fs[m][1][1];
            if (kStore) {
                const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)n * gridDim.x) * kTileRows;
                double* out = args.sumtable + (row0 + m * 8 + g) * kRow + c * kStates + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
                    if (nt < 2 || t < 2) *reinterpret_cast<double2*>(out + nt * 8) = make_double2(accA[m][nt][0], accA[m][nt][1]);
            }
        }
        const int rslot = n % kRedSlots;
        const long long tk4 = tr ? clock64() : 0;
        mbar_wait(red_empty + rslot, ((n / kRedSlots) & 1) ^ 1);
        const long long tk5 = tr ? clock64() : 0;
        double* red = s_red + ((size_t)rslot * kCats + c) * kTileRows * 3;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double* dst = red + (m * 8 + g) * 3;
            if (t == 0) {
                dst[0] = fs[m][0][0];
                dst[1] = fs[m][0][1];
            } else if (t == 1) {
                dst[2] = fs[m][0][0];
            }
        }
        if (c == 0 && lane < kTileRows) s_side[rslot * kTileRows + lane] = side;
        __syncwarp();
        if (lane == 0) mbar_arrive(red_full + rslot);
        if (tr) {  // phases: wait data | fragments + wait turn | MMAs | contraction | wait slot | row sums out | tiles
            long long* row = args.trace + warp * 8;
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
}

__global__ void k_publish(const double* result, Publish pub) {
    const double r[3] = {result[0], result[1], result[2]};
    publish_result(pub, r, result[3], false, publish_prefetch(pub, result[3]));
}

template <int KA, bool kStore>
void launch_one(const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const int ntiles = (int)(np / kTileRows);
    const int grid = ntiles < sms ? ntiles : sms;
    launch_pdl(k_branch_mma<KA, kStore>, grid, kThreadsBranch, BranchPlan<KA>::kBytes, stream, args, ntiles);
}

template <int KA>
void configure_one() {
    cudaFuncSetAttribute(k_branch_mma<KA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BranchPlan<KA>::kBytes);
    cudaFuncSetAttribute(k_branch_mma<KA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BranchPlan<KA>::kBytes);
}

}  // namespace

void configure_branch_kernels() {
    configure_one<kSideInner>();
    configure_one<kSideTip>();
    configure_one<kSideCherry>();
}

// args.result[0..2] = lnL, dlnL/dt, d2lnL/dt2 of this rank's patterns.  np must be a multiple of 16.
void launch_branch_mma(const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const int ka = side_kind(args.a);
    const bool store = args.sumtable != nullptr;
    if (ka == kSideTip) store ? launch_one<kSideTip, true>(args, np, sms, stream) : launch_one<kSideTip, false>(args, np, sms, stream);
    else if (ka == kSideCherry) store ? launch_one<kSideCherry, true>(args, np, sms, stream) : launch_one<kSideCherry, false>(args, np, sms, stream);
    else store ? launch_one<kSideInner, true>(args, np, sms, stream) : launch_one<kSideInner, false>(args, np, sms, stream);
}

void launch_publish(const double* result, const Publish& pub, cudaStream_t stream) { k_publish<<<1, 1, 0, stream>>>(result, pub); }

}  // namespace pml
