// Fused CLV update + branch pass: the last newview of a branch visit and the pass over that branch in ONE launch.
//
// In a smoothing sweep every visit of a branch (x, y) ends with "recompute the CLV of x looking at y" followed by "one pass
// over the CLVs at both ends of (x, y)" (raxmlHPC: newviewGeneric + makenewzGeneric).  As two kernels the fresh CLV goes out
// to HBM and comes straight back, and the visit pays two prologues and two drains.  Here a 16-pattern tile of the new CLV
// is consumed from the shared-memory slot it is stored from:
//   warps 0-3   NV group (one warp per rate category): children tiles x P  ->  products  ->  product slot   (newview_mma.cu)
//   warps 4-7   BR group: product slot (the x end) and the tile of the y end  ->  eigen-space transforms, contraction with
//               exp(lambda r t) {1, lambda r, (lambda r)^2}  ->  row sums                                      (branch_mma.cu)
//   warps 8-9   helpers: rescale test + bulk store of the new CLV tile, and the finishing of the row sums (log, weights)
//   warp 10     producer (TMA bulk copies for both groups), warp 11 only helps with the prologue
// The two groups alternate on the FP64 tensor pipe, the BR group one tile behind the NV group so that it never waits for
// products.  The BR group reads the products BEFORE the (rare) x2^256 rescale is applied in place, and uses the children's
// scaling counts without the increment: the same lnL, bit for bit the same scaled CLV in HBM.
// The tail is the branch kernel's: fixed-order sum over CTAs, sum over ranks through NVLink mailboxes, guarded NR step on
// the device, publication.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "mma_common.cuh"
#include "pmatrix.cuh"

namespace pml {

namespace {

using namespace mma;

constexpr double kLogMinLik = -177.445678223345993274;  // ln 2^-256
constexpr int kTipVecPad = 22;
constexpr int kThreadsFused = 384;
constexpr int kStagers = kThreadsFused - 32;
constexpr int kFDepth = 4;      // input stages of either group
constexpr int kSlots = 4;       // product slots and row-sum slots (tile n -> n % 4)
constexpr int kStageBarrier = 2;
constexpr int kFinishBarrier = 3;

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

// KL, KR: the children of the update (tip / cherry / inner, ordered tip < cherry < inner, at least one not a tip);
// KY: the far end of the branch.  A folded cherry (Side in kernels.h) is formed in registers from two tip look-ups.
constexpr int kTableDoubles = kCodes * kTipPad;
template <int KL, int KR, int KY>
struct FusedPlan {
    static constexpr int kInner = (KL == kSideInner ? 1 : 0) + (KR == kSideInner ? 1 : 0);
    static constexpr int kTablesL = KL == kSideTip ? 1 : (KL == kSideCherry ? 2 : 0);
    static constexpr int kTablesR = KR == kSideTip ? 1 : (KR == kSideCherry ? 2 : 0);
    static constexpr int kTablesY = KY == kSideCherry ? 2 : 0;
    // branches whose P matrices the prologue builds: the two of the update, then the tips of the cherries (left, right, far end)
    static constexpr int kBranches = 2 + (KL == kSideCherry ? 2 : 0) + (KR == kSideCherry ? 2 : 0) + (KY == kSideCherry ? 2 : 0);
    __host__ __device__ static constexpr int branch_id(int pos) {  // 0, 1: the update's; 2, 3 / 4, 5 / 6, 7: tips of the left / right / far cherry
        if (pos < 2) return pos;
        int p = pos - 2;
        if (KL == kSideCherry) {
            if (p < 2) return 2 + p;
            p -= 2;
        }
        if (KR == kSideCherry) {
            if (p < 2) return 4 + p;
            p -= 2;
        }
        return 6 + p;
    }
    // NV stage: CLV tiles of the inner children + [0,64) / [64,128) their scaling counts, [128,160) / [160,192) residue codes of
    // the left / right child (tip: 16 B, cherry: 2 x 16 B)
    static constexpr int kNvStageDoubles = kInner * kTileDoubles + 32;
    // Y stage: CLV tile of the far end (inner node only) + [0,64) its scaling counts, [64,128) pattern weights, [128,160) its codes
    static constexpr int kYStageDoubles = (KY == kSideInner ? kTileDoubles : 0) + 32;
    static constexpr int kModelDoubles = 3 * pmat::kMat;          // V, Vinv, pi V
    static constexpr int kExpDoubles = 3 * kCats * 24;
    static constexpr int kTipDoubles = (kTablesL + kTablesR + kTablesY) * kTableDoubles;
    static constexpr int kTipYDoubles = KY == kSideTip ? kCodes * kTipVecPad : 0;
    static constexpr int kRedDoubles = kSlots * kCats * kTileRows * 3;
    static constexpr int kInts = kSlots * kCats * kTileRows /* max */ + kSlots * kTileRows /* sc */ + kSlots * kTileRows * 2 /* side */;
    static constexpr size_t kBarBytes = 512;
    static constexpr size_t kBytes = kBarBytes +
                                     sizeof(double) * (size_t)(kModelDoubles + kExpDoubles + kTipDoubles + kTipYDoubles + kRedDoubles + 8) +
                                     sizeof(int) * kInts +
                                     sizeof(double) * (size_t)(kSlots * kTileDoubles + kFDepth * (kNvStageDoubles + kYStageDoubles));
    static_assert(kSlots * kTileDoubles >= 8 * pmat::kFragSlotDoubles, "the product slots double as the fragment exchange area");
    static_assert(kBytes <= 227 * 1024, "shared memory budget of one CTA");
};

template <int KL, int KR, int KY>
__global__ void __launch_bounds__(kThreadsFused, 1) k_fused(NewviewOp op, BranchArgs args, int ntiles) {
    using Plan = FusedPlan<KL, KR, KY>;
    constexpr bool kInnerL = KL == kSideInner, kInnerR = KR == kSideInner, kInnerY = KY == kSideInner;
    constexpr bool kTipL = KL == kSideTip, kTipR = KR == kSideTip, kTipY = KY == kSideTip;
    constexpr bool kChL = KL == kSideCherry, kChR = KR == kSideCherry, kChY = KY == kSideCherry;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* nv_full = reinterpret_cast<uint64_t*>(smem_raw);  // [kFDepth]
    uint64_t* nv_empty = nv_full + kFDepth;                      // [kFDepth], 4 arrivals (NV warps)
    uint64_t* y_full = nv_empty + kFDepth;                       // [kFDepth]
    uint64_t* y_empty = y_full + kFDepth;                        // [kFDepth], 4 arrivals (BR warps)
    uint64_t* prod_full = y_empty + kFDepth;                     // [kSlots], 4 arrivals (NV warps)
    uint64_t* prod_read = prod_full + kSlots;                    // [kSlots], 4 arrivals (BR warps have their fragments)
    uint64_t* prod_empty = prod_read + kSlots;                   // [kSlots], 5 arrivals (BR warps + the storing helper)
    uint64_t* red_full = prod_empty + kSlots;                    // [kSlots], 4 arrivals (BR warps)
    uint64_t* red_empty = red_full + kSlots;                     // [kSlots], 1 arrival (finishing helper)
    double* s_model = reinterpret_cast<double*>(smem_raw + Plan::kBarBytes);  // V, Vinv (pmatrix.cuh layout), then pi V
    double* s_vinv = s_model + pmat::kMat;
    double* s_piv = s_model + 2 * pmat::kMat;
    double* s_exp = s_model + Plan::kModelDoubles;               // [3][kCats][24]
    double* s_tabL = s_exp + Plan::kExpDoubles;                  // 23 x 80 look-ups of the left child (tip: 1, cherry: 2)
    double* s_tabR = s_tabL + Plan::kTablesL * kTableDoubles;    // ... of the right child
    double* s_tabY = s_tabR + Plan::kTablesR * kTableDoubles;    // ... of a cherry at the far end
    double* s_tipy = s_tabY + Plan::kTablesY * kTableDoubles;    // tip far end: 23 x 20 lookup (pi V sums)
    double* s_red = s_tipy + Plan::kTipYDoubles;                 // [kSlots][kCats][16][3]
    double* s_fin = s_red + Plan::kRedDoubles;                   // [2][3] (+2)
    int* s_max = reinterpret_cast<int*>(s_fin + 8);              // [kSlots][kCats][16]
    int* s_sc = s_max + kSlots * kCats * kTileRows;              // [kSlots][16] scaling counts of the children
    int2* s_side = reinterpret_cast<int2*>(s_sc + kSlots * kTileRows);  // [kSlots][16] {total scaling count, weight}
    double* s_prod = reinterpret_cast<double*>(s_side + kSlots * kTileRows);
    double* s_nvstage = s_prod + kSlots * kTileDoubles;
    double* s_ystage = s_nvstage + kFDepth * Plan::kNvStageDoubles;
    const DeviceModel* dm = args.dm;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stid = warp < kProducerWarp ? threadIdx.x : threadIdx.x - 32;  // rank among the staging threads
    const bool tl = op.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    if (tl) op.timeline[0] = global_timer_ns();
    pdl_launch_dependents();
    // ---- static model constants first (before the dependency wait and before any bulk load is queued) ----------------
    double pre[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    double pre_lambda = 0.0, pre_rate = 0.0, lr = 0.0, pre_tip[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    pmat::ExtraB xb{};
    const int c_p = warp & 3;  // MMA warp w builds category w & 3 of the branches at positions (w >> 2), (w >> 2) + 2, ... (Plan::branch_id)
    if (warp != kProducerWarp) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = stid + q * kStagers;
            if (idx < pmat::kMat) {
                pre[0][q] = (&dm->V[0][0])[idx];
                pre[1][q] = (&dm->Vinv[0][0])[idx];
                pre[2][q] = (&dm->piV[0][0])[idx];
            }
        }
        // the 96 threads that build no matrix (helper warps, warp 11) take the exponentials and the tip vector: on the MMA warps
        // those exponentials (FP64 library code next to the other warps' DMMAs) delayed three of the eight matrix builders
        if (stid >= kMmaWarps * 32) {
            const int j = stid - kMmaWarps * 32;
            if (j % 24 < kStates) {
                pre_lambda = dm->lambda[j % 24];
                pre_rate = dm->rates[j / 24];
            }
            if (kTipY) {
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int idx = j + q * 96;
                    pre_tip[q] = idx < kCodes * kStates ? (&dm->tipvec[0][0])[idx] : 0.0;
                }
            }
        }
        if (warp < kMmaWarps && lane < kStates) lr = dm->lambda[lane] * dm->rates[c_p];
        if (warp < kMmaWarps && Plan::kTablesL + Plan::kTablesR + Plan::kTablesY > 0) xb = pmat::extra_b(dm, lane);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kFDepth; ++i) {
            mbar_init(nv_full + i, 1);
            mbar_init(nv_empty + i, 4);
            mbar_init(y_full + i, 1);
            mbar_init(y_empty + i, 4);
        }
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(prod_full + i, 4);
            mbar_init(prod_read + i, 4);
            mbar_init(prod_empty + i, 5);
            mbar_init(red_full + i, 4);
            mbar_init(red_empty + i, 1);
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
        const int id = Plan::branch_id(2 * r + (warp >> 2));
        len_src[r] = id == 0 ? op.len_left : id == 1 ? op.len_right : id == 2 ? op.left.len1 : id == 3 ? op.left.len2
                     : id == 4 ? op.right.len1 : id == 5 ? op.right.len2 : id == 6 ? args.a.len1 : args.a.len2;
        len_mul[r] = id < 2 ? op.len_scale : 1.0;
        asm volatile("" : "+l"(len_src[r]), "+d"(len_mul[r]));
    }
    const double* t_src = args.t_ptr;
    asm volatile("" : "+l"(t_src));
    pdl_wait();  // from here on the kernel touches what its predecessors wrote: branch lengths, CLVs, scaling counts
    if (tl) op.timeline[1] = global_timer_ns();
    // lengths are requested before the model constants (in flight since the kernel started) are consumed, and nothing ahead of
    // the CTA barrier waits for them: one latency, not two
    double my_len[kRounds];
    if (warp < kMmaWarps) {
#pragma unroll
        for (int r = 0; r < kRounds; ++r) my_len[r] = *len_src[r];
    }
    // the length the sums are taken at is needed by the warps that build no matrix only (exponentials, publication): the MMA
    // warps do not wait for its load
    const bool wants_t = warp >= kMmaWarps && warp != kProducerWarp;
    const double t_raw = (t_src && wants_t) ? *t_src : args.t;
    if (warp != kProducerWarp) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = stid + q * kStagers;
            if (idx < pmat::kMat) {
                s_model[idx] = pre[0][q];
                s_vinv[idx] = pre[1][q];
                s_piv[idx] = pre[2][q];
            }
        }
    }
    __syncthreads();
    if (warp < kMmaWarps) {
#pragma unroll
        for (int r = 0; r < kRounds; ++r) my_len[r] *= len_mul[r];
    }
    // the length the sums are taken at: the device copy of the branch length is brought into the NR range first
    const double tt = (t_src && wants_t) ? nr_clamp_length(t_raw) : t_raw;

    const int cta_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == kProducerWarp) {
        // ---------------------------------------------------------------------------------------------- producer
        constexpr uint32_t bytes = kTileDoubles * sizeof(double), int_bytes = kTileRows * sizeof(int32_t);
        for (int n = 0; n < cta_tiles; ++n) {
            const int slot = n % kFDepth;
            const uint32_t par = ((n / kFDepth) & 1) ^ 1;
            if (lane == 0) {
                mbar_wait(nv_empty + slot, par);
                mbar_expect_tx(nv_full + slot, Plan::kInner * (bytes + int_bytes) + (kTipL ? kTileRows : 0) + (kChL ? 2 * kTileRows : 0) +
                                                   (kTipR ? kTileRows : 0) + (kChR ? 2 * kTileRows : 0));
            } else if (lane == 4) {
                mbar_wait(y_empty + slot, par);
                mbar_expect_tx(y_full + slot, (kInnerY ? bytes + int_bytes : (kTipY ? kTileRows : 2 * kTileRows)) + int_bytes);
            }
            __syncwarp();
            const size_t tile = (size_t)blockIdx.x + (size_t)n * gridDim.x;
            const size_t goff = tile * kTileDoubles;
            double* dst = s_nvstage + (size_t)slot * Plan::kNvStageDoubles;
            unsigned char* aux = reinterpret_cast<unsigned char*>(dst + Plan::kInner * kTileDoubles);
            double* ydst = s_ystage + (size_t)slot * Plan::kYStageDoubles;
            unsigned char* yaux = reinterpret_cast<unsigned char*>(ydst + (kInnerY ? kTileDoubles : 0));
            if (lane == 0) {
                if (kInnerL) bulk_g2s(dst, op.left.clv + goff, bytes, nv_full + slot);
                else bulk_g2s(aux + 128, op.left.codes + tile * kTileRows, kTileRows, nv_full + slot);
            } else if (lane == 1) {
                if (kInnerR) bulk_g2s(dst + (kInnerL ? kTileDoubles : 0), op.right.clv + goff, bytes, nv_full + slot);
                else bulk_g2s(aux + 160, op.right.codes + tile * kTileRows, kTileRows, nv_full + slot);
            } else if (lane == 2) {
                if (kInnerL) bulk_g2s(aux, op.left.scale + tile * kTileRows, int_bytes, nv_full + slot);
                else if (kChL) bulk_g2s(aux + 144, op.left.codes2 + tile * kTileRows, kTileRows, nv_full + slot);
            } else if (lane == 3) {
                if (kInnerR) bulk_g2s(aux + int_bytes, op.right.scale + tile * kTileRows, int_bytes, nv_full + slot);
                else if (kChR) bulk_g2s(aux + 176, op.right.codes2 + tile * kTileRows, kTileRows, nv_full + slot);
            } else if (lane == 4) {
                if (kInnerY) bulk_g2s(ydst, args.a.clv + goff, bytes, y_full + slot);
                else bulk_g2s(yaux + 128, args.a.codes + tile * kTileRows, kTileRows, y_full + slot);
            } else if (lane == 5) {
                bulk_g2s(yaux + 64, args.weights + tile * kTileRows, int_bytes, y_full + slot);
            } else if (lane == 6) {
                if (kInnerY) bulk_g2s(yaux, args.a.scale + tile * kTileRows, int_bytes, y_full + slot);
                else if (kChY) bulk_g2s(yaux + 144, args.a.codes2 + tile * kTileRows, kTileRows, y_full + slot);
            }
        }
        return;
    }

    // ---- prologue of every other warp ----------------------------------------------------------------------------------
    if (stid >= kMmaWarps * 32) {
        const int j = stid - kMmaWarps * 32;   // 0 .. 95 = kCats * 24
        const double a = pre_lambda * pre_rate;
        const double e = j % 24 < kStates ? pmat::exp_neg(a * tt) : 0.0;
        s_exp[j] = e;
        s_exp[kCats * 24 + j] = a * e;
        s_exp[2 * kCats * 24 + j] = a * a * e;
        if (kTipY) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int idx = j + q * 96;
                if (idx < kCodes * kStates) s_tipy[(idx / kStates) * kTipVecPad + idx % kStates] = pre_tip[q];
            }
        }
    }
    double fragL[3][5], fragR[3][5];  // NV group: P fragments of the two children;  BR group: pi V (y end) and Vinv (x end)
    double* s_x = s_prod;             // fragment exchange: [branch 0 / 1][category][15][32]
    if (warp < kMmaWarps) {
        // an inner or cherry child's matrix leaves the accumulators as B fragments for the NV warp of its category; every tip's
        // matrix -- a tip child's or a cherry's -- goes straight into its look-up table
        // the exponentials of all rounds first: their dependency chains run side by side instead of one per round
        double e_round[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) e_round[r] = pmat::exp_neg(lr * my_len[r]);
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int id = Plan::branch_id(2 * r + (warp >> 2));
            double acc[3][3][2];
            pmat::build_p_tiles(s_model, e_round[r], lane, xb, acc);
            double* table = nullptr;
            if (id == 0 && kTipL) table = s_tabL;
            else if (id == 1 && kTipR) table = s_tabR;
            else if (id == 2 || id == 3) table = s_tabL + (id - 2) * kTableDoubles;
            else if (id == 4 || id == 5) table = s_tabR + (id - 4) * kTableDoubles;
            else if (id >= 6) table = s_tabY + (id - 6) * kTableDoubles;
            if (table) pmat::tiles_to_lookup(acc, lane, c_p, table, kTipPad);
            else {
                double frag[3][5];
                pmat::tiles_to_fragments(acc, lane, frag);
                pmat::fragments_to_smem(frag, lane, s_x + (id * kCats + c_p) * pmat::kFragSlotDoubles);
            }
        }
    }
    named_barrier(kStageBarrier, kStagers);
    double efrag[3][2];
    if (warp < 4) {  // NV group: the fragments of both children come from the exchange area
        if (!kTipL) pmat::fragments_from_smem(fragL, lane, s_x + c_p * pmat::kFragSlotDoubles);
        if (!kTipR) pmat::fragments_from_smem(fragR, lane, s_x + (kCats + c_p) * pmat::kFragSlotDoubles);
    } else if (warp < kMmaWarps) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            const int k = nt * 8 + g;
#pragma unroll
            for (int kt = 0; kt < 5; ++kt) {
                fragL[nt][kt] = (!kTipY && k < kStates) ? s_piv[kmap(kt, t) * kStates + k] : 0.0;   // y end: x pi V
                fragR[nt][kt] = k < kStates ? s_vinv[k * kStates + kmap(kt, t)] : 0.0;               // x end: Vinv x
            }
        }
    }
    named_barrier(kStageBarrier, kStagers);  // exponentials, lookups are in place; the product slots are free for their purpose
    if (warp >= 4 && warp < kMmaWarps) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) efrag[nt][jj] = g < 3 ? s_exp[(g * kCats + c_p) * 24 + nt * 8 + 2 * t + jj] : 0.0;
    }
    if (warp > kProducerWarp) return;

    if (warp >= kMmaWarps) {
        // ---------------------------------------------------------------------------------------------- helpers
        // warp e stores the tiles n = e, e+2, ... of the new CLV and finishes the row sums of the tile pairs q = e, e+2, ...
        const int e = warp - kMmaWarps;
        double sum_l = 0.0, sum_d1 = 0.0, sum_d2 = 0.0;
        const PublishEarly early = publish_prefetch(args.pub, tt);  // while the tiles are still streaming
        int pending[2] = {-1, -1};  // product slots whose bulk store may still be reading shared memory
        auto release_stores = [&]() {
            if (lane == 0 && (pending[0] >= 0 || pending[1] >= 0)) {
                bulk_wait_read<0>();
                if (pending[0] >= 0) mbar_arrive(prod_empty + pending[0]);
                if (pending[1] >= 0) mbar_arrive(prod_empty + pending[1]);
            }
            pending[0] = pending[1] = -1;
        };
        auto store_tile = [&](int n, int which) {
            const int slot = n % kSlots, r = lane & 15;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)n * gridDim.x;
            mbar_wait(prod_full + slot, (n / kSlots) & 1);
            const int* mx = s_max + slot * kCats * kTileRows;
            const int big = max(max(mx[r], mx[kTileRows + r]), max(mx[2 * kTileRows + r], mx[3 * kTileRows + r]));
            const bool rescale = big < kMinLikHi;
            const unsigned flagged = __ballot_sync(0xffffffffu, rescale && lane < kTileRows);
            double* prod = s_prod + (size_t)slot * kTileDoubles;
            if (flagged) {  // rare: the BR group must have its fragments before the rows change in place
                mbar_wait(prod_read + slot, (n / kSlots) & 1);
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
            pending[which] = slot;
        };
        auto finish_pair = [&](int q) {
            const int half = lane >> 4, r = lane & 15;
            const int n0 = 2 * q, n1 = n0 + 1, n = n0 + half;
            mbar_wait(red_full + n0 % kSlots, (n0 / kSlots) & 1);
            if (n1 < cta_tiles) mbar_wait(red_full + n1 % kSlots, (n1 / kSlots) & 1);
            if (n < cta_tiles) {
                const int slot = n % kSlots;
                const double* red = s_red + (size_t)slot * kCats * kTileRows * 3 + r * 3;
                constexpr int cs = kTileRows * 3;
                const double f = (red[0] + red[cs]) + (red[2 * cs] + red[3 * cs]);
                const double f1 = (red[1] + red[cs + 1]) + (red[2 * cs + 1] + red[3 * cs + 1]);
                const double f2 = (red[2] + red[cs + 2]) + (red[2 * cs + 2] + red[3 * cs + 2]);
                const int2 side = s_side[slot * kTileRows + r];
                const double w = (double)side.y;
                if (args.want_lnl) {
                    const double lnl = log(0.25 * f) + side.x * kLogMinLik;
                    if (args.site_lnl) args.site_lnl[((int64_t)blockIdx.x + (int64_t)n * gridDim.x) * kTileRows + r] = lnl;
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
                mbar_arrive(red_empty + n0 % kSlots);
                if (n1 < cta_tiles) mbar_arrive(red_empty + n1 % kSlots);
            }
        };
        const int pairs = (cta_tiles + 1) / 2;
        for (int j = 0; 4 * j + e < cta_tiles || 2 * j + e < pairs; ++j) {
            const int ta = 4 * j + e, tb = 4 * j + 2 + e, q = 2 * j + e;
            if (ta < cta_tiles) store_tile(ta, 0);
            if (tb < cta_tiles) store_tile(tb, 1);
            release_stores();
            if (q < pairs) finish_pair(q);
        }
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
        // CTA partials, then the CTA that draws the last ticket adds all of them in a fixed order (branch_mma.cu)
        if (lane < 3) {
            args.partials[(int64_t)lane * gridDim.x + blockIdx.x] = s_fin[lane] + s_fin[3 + lane];
            fence_acq_rel_gpu();
        }
        __syncwarp();
        unsigned int ticket = 0;
        if (lane == 0) ticket = atomicAdd(args.ticket, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket != gridDim.x - 1) return;
        if (op.timeline && lane == 0) op.timeline[4] = global_timer_ns();
        fence_acq_rel_gpu();
        double r3[3] = {0.0, 0.0, 0.0};
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
            if (op.timeline) op.timeline[5] = global_timer_ns();
        }
        return;
    }

    const int c = warp & 3, g = lane >> 2, t = lane & 3;
    if (warp < 4) {
        // -------------------------------------------------------------------------------------------- NV group
        const bool tr = op.trace != nullptr && blockIdx.x == 0 && lane == 0;
        mma_turn_init(0);
        if (tl) op.timeline[2] = global_timer_ns();
        for (int n = 0; n <= cta_tiles; ++n) {
            if (n == cta_tiles) {  // one empty turn at the end: the BR group is one tile behind
                mma_turn_begin(0);
                mma_turn_end(0);
                if (tl) op.timeline[3] = global_timer_ns();
                break;
            }
            const int slot = n % kFDepth;
            const long long k0 = tr ? clock64() : 0;
            mbar_wait(nv_full + slot, (n / kFDepth) & 1);
            const long long k1 = tr ? clock64() : 0;
            const double* stage = s_nvstage + (size_t)slot * Plan::kNvStageDoubles;
            const unsigned char* aux = reinterpret_cast<const unsigned char*>(stage + Plan::kInner * kTileDoubles);
            int32_t sc_sum = 0;
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
            mma_turn_begin(0);
            const long long k2 = tr ? clock64() : 0;
            const int multiplied = children_mma<KL, KR>(aL, aR, cL, cR, fragL, fragR, accL, accR, t);
            const long long k3 = tr ? clock64() : 0;
            mma_turn_end(0);
            __syncwarp();
            if (lane == 0) mbar_arrive(nv_empty + slot);
            const int pslot = n % kSlots;
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
            const long long k4 = tr ? clock64() : 0;
            mbar_wait(prod_empty + pslot, ((n / kSlots) & 1) ^ 1);
            const long long k5 = tr ? clock64() : 0;
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
            if (tr) {  // wait data | fragments + wait turn | MMAs | products | wait slot | store | tiles
                long long* row = op.trace + warp * 8;
                const long long k6 = clock64();
                row[0] += k1 - k0;
                row[1] += k2 - k1;
                row[2] += k3 - k2;
                row[3] += k4 - k3;
                row[4] += k5 - k4;
                row[5] += k6 - k5;
                row[6] += 1;
            }
        }
        return;
    }

    // ------------------------------------------------------------------------------------------------ BR group
    // The contraction of a tile's products with the exponentials (12 DMMA in four chains of three) is issued one turn
    // LATER, between the transforms of the next tile: issued right behind its own transforms it waits for their results and
    // then for each link of its chains while the group holds the pipe (measured 1,750 instead of 1,152 clk per turn).
    double sv[2][3][2];        // products of the previous tile, waiting for their contraction
    int2 side_prev = make_int2(0, 0);
    int prev_n = -1;
    // two links per block and k step; a block's chain advances by two DMMAs that sit a whole k step of transforms apart from
    // the next pair, so one accumulator per block is enough (no partial sums to add afterwards)
    auto contraction_step = [&](int nt, double (&fs)[2][2]) {
#pragma unroll
        for (int m = 0; m < 2; ++m) dmma(fs[m][0], fs[m][1], sv[m][nt][0], efrag[nt][0]);
#pragma unroll
        for (int m = 0; m < 2; ++m) dmma(fs[m][0], fs[m][1], sv[m][nt][1], efrag[nt][1]);
    };
    auto emit_row_sums = [&](int n, const double (&fs)[2][2], int2 side) {
        const int rslot = n % kSlots;
        mbar_wait(red_empty + rslot, ((n / kSlots) & 1) ^ 1);
        double* red = s_red + ((size_t)rslot * kCats + c) * kTileRows * 3;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double* dst = red + (m * 8 + g) * 3;
            if (t == 0) {
                dst[0] = fs[m][0];
                dst[1] = fs[m][1];
            } else if (t == 1) {
                dst[2] = fs[m][0];
            }
        }
        if (c == 0 && lane < kTileRows) s_side[rslot * kTileRows + lane] = side;
        __syncwarp();
        if (lane == 0) mbar_arrive(red_full + rslot);
    };
    mma_turn_init(1);
    mma_turn_begin(1);  // one empty turn at the start: the NV group runs one tile ahead
    mma_turn_end(1);
    const bool trb = op.trace != nullptr && blockIdx.x == 0 && lane == 0;
    for (int n = 0; n < cta_tiles; ++n) {
        const int pslot = n % kSlots, yslot = n % kFDepth;
        const long long b0 = trb ? clock64() : 0;
        mbar_wait(prod_full + pslot, (n / kSlots) & 1);
        const long long b1 = trb ? clock64() : 0;
        mbar_wait(y_full + yslot, (n / kFDepth) & 1);
        const long long b2 = trb ? clock64() : 0;
        const double* prod = s_prod + (size_t)pslot * kTileDoubles;
        const double* ystage = s_ystage + (size_t)yslot * Plan::kYStageDoubles;
        const unsigned char* yaux = reinterpret_cast<const unsigned char*>(ystage + (kInnerY ? kTileDoubles : 0));
        int2 side = make_int2(0, 0);
        if (c == 0 && lane < kTileRows) {
            const int32_t* ai = reinterpret_cast<const int32_t*>(yaux);
            side.x = s_sc[pslot * kTileRows + lane] + (kInnerY ? ai[lane] : 0);
            side.y = ai[kTileRows + lane];
        }
        AFrag fx[2], fy[2];
        CherryIn cY{};
        if (kChY) cY = cherry_begin(s_tabY, s_tabY + kTableDoubles, yaux + 128, yaux + 144, g, c, t);  // formed inside the turn (mma_common.cuh)
        double accY[2][3][2], accX[2][3][2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            fx[m] = load_a(prod + m * kBlockDoubles, c, lane);
            if (kInnerY) fy[m] = load_a(ystage + m * kBlockDoubles, c, lane);
            const int code = kTipY ? yaux[128 + m * 8 + g] : 0;
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                accX[m][nt][0] = accX[m][nt][1] = 0.0;
                if (kTipY) {
                    const bool ok = nt < 2 || t < 2;
                    const double* row = s_tipy + code * kTipVecPad + nt * 8 + 2 * t;
                    accY[m][nt][0] = ok ? row[0] : 0.0;
                    accY[m][nt][1] = ok ? row[1] : 0.0;
                } else {
                    accY[m][nt][0] = accY[m][nt][1] = 0.0;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {  // the products are in registers: the slot may be rescaled, stored and reused
            mbar_arrive(prod_read + pslot);
            mbar_arrive(prod_empty + pslot);
        }
        double fs[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const long long b3 = trb ? clock64() : 0;
        mma_turn_begin(1);
        const long long b4 = trb ? clock64() : 0;
        // previous tile's contraction: one link of each chain ahead of the first three k steps of block 0
        const bool have_prev = prev_n >= 0;
        auto slip_in = [&](int kt) {
            if (kt < 3 && have_prev) contraction_step(kt, fs);
        };
        const int multiplied = children_mma<KY, kSideInner>(fy, fx, cY, cY, fragL, fragR, accY, accX, t, slip_in);
        const long long b5 = trb ? clock64() : 0;
        mma_turn_end(1);
        __syncwarp();
        if (lane == 0) mbar_arrive(y_empty + yslot);
        if (prev_n >= 0) emit_row_sums(prev_n, fs, side_prev);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                sv[m][nt][0] = (multiplied >> m & 1) ? accY[m][nt][0] : accY[m][nt][0] * accX[m][nt][0];
                sv[m][nt][1] = (multiplied >> m & 1) ? accY[m][nt][1] : accY[m][nt][1] * accX[m][nt][1];
            }
        side_prev = side;
        prev_n = n;
        if (trb) {  // wait products | wait far-end tile | fragments | wait turn | MMAs | row sums + products | tiles
            long long* row = op.trace + warp * 8;
            const long long b6 = clock64();
            row[0] += b1 - b0;
            row[1] += b2 - b1;
            row[2] += b3 - b2;
            row[3] += b4 - b3;
            row[4] += b5 - b4;
            row[5] += b6 - b5;
            row[6] += 1;
        }
    }
    if (prev_n >= 0) {  // the last tile's contraction: nobody else needs the pipe any more
        double fs[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) contraction_step(nt, fs);
        emit_row_sums(prev_n, fs, side_prev);
    }
}

template <int KL, int KR, int KY>
void launch_one(const NewviewOp& op, const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const int ntiles = (int)(np / kTileRows);
    const int grid = ntiles < sms ? ntiles : sms;
    launch_pdl(k_fused<KL, KR, KY>, grid, kThreadsFused, FusedPlan<KL, KR, KY>::kBytes, stream, op, args, ntiles);
}
template <int KL, int KR>
void launch_far(const NewviewOp& op, const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const int ky = side_kind(args.a);
    if (ky == kSideInner) launch_one<KL, KR, kSideInner>(op, args, np, sms, stream);
    else if (ky == kSideTip) launch_one<KL, KR, kSideTip>(op, args, np, sms, stream);
    else launch_one<KL, KR, kSideCherry>(op, args, np, sms, stream);
}
template <int KL, int KR>
void configure_pair() {
    cudaFuncSetAttribute(k_fused<KL, KR, kSideInner>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedPlan<KL, KR, kSideInner>::kBytes);
    cudaFuncSetAttribute(k_fused<KL, KR, kSideTip>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedPlan<KL, KR, kSideTip>::kBytes);
    cudaFuncSetAttribute(k_fused<KL, KR, kSideCherry>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedPlan<KL, KR, kSideCherry>::kBytes);
}

}  // namespace

void configure_fused_kernels() {
    configure_pair<kSideTip, kSideInner>();
    configure_pair<kSideInner, kSideInner>();
    configure_pair<kSideTip, kSideCherry>();
    configure_pair<kSideCherry, kSideCherry>();
    configure_pair<kSideCherry, kSideInner>();
}

// op: the CLV update of the branch's x end (children not both tips); args.a: the y end (inner, tip or folded cherry), args.b is
// ignored (the x end is op.out).  No product table (args.sumtable must be null).  np must be a multiple of 16.
void launch_fused(const NewviewOp& in, const BranchArgs& args, int64_t np, int sms, cudaStream_t stream) {
    const NewviewOp op = canonical_children(in);
    const int kl = side_kind(op.left), kr = side_kind(op.right);
    if (kl == kSideTip) kr == kSideCherry ? launch_far<kSideTip, kSideCherry>(op, args, np, sms, stream) : launch_far<kSideTip, kSideInner>(op, args, np, sms, stream);
    else if (kl == kSideCherry) kr == kSideCherry ? launch_far<kSideCherry, kSideCherry>(op, args, np, sms, stream) : launch_far<kSideCherry, kSideInner>(op, args, np, sms, stream);
    else launch_far<kSideInner, kSideInner>(op, args, np, sms, stream);
}

}  // namespace pml
