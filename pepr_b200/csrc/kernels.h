// Device-side interface of the engine: every function enqueues work on `stream` and returns immediately.
// The transition matrices P(t) are built inside the CLV kernels (pmatrix.cuh) from the eigensystem below and the branch
// lengths in the tree's device array.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "model.h"

namespace pml {

// model constants in device memory, one copy per alignment (eigenvalues, V, Vinv, pi, rates)
struct DeviceModel {
    double lambda[kStates];
    double V[kStates][kStates];
    double Vinv[kStates][kStates];
    double pi[kStates];
    double rates[kCats];
    double piV[kStates][kStates];   // pi_i * V[i][k]
    double tipvec[kCodes][kStates]; // sum over the residues a code allows of pi_i V[i][k]: the eigen-space vector of a tip end
    // Vinv[k][.] summed over the residues of the ambiguity codes 20 (B = N|D), 21 (Z = Q|E), 22 (undetermined), then 0: three more
    // columns for the matrix product that builds P(t), which then delivers a tip's look-up rows of those codes by itself
    double vinv_codes[kStates][4];
};

// one side of a branch: an inner node's CLV (+ cumulative scaling counts), a tip's residue codes, or a CHERRY -- an inner
// node both of whose children on that side are tips.  A cherry's CLV is the product of two tip look-ups and is never
// stored: every consumer forms it on the fly from the two code rows and the two tip branch lengths (no 640 B/pattern round
// trip through HBM, no launch).  Its scaling count is 0: with both lengths inside the Newton-Raphson range the
// fastest rate category alone keeps one of the 80 products far above 2^-256.
struct Side {
    const double* clv;      // inner node only
    const int32_t* scale;   // inner node only
    const uint8_t* codes;   // tip: its residue codes (0..22); cherry: codes of its first tip
    const uint8_t* codes2;  // cherry only: codes of its second tip
    const double* len1;     // cherry only: where the branch lengths of its two tips live on the device
    const double* len2;
};
enum SideKind : int { kSideInner = 0, kSideTip = 1, kSideCherry = 2 };
inline int side_kind(const Side& s) { return s.clv ? kSideInner : (s.codes2 ? kSideCherry : kSideTip); }

struct NewviewOp {
    Side left, right;
    const double* len_left;   // the two branch lengths, read on the device (tree's length array)
    const double* len_right;
    double len_scale;         // both lengths are multiplied by this (1; 0.5 for the halves of a branch a subtree is inserted into)
    const DeviceModel* dm;
    double* out;
    int32_t* out_scale;
    long long* trace;  // optional (profiling aid): per warp of CTA 0, cycles spent in each phase of the pipeline
    // optional (profiling aid, pml_timeline_*): six %globaltimer stamps of this launch -- [0] CTA 0 enters, [1] its dependency
    // wait returns, [2] its first MMA turn, [3] its last tile is through the MMA warps, [4] fused kernel: the last CTA has drawn
    // its ticket, [5] fused kernel: the result is published
    unsigned long long* timeline;
};

// the children ordered tip < cherry < inner (their product commutes exactly): the CLV kernels exist for that order only
NewviewOp canonical_children(const NewviewOp& op);
// CLV update on the FP64 tensor path (TMA-fed DMMA); np must be a multiple of 128 and all buffers hold np rows
void launch_newview_mma(const NewviewOp& op, int64_t np, int sms, cudaStream_t stream);
// raises the dynamic shared-memory limit of the tensor-path kernels on the current device (once per context)
void configure_mma_kernels();

// ---- guarded Newton-Raphson step on one branch (raxmlHPC topLevelMakenewz / makenewzGeneric) ----------
// z = exp(-t) is updated in log z; d1t, d2t are dlnL/dt and d2lnL/dt2 at length t.
//   bad curvature (d2 >= 0 in log z, z < zmax)  ->  z = 0.37 z + 0.63, the derivatives must be taken again (kNrRetry)
//   otherwise z *= exp(-d1/d2) when that exponent is < 100, capped at 0.25 z + 0.75 and at zmax (kNrDone)
constexpr double kZmin = 1.0e-15, kZmax = 1.0 - 1.0e-6;
enum NrStatus : int { kNrNone = 0, kNrDone = 1, kNrRetry = 2, kNrSkipped = 3 };
// the same bounds on t = -log z (kTmax = -log kZmin, kTmin = -log kZmax).  Clamping t itself is what clamping z and taking the
// logarithm again does, without the exponential and the logarithm (3 k clk at the head of every branch pass) and without
// moving a length that is already inside the range by a rounding error.
constexpr double kTmax = 34.538776394910684, kTmin = 1.000000500029089e-06;
__host__ __device__ inline double nr_clamp_length(double t) { return t > kTmax ? kTmax : (t < kTmin ? kTmin : t); }
// z of a length, brought into the NR range (the device takes it while the pass is still running: it needs the length only)
__host__ __device__ inline double nr_z(double t) {
    const double z = exp(-t);
    return z < kZmin ? kZmin : (z > kZmax ? kZmax : z);
}
__host__ __device__ inline int nr_step_z(double z, double d1t, double d2t, double* t_new);
__host__ __device__ inline int nr_step(double t, double d1t, double d2t, double* t_new) { return nr_step_z(nr_z(t), d1t, d2t, t_new); }
__host__ __device__ inline int nr_step_z(double z, double d1t, double d2t, double* t_new) {
    const double d1 = -d1t, d2 = d2t;  // derivatives in lz = log z = -t
    if (d2 >= 0.0 && z < kZmax) {
        *t_new = -log(0.37 * z + 0.63);
        return kNrRetry;
    }
    const double zprev = z;
    if (d2 < 0.0) {
        const double step = -d1 / d2;
        if (step < 100.0) {
            z *= exp(step);
            z = z < kZmin ? kZmin : z;
            const double cap = 0.25 * zprev + 0.75;
            z = z > cap ? cap : z;
        } else {
            z = 0.25 * zprev + 0.75;
        }
    }
    z = z > kZmax ? kZmax : z;
    *t_new = -log(z);
    return kNrDone;
}

// ---- flag-in-data words ("LL" encoding) --------------------------------------------------------------------------------
// Every double that crosses a coherence domain without a fence (device -> mapped host memory, device -> peer device over
// NVLink) travels as TWO 8-byte words, each carrying 32 payload bits and a 32-bit flag derived from the pass's sequence
// number.  An aligned 8-byte store / load is single-copy atomic on both sides, so a word whose flag matches carries its
// payload: no fence, no reliance on 16-byte vector stores arriving whole (PTX models them as separate scalar accesses).
// The flag is never 0 (fresh memory) and repeats only after 2^31 passes, far beyond the depth of any ring it is stored in.
__host__ __device__ inline uint32_t ll_flag(double seq) { return ((uint32_t)(long long)seq & 0x7fffffffu) | 0x80000000u; }
__host__ __device__ inline unsigned long long ll_word(uint32_t payload, uint32_t flag) {
    return ((unsigned long long)flag << 32) | payload;
}
// host side: returns true and the value once both words of the pair at p carry `flag`
inline bool ll_try_read_host(const volatile unsigned long long* p, uint32_t flag, double* v) {
    const unsigned long long lo = p[0], hi = p[1];
    if ((uint32_t)(lo >> 32) != flag || (uint32_t)(hi >> 32) != flag) return false;
    const unsigned long long bits = (hi << 32) | (lo & 0xffffffffull);
    __builtin_memcpy(v, &bits, sizeof bits);
    return true;
}
#ifdef __CUDACC__
__device__ __forceinline__ void ll_store_sys(unsigned long long* p, double v, uint32_t flag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long lo = ll_word((uint32_t)bits, flag), hi = ll_word((uint32_t)(bits >> 32), flag);
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ bool ll_try_load_sys(const unsigned long long* p, uint32_t flag, double* v) {
    unsigned long long lo, hi;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    if ((uint32_t)(lo >> 32) != flag || (uint32_t)(hi >> 32) != flag) return false;
    *v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
    return true;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

// Where a branch pass leaves its outcome.  `slot` is mapped pinned host memory the host polls: five doubles in the LL
// encoding above (10 words):
//   0..2 = lnL, dlnL/dt, d2lnL/dt2 (summed over ranks), 3 = the branch length after the step, 4 = NrStatus.
// With `len` set the guarded NR step is done on the device and stored to *len unless *poison is set; a step that ends in
// kNrRetry sets *poison, so that work queued behind it cannot move other branches before the host has dealt with the retry.
// kNrCommLost: the sum over the ranks did not complete (a peer never delivered its part, see PeerReduce); the three sums are
// this rank's own, no step is taken, *poison is set, and the host reports PML_ECOMM.
constexpr int kSlotDoubles = 10;
enum : int { kNrCommLost = 4 };
struct Publish {
    double* slot;   // nullptr: nothing is published
    double seq;
    double* len;    // nullptr: no NR step
    int* poison;
};
// What the publishing thread can fetch while the pass is still running (after the dependency wait): the poison flag -- only a
// launch that has completed can have raised it -- and z of the length the sums are taken at.  Both used to sit, a dependent
// global load and an exponential, between the last CTA's sum and the store of the new length.
// ... and, since the step is taken in t = -log z (below), the three lengths it can end on without knowing the derivatives: the
// exponential and the logarithms of the guarded step leave the tail of the pass (2.8 k clk on one thread behind the last CTA's
// sum) for the time the tiles stream.
struct PublishEarly {
    int poison;
    double z;        // z of the length the sums are taken at
    double t;        // that length (already inside the NR range)
    double t_lo;     // -log min(0.25 z + 0.75, zmax): the shortest length the step may end on
    double t_retry;  // -log(0.37 z + 0.63): where the derivatives are taken again after a bad curvature
};
__host__ __device__ inline PublishEarly nr_early(double t) {
    PublishEarly e{0, nr_z(t), t, 0.0, 0.0};
    const double cap = 0.25 * e.z + 0.75;
    e.t_lo = -log(cap < kZmax ? cap : kZmax);
    e.t_retry = -log(0.37 * e.z + 0.63);
    return e;
}
// nr_step_z in t = -log z: z exp(step) is t - step, the clamps of z become clamps of t (same decisions; the new length differs
// from -log(z exp(step)) by rounding only)
__host__ __device__ inline int nr_step_t(const PublishEarly& e, double d1t, double d2t, double* t_new) {
    const double d1 = -d1t, d2 = d2t;  // derivatives in lz = log z = -t
    if (d2 >= 0.0 && e.z < kZmax) {
        *t_new = e.t_retry;
        return kNrRetry;
    }
    double t = e.t;
    if (d2 < 0.0) {
        const double step = -d1 / d2;
        if (step < 100.0) {
            t = e.t - step;
            t = t > kTmax ? kTmax : t;
            t = t < e.t_lo ? e.t_lo : t;
        } else {
            t = e.t_lo;
        }
    }
    *t_new = t;
    return kNrDone;
}
#ifdef __CUDACC__
__device__ __forceinline__ PublishEarly publish_prefetch(const Publish& pub, double t) {
    PublishEarly e{0, 0.0, t, 0.0, 0.0};
    if (pub.len) {
        e = nr_early(t);
        e.poison = *reinterpret_cast<volatile int*>(pub.poison);
    }
    return e;
}
// release / acquire around the ticket of the last-CTA sum (cheaper than the sequentially consistent __threadfence)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// tail of a branch pass on one thread: NR step (optional) and publication; r = {lnL, d1, d2}, t = the length they were taken at
__device__ __forceinline__ void publish_result(const Publish& pub, const double r[3], double t, bool comm_lost, const PublishEarly& early) {
    double t_new = t;
    int status = kNrNone;
    if (comm_lost) {
        status = kNrCommLost;
        if (pub.poison) *pub.poison = 1;
    } else if (pub.len) {
        if (early.poison) status = kNrSkipped;
        else {
            status = nr_step_t(early, r[1], r[2], &t_new);
            *pub.len = t_new;
            if (status == kNrRetry) *pub.poison = 1;
        }
    }
    if (pub.slot) {
        unsigned long long* w = reinterpret_cast<unsigned long long*>(pub.slot);
        const uint32_t flag = ll_flag(pub.seq);
        ll_store_sys(w + 0, r[0], flag);
        ll_store_sys(w + 2, r[1], flag);
        ll_store_sys(w + 4, r[2], flag);
        ll_store_sys(w + 6, t_new, flag);
        ll_store_sys(w + 8, (double)status, flag);
    }
}
#endif

// Sum of the three doubles of a branch pass over the ranks of a site-sharded group, done INSIDE the branch kernel's tail over
// NVLink peer memory instead of a separate NCCL launch (which costs ~20 us per branch visit in launch gaps and latency).
// Every rank owns a mailbox [kPeerRing][nranks][3] of LL pairs that all peers can write: the last CTA stores this rank's
// three sums into slot (seq mod kPeerRing, rank) of EVERY mailbox with system-scope stores, then polls its own mailbox until
// all nranks entries carry the flag of seq, and adds them in rank order -- identical bits on every rank.  Ranks run the
// same passes in lock step and can be at most one pass apart, so a ring of 4 is never overwritten early.
// The wait is BOUNDED: a peer that died, failed before its launch or fell out of step never delivers; after timeout_ns
// (globaltimer) the pass gives up, raises *lost (sticky: later passes do not wait at all) and reports kNrCommLost.
constexpr int kPeerRing = 4;
constexpr int kMaxPeers = 16;
struct PeerReduce {
    double* const* mail;  // device array of nranks mailbox pointers (own included); nullptr: single rank or NCCL path
    int rank, nranks;
    int* lost;                       // device flag, sticky
    unsigned long long timeout_ns;
};
#ifdef __CUDACC__
// called by one full warp; r3 = this rank's sums in, the group's sums out (same in every lane); false = gave up
__device__ __forceinline__ bool peer_allreduce3(const PeerReduce& pr, double seq, double r3[3]) {
    const int lane = threadIdx.x & 31;
    const size_t slot = (size_t)((long long)seq % kPeerRing) * pr.nranks;
    const uint32_t flag = ll_flag(seq);
    if (lane < pr.nranks) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(pr.mail[lane]) + (slot + pr.rank) * 6;
#pragma unroll
        for (int v = 0; v < 3; ++v) ll_store_sys(dst + 2 * v, r3[v], flag);
    }
    double mine[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    if (lane < pr.nranks) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(pr.mail[pr.rank]) + (slot + lane) * 6;
        const bool dead = *reinterpret_cast<volatile int*>(pr.lost) != 0;
        const unsigned long long t0 = global_timer_ns();
#pragma unroll
        for (int v = 0; v < 3 && ok; ++v) {
            unsigned spins = 0;
            while (!ll_try_load_sys(src + 2 * v, flag, &mine[v])) {
                if (dead || ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > pr.timeout_ns)) {
                    ok = false;
                    break;
                }
            }
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) {
        if (lane == 0) *pr.lost = 1;
        return false;
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        double acc = 0.0;
        for (int r = 0; r < pr.nranks; ++r) acc += __shfl_sync(0xffffffffu, mine[v], r);
        r3[v] = acc;
    }
    return true;
}
#endif

// One pass over the two ends of a branch (b inner, a inner or tip): lnL, dlnL/dt, d2lnL/dt2 at the given length
// (+ per-pattern lnL when site_lnl != nullptr, + the eigen-space product table when sumtable != nullptr).
struct BranchArgs {
    Side a, b;
    const DeviceModel* dm;
    const int32_t* weights;
    double t;             // branch length the sums are evaluated at ...
    const double* t_ptr;  // ... or, when set, where to read it on the device (clamped to the NR range)
    double* site_lnl;     // optional
    int want_lnl, want_derivs;  // which of the sums the caller will look at (the others come back as 0)
    double* sumtable;     // optional: np x 80
    int32_t* sum_scale;   // with sumtable
    double* partials;     // 3 x grid doubles
    unsigned int* ticket; // zero-initialised; the CTA drawing the last ticket adds the partials and resets it
    double* result;       // 4 doubles in device memory: lnL, d1, d2 of this rank's patterns (input of the allreduce), length used
    Publish pub;          // the kernel itself publishes (single rank, or ranks joined by `peer`); otherwise launch_publish
    PeerReduce peer;      // ranks > 1 with peer access: the sums are added over NVLink inside the kernel
    long long* trace;     // optional (profiling aid), as NewviewOp::trace
};
void launch_branch_mma(const BranchArgs& args, int64_t np, int sms, cudaStream_t stream);
void configure_branch_kernels();
// fused_mma.cu: the CLV update `op` (at least one inner child) and the branch pass between its result (the x end) and
// args.a (the y end, inner or tip) in one launch; args.b is ignored, args.sumtable must be null
void launch_fused(const NewviewOp& op, const BranchArgs& args, int64_t np, int sms, cudaStream_t stream);
void configure_fused_kernels();
// NR step + publication from result[0..3] (after the allreduce of result[0..2] when ranks > 1)
void launch_publish(const double* result, const Publish& pub, cudaStream_t stream);

// result[0..2] = lnL, dlnL/dt, d2lnL/dt2 at branch length *d_t (device scalar) from a sumtable
void launch_core(const DeviceModel* dm, const double* sumtable, const int32_t* sum_scale, const int32_t* weights, int64_t np, double t,
                 double* partials, double* result, cudaStream_t stream);

// lnl[r] = sum_p W[r][p] * site_lnl[p]
void launch_replicate_lnl(const int32_t* W, int nrep, int64_t np, int64_t ldw, const double* site_lnl, double* lnl,
                          cudaStream_t stream);

// one Fitch parsimony scan over this rank's patterns (parsimony.cu); out[0] += score, out[1 + i] += cost of attaching
// next_taxon above pre[i] (all weighted, exact integers); down / up: [nnodes][npad] scratch
struct ParsimonyArgs {
    const uint8_t* codes;     // [ntax][npad]
    const int32_t* weights;   // [npad]
    int64_t npad, nloc;
    const int4* nodes;        // per node: left, right, taxon, parent
    const int* pre;           // nodes below the root tip, parents first
    int npre, top, root_taxon, next_taxon;
    uint32_t* down;
    uint32_t* up;
    unsigned long long* out;  // 1 + npre, zeroed by the caller
};
void launch_parsimony_scan(const ParsimonyArgs& a, cudaStream_t stream);

int64_t reduce_partials_capacity(int64_t np);  // doubles needed in `partials`

}  // namespace pml
