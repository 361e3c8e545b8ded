// Shared device helpers of the tensor-path kernels: mbarrier / TMA bulk-copy wrappers, the DMMA m8n8k4 wrapper and the
// fragment conventions (which state a lane feeds into which k-tile).
//
// CLV layout in HBM ("blocked"): patterns are grouped in blocks of 8 rows; one block is 640 contiguous doubles laid out
// [category c][chunk], where the chunks of a category are  states 0-7 as 8 rows x 8,  states 8-15 as 8 rows x 8,
// states 16-19 as 8 rows x 4  (160 doubles per category).  With lane = 4*g + t (g = row in block) this makes
//   - a 16-row tile one contiguous 10,240 B piece: ONE TMA bulk copy per child and stage, no row padding;
//   - the A-fragment reads (states {2t,2t+1}, {8+2t,9+2t}, {16+t}) two 128-bit and one 64-bit shared-memory loads whose
//     32 lanes touch consecutive addresses -> conflict free;
//   - the D-fragment writes (states nt*8+2t+{0,1}) 128-bit stores to consecutive addresses -> fully coalesced.
// Only the engine's own kernels ever read CLVs, so the layout is private to this directory.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace pml {
namespace mma {

constexpr int kTileRows = 16;                // patterns per pipeline stage (one stage = one group iteration)
constexpr int kBlockRows = 8;                // rows of one layout block = rows of one MMA m-tile
constexpr int kBlockDoubles = kBlockRows * kRow;  // 640
constexpr int kCatDoubles = kBlockRows * kStates; // 160
constexpr int kTileDoubles = kTileRows * kRow;    // one child's share of a stage
constexpr int kTipPad = 82;                  // doubles per padded tip-table row
// warp roles of the streaming kernels (newview_mma.cu, branch_mma.cu): 8 MMA warps in two groups of four (one warp per
// rate category), 2 epilogue / finishing warps, 1 producer warp
constexpr int kMmaGroups = 2;
constexpr int kMmaWarps = 4 * kMmaGroups;
constexpr int kEpiWarps = 2;
constexpr int kProducerWarp = kMmaWarps + kEpiWarps;
constexpr int kDepth = 3;                    // stages in each group's private input ring (newview)
constexpr double kTwo256 = 1.157920892373161954235709850086879078532699846656405640394575840079131296399e77;
constexpr double kMinLik = 8.636168555094444625386351862800399571116000364436281385023703470168591803162e-78;
constexpr int kMinLikHi = 0x2FF00000;        // high word of 2^-256 (biased exponent 1023 - 256 = 0x2FF, mantissa 0)

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
__device__ __forceinline__ void named_barrier_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Programmatic dependent launch: the streaming kernels are launched with programmatic stream serialisation, so a kernel's
// CTAs may start (barrier set-up, loads of the static model constants) while the previous kernel on the stream is still
// draining; pdl_wait() returns once that kernel has completed and its writes are visible.  Every kernel triggers its own
// dependents right away -- they cannot take an SM before this grid's CTA there has exited (shared memory), so the trigger
// only removes the launch latency between two dependent kernels.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename Kernel, typename... Args>
inline void launch_pdl(Kernel kernel, int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

// MMA turn-taking between the two groups of a CTA.  The FP64 tensor pipe is one unit per SM (4 clk per DMMA.8x8x4) and ONE
// group -- four warps, one per SM sub-partition, each issuing a DMMA every 16 clk -- already saturates it.  Left to the
// round-robin warp scheduler all MMA warps crawl through their bursts together and then do their other work together,
// leaving the pipe idle (measured: 56 % busy).  With the token below the bursts of the groups follow each other and
// every group's loads, products and stores run under the other group's MMAs (the "ping-pong" schedule of warp-specialised
// GEMMs).  Barrier kTurnBarrier + g is shared by group g (sync) and the other group (arrive).
constexpr int kTurnBarrier = 4;
constexpr int kTurnThreads = 2 * 4 * 32;
__device__ __forceinline__ void mma_turn_begin(int grp) { named_barrier(kTurnBarrier + grp, kTurnThreads); }
__device__ __forceinline__ void mma_turn_end(int grp) { named_barrier_arrive(kTurnBarrier + (grp + 1) % kMmaGroups, kTurnThreads); }
__device__ __forceinline__ void mma_turn_init(int grp) {
    if (grp == kMmaGroups - 1) named_barrier_arrive(kTurnBarrier, kTurnThreads);  // group 0 may start
}

// TMA engine bulk copy shared -> global (bulk async-group completion) and the proxy fence a generic-proxy writer needs
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// state index that lane t feeds into k-tile kt: pairs of k-tiles share one 128-bit shared-memory load
__device__ __forceinline__ int kmap(int kt, int t) { return kt < 4 ? (kt >> 1) * 8 + 2 * t + (kt & 1) : 16 + t; }

// B fragments of P_c^T for one child from P_c[i][j] (row-major 20 x 20, shared memory):
// frag[nt][kt] = P_c[nt*8 + g][kmap(kt, t)] (0 beyond state 19)
__device__ __forceinline__ void load_p_fragments(const double* Pc, int g, int t, double (&frag)[3][5]) {
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int i = nt * 8 + g;
#pragma unroll
        for (int kt = 0; kt < 5; ++kt) frag[nt][kt] = i < kStates ? Pc[i * kStates + kmap(kt, t)] : 0.0;
    }
}

// A fragments of one 8-row block: the lane's five state values x[row g][c][kmap(kt, t)]
struct AFrag {
    double v[5];
};
__device__ __forceinline__ AFrag load_a(const double* block, int c, int lane) {
    const double* x = block + c * kCatDoubles;
    const double2 a01 = reinterpret_cast<const double2*>(x)[lane];
    const double2 a23 = reinterpret_cast<const double2*>(x + 64)[lane];
    AFrag f;
    f.v[0] = a01.x;
    f.v[1] = a01.y;
    f.v[2] = a23.x;
    f.v[3] = a23.y;
    f.v[4] = x[128 + lane];
    return f;
}
// ---- folded cherries (Side in kernels.h) ---------------------------------------------------------------------------------
// The CLV a tip-tip update would have stored is the product of the two tips' look-up entries (23 x 80 tables, rows padded to
// kTipPad doubles).  A consumer forms the five state values a lane feeds into the k-tiles JUST IN TIME inside its own MMA
// turn: the raw table entries of the first k-tile pair are fetched before the turn, every later pair while the DMMAs of the
// previous one issue, and the multiplications sit between this warp's own DMMAs.  (Formed ahead of the turn, those FP64
// multiplications queued behind the other group's DMMAs: +300 clk on every tile's critical path.)
struct CherryRows {
    const double* a[2];   // per 8-row block m: row of the first / second tip's look-up for this lane's pattern, category applied
    const double* b[2];
};
__device__ __forceinline__ CherryRows cherry_rows(const double* tab_a, const double* tab_b, int code_a0, int code_b0, int code_a1, int code_b1, int c) {
    CherryRows r;
    r.a[0] = tab_a + code_a0 * kTipPad + c * kStates;
    r.b[0] = tab_b + code_b0 * kTipPad + c * kStates;
    r.a[1] = tab_a + code_a1 * kTipPad + c * kStates;
    r.b[1] = tab_b + code_b1 * kTipPad + c * kStates;
    return r;
}
// raw entries of block m, chunk ch (k-tiles 2ch, 2ch+1: states {2t, 2t+1} of the first / second octet; chunk 2: state 16 + t)
struct CherryRaw {
    double2 a, b;
};
__device__ __forceinline__ CherryRaw cherry_fetch(const CherryRows& r, int m, int ch, int t) {
    CherryRaw v;
    if (ch < 2) {
        v.a = *reinterpret_cast<const double2*>(r.a[m] + 8 * ch + 2 * t);
        v.b = *reinterpret_cast<const double2*>(r.b[m] + 8 * ch + 2 * t);
    } else {
        v.a = make_double2(r.a[m][16 + t], 0.0);
        v.b = make_double2(r.b[m][16 + t], 0.0);
    }
    return v;
}
// what a cherry side carries into the turn: its rows and the raw entries of the first two chunks (block 0), fetched before the turn
struct CherryIn {
    CherryRows rows;
    CherryRaw r0, r1;
};
__device__ __forceinline__ CherryIn cherry_begin(const double* tab_a, const double* tab_b, const unsigned char* codes_a, const unsigned char* codes_b,
                                                 int g, int c, int t) {
    CherryIn in;
    in.rows = cherry_rows(tab_a, tab_b, codes_a[g], codes_b[g], codes_a[8 + g], codes_b[8 + g], c);
    in.r0 = cherry_fetch(in.rows, 0, 0, t);
    in.r1 = cherry_fetch(in.rows, 0, 1, t);
    return in;
}
constexpr int kSideInnerK = 0, kSideTipK = 1, kSideCherryK = 2;  // = SideKind (kernels.h)

// Where the element-wise product of the two children is taken (12 FP64 multiplications per warp and tile).  Issued after the
// turn has been handed over they queue behind the other group's DMMAs (~28 clk each, in-order issue: +300 clk on the group's
// cycle); issued inside the turn they cost the tensor pipe the latency of the last DMMAs.  kProductsInTurn: block 0's products
// go out under block 1's DMMAs, block 1's right behind its last DMMA, still inside the turn.
#ifndef PML_PRODUCTS_IN_TURN
#define PML_PRODUCTS_IN_TURN 2
#endif

// The MMA turn of a CLV update, block-major: acc?[m][nt] += A?[m][kt] * frag?[nt][kt] over the five k-tiles for both children, then
// (PML_PRODUCTS_IN_TURN) accL[m] *= accR[m].  A child that is an inner node brings its A fragments (a?), a tip takes no part (its
// look-up rows already sit in the accumulators), a cherry its rows and first raw entries (c?).  A cherry's products for a
// chunk are issued one chunk AHEAD of the DMMAs that use them, its raw entries fetched two chunks ahead.
// Returns the blocks whose products have been taken (bit m).
struct NoExtra {
    __device__ __forceinline__ void operator()(int) const {}
};
// extra(kt) is called ahead of the DMMAs of k-tile kt of block 0 (the fused kernel slips the previous tile's contraction in there)
template <int KL, int KR, typename Extra = NoExtra>
__device__ __forceinline__ int children_mma(const AFrag (&aL)[2], const AFrag (&aR)[2], const CherryIn& cL, const CherryIn& cR,
                                            const double (&fragL)[3][5], const double (&fragR)[3][5], double (&accL)[2][3][2],
                                            double (&accR)[2][3][2], int t, Extra extra = Extra()) {
    constexpr bool kCL = KL == kSideCherryK, kCR = KR == kSideCherryK;
    CherryRaw rawL = cL.r1, rawR = cR.r1;       // raw entries of the chunk after the next one to be multiplied
    double vL[2] = {0.0, 0.0}, vR[2] = {0.0, 0.0};  // products of the chunk whose DMMAs come next
    if (kCL) {
        vL[0] = cL.r0.a.x * cL.r0.b.x;
        vL[1] = cL.r0.a.y * cL.r0.b.y;
    }
    if (kCR) {
        vR[0] = cR.r0.a.x * cR.r0.b.x;
        vR[1] = cR.r0.a.y * cR.r0.b.y;
    }
    int done = 0;
#pragma unroll
    for (int step = 0; step < 6; ++step) {  // (block, chunk) pairs in order
        const int m = step / 3, ch = step % 3;
        // products of the NEXT chunk (their raw entries arrived a chunk ago), then the fetch for the one after it
        double nL[2] = {0.0, 0.0}, nR[2] = {0.0, 0.0};
        if (step < 5) {
            const int ch1 = (step + 1) % 3;
            if (kCL) {
                nL[0] = rawL.a.x * rawL.b.x;
                if (ch1 < 2) nL[1] = rawL.a.y * rawL.b.y;
            }
            if (kCR) {
                nR[0] = rawR.a.x * rawR.b.x;
                if (ch1 < 2) nR[1] = rawR.a.y * rawR.b.y;
            }
        }
        if (step < 4) {
            const int m2 = (step + 2) / 3, ch2 = (step + 2) % 3;
            if (kCL) rawL = cherry_fetch(cL.rows, m2, ch2, t);
            if (kCR) rawR = cherry_fetch(cR.rows, m2, ch2, t);
        }
#pragma unroll
        for (int kk = 0; kk < (ch < 2 ? 2 : 1); ++kk) {
            const int kt = 2 * ch + kk;
            if (m == 0) extra(kt);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                if (KL != kSideTipK) dmma(accL[m][nt][0], accL[m][nt][1], kCL ? vL[kk] : aL[m].v[kt], fragL[nt][kt]);
                if (KR != kSideTipK) dmma(accR[m][nt][0], accR[m][nt][1], kCR ? vR[kk] : aR[m].v[kt], fragR[nt][kt]);
            }
        }
        vL[0] = nL[0];
        vL[1] = nL[1];
        vR[0] = nR[0];
        vR[1] = nR[1];
        if (PML_PRODUCTS_IN_TURN >= 1 && step == 3) {  // block 0 has long left the pipe
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                accL[0][nt][0] *= accR[0][nt][0];
                accL[0][nt][1] *= accR[0][nt][1];
            }
            done |= 1;
        }
    }
    if (PML_PRODUCTS_IN_TURN >= 2) {
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            accL[1][nt][0] *= accR[1][nt][0];
            accL[1][nt][1] *= accR[1][nt][1];
        }
        done |= 2;
    }
    return done;
}
// D fragments (states nt*8 + 2t + {0,1}; nt = 2 only for t < 2) of one 8-row block back into the blocked layout
__device__ __forceinline__ void store_d(double* block, int c, int lane, const double (&d)[3][2]) {
    double* x = block + c * kCatDoubles;
    reinterpret_cast<double2*>(x)[lane] = make_double2(d[0][0], d[0][1]);
    reinterpret_cast<double2*>(x + 64)[lane] = make_double2(d[1][0], d[1][1]);
    if ((lane & 3) < 2) *reinterpret_cast<double2*>(x + 128 + (lane >> 2) * 4 + 2 * (lane & 3)) = make_double2(d[2][0], d[2][1]);
}

}  // namespace mma
}  // namespace pml
