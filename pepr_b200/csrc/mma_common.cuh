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
// A fragments of a cherry (Side in kernels.h): the lane's five state values of row g are products of the two tips' look-up
// entries (23 x 80 tables, rows padded to kTipPad doubles) -- the CLV a tip-tip update would have stored, never materialised
__device__ __forceinline__ AFrag cherry_a(const double* tab_a, const double* tab_b, int code_a, int code_b, int c, int t) {
    const double* ra = tab_a + code_a * kTipPad + c * kStates;
    const double* rb = tab_b + code_b * kTipPad + c * kStates;
    const double2 a0 = *reinterpret_cast<const double2*>(ra + 2 * t), b0 = *reinterpret_cast<const double2*>(rb + 2 * t);
    const double2 a1 = *reinterpret_cast<const double2*>(ra + 8 + 2 * t), b1 = *reinterpret_cast<const double2*>(rb + 8 + 2 * t);
    AFrag f;
    f.v[0] = a0.x * b0.x;
    f.v[1] = a0.y * b0.y;
    f.v[2] = a1.x * b1.x;
    f.v[3] = a1.y * b1.y;
    f.v[4] = ra[16 + t] * rb[16 + t];
    return f;
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
