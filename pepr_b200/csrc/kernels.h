// Device-side interface of the engine: every function enqueues work on `stream` and returns immediately.
// One "PBlock" per (traversal entry, child) carries the four 20x20 transition matrices of that branch and, for tip
// children, the 23 x 80 lookup of P applied to each residue code's indicator vector.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "model.h"

namespace pml {

struct PBlock {
    double P[kCats][kStates][kStates];  // P_c(i->j)
    double tip[kCodes][kRow];           // tip[code][c*20+i] = sum_j P_c[i][j] * indicator(code)[j]
};

// model constants in device memory, one copy per alignment (eigenvalues, V, Vinv, pi, rates)
struct DeviceModel {
    double lambda[kStates];
    double V[kStates][kStates];
    double Vinv[kStates][kStates];
    double pi[kStates];
    double rates[kCats];
    double piV[kStates][kStates];   // pi_i * V[i][k]
};

// one side of a branch: an inner node's CLV (+ cumulative scaling counts) or a tip's residue codes
struct Side {
    const double* clv;      // nullptr for a tip
    const int32_t* scale;   // nullptr for a tip
    const uint8_t* codes;   // tip residue codes (0..22), nullptr for inner
};

struct NewviewOp {
    Side left, right;
    const PBlock* pleft;
    const PBlock* pright;
    double* out;
    int32_t* out_scale;
    long long* trace;  // optional (profiling aid): per warp of CTA 0, cycles spent in each phase of the pipeline
};

// P(t) for `nblocks` branches: lengths[b] in expected substitutions per site; tips[b] != 0 also fills PBlock::tip
// small batches travel as kernel arguments (no staging copy): up to kMakePInline branches
constexpr int kMakePInline = 32;
struct MakePInline {
    double length[kMakePInline];
    uint8_t want_tip[kMakePInline];
};
void launch_make_p_inline(const DeviceModel* dm, const MakePInline& batch, PBlock* d_blocks, int nblocks, cudaStream_t stream);
void launch_make_p(const DeviceModel* dm, const double* d_lengths, const uint8_t* d_want_tip, PBlock* d_blocks, int nblocks, cudaStream_t stream);

// CLV update on the FP64 tensor path (TMA-fed DMMA); np must be a multiple of 64 and all buffers hold np rows
void launch_newview_mma(const NewviewOp& op, int64_t np, int sms, cudaStream_t stream);
// raises the dynamic shared-memory limit of the tensor-path kernels on the current device (once per context)
void configure_mma_kernels();

// One pass over the two ends of a branch (b inner, a inner or tip): lnL, dlnL/dt, d2lnL/dt2 at length *d_t as CTA partials
// (+ per-pattern lnL when site_lnl != nullptr, + the eigen-space product table when sumtable != nullptr).
struct BranchArgs {
    Side a, b;
    const DeviceModel* dm;
    const int32_t* weights;
    double t;             // branch length the sums are evaluated at
    double* site_lnl;     // optional
    double* sumtable;     // optional: np x 80
    int32_t* sum_scale;   // with sumtable
    double* rowsum;       // np x 3 scratch: f, f', f'' per pattern
    double* partials;     // 3 x grid doubles
    unsigned int* ticket; // zero-initialised; the CTA drawing the last ticket adds the partials and resets it
    double* result;       // 3 doubles in device memory (input of the NCCL allreduce when ranks > 1)
    volatile double* host_result;  // optional: 4 doubles in mapped pinned memory; [3] receives `sequence` after [0..2]
    double sequence;
};
void launch_branch_mma(const BranchArgs& args, int64_t np, int sms, cudaStream_t stream);
void configure_branch_kernels();
// fixed-order sum of `nblocks` partials for each of `nvals` values
void launch_reduce(const double* partials, int nblocks, int nvals, double* result, cudaStream_t stream);

// result[0..2] = lnL, dlnL/dt, d2lnL/dt2 at branch length *d_t (device scalar) from a sumtable
void launch_core(const DeviceModel* dm, const double* sumtable, const int32_t* sum_scale, const int32_t* weights, int64_t np, double t,
                 double* partials, double* result, cudaStream_t stream);

// lnl[r] = sum_p W[r][p] * site_lnl[p]
void launch_replicate_lnl(const int32_t* W, int nrep, int64_t np, int64_t ldw, const double* site_lnl, double* lnl,
                          cudaStream_t stream);

int64_t reduce_partials_capacity(int64_t np);  // doubles needed in `partials`

}  // namespace pml
