// Fitch parsimony scan of one stepwise-addition step on the device (raxmlHPC makeParsimonyTree's inner loops; host control
// flow in host.cpp: parsimony_start_tree).  One thread per alignment pattern walks the current tree: children-first for
// the state sets below every node (and the tree's score), parents-first for the sets above, and right there the cost of
// attaching the next taxon to the branch above each node.  Sets are 20-bit masks in 32-bit words, [node][pattern] in HBM,
// so every access of a warp is one coalesced 128-byte line; the per-node costs are integer sums (bit-exact against the
// host implementation): warp reduction, then one 64-bit atomic per warp and node.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace pml {

namespace {

__device__ __forceinline__ uint32_t code_mask(int code) {
    if (code < 20) return 1u << code;
    if (code == 20) return (1u << 2) | (1u << 3);
    if (code == 21) return (1u << 5) | (1u << 6);
    return 0xFFFFFu;
}

__global__ void __launch_bounds__(256) k_parsimony_scan(ParsimonyArgs a) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < a.nloc;
    const int64_t col = live ? s : 0;
    const unsigned w = live ? (unsigned)a.weights[col] : 0u;
    const int lane = threadIdx.x & 31;
    unsigned score = 0;
    // children first: the list is parents-first, so walk it backwards
    for (int i = a.npre - 1; i >= 0; --i) {
        const int v = a.pre[i];
        const int4 nd = a.nodes[v];  // left, right, taxon, parent
        uint32_t d;
        if (nd.x < 0) d = code_mask(a.codes[(int64_t)nd.z * a.npad + col]);
        else {
            const uint32_t A = a.down[(int64_t)nd.x * a.npad + col], B = a.down[(int64_t)nd.y * a.npad + col];
            d = A & B;
            if (!d) {
                d = A | B;
                score += w;
            }
        }
        a.down[(int64_t)v * a.npad + col] = d;
    }
    const uint32_t rootmask = code_mask(a.codes[(int64_t)a.root_taxon * a.npad + col]);
    if (!(a.down[(int64_t)a.top * a.npad + col] & rootmask)) score += w;
    score = __reduce_add_sync(0xffffffffu, score);
    if (lane == 0 && score) atomicAdd(a.out, (unsigned long long)score);
    if (a.next_taxon < 0) return;
    const uint32_t X = code_mask(a.codes[(int64_t)a.next_taxon * a.npad + col]);
    // parents first: the set above a node comes from its parent's set above and its sibling's set below
    for (int i = 0; i < a.npre; ++i) {
        const int v = a.pre[i];
        const int4 nd = a.nodes[v];
        uint32_t U;
        if (v == a.top) U = rootmask;
        else {
            const int4 pn = a.nodes[nd.w];
            const int sib = pn.x == v ? pn.y : pn.x;
            const uint32_t PU = a.up[(int64_t)nd.w * a.npad + col], S = a.down[(int64_t)sib * a.npad + col];
            U = PU & S;
            if (!U) U = PU | S;
        }
        if (nd.x >= 0) a.up[(int64_t)v * a.npad + col] = U;  // only inner nodes are somebody's parent
        const uint32_t D = a.down[(int64_t)v * a.npad + col];
        uint32_t e = U & D;
        if (!e) e = U | D;
        unsigned c = (e & X) ? 0u : w;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(a.out + 1 + i, (unsigned long long)c);
    }
}

}  // namespace

void launch_parsimony_scan(const ParsimonyArgs& a, cudaStream_t stream) {
    const int grid = (int)((a.nloc + 255) / 256);
    if (grid > 0) k_parsimony_scan<<<grid, 256, 0, stream>>>(a);
}

}  // namespace pml
