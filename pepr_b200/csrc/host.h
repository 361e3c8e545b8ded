// Host-side, device-independent pieces of the engine: alignment pattern crunch, unrooted tree + newick I/O,
// traversal planning with CLV orientation tracking, bootstrap weight stream, bipartition support counting.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace pml {

// ---- alignment -------------------------------------------------------------------------------------------
struct Patterns {
    int ntax = 0;
    int64_t nsites = 0, npat = 0;
    std::vector<std::string> names;
    std::vector<uint8_t> codes;        // ntax x npat, row-major: residue codes of each pattern
    std::vector<int32_t> weight;       // npat
    std::vector<int64_t> site_to_pat;  // nsites, -1 for dropped columns
};
// column sort + duplicate merge in the reference's order (raxmlHPC sitesort/sitecombcrunch: lexicographic by taxon row)
void crunch_patterns(int ntax, int64_t nsites, const uint8_t* chars, const int32_t* site_w, Patterns& out);
bool read_phylip(const std::string& path, std::vector<std::string>& names, std::vector<uint8_t>& chars, int64_t& nsites,
                 std::string& err);

// ---- tree ------------------------------------------------------------------------------------------------
// Unrooted binary tree. Nodes 0..ntax-1 are tips (index = taxon index), ntax..2ntax-3 are inner nodes.
struct Topology {
    int ntax = 0;
    std::vector<std::array<int, 3>> nbr;   // neighbour node per slot (-1 unused; tips use slot 0 only)
    std::vector<std::array<int, 3>> edge;  // branch id per slot
    std::vector<int> ea, eb;               // endpoints of each branch
    std::vector<double> len;               // expected substitutions per site
    int nnodes() const { return 2 * ntax - 2; }
    int nedges() const { return 2 * ntax - 3; }
    bool is_tip(int v) const { return v < ntax; }
    int slot_of(int v, int neighbour) const {
        for (int s = 0; s < 3; ++s)
            if (nbr[v][s] == neighbour) return s;
        return -1;
    }
};
bool parse_newick(const std::string& text, const std::vector<std::string>& names, double default_len, Topology& out,
                  std::string& err);
// RAxML_result style: trifurcation at the inner node adjacent to taxon 0, taxon 0 printed last, 20 decimals, ":0.0;"
std::string write_newick_result(const Topology& t, const std::vector<std::string>& names);

// one entry of a traversal descriptor: recompute the CLV of inner node `node` looking towards slot `toward`
struct ViewOp {
    int node, toward;
    int child[2];   // node ids (tip or inner)
    int cedge[2];   // branch ids
};
// CLV orientation book-keeping (what raxmlHPC tracks with its x-flags): orient[v - ntax] = slot the stored CLV of v
// faces, -1 = stale.
struct ViewState {
    std::vector<int> orient;
    void reset(const Topology& t) { orient.assign(t.ntax - 2 > 0 ? t.ntax - 2 : 0, -1); }
    // appends the ops needed so that node v holds a CLV summarising everything except the subtree behind `toward`
    void plan(const Topology& t, int v, int toward_node, std::vector<ViewOp>& ops);
    // a branch length changed: every stored CLV whose subtree contains that branch becomes stale
    void branch_changed(const Topology& t, int e);
};

// ---- bootstrap weights -----------------------------------------------------------------------------------
double randum(int64_t* seed);  // raxmlHPC randum(): 36-bit LCG in three 12-bit limbs
void bootstrap_replicates(int64_t* seed, const std::vector<int32_t>& pattern_weight, int nrep, int32_t* out);

// ---- support ---------------------------------------------------------------------------------------------
// returns "" and sets err on failure.  counts (optional) receives one entry per labelled inner node, output order.
std::string support_tree(const std::string& main_newick, const std::vector<std::string>& trees, bool as_percent,
                         std::vector<int32_t>* counts, std::string& err);

}  // namespace pml
