// Host-side, device-independent pieces of the engine: alignment pattern crunch, unrooted tree + newick I/O,
// traversal planning with CLV orientation tracking, bootstrap weight stream, bipartition support counting.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

namespace pml {

// ---- alignment -------------------------------------------------------------------------------------------
struct Patterns {
    int ntax = 0;
    int64_t nsites = 0, npat = 0;
    std::vector<std::string> names;
    std::vector<uint8_t> codes;        // ntax x codes_n, row-major: residue codes of patterns [codes_p0, codes_p0 + codes_n)
    int64_t codes_p0 = 0, codes_n = 0; // (all patterns unless a rank asked for its own block only)
    std::vector<int32_t> weight;       // npat
    std::vector<int64_t> site_to_pat;  // nsites, -1 for dropped columns
    std::vector<int64_t> first;        // npat: a representative column of every pattern
    // A rank that shares the sort with others (crunch_patterns(..., lazy)) finishes only ITS pattern block at load time:
    // own_weight / own_first cover patterns [codes_p0, codes_p0 + codes_n), weight / site_to_pat / first stay empty (full ==
    // false) until somebody asks for them (finish_patterns: per-site output, bootstrap weights, pml_aln_patterns) -- the passes
    // over ALL columns are what every one of N processes would otherwise repeat at every load.
    bool full = true;
    std::vector<int32_t> own_weight;
    std::vector<int64_t> own_first;
    std::vector<int64_t> sorted_cols;   // kept for finish_patterns
    std::vector<uint8_t> sorted_fresh;
    std::vector<int32_t> col_weight;    // the caller's column weights (empty: all 1)
};
// column sort + duplicate merge in the reference's order (raxmlHPC sitesort/sitecombcrunch: lexicographic by taxon row)
// rank / nranks: keep the residue codes of that rank's contiguous pattern block only (the sort itself is always global)
// share (optional, nranks > 1): the ranks split the radix sort between them -- after the common first level every rank
// sorts only the first-residue buckets assigned to it (largest bucket first to the least loaded rank), leaves zeros
// elsewhere, and share(order, fresh, n) must return with the element-wise SUM over all ranks in both arrays (the engine
// does it with one NCCL allreduce each); everything after the sort is cheap and identical on every rank.
using CrunchShare = std::function<bool(int64_t* order, uint8_t* fresh, int64_t n)>;
void crunch_patterns(int ntax, int64_t nsites, const uint8_t* chars, const int32_t* site_w, Patterns& out, int rank = 0,
                     int nranks = 1, const CrunchShare* share = nullptr, bool lazy = false);
// fills weight / site_to_pat / first of a lazily crunched alignment (no-op when already full)
void finish_patterns(Patterns& p);
// the code rows of patterns [p0, p0 + n) of an alignment crunched elsewhere (first: representative column per pattern)
void gather_codes(int ntax, int64_t nsites, const uint8_t* chars, const std::vector<int64_t>& first, int64_t p0, int64_t n,
                  std::vector<uint8_t>& codes);
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
    // Cherry folding: the view of an inner node whose two other neighbours are tips is never stored -- every consumer forms it
    // from the two tips (Side in kernels.h) -- so planning emits no op for it and it can never be stale.
    bool fold_cherries = false;
    static bool is_cherry_view(const Topology& t, int v, int toward_node) {
        if (t.is_tip(v)) return false;
        int tips = 0, others = 0;
        for (int s = 0; s < 3; ++s) {
            const int nb = t.nbr[v][s];
            if (nb == toward_node) continue;
            ++others;
            if (nb >= 0 && t.is_tip(nb)) ++tips;
        }
        return others == 2 && tips == 2;
    }
    // A folded view still counts as one CLV update whenever a stored one would have been recomputed: fresh[v] is what
    // orient[v] == slot would say if the cherry were stored (cleared by branch_changed / topology edits), `folded` counts the
    // updates that planning skipped.
    std::vector<char> fresh;
    int64_t folded = 0;
    void touch(const Topology& t, int v) {
        if (v >= t.ntax && (size_t)(v - t.ntax) < fresh.size()) fresh[v - t.ntax] = 0;
    }
    void reset(const Topology& t) {
        orient.assign(t.ntax - 2 > 0 ? t.ntax - 2 : 0, -1);
        fresh.assign(orient.size(), 0);
    }
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

namespace pml {

// ---- topology editing (subtree pruning and regrafting) ------------------------------------------------------
// Moves inner node p together with the subtree behind its neighbour s: the other two neighbours q and r of p are joined
// by one branch (length = sum), and p is inserted into `target` (each half gets half of its length).  Views whose subtree
// changed are marked stale.  `undo` restores node links, branch ids and lengths exactly.
struct SprMove {
    int p = -1, s = -1, q = -1, r = -1, a = -1, b = -1;
    int e_s = -1, e_q = -1, e_r = -1, e_t = -1;
    int slot_q = -1, slot_r = -1;  // p's own link slots (q and r may coincide with a or b, so p is never searched by value)
    double len_q = 0, len_r = 0, len_t = 0, len_s = 0;  // lengths before the move (len_s: callers polish e_s after applying)
};
bool spr_apply(Topology& T, ViewState& V, int p, int s, int target, SprMove& mv);
void spr_undo(Topology& T, ViewState& V, const SprMove& mv);
// Pruned state for a scan over regraft points: p (with the subtree behind s) is detached, q -- r joined (lengths added), and
// the remaining tree stays put while candidates are scored by a virtual insertion (engine.cu: score_candidates).  Only
// traversal planning and branch orientation may be used on the tree until spr_unprune has restored it.
bool spr_prune(Topology& T, ViewState& V, int p, int s, SprMove& mv);
void spr_unprune(Topology& T, ViewState& V, const SprMove& mv);
// branches of the tree that remains after pruning (p, s), at 1..radius steps from the pruning point, excluding the subtree
// behind s and the two branches next to p (re-inserting there gives the same topology)
std::vector<int> spr_targets(const Topology& T, int p, int s, int radius);

// ---- parsimony starting tree ---------------------------------------------------------------------------------
// randomised stepwise addition under Fitch parsimony on the weighted patterns (raxmlHPC makeParsimonyTree's role);
// taxa are added in a random order drawn from randum(seed); all branch lengths are set to default_len.
// The tree grows rooted at the first taxon: node 0 = root tip; every other node has a parent; inner nodes have two children.
struct GrowTree {
    std::vector<int> parent, left, right, taxon;
    int add_node(int tx) {
        parent.push_back(-1);
        left.push_back(-1);
        right.push_back(-1);
        taxon.push_back(tx);
        return (int)parent.size() - 1;
    }
};
// One addition step over all patterns: `pre` lists the nodes below the root tip parents-first.  Returns the weighted Fitch
// score of the current tree and, when next_taxon >= 0, cost[i] = weighted number of patterns that gain a change if
// next_taxon is attached to the branch above pre[i].  The engine supplies a GPU implementation (csrc/parsimony.cu).
using ParsimonyScan = std::function<bool(const GrowTree& g, const std::vector<int>& pre, int next_taxon, int64_t& score, std::vector<int64_t>& cost)>;

// ---- topological constraints (FastTree -constraints, FastTreeRunner.java:53-83, 243-273) --------------------------------------
// A constraint is a split of (some of) the taxa: the tree must have a branch with all `one` taxa on one side and all `zero`
// taxa on the other; taxa in neither set are free.  Bits are taxon indices of the alignment.
struct SplitConstraint {
    std::vector<uint64_t> one, zero;
};
struct Constraints {
    int ntax = 0;
    std::vector<SplitConstraint> splits;   // only splits that can fail (two or more taxa on either side)
    bool empty() const { return splits.empty(); }
};
// FastTree's constraint alignment: ">name" lines followed by a row of 0 / 1 / - (one column per split); names must be taxa of the
// alignment, taxa that are not listed are unconstrained
bool parse_constraints(const std::string& text, const std::vector<std::string>& names, Constraints& out, std::string& err);
// the constraint alignment of a tree as FastTreeRunner.getFastTreeConstraintsForTree writes it: taxa sorted, one column per
// node of the tree (1 = the taxon descends from that node)
std::string constraints_from_tree(const std::string& newick, std::string& err);
bool satisfies(const Topology& T, const Constraints& C);
// stepwise addition: may next_taxon be attached to the branch above pre[i] without making a split impossible among the taxa
// present afterwards?
std::vector<char> allowed_insertions(const GrowTree& g, const std::vector<int>& pre, int next_taxon, const Constraints& C);

bool parsimony_start_tree(const Patterns& pat, int64_t seed, double default_len, Topology& out, int64_t* score,
                          const ParsimonyScan* scan = nullptr, const Constraints* constraints = nullptr);
inline uint32_t parsimony_code_mask(int code) {
    if (code < 20) return 1u << code;
    if (code == 20) return (1u << 2) | (1u << 3);
    if (code == 21) return (1u << 5) | (1u << 6);
    return 0xFFFFFu;
}

}  // namespace pml
