#include "host.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <numeric>
#include <sstream>
#include <thread>

#include "model.h"

namespace pml {

// =========================================================================================== alignment ======
namespace {

int crunch_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(8, std::max<unsigned>(1, hw));
}

// runs fn(i) for i = 0 .. n-1 on up to `threads` host threads (work stealing through an atomic counter)
template <typename F>
void parallel_for(int64_t n, int threads, F fn) {
    if (threads <= 1 || n <= 1) {
        for (int64_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    const int nt = (int)std::min<int64_t>(threads, n);
    for (int k = 0; k < nt; ++k)
        pool.emplace_back([&] {
            for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto& th : pool) th.join();
}

// Lexicographic order of alignment columns by residue code in taxon order (ties: column index), as an MSD radix sort that
// works on the caller's taxon-major letters directly: level t partitions a bucket by the codes of row t, and row t (one
// byte per site) stays cache resident while it is gathered from.  Random protein columns separate after ~5 levels.
struct ColumnSorter {
    int ntax;
    int64_t nsites;
    const uint8_t* chars;
    uint8_t lut[256];
    int code(int t, int64_t s) const { return lut[chars[(size_t)t * nsites + s]]; }
    // first taxon >= from at which the two columns differ (ntax if none)
    int first_difference(int64_t a, int64_t b, int from) const {
        for (int t = from; t < ntax; ++t)
            if (code(t, a) != code(t, b)) return t;
        return ntax;
    }
    bool less(int64_t a, int64_t b, int from) const {
        const int t = first_difference(a, b, from);
        return t < ntax ? code(t, a) < code(t, b) : a < b;
    }
    // sorts order[lo, hi), whose columns agree on taxa < level; tmp is scratch of the same extent.  fresh[k] (preset to 1)
    // is cleared wherever sorted column k equals column k-1: such pairs end up together in a bucket that survived all levels
    // or in a small bucket finished by insertion sort, so the pattern boundaries cost no extra pass over the alignment.
    void sort_range(int64_t* order, int64_t* tmp, uint8_t* fresh, int64_t lo, int64_t hi, int level) const {
        struct Job { int64_t lo, hi; int level; };
        std::vector<Job> stack{{lo, hi, level}};
        while (!stack.empty()) {
            const Job j = stack.back();
            stack.pop_back();
            const int64_t n = j.hi - j.lo;
            if (n <= 1) continue;
            if (j.level >= ntax) {  // identical columns keep their index order (every pass is stable)
                for (int64_t k = j.lo + 1; k < j.hi; ++k) fresh[k] = 0;
                continue;
            }
            if (n <= 24) {
                for (int64_t i = j.lo + 1; i < j.hi; ++i) {
                    const int64_t v = order[i];
                    int64_t k = i;
                    while (k > j.lo && less(v, order[k - 1], j.level)) {
                        order[k] = order[k - 1];
                        --k;
                    }
                    order[k] = v;
                }
                for (int64_t k = j.lo + 1; k < j.hi; ++k) fresh[k] = first_difference(order[k], order[k - 1], j.level) < ntax;
                continue;
            }
            const uint8_t* row = chars + (size_t)j.level * nsites;
            int64_t count[kCodes + 1] = {0};
            for (int64_t i = j.lo; i < j.hi; ++i) ++count[lut[row[order[i]]] + 1];
            int used = 0;
            for (int c = 0; c < kCodes; ++c) used += count[c + 1] != 0;
            if (used == 1) {  // nothing to move
                stack.push_back({j.lo, j.hi, j.level + 1});
                continue;
            }
            for (int c = 0; c < kCodes; ++c) count[c + 1] += count[c];
            int64_t pos[kCodes];
            for (int c = 0; c < kCodes; ++c) pos[c] = j.lo + count[c];
            for (int64_t i = j.lo; i < j.hi; ++i) tmp[pos[lut[row[order[i]]]]++] = order[i];
            std::memcpy(order + j.lo, tmp + j.lo, sizeof(int64_t) * (size_t)n);
            for (int c = 0; c < kCodes; ++c)
                if (count[c + 1] - count[c] > 1) stack.push_back({j.lo + count[c], j.lo + count[c + 1], j.level + 1});
        }
    }
};

}  // namespace

namespace {
// pattern index of every sorted column = running count of boundaries; weights, site -> pattern, representatives (two parallel
// passes over chunks of the sorted column list)
void fill_patterns(Patterns& out, const std::vector<int64_t>& order, const std::vector<uint8_t>& fresh, const int32_t* site_w, int threads) {
    const int64_t n = (int64_t)order.size();
    const int64_t chunk = 1 << 15, nchunks = (n + chunk - 1) / chunk;
    std::vector<int64_t> base(nchunks + 1, 0);
    parallel_for(nchunks, threads, [&](int64_t b) {
        int64_t cnt = 0;
        for (int64_t k = b * chunk; k < std::min(n, (b + 1) * chunk); ++k) cnt += fresh[k];
        base[b + 1] = cnt;
    });
    for (int64_t b = 0; b < nchunks; ++b) base[b + 1] += base[b];
    out.npat = base[nchunks];
    out.site_to_pat.assign(out.nsites, -1);
    out.first.assign(out.npat, 0);
    out.weight.assign(out.npat, 0);
    std::vector<int64_t>& first = out.first;
    parallel_for(nchunks, threads, [&](int64_t b) {
        int64_t p = base[b] - 1;
        const int64_t lo = b * chunk, hi = std::min(n, (b + 1) * chunk);
        // a pattern may straddle chunk boundaries: every chunk adds its share of a pattern's weight with one atomic add
        int64_t acc = 0;
        for (int64_t k = lo; k < hi; ++k) {
            const int64_t s = order[k];
            if (fresh[k]) {
                if (acc) __atomic_fetch_add(&out.weight[p], (int32_t)acc, __ATOMIC_RELAXED);
                acc = 0;
                first[++p] = s;
            }
            acc += site_w ? site_w[s] : 1;
            out.site_to_pat[s] = p;
        }
        if (acc) __atomic_fetch_add(&out.weight[p], (int32_t)acc, __ATOMIC_RELAXED);
    });
    out.full = true;
}
}  // namespace

void finish_patterns(Patterns& p) {
    if (p.full) return;
    fill_patterns(p, p.sorted_cols, p.sorted_fresh, p.col_weight.empty() ? nullptr : p.col_weight.data(), p.nsites >= 20000 ? crunch_threads() : 1);
    std::vector<int64_t>().swap(p.sorted_cols);
    std::vector<uint8_t>().swap(p.sorted_fresh);
}

void crunch_patterns(int ntax, int64_t nsites, const uint8_t* chars, const int32_t* site_w, Patterns& out, int rank, int nranks,
                     const CrunchShare* share, bool lazy) {
    out.ntax = ntax;
    out.nsites = nsites;
    ColumnSorter cs{ntax, nsites, chars, {}};
    for (int ch = 0; ch < 256; ++ch) cs.lut[ch] = (uint8_t)residue_code((unsigned char)ch);
    // ranks that share the sort are processes of one box: together they should not ask for more threads than it has cores
    int threads = nsites >= 20000 ? crunch_threads() : 1;
    if (share != nullptr && nranks > 1) threads = std::max(1, std::min(threads, (int)std::thread::hardware_concurrency() / nranks));
    const bool timing = getenv("PEPRML_CRUNCH_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    auto lap = [&](const char* what) {
        if (timing) {
            const double t1 = now();
            fprintf(stderr, "crunch %-12s %.1f ms\n", what, t1 - t0);
            t0 = t1;
        }
    };
    // level 0 by hand, so that its (up to 23) buckets can be sorted by different threads: the columns with positive weight are
    // counted per chunk and first residue, then every chunk scatters its columns to its own places (stable: chunk order = column
    // order) -- two parallel passes over row 0 instead of building an index list first
    const int64_t chunk0 = 1 << 16, nchunk0 = (nsites + chunk0 - 1) / chunk0;
    std::vector<int64_t> cnt0((size_t)nchunk0 * kCodes, 0);
    parallel_for(nchunk0, threads, [&](int64_t b) {
        int64_t* cnt = cnt0.data() + (size_t)b * kCodes;
        for (int64_t s = b * chunk0; s < std::min(nsites, (b + 1) * chunk0); ++s)
            if (!site_w || site_w[s] > 0) ++cnt[cs.lut[chars[s]]];
    });
    int64_t count[kCodes + 1] = {0};
    for (int64_t b = 0; b < nchunk0; ++b)
        for (int c = 0; c < kCodes; ++c) count[c + 1] += cnt0[(size_t)b * kCodes + c];
    for (int c = 0; c < kCodes; ++c) count[c + 1] += count[c];
    const int64_t n = count[kCodes];
    {
        int64_t run[kCodes];
        for (int c = 0; c < kCodes; ++c) run[c] = count[c];
        for (int64_t b = 0; b < nchunk0; ++b)
            for (int c = 0; c < kCodes; ++c) {
                const int64_t k = cnt0[(size_t)b * kCodes + c];
                cnt0[(size_t)b * kCodes + c] = run[c];
                run[c] += k;
            }
    }
    std::vector<int64_t> order((size_t)n), tmp((size_t)n);
    std::vector<uint8_t> fresh((size_t)n, 1);
    lap("setup");
    {
        parallel_for(nchunk0, threads, [&](int64_t b) {
            int64_t* pos = cnt0.data() + (size_t)b * kCodes;
            for (int64_t s = b * chunk0; s < std::min(nsites, (b + 1) * chunk0); ++s)
                if (!site_w || site_w[s] > 0) order[pos[cs.lut[chars[s]]]++] = s;
        });
        lap("level 0");
        int64_t pos[kCodes];
        // buckets of this rank: all of them, or -- when the ranks share the sort -- its part of a greedy balanced split
        bool mine[kCodes];
        for (int c = 0; c < kCodes; ++c) mine[c] = true;
        const bool shared = share != nullptr && nranks > 1;
        if (shared) {
            int by_size[kCodes];
            for (int c = 0; c < kCodes; ++c) by_size[c] = c;
            std::stable_sort(by_size, by_size + kCodes, [&](int a, int b) { return count[a + 1] - count[a] > count[b + 1] - count[b]; });
            std::vector<int64_t> load(nranks, 0);
            for (int i = 0; i < kCodes; ++i) {
                const int c = by_size[i];
                const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                load[r] += count[c + 1] - count[c];
                mine[c] = r == rank;
            }
        }
        parallel_for(kCodes, threads, [&](int64_t c) {
            if (mine[c]) cs.sort_range(order.data(), tmp.data(), fresh.data(), count[c], count[c + 1], 1);
            else {
                std::fill(order.begin() + count[c], order.begin() + count[c + 1], 0);
                std::fill(fresh.begin() + count[c], fresh.begin() + count[c + 1], 0);
            }
        });
        lap("radix");
        if (shared && !(*share)(order.data(), fresh.data(), n)) {
            // the exchange failed: finish alone (costs time, not correctness).  Every rank sees the same failure.
            std::vector<int64_t> again;
            again.reserve(n);
            for (int64_t s = 0; s < nsites; ++s)
                if (!site_w || site_w[s] > 0) again.push_back(s);
            for (int c = 0; c < kCodes; ++c) pos[c] = count[c];
            for (int64_t i = 0; i < n; ++i) order[pos[cs.lut[chars[again[i]]]]++] = again[i];
            std::fill(fresh.begin(), fresh.end(), 1);
            parallel_for(kCodes, threads, [&](int64_t c) { cs.sort_range(order.data(), tmp.data(), fresh.data(), count[c], count[c + 1], 1); });
        }
    }
    lap("exchange");
    if (lazy && nranks > 1) {
        // this rank's pattern block only: count the boundaries (parallel), then one walk over the block's own columns
        const int64_t chunk = 1 << 15, nchunks = (n + chunk - 1) / chunk;
        std::vector<int64_t> base(nchunks + 1, 0);
        parallel_for(nchunks, threads, [&](int64_t b) {
            int64_t cnt = 0;
            for (int64_t k = b * chunk; k < std::min(n, (b + 1) * chunk); ++k) cnt += fresh[k];
            base[b + 1] = cnt;
        });
        for (int64_t b = 0; b < nchunks; ++b) base[b + 1] += base[b];
        out.npat = base[nchunks];
        out.codes_p0 = out.npat * rank / nranks;
        out.codes_n = out.npat * (rank + 1) / nranks - out.codes_p0;
        out.own_weight.assign((size_t)out.codes_n, 0);
        out.own_first.assign((size_t)out.codes_n, 0);
        if (out.codes_n > 0) {
            // the chunk in which pattern codes_p0 starts, then forward to its first column
            int64_t b = std::upper_bound(base.begin(), base.end(), out.codes_p0) - base.begin() - 1;
            int64_t k = b * chunk, p = base[b] - 1;
            for (; k < n; ++k) {
                if (fresh[k] && ++p == out.codes_p0) break;
            }
            --p;
            for (; k < n; ++k) {
                if (fresh[k]) {
                    if (++p >= out.codes_p0 + out.codes_n) break;
                    out.own_first[(size_t)(p - out.codes_p0)] = order[k];
                }
                out.own_weight[(size_t)(p - out.codes_p0)] += site_w ? site_w[order[k]] : 1;
            }
        }
        lap("own block");
        gather_codes(ntax, nsites, chars, out.own_first, 0, out.codes_n, out.codes);
        lap("codes");
        out.full = false;
        out.weight.clear();
        out.site_to_pat.clear();
        out.first.clear();
        out.sorted_cols.swap(order);
        out.sorted_fresh.swap(fresh);
        if (site_w) out.col_weight.assign(site_w, site_w + nsites);
        else out.col_weight.clear();
        return;
    }
    fill_patterns(out, order, fresh, site_w, threads);
    lap("weights");
    out.codes_p0 = out.npat * rank / nranks;
    out.codes_n = out.npat * (rank + 1) / nranks - out.codes_p0;
    gather_codes(ntax, nsites, chars, out.first, out.codes_p0, out.codes_n, out.codes);
    lap("codes");
}

void gather_codes(int ntax, int64_t nsites, const uint8_t* chars, const std::vector<int64_t>& first, int64_t p0, int64_t n,
                  std::vector<uint8_t>& codes) {
    uint8_t lut[256];
    for (int ch = 0; ch < 256; ++ch) lut[ch] = (uint8_t)residue_code((unsigned char)ch);
    const int threads = nsites >= 20000 ? crunch_threads() : 1;
    codes.assign((size_t)ntax * n, 22);
    // gather of the representatives' codes, two taxon rows per pass over the (32-bit) column list
    std::vector<uint32_t> first32;
    const bool narrow = nsites < (int64_t)1 << 32;
    if (narrow) first32.assign(first.begin() + p0, first.begin() + p0 + n);
    parallel_for((ntax + 1) / 2, threads, [&](int64_t pair) {
        const int t0r = (int)(2 * pair), t1r = std::min(ntax - 1, t0r + 1);
        const uint8_t* row0 = chars + (size_t)t0r * nsites;
        const uint8_t* row1 = chars + (size_t)t1r * nsites;
        uint8_t* dst0 = codes.data() + (size_t)t0r * n;
        uint8_t* dst1 = codes.data() + (size_t)t1r * n;
        for (int64_t p = 0; p < n; ++p) {
            const size_t s = narrow ? (size_t)first32[p] : (size_t)first[p0 + p];
            dst0[p] = lut[row0[s]];
            dst1[p] = lut[row1[s]];
        }
    });
}

bool read_phylip(const std::string& path, std::vector<std::string>& names, std::vector<uint8_t>& chars, int64_t& nsites,
                 std::string& err) {
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        err = "cannot open alignment file " + path;
        return false;
    }
    long ntax = 0, len = 0;
    if (!(in >> ntax >> len) || ntax < 1 || len < 1) {
        err = "bad phylip header in " + path;
        return false;
    }
    nsites = len;
    names.clear();
    chars.assign((size_t)ntax * len, '?');
    for (long t = 0; t < ntax; ++t) {
        std::string name;
        if (!(in >> name)) {
            err = "phylip: missing taxon " + std::to_string(t + 1);
            return false;
        }
        names.push_back(name);
        int64_t got = 0;
        char ch;
        while (got < len && in.get(ch)) {
            if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') continue;
            chars[(size_t)t * len + got++] = (uint8_t)ch;
        }
        if (got != len) {
            err = "phylip: sequence of " + name + " is shorter than " + std::to_string(len);
            return false;
        }
    }
    return true;
}

// =========================================================================================== newick =========
namespace {

struct RawNode {
    std::vector<int> kids;
    std::string label;
    std::string length;  // textual, empty when absent
    int parent = -1;
};

struct RawTree {
    std::vector<RawNode> nodes;
    int top = -1;
};

bool parse_raw(const std::string& s, RawTree& out, std::string& err) {
    size_t pos = 0;
    auto skip = [&] {
        while (pos < s.size() && (s[pos] == ' ' || s[pos] == '\n' || s[pos] == '\t' || s[pos] == '\r')) ++pos;
    };
    std::function<int()> node = [&]() -> int {
        skip();
        const int id = (int)out.nodes.size();
        out.nodes.emplace_back();
        if (pos < s.size() && s[pos] == '(') {
            ++pos;
            for (;;) {
                const int k = node();
                if (k < 0) return -1;
                out.nodes[k].parent = id;
                out.nodes[id].kids.push_back(k);
                skip();
                if (pos < s.size() && s[pos] == ',') {
                    ++pos;
                    continue;
                }
                if (pos < s.size() && s[pos] == ')') {
                    ++pos;
                    break;
                }
                err = "newick: expected ',' or ')' at offset " + std::to_string(pos);
                return -1;
            }
        }
        skip();
        size_t st = pos;
        while (pos < s.size() && !std::strchr(":,();[ \n\t\r", s[pos])) ++pos;
        out.nodes[id].label = s.substr(st, pos - st);
        skip();
        if (pos < s.size() && s[pos] == ':') {
            ++pos;
            skip();
            st = pos;
            while (pos < s.size() && !std::strchr(",();[ \n\t\r", s[pos])) ++pos;
            out.nodes[id].length = s.substr(st, pos - st);
            skip();
        }
        if (pos < s.size() && s[pos] == '[') {  // `:len[support]` form
            const size_t close = s.find(']', pos);
            if (close == std::string::npos) {
                err = "newick: unterminated '['";
                return -1;
            }
            if (out.nodes[id].label.empty()) out.nodes[id].label = s.substr(pos + 1, close - pos - 1);
            pos = close + 1;
        }
        return id;
    };
    out.nodes.clear();
    out.top = node();
    if (out.top < 0) return false;
    if (out.nodes[out.top].kids.empty()) {
        err = "newick: no tree found";
        return false;
    }
    return true;
}

std::string fixed20(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.20f", v);
    return buf;
}

// shortest decimal that round-trips, printed the way Java's Double.toString lays it out
std::string java_double(double v) {
    if (v == 0.0) return "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 17; ++prec) {
        std::snprintf(buf, sizeof buf, "%.*e", prec - 1, v);
        if (std::strtod(buf, nullptr) == v) break;
    }
    std::string m(buf);
    const size_t epos = m.find('e');
    int ex = std::atoi(m.c_str() + epos + 1);
    std::string digits;
    bool neg = false;
    for (size_t i = 0; i < epos; ++i) {
        if (m[i] == '-') neg = true;
        else if (m[i] != '.') digits.push_back(m[i]);
    }
    std::string out = neg ? "-" : "";
    if (ex >= -3 && ex < 7) {
        if (ex >= 0) {
            while ((int)digits.size() < ex + 2) digits.push_back('0');
            out += digits.substr(0, ex + 1) + "." + digits.substr(ex + 1);
        } else {
            out += "0." + std::string(-ex - 1, '0') + digits;
        }
    } else {
        out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "E" + std::to_string(ex);
    }
    return out;
}

}  // namespace

bool parse_newick(const std::string& text, const std::vector<std::string>& names, double default_len, Topology& T,
                  std::string& err) {
    RawTree raw;
    if (!parse_raw(text, raw, err)) return false;
    const int ntax = (int)names.size();
    if (ntax < 3) {
        err = "need at least 3 taxa";
        return false;
    }
    std::unordered_map<std::string, int> index;
    for (int i = 0; i < ntax; ++i) index[names[i]] = i;
    T = Topology();
    T.ntax = ntax;
    T.nbr.assign(T.nnodes(), {-1, -1, -1});
    T.edge.assign(T.nnodes(), {-1, -1, -1});
    std::vector<int> deg(T.nnodes(), 0), seen(ntax, 0);
    int next_inner = ntax;
    bool ok = true;
    auto connect = [&](int a, int b, double l) {
        if (a < 0 || b < 0) return;
        if (deg[a] >= 3 || deg[b] >= 3 || (int)T.ea.size() >= T.nedges()) {
            ok = false;
            return;
        }
        const int e = (int)T.ea.size();
        T.ea.push_back(a);
        T.eb.push_back(b);
        T.len.push_back(l);
        T.nbr[a][deg[a]] = b;
        T.edge[a][deg[a]++] = e;
        T.nbr[b][deg[b]] = a;
        T.edge[b][deg[b]++] = e;
    };
    auto length_of = [&](const RawNode& n, bool& has) {
        has = !n.length.empty();
        return has ? std::strtod(n.length.c_str(), nullptr) : 0.0;
    };
    // returns the engine node id of raw node r (below the top)
    std::function<int(int)> build = [&](int r) -> int {
        const RawNode& n = raw.nodes[r];
        if (n.kids.empty()) {
            auto it = index.find(n.label);
            if (it == index.end()) {
                err = "newick: taxon '" + n.label + "' is not in the alignment";
                ok = false;
                return -1;
            }
            if (seen[it->second]++) {
                err = "newick: taxon '" + n.label + "' appears twice";
                ok = false;
                return -1;
            }
            return it->second;
        }
        if (n.kids.size() != 2) {
            err = "newick: only bifurcating trees are supported below the top node";
            ok = false;
            return -1;
        }
        if (next_inner >= T.nnodes()) {
            ok = false;
            return -1;
        }
        const int v = next_inner++;
        for (int k : n.kids) {
            const int c = build(k);
            if (!ok) return -1;
            bool has;
            const double l = length_of(raw.nodes[k], has);
            connect(v, c, has ? l : default_len);
        }
        return v;
    };
    const RawNode& top = raw.nodes[raw.top];
    if (top.kids.size() == 3) {
        const int v = next_inner++;
        for (int k : top.kids) {
            const int c = build(k);
            if (!ok) break;
            bool has;
            const double l = length_of(raw.nodes[k], has);
            connect(v, c, has ? l : default_len);
        }
    } else if (top.kids.size() == 2) {
        const int a = build(top.kids[0]);
        const int b = ok ? build(top.kids[1]) : -1;
        bool ha, hb;
        const double la = length_of(raw.nodes[top.kids[0]], ha), lb = length_of(raw.nodes[top.kids[1]], hb);
        if (ok) connect(a, b, (ha && hb) ? la + lb : (ha ? la : (hb ? lb : default_len)));
    } else {
        err = "newick: top node must have 2 or 3 children";
        return false;
    }
    if (!ok || next_inner != T.nnodes() || (int)T.ea.size() != T.nedges()) {
        if (err.empty()) err = "newick: tree is not a binary tree over exactly the alignment's taxa";
        return false;
    }
    for (int i = 0; i < ntax; ++i)
        if (!seen[i]) {
            err = "newick: taxon '" + names[i] + "' is missing from the tree";
            return false;
        }
    for (double& l : T.len)
        if (!(l >= 0.0)) l = default_len;
    return true;
}

std::string write_newick_result(const Topology& T, const std::vector<std::string>& names) {
    std::string out;
    std::function<void(int, int)> sub = [&](int v, int from) {
        if (T.is_tip(v)) {
            out += names[v];
            return;
        }
        out += "(";
        bool first = true;
        for (int s = 0; s < 3; ++s) {
            if (T.nbr[v][s] == from) continue;
            if (!first) out += ",";
            first = false;
            sub(T.nbr[v][s], v);
            out += ":" + fixed20(T.len[T.edge[v][s]]);
        }
        out += ")";
    };
    const int root = T.nbr[0][0];
    out += "(";
    bool first = true;
    for (int s = 0; s < 3; ++s) {
        if (T.nbr[root][s] == 0) continue;
        if (!first) out += ",";
        first = false;
        sub(T.nbr[root][s], root);
        out += ":" + fixed20(T.len[T.edge[root][s]]);
    }
    out += "," + names[0] + ":" + fixed20(T.len[T.edge[0][0]]) + "):0.0;";
    return out;
}

// =========================================================================================== traversal ======
void ViewState::plan(const Topology& T, int v, int toward_node, std::vector<ViewOp>& ops) {
    if (T.is_tip(v)) return;
    if (fold_cherries && is_cherry_view(T, v, toward_node)) {
        if (!fresh[v - T.ntax]) {
            fresh[v - T.ntax] = 1;
            ++folded;
        }
        return;
    }
    const int s = T.slot_of(v, toward_node);
    if (orient[v - T.ntax] == s) return;
    ViewOp op{};
    op.node = v;
    op.toward = s;
    int k = 0;
    for (int q = 0; q < 3; ++q) {
        if (q == s) continue;
        op.child[k] = T.nbr[v][q];
        op.cedge[k] = T.edge[v][q];
        ++k;
    }
    plan(T, op.child[0], v, ops);
    plan(T, op.child[1], v, ops);
    ops.push_back(op);
    orient[v - T.ntax] = s;
}

void ViewState::branch_changed(const Topology& T, int e) {
    // walk away from the branch on both sides; a stored CLV at x (reached from y) contains the branch unless it faces y
    if (T.is_tip(T.ea[e])) touch(T, T.eb[e]);  // a folded cherry contains exactly the branches of its two tips
    if (T.is_tip(T.eb[e])) touch(T, T.ea[e]);
    std::vector<std::pair<int, int>> stack{{T.ea[e], T.eb[e]}, {T.eb[e], T.ea[e]}};
    while (!stack.empty()) {
        const auto [x, y] = stack.back();
        stack.pop_back();
        if (T.is_tip(x)) continue;
        int& o = orient[x - T.ntax];
        if (o >= 0 && T.nbr[x][o] != y) o = -1;
        for (int s = 0; s < 3; ++s)
            if (T.nbr[x][s] != y) stack.push_back({T.nbr[x][s], x});
    }
}

// =========================================================================================== bootstrap ======
double randum(int64_t* seed) {
    // 36-bit multiplicative congruential generator held in limbs of 12/12/12(8) bits; multiplier limbs 1549 and 406
    const int64_t lo = *seed & 0xFFF, mid = (*seed >> 12) & 0xFFF, hi = (*seed >> 24) & 0xFF;
    int64_t acc = 1549 * lo;
    const int64_t nlo = acc & 0xFFF;
    acc = (acc >> 12) + 1549 * mid + 406 * lo;
    const int64_t nmid = acc & 0xFFF;
    acc = (acc >> 12) + 1549 * hi + 406 * mid;
    const int64_t nhi = acc & 0xFF;
    *seed = (nhi << 24) | (nmid << 12) | nlo;
    return 0.00390625 * ((double)nhi + 0.000244140625 * ((double)nmid + 0.000244140625 * (double)nlo));
}

void bootstrap_replicates(int64_t* seed, const std::vector<int32_t>& pw, int nrep, int32_t* out) {
    const int64_t npat = (int64_t)pw.size();
    int64_t total = 0;
    for (int32_t w : pw) total += w;
    // draws land on expanded-site slots; slot -> pattern through a prefix table
    std::vector<int64_t> slot_pat((size_t)total);
    {
        int64_t pos = 0;
        for (int64_t p = 0; p < npat; ++p)
            for (int32_t k = 0; k < pw[p]; ++k) slot_pat[pos++] = p;
    }
    for (int r = 0; r < nrep; ++r) {
        int32_t* w = out + (int64_t)r * npat;
        std::fill(w, w + npat, 0);
        for (int64_t j = 0; j < total; ++j) {
            const int64_t slot = (int64_t)((double)total * randum(seed));
            ++w[slot_pat[slot]];
        }
    }
}

// =========================================================================================== support ========
namespace {

using Bits = std::vector<uint64_t>;
struct BitsHash {
    size_t operator()(const Bits& b) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint64_t w : b) h = (h ^ w) * 0xff51afd7ed558ccdull + (h >> 29);
        return (size_t)h;
    }
};

int popcount_bits(const Bits& b) {
    int c = 0;
    for (uint64_t w : b) c += __builtin_popcountll(w);
    return c;
}
int lowest_bit(const Bits& b) {
    for (size_t i = 0; i < b.size(); ++i)
        if (b[i]) return (int)(i * 64 + __builtin_ctzll(b[i]));
    return 1 << 30;
}

// canonical side of a split: the smaller one; on a tie the side holding the lowest taxon index (Bipartition.java:41-64;
// when the given side does not own the lowest index the complement is taken, matching the `else` branch there)
Bits canonical(const Bits& side, int ntax) {
    Bits comp(side.size());
    for (size_t i = 0; i < side.size(); ++i) comp[i] = ~side[i];
    if (ntax % 64) comp.back() &= (~0ull) >> (64 - ntax % 64);
    const int a = popcount_bits(side), b = ntax - a;
    if (a < b) return side;
    if (a > b) return comp;
    return lowest_bit(side) < lowest_bit(comp) ? side : comp;
}

// PEPR's unroot (BasicTree.java:669-717): a bifurcating top is dissolved, its first non-leaf child becomes the top and
// adopts the other child, whose branch absorbs the new top's former branch.  Returns true if the tree was rooted.
bool unroot(RawTree& t) {
    RawNode& top = t.nodes[t.top];
    if (top.kids.size() != 2) return false;
    int keep = top.kids[0], move = top.kids[1];
    if (t.nodes[keep].kids.size() < 2) std::swap(keep, move);
    const double sum = std::strtod(t.nodes[move].length.c_str(), nullptr) + std::strtod(t.nodes[keep].length.c_str(), nullptr);
    if (!t.nodes[move].length.empty() || !t.nodes[keep].length.empty()) t.nodes[move].length = java_double(sum);
    t.nodes[keep].length.clear();
    t.nodes[keep].kids.push_back(move);
    t.nodes[move].parent = keep;
    t.nodes[keep].parent = -1;
    top.kids.clear();
    t.top = keep;
    return true;
}

}  // namespace

std::string support_tree(const std::string& main_newick, const std::vector<std::string>& trees, bool as_percent,
                         std::vector<int32_t>* counts, std::string& err) {
    RawTree main;
    if (!parse_raw(main_newick, main, err)) return "";
    unroot(main);
    std::vector<std::string> taxa;
    for (const RawNode& n : main.nodes)
        if (n.kids.empty() && !n.label.empty()) taxa.push_back(n.label);
    std::sort(taxa.begin(), taxa.end());
    const int ntax = (int)taxa.size();
    const size_t words = (size_t)(ntax + 63) / 64;
    auto taxon_index = [&](const std::string& s) {
        auto it = std::lower_bound(taxa.begin(), taxa.end(), s);
        return (it != taxa.end() && *it == s) ? (int)(it - taxa.begin()) : -1;
    };
    // leaf set of every node of a tree (children before parents because ids are assigned in preorder)
    auto leafsets = [&](const RawTree& t) {
        std::vector<Bits> sets(t.nodes.size(), Bits(words, 0));
        for (int v = (int)t.nodes.size() - 1; v >= 0; --v) {
            const RawNode& n = t.nodes[v];
            if (n.kids.empty()) {
                const int ix = n.label.empty() ? -1 : taxon_index(n.label);
                if (ix >= 0) sets[v][ix / 64] |= 1ull << (ix % 64);
            }
            for (int k : n.kids)
                for (size_t w = 0; w < words; ++w) sets[v][w] |= sets[k][w];
        }
        return sets;
    };
    std::unordered_map<Bits, int32_t, BitsHash> tally;
    for (const std::string& text : trees) {
        RawTree st;
        if (!parse_raw(text, st, err)) return "";
        unroot(st);
        // every node of the support tree adds its split once, the dissolved root included (it contributes the empty set),
        // exactly as TreeSupportDecorator.java:124-140 does
        for (const Bits& s : leafsets(st)) ++tally[canonical(s, ntax)];
    }
    const std::vector<Bits> msets = leafsets(main);
    const int ntrees = (int)trees.size();
    std::string out;
    if (counts) counts->clear();
    std::function<void(int)> emit = [&](int v) {
        const RawNode& n = main.nodes[v];
        if (n.kids.empty()) {
            out += n.label;
            return;
        }
        if (n.kids.size() > 1) out += "(";
        for (size_t k = 0; k < n.kids.size(); ++k) {
            if (k) out += ",";
            emit(n.kids[k]);
            const std::string& l = main.nodes[n.kids[k]].length;
            if (as_percent) out += ":" + (l.empty() ? std::string("0.0") : l);
            else out += ":" + java_double(l.empty() ? 0.0 : std::strtod(l.c_str(), nullptr));
        }
        if (n.kids.size() > 1) out += ")";
        auto it = tally.find(canonical(msets[v], ntax));
        int32_t c = it == tally.end() ? 0 : it->second;
        if (as_percent) {
            if (v == main.top) return;  // raxmlHPC -f b leaves the top node unlabelled
            c = ntrees > 0 ? (int32_t)std::floor(100.0 * c / ntrees + 0.5) : 0;
        }
        if (counts && v != main.top) counts->push_back(c);
        out += std::to_string(c);
    };
    emit(main.top);
    out += ";";
    return out;
}

}  // namespace pml

// =========================================================================================== SPR ============
namespace pml {

namespace {
void set_link(Topology& T, int v, int old_nb, int new_nb, int e) {
    const int s = T.slot_of(v, old_nb);
    T.nbr[v][s] = new_nb;
    T.edge[v][s] = e;
}
}  // namespace

bool spr_apply(Topology& T, ViewState& V, int p, int s, int target, SprMove& mv) {
    if (T.is_tip(p)) return false;
    mv = SprMove();
    mv.p = p;
    mv.s = s;
    for (int k = 0; k < 3; ++k) {
        const int nb = T.nbr[p][k];
        if (nb == s) mv.e_s = T.edge[p][k];
        else if (mv.q < 0) {
            mv.q = nb;
            mv.e_q = T.edge[p][k];
            mv.slot_q = k;
        } else {
            mv.r = nb;
            mv.e_r = T.edge[p][k];
            mv.slot_r = k;
        }
    }
    if (mv.e_s < 0 || mv.r < 0 || target == mv.e_q || target == mv.e_r || target == mv.e_s) return false;
    mv.e_t = target;
    mv.a = T.ea[target];
    mv.b = T.eb[target];
    mv.len_q = T.len[mv.e_q];
    mv.len_r = T.len[mv.e_r];
    mv.len_t = T.len[target];
    mv.len_s = T.len[mv.e_s];
    // prune: q -- r through branch e_q
    set_link(T, mv.q, p, mv.r, mv.e_q);
    set_link(T, mv.r, p, mv.q, mv.e_q);
    T.ea[mv.e_q] = mv.q;
    T.eb[mv.e_q] = mv.r;
    T.len[mv.e_q] = mv.len_q + mv.len_r;
    // regraft: a -- p through e_t, p -- b through e_r
    set_link(T, mv.a, mv.b, p, mv.e_t);
    set_link(T, mv.b, mv.a, p, mv.e_r);
    T.nbr[p][mv.slot_q] = mv.a;
    T.edge[p][mv.slot_q] = mv.e_t;
    T.nbr[p][mv.slot_r] = mv.b;
    T.edge[p][mv.slot_r] = mv.e_r;
    T.ea[mv.e_t] = mv.a;
    T.eb[mv.e_t] = p;
    T.ea[mv.e_r] = p;
    T.eb[mv.e_r] = mv.b;
    T.len[mv.e_t] = T.len[mv.e_r] = 0.5 * mv.len_t;
    V.orient[p - T.ntax] = -1;
    for (int v : {p, mv.q, mv.r, mv.a, mv.b}) V.touch(T, v);
    V.branch_changed(T, mv.e_q);
    V.branch_changed(T, mv.e_t);
    V.branch_changed(T, mv.e_r);
    return true;
}

bool spr_prune(Topology& T, ViewState& V, int p, int s, SprMove& mv) {
    if (T.is_tip(p)) return false;
    mv = SprMove();
    mv.p = p;
    mv.s = s;
    for (int k = 0; k < 3; ++k) {
        const int nb = T.nbr[p][k];
        if (nb == s) mv.e_s = T.edge[p][k];
        else if (mv.q < 0) {
            mv.q = nb;
            mv.e_q = T.edge[p][k];
            mv.slot_q = k;
        } else {
            mv.r = nb;
            mv.e_r = T.edge[p][k];
            mv.slot_r = k;
        }
    }
    if (mv.e_s < 0 || mv.r < 0) return false;
    mv.len_q = T.len[mv.e_q];
    mv.len_r = T.len[mv.e_r];
    // q -- r through branch e_q; p keeps only its link to s, branch e_r is parked
    set_link(T, mv.q, p, mv.r, mv.e_q);
    set_link(T, mv.r, p, mv.q, mv.e_q);
    T.ea[mv.e_q] = mv.q;
    T.eb[mv.e_q] = mv.r;
    T.len[mv.e_q] = mv.len_q + mv.len_r;
    T.nbr[p][mv.slot_q] = T.nbr[p][mv.slot_r] = -1;
    V.orient[p - T.ntax] = -1;
    for (int v : {p, mv.q, mv.r}) V.touch(T, v);
    V.branch_changed(T, mv.e_q);
    return true;
}

void spr_unprune(Topology& T, ViewState& V, const SprMove& mv) {
    const int p = mv.p;
    set_link(T, mv.q, mv.r, p, mv.e_q);
    set_link(T, mv.r, mv.q, p, mv.e_r);
    T.nbr[p][mv.slot_q] = mv.q;
    T.edge[p][mv.slot_q] = mv.e_q;
    T.nbr[p][mv.slot_r] = mv.r;
    T.edge[p][mv.slot_r] = mv.e_r;
    T.ea[mv.e_q] = p;
    T.eb[mv.e_q] = mv.q;
    T.ea[mv.e_r] = p;
    T.eb[mv.e_r] = mv.r;
    T.len[mv.e_q] = mv.len_q;
    T.len[mv.e_r] = mv.len_r;
    V.orient[p - T.ntax] = -1;
    for (int v : {p, mv.q, mv.r}) V.touch(T, v);
    // views computed on the pruned tree that look across the place where p belongs lack the subtree behind s
    V.branch_changed(T, mv.e_q);
    V.branch_changed(T, mv.e_r);
}

void spr_undo(Topology& T, ViewState& V, const SprMove& mv) {
    const int p = mv.p;
    // detach from a, b
    set_link(T, mv.a, p, mv.b, mv.e_t);
    set_link(T, mv.b, p, mv.a, mv.e_t);
    T.ea[mv.e_t] = mv.a;
    T.eb[mv.e_t] = mv.b;
    T.len[mv.e_t] = mv.len_t;
    // back between q and r
    set_link(T, mv.q, mv.r, p, mv.e_q);
    set_link(T, mv.r, mv.q, p, mv.e_r);
    T.nbr[p][mv.slot_q] = mv.q;
    T.edge[p][mv.slot_q] = mv.e_q;
    T.nbr[p][mv.slot_r] = mv.r;
    T.edge[p][mv.slot_r] = mv.e_r;
    T.ea[mv.e_q] = p;
    T.eb[mv.e_q] = mv.q;
    T.ea[mv.e_r] = p;
    T.eb[mv.e_r] = mv.r;
    T.len[mv.e_q] = mv.len_q;
    T.len[mv.e_r] = mv.len_r;
    V.orient[p - T.ntax] = -1;
    for (int v : {p, mv.q, mv.r, mv.a, mv.b}) V.touch(T, v);
    V.branch_changed(T, mv.e_q);
    V.branch_changed(T, mv.e_t);
    V.branch_changed(T, mv.e_r);
    if (T.len[mv.e_s] != mv.len_s) {  // the caller re-optimised the subtree's branch on the rejected topology
        T.len[mv.e_s] = mv.len_s;
        V.branch_changed(T, mv.e_s);
    }
}

std::vector<int> spr_targets(const Topology& T, int p, int s, int radius) {
    std::vector<int> out;
    if (T.is_tip(p)) return out;
    int q = -1, r = -1;
    for (int k = 0; k < 3; ++k) {
        const int nb = T.nbr[p][k];
        if (nb == s) continue;
        (q < 0 ? q : r) = nb;
    }
    // walk away from p through q and through r; depth counts branches beyond the two next to p
    struct Item { int node, from, depth; };
    std::vector<Item> stack{{q, p, 0}, {r, p, 0}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        if (T.is_tip(it.node) || it.depth >= radius) continue;
        for (int k = 0; k < 3; ++k) {
            const int nb = T.nbr[it.node][k];
            if (nb == it.from) continue;
            out.push_back(T.edge[it.node][k]);
            stack.push_back({nb, it.node, it.depth + 1});
        }
    }
    return out;
}

// =========================================================================================== parsimony ======
namespace {

// reference implementation of ParsimonyScan on the host (pml_parsimony_tree, which has no context and no GPU)
struct HostParsimony {
    const Patterns& pat;
    std::vector<std::vector<uint32_t>> tipmask, down, up;
    explicit HostParsimony(const Patterns& p) : pat(p), tipmask(p.ntax, std::vector<uint32_t>((size_t)p.npat)) {  // needs all codes
        for (int t = 0; t < pat.ntax; ++t)
            for (int64_t s = 0; s < pat.npat; ++s) tipmask[t][s] = parsimony_code_mask(pat.codes[(size_t)t * pat.npat + s]);
    }
    bool operator()(const GrowTree& g, const std::vector<int>& pre, int next_taxon, int64_t& score_out, std::vector<int64_t>& cost) {
        const int64_t P = pat.npat;
        const int N = (int)g.parent.size(), root = 0;
        down.assign(N, {});
        up.assign(N, {});
        int64_t score = 0;
        for (int i = (int)pre.size() - 1; i >= 0; --i) {  // children first
            const int v = pre[i];
            if (g.left[v] < 0) {
                down[v] = tipmask[g.taxon[v]];
                continue;
            }
            const auto &A = down[g.left[v]], &B = down[g.right[v]];
            down[v].resize((size_t)P);
            for (int64_t s = 0; s < P; ++s) {
                const uint32_t x = A[s] & B[s];
                if (x) down[v][s] = x;
                else {
                    down[v][s] = A[s] | B[s];
                    score += pat.weight[s];
                }
            }
        }
        {
            const auto &A = down[g.left[root]], &B = tipmask[g.taxon[root]];
            for (int64_t s = 0; s < P; ++s)
                if (!(A[s] & B[s])) score += pat.weight[s];
        }
        score_out = score;
        if (next_taxon < 0) return true;
        // parents first: set above each node
        up[g.left[root]] = tipmask[g.taxon[root]];
        for (int v : pre) {
            if (g.left[v] < 0) continue;
            for (int side = 0; side < 2; ++side) {
                const int c = side ? g.right[v] : g.left[v], sib = side ? g.left[v] : g.right[v];
                up[c].resize((size_t)P);
                const auto &U = up[v], &S = down[sib];
                for (int64_t s = 0; s < P; ++s) {
                    const uint32_t x = U[s] & S[s];
                    up[c][s] = x ? x : (U[s] | S[s]);
                }
            }
        }
        const std::vector<uint32_t>& X = tipmask[next_taxon];
        cost.assign(pre.size(), 0);
        for (size_t i = 0; i < pre.size(); ++i) {
            const auto &U = up[pre[i]], &D = down[pre[i]];
            int64_t c = 0;
            for (int64_t s = 0; s < P; ++s) {
                uint32_t e = U[s] & D[s];
                if (!e) e = U[s] | D[s];
                if (!(e & X[s])) c += pat.weight[s];
            }
            cost[i] = c;
        }
        return true;
    }
};

}  // namespace

bool parsimony_start_tree(const Patterns& pat, int64_t seed, double default_len, Topology& out, int64_t* score_out,
                          const ParsimonyScan* scan, const Constraints* constraints) {
    const int n = pat.ntax;
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    for (int i = n - 1; i > 0; --i) {  // Fisher-Yates on the randum stream
        const int j = (int)((double)(i + 1) * randum(&seed));
        std::swap(order[i], order[j < 0 ? 0 : (j > i ? i : j)]);
    }
    std::unique_ptr<HostParsimony> host;
    ParsimonyScan host_scan;
    if (!scan) {
        host = std::make_unique<HostParsimony>(pat);
        host_scan = [&](const GrowTree& g, const std::vector<int>& pre, int next, int64_t& sc, std::vector<int64_t>& cost) {
            return (*host)(g, pre, next, sc, cost);
        };
        scan = &host_scan;
    }

    GrowTree g;
    const int root = g.add_node(order[0]);  // the root tip hangs above the single top node
    int top = g.add_node(-1);
    {
        const int a = g.add_node(order[1]), b = g.add_node(order[2]);
        g.left[top] = a;
        g.right[top] = b;
        g.parent[a] = g.parent[b] = top;
        g.parent[top] = root;
        g.left[root] = top;
    }
    int64_t total_score = 0;
    std::vector<int64_t> cost;
    for (int k = 3; k <= n; ++k) {
        // nodes below the root tip, parents first
        std::vector<int> pre, stack{g.left[root]};
        while (!stack.empty()) {
            const int v = stack.back();
            stack.pop_back();
            pre.push_back(v);
            if (g.left[v] >= 0) {
                stack.push_back(g.left[v]);
                stack.push_back(g.right[v]);
            }
        }
        if (!(*scan)(g, pre, k < n ? order[k] : -1, total_score, cost)) return false;
        if (k == n) break;
        // cheapest branch (above node v) for the next taxon; the first one in `pre` wins a tie.  Under topological constraints
        // only the branches that keep every split possible among the taxa present are candidates (there always is one as long
        // as the splits are compatible with each other)
        std::vector<char> allowed;
        if (constraints && !constraints->empty()) allowed = allowed_insertions(g, pre, order[k], *constraints);
        int best_v = -1;
        int64_t best_cost = 0;
        for (size_t i = 0; i < pre.size(); ++i)
            if ((allowed.empty() || allowed[i]) && (best_v < 0 || cost[i] < best_cost)) {
                best_v = pre[i];
                best_cost = cost[i];
            }
        if (best_v < 0) return false;  // contradictory constraints
        // insert a new inner node above best_v
        const int w = g.add_node(-1), x = g.add_node(order[k]);
        const int par = g.parent[best_v];
        g.parent[w] = par;
        if (g.left[par] == best_v) g.left[par] = w; else g.right[par] = w;
        g.left[w] = best_v;
        g.right[w] = x;
        g.parent[best_v] = g.parent[x] = w;
    }
    if (score_out) *score_out = total_score;
    // to newick (root tip joined to the top node as a trifurcation), then through the ordinary parser
    std::function<std::string(int)> nw = [&](int v) -> std::string {
        if (g.left[v] < 0) return pat.names[g.taxon[v]];
        return "(" + nw(g.left[v]) + "," + nw(g.right[v]) + ")";
    };
    const int t0 = g.left[root];
    std::string text;
    if (g.left[t0] >= 0) text = "(" + pat.names[g.taxon[root]] + "," + nw(g.left[t0]) + "," + nw(g.right[t0]) + ");";
    else text = "(" + pat.names[g.taxon[root]] + "," + nw(t0) + ");";
    std::string err;
    return parse_newick(text, pat.names, default_len, out, err);
}

// =========================================================================================== constraints ====
namespace {
inline bool bit(const std::vector<uint64_t>& b, int i) { return (b[(size_t)i >> 6] >> (i & 63)) & 1ull; }
inline void set_bit(std::vector<uint64_t>& b, int i) { b[(size_t)i >> 6] |= 1ull << (i & 63); }
inline bool subset(const std::vector<uint64_t>& a, const std::vector<uint64_t>& of) {
    for (size_t w = 0; w < a.size(); ++w)
        if (a[w] & ~of[w]) return false;
    return true;
}
inline bool disjoint(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
    for (size_t w = 0; w < a.size(); ++w)
        if (a[w] & b[w]) return false;
    return true;
}
inline int count_bits(const std::vector<uint64_t>& a) {
    int c = 0;
    for (uint64_t w : a) c += __builtin_popcountll(w);
    return c;
}
}  // namespace

bool parse_constraints(const std::string& text, const std::vector<std::string>& names, Constraints& out, std::string& err) {
    out = Constraints();
    out.ntax = (int)names.size();
    const size_t words = ((size_t)out.ntax + 63) / 64;
    std::unordered_map<std::string, int> index;
    for (int i = 0; i < out.ntax; ++i) index[names[i]] = i;
    std::vector<std::pair<int, std::string>> rows;
    std::istringstream in(text);
    std::string line;
    int cur = -1;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            std::string nm = line.substr(1);
            const size_t sp = nm.find_first_of(" \t");
            if (sp != std::string::npos) nm.resize(sp);
            auto it = index.find(nm);
            if (it == index.end()) {
                err = "constraints name a taxon that is not in the alignment: " + nm;
                return false;
            }
            cur = it->second;
            rows.push_back({cur, std::string()});
        } else if (cur >= 0) rows.back().second += line;
    }
    size_t ncol = 0;
    for (auto& r : rows) ncol = std::max(ncol, r.second.size());
    for (size_t col = 0; col < ncol; ++col) {
        SplitConstraint sc{std::vector<uint64_t>(words, 0), std::vector<uint64_t>(words, 0)};
        for (auto& r : rows) {
            const char ch = col < r.second.size() ? r.second[col] : '-';
            if (ch == '1') set_bit(sc.one, r.first);
            else if (ch == '0') set_bit(sc.zero, r.first);
        }
        if (count_bits(sc.one) >= 2 && count_bits(sc.zero) >= 2) out.splits.push_back(std::move(sc));  // the others hold in every tree
    }
    return true;
}

std::string constraints_from_tree(const std::string& newick, std::string& err) {
    RawTree t;
    if (!parse_raw(newick, t, err)) return "";
    std::vector<std::string> taxa;
    for (const RawNode& n : t.nodes)
        if (n.kids.empty() && !n.label.empty()) taxa.push_back(n.label);
    std::sort(taxa.begin(), taxa.end());
    const int ntax = (int)taxa.size(), nnodes = (int)t.nodes.size();
    // leaf set of every node (node ids are in preorder: children after parents)
    std::vector<std::vector<char>> has((size_t)nnodes, std::vector<char>((size_t)ntax, 0));
    for (int v = nnodes - 1; v >= 0; --v) {
        const RawNode& n = t.nodes[v];
        if (n.kids.empty() && !n.label.empty()) has[v][std::lower_bound(taxa.begin(), taxa.end(), n.label) - taxa.begin()] = 1;
        for (int k : n.kids)
            for (int i = 0; i < ntax; ++i) has[v][i] |= has[k][i];
    }
    std::string out;
    for (int i = 0; i < ntax; ++i) {
        out += ">" + taxa[i] + "\n";
        for (int v = 0; v < nnodes; ++v) out += has[v][i] ? '1' : '0';
        out += "\n";
    }
    return out;
}

bool satisfies(const Topology& T, const Constraints& C) {
    if (C.empty()) return true;
    const size_t words = ((size_t)T.ntax + 63) / 64;
    // taxa behind every branch, seen from taxon 0 (depth-first from its neighbour; children before parents on the way back)
    std::vector<std::vector<uint64_t>> below;  // one set per directed edge away from taxon 0, in completion order
    std::vector<std::vector<uint64_t>> at((size_t)T.nnodes(), std::vector<uint64_t>(words, 0));
    struct Item { int v, from; bool done; };
    std::vector<Item> stack{{T.nbr[0][0], 0, false}};
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (T.is_tip(it.v)) {
            set_bit(at[it.v], it.v);
            below.push_back(at[it.v]);
            continue;
        }
        if (!it.done) {
            stack.push_back({it.v, it.from, true});
            for (int s = 0; s < 3; ++s)
                if (T.nbr[it.v][s] != it.from && T.nbr[it.v][s] >= 0) stack.push_back({T.nbr[it.v][s], it.v, false});
        } else {
            for (int s = 0; s < 3; ++s) {
                const int nb = T.nbr[it.v][s];
                if (nb == it.from || nb < 0) continue;
                for (size_t w = 0; w < words; ++w) at[it.v][w] |= at[nb][w];
            }
            below.push_back(at[it.v]);
        }
    }
    for (const SplitConstraint& sc : C.splits) {
        bool ok = false;
        for (const auto& S : below) {
            if ((subset(sc.one, S) && disjoint(sc.zero, S)) || (subset(sc.zero, S) && disjoint(sc.one, S))) {
                ok = true;
                break;
            }
        }
        if (!ok) return false;
    }
    return true;
}

std::vector<char> allowed_insertions(const GrowTree& g, const std::vector<int>& pre, int next_taxon, const Constraints& C) {
    const int N = (int)g.parent.size();
    const size_t words = ((size_t)C.ntax + 63) / 64;
    std::vector<char> allowed(pre.size(), 1);
    // taxa below every node (the root tip, node 0, is above everything and belongs to no set); `pre` lists parents first
    std::vector<std::vector<uint64_t>> desc((size_t)N, std::vector<uint64_t>(words, 0));
    std::vector<uint64_t> present(words, 0);
    set_bit(present, g.taxon[0]);
    for (size_t i = pre.size(); i-- > 0;) {
        const int v = pre[i];
        if (g.left[v] < 0) {
            set_bit(desc[v], g.taxon[v]);
            set_bit(present, g.taxon[v]);
        } else
            for (size_t w = 0; w < words; ++w) desc[v][w] = desc[g.left[v]][w] | desc[g.right[v]][w];
    }
    std::vector<int> anc_b((size_t)N, 0);
    std::vector<char> anc_a((size_t)N, 0), wit_a((size_t)N, 0), wit_b((size_t)N, 0);
    for (const SplitConstraint& sc : C.splits) {
        const bool in_one = bit(sc.one, next_taxon), in_zero = bit(sc.zero, next_taxon);
        if (!in_one && !in_zero) continue;  // the new taxon is free in this split: nothing it does can break it
        // A = the side the new taxon belongs to, B = the other side, both restricted to the taxa already in the tree
        std::vector<uint64_t> A(words), B(words);
        for (size_t w = 0; w < words; ++w) {
            A[w] = (in_one ? sc.one[w] : sc.zero[w]) & present[w];
            B[w] = (in_one ? sc.zero[w] : sc.one[w]) & present[w];
        }
        if (count_bits(A) == 0 || count_bits(B) == 0) continue;  // first of its side, or nobody to be separated from
        // witnesses of the split in the current tree: nodes holding all of A and none of B (they must take the new taxon in),
        // nodes holding all of B and none of A (they must not)
        int total_b = 0;
        for (int v : pre) {
            wit_a[v] = subset(A, desc[v]) && disjoint(B, desc[v]);
            wit_b[v] = subset(B, desc[v]) && disjoint(A, desc[v]);
            total_b += wit_b[v];
        }
        for (size_t i = 0; i < pre.size(); ++i) {  // parents first: what the strict ancestors offer
            const int v = pre[i], par = g.parent[v];
            const bool has_par = par > 0;  // node 0 is the root tip
            anc_a[v] = has_par && (anc_a[par] || wit_a[par]);
            anc_b[v] = has_par ? anc_b[par] + wit_b[par] : 0;
            // attached above v the new taxon joins v's set (in the new node) and every ancestor's
            const bool ok = wit_a[v] || anc_a[v] || (total_b - anc_b[v]) > 0;
            if (!ok) allowed[i] = 0;
        }
    }
    return allowed;
}

}  // namespace pml
