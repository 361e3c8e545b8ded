// peprml -- command-line stand-in for the raxmlHPC / raxmlHPC-PTHREADS executions PEPR performs.
//
// PEPR finds its tools through JVM system properties named after the executable (ExecUtilities.getCommandPath,
// src/edu/vt/vbi/ci/pepr/util/ExecUtilities.java:168-190; PhyloPipeline.setCommandPaths, PhyloPipeline.java:824-870), so
// pointing `-DraxmlHPC-PTHREADS=<this file>` (or installing it under that name) swaps the engine in with zero Java change.
// It speaks the file protocol of SURVEY.md section 8b: reads relaxed phylip / newick from the CWD, writes RAxML_info.<n>,
// RAxML_result.<n>, RAxML_log.<n>, RAxML_perSiteLLs.<n>, RAxML_bipartitions.<n>, RAxML_bipartitionsBranchLabels.<n>,
// RAxML_bestTree.<n>, RAxML_parsimonyTree.<n>, RAxML_bootstrap.<n>, <aln>.BS<k>.  Flags honoured: -f d|a|e|g|n|b|j,
// -m PROTGAMMAWAG, -s, -n, -t, -z, -a, -b, -x, -p, -#/-N, -e, -y, -T (accepted: the pattern-parallel workers of -T are the
// GPU's SMs here), -w.  PEPRML_GPUS=n spreads ONE call over n GPUs of the box the way `-T n` spreads raxmlHPC-PTHREADS over n
// cores (RAxMLRunner.java:130-132): the alignment patterns are sharded over a pml_group_create group, one host thread per GPU
// makes the same calls, rank 0 writes the files; results are the single-GPU results (sums are added in rank order).  `-f d` = parsimony start tree + lazy-SPR hill climbing + model optimisation; `-f a -x seed -N k` =
// k bootstrap replicates (raxmlHPC's weight stream, a quick search each) + ML search + supports drawn on the ML tree.
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "peprml.h"

namespace {

struct Args {
    std::string f = "d", model, aln, name, tree, trees, weights, workdir;
    long long bseed = 0, pseed = 12345;
    bool parsimony_only = false;
    int reps = 1, threads = 1;
    double eps = 0.1;
};

[[noreturn]] void die(const std::string& msg, int rc = 1) {
    std::fprintf(stderr, "peprml: %s\n", msg.c_str());
    std::exit(rc);
}

std::string slurp(const std::string& path) {
    std::ifstream in(path);
    if (!in) die("cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

std::vector<std::string> read_trees(const std::string& path) {
    std::vector<std::string> out;
    std::string all = slurp(path), cur;
    for (char ch : all) {
        if (ch == '\n' || ch == '\r') continue;
        cur.push_back(ch);
        if (ch == ';') {
            out.push_back(cur);
            cur.clear();
        }
    }
    return out;
}

bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

void check(pml_ctx* ctx, int rc, const char* what) {
    if (rc != PML_OK) die(std::string(what) + ": " + pml_last_error(ctx), 3);
}

std::string tree_string(pml_tree* t) {
    const int64_t n = pml_tree_newick(t, nullptr, 0);
    std::string s((size_t)n, '\0');
    pml_tree_newick(t, s.data(), (size_t)n);
    s.resize(std::strlen(s.c_str()));
    return s;
}

// rendezvous of the rank threads of one call (PEPRML_GPUS > 1)
class Meeting {
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0, generation = 0;
    const int n;

public:
    explicit Meeting(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lock(mu);
        const int gen = generation;
        if (++waiting == n) {
            waiting = 0;
            ++generation;
            cv.notify_all();
        } else cv.wait(lock, [&] { return generation != gen; });
    }
};

double tree_length(pml_tree* t) {
    double sum = 0.0;
    for (int e = 0; e < pml_tree_num_branches(t); ++e) {
        double l;
        pml_tree_branch(t, e, nullptr, nullptr, &l);
        sum += l;
    }
    return sum;
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) die("missing value after " + k);
            return argv[++i];
        };
        if (k == "-f") a.f = val();
        else if (k == "-m") a.model = val();
        else if (k == "-s") a.aln = val();
        else if (k == "-n") a.name = val();
        else if (k == "-t") a.tree = val();
        else if (k == "-z") a.trees = val();
        else if (k == "-a") a.weights = val();
        else if (k == "-w") a.workdir = val();
        else if (k == "-b" || k == "-x") a.bseed = std::atoll(val().c_str());
        else if (k == "-p") a.pseed = std::atoll(val().c_str());
        else if (k == "-y") a.parsimony_only = true;
        else if (k == "-#" || k == "-N") a.reps = std::atoi(val().c_str());
        else if (k == "-T") a.threads = std::atoi(val().c_str());
        else if (k == "-e") a.eps = std::atof(val().c_str());
        else if (k == "-v") {
            std::printf("%s (raxmlHPC-compatible front end)\n", pml_version());
            return 0;
        } else if (k == "-Y" || k == "-k" || k == "-d" || k == "-D" || k == "-F" || k == "-j" || k == "-M") {
            // accepted, no effect on the modes implemented here
        } else die("unknown option " + k);
    }
    if (a.name.empty()) die("-n runName is required");
    if (a.model.empty()) die("-m model is required");
    const std::string dir = a.workdir.empty() ? std::string() : a.workdir + "/";
    const std::string info_path = dir + "RAxML_info." + a.name;
    if (exists(info_path)) die("RAxML output files with the run ID <" + a.name + "> already exist", 1);  // raxmlHPC refuses too

    const auto t_start = std::chrono::steady_clock::now();
    std::ofstream info_file(info_path);
    std::ostream& info = info_file;
    info << "\n\nThis is " << pml_version() << ", answering for RAxML's command line.\n\n";

    if (a.f == "b") {  // draw bipartition support of the -z trees on the -t tree; host-only integer path
        if (a.tree.empty() || a.trees.empty()) die("-f b needs -t and -z");
        std::vector<std::string> sup = read_trees(a.trees);
        std::vector<const char*> ptr;
        for (auto& s : sup) ptr.push_back(s.c_str());
        const std::string main_tree = read_trees(a.tree).at(0);
        const int64_t n = pml_support_tree(main_tree.c_str(), ptr.data(), (int)ptr.size(), 1, nullptr, 0);
        if (n < 0) die(std::string("-f b: ") + pml_last_error(nullptr), 3);
        std::string out((size_t)n, '\0');
        pml_support_tree(main_tree.c_str(), ptr.data(), (int)ptr.size(), 1, out.data(), (size_t)n);
        out.resize(std::strlen(out.c_str()));
        std::ofstream(dir + "RAxML_bipartitions." + a.name) << out << "\n";
        // branch-label flavour: `:len[support]`
        std::string lab;
        for (size_t i = 0; i < out.size(); ++i) {
            if (out[i] == ')' && i + 1 < out.size() && std::isdigit((unsigned char)out[i + 1])) {
                size_t j = i + 1;
                while (j < out.size() && std::isdigit((unsigned char)out[j])) ++j;
                const std::string support = out.substr(i + 1, j - i - 1);
                size_t k = j;
                if (k < out.size() && out[k] == ':') {
                    ++k;
                    while (k < out.size() && !std::strchr(",);", out[k])) ++k;
                }
                lab += ")" + out.substr(j, k - j) + "[" + support + "]";
                i = k - 1;
            } else lab.push_back(out[i]);
        }
        std::ofstream(dir + "RAxML_bipartitionsBranchLabels." + a.name) << lab << "\n";
        info << "Found " << sup.size() << " trees in File " << a.trees << "\n";
        return 0;
    }

    if (a.aln.empty()) die("-s alignment is required");
    int ngpu = 1;
    if (const char* g = std::getenv("PEPRML_GPUS")) ngpu = std::max(1, std::atoi(g));
    std::vector<pml_ctx*> ctxs((size_t)ngpu, nullptr);
    if (ngpu == 1) {
        if (pml_ctx_create(0, 0, 1, nullptr, &ctxs[0]) != PML_OK) die(std::string("cannot create GPU context: ") + pml_last_error(nullptr), 3);
    } else {
        std::vector<int> ids((size_t)ngpu);
        for (int i = 0; i < ngpu; ++i) ids[i] = i;
        if (pml_group_create(ids.data(), ngpu, ctxs.data()) != PML_OK)
            die(std::string("cannot create a group of ") + std::to_string(ngpu) + " GPUs: " + pml_last_error(nullptr), 3);
    }
    std::ofstream nowhere;  // never opened: swallows what the ranks other than 0 would write
    Meeting meeting(ngpu);
    std::vector<std::vector<double>> per_site_of((size_t)ngpu);  // -f g: every rank holds the values of its own patterns' columns
    // one rank of the call: every rank makes the same engine calls (a branch pass waits for its peers' sums), rank 0 owns the files
    auto work = [&](int rank) {
    pml_ctx* ctx = ctxs[(size_t)rank];
    const bool lead = rank == 0;
    std::ostream& info = lead ? static_cast<std::ostream&>(info_file) : nowhere;
    auto out_file = [&](const std::string& path) { return lead ? std::ofstream(path) : std::ofstream(); };
    pml_aln* aln = nullptr;
    check(ctx, pml_aln_load_phylip(ctx, a.aln.c_str(), a.weights.empty() ? nullptr : a.weights.c_str(), &aln), "alignment");
    int ntax;
    int64_t nsites, npat;
    pml_aln_dims(aln, &ntax, &nsites, &npat, nullptr);
    check(ctx, pml_model_set(aln, a.model.c_str(), 1.0), "model");
    info << "Alignment has " << npat << " distinct alignment patterns\n\n"
         << "RAxML was called as follows:\n\n";
    for (int i = 0; i < argc; ++i) info << argv[i] << " ";
    info << "\n\n";

    if (a.f == "j") {  // bootstrapped alignments, columns in sorted-pattern order, as raxmlHPC -f j -b seed -# n writes them
        std::vector<int32_t> pw((size_t)npat), W((size_t)npat * a.reps);
        std::vector<int64_t> s2p((size_t)nsites);
        pml_aln_patterns(aln, pw.data(), s2p.data());
        int64_t seed = a.bseed;
        check(ctx, pml_bootstrap_weights(aln, &seed, a.reps, W.data()), "bootstrap weights");
        std::vector<int64_t> rep_col((size_t)npat, -1);  // a representative original column of each pattern
        for (int64_t s = 0; s < nsites; ++s)
            if (s2p[s] >= 0 && rep_col[s2p[s]] < 0) rep_col[s2p[s]] = s;
        std::vector<std::string> rows;
        {
            std::ifstream in(a.aln);
            long nt, ns;
            in >> nt >> ns;
            for (int t = 0; t < ntax; ++t) {
                std::string nm, sq, piece;
                in >> nm;
                while ((int64_t)sq.size() < nsites && in >> piece) sq += piece;
                rows.push_back(sq);
            }
        }
        for (int r = 0; r < a.reps; ++r) {
            std::ofstream out = out_file(a.aln + ".BS" + std::to_string(r));
            out << ntax << " " << nsites << "\n";
            for (int t = 0; t < ntax; ++t) {
                out << pml_aln_name(aln, t) << " ";
                for (int64_t p = 0; p < npat; ++p)
                    for (int32_t k = 0; k < W[(size_t)r * npat + p]; ++k) out << rows[t][rep_col[p]];
                out << "\n";
            }
        }
        return;
    }

    if (a.f == "d" || a.f == "a" || a.f == "o") {
        // ---- ML tree search (RAxMLRunner.run: `-f d`, or `-f a -x seed -N reps` when bootstrapReps > 0) ----------------
        auto search_tree = [&](const int32_t* w, int64_t pseed, int rounds, bool final_opt, double* lnl_out, double* alpha_out) {
            pml_tree* t = nullptr;
            check(ctx, pml_model_set(aln, a.model.c_str(), 1.0), "model");
            check(ctx, pml_tree_start_parsimony(aln, pseed, w, &t), "parsimony start tree");
            double lnl = 0.0, alpha = 1.0;
            check(ctx, pml_optimize(t, 1, 5.0, w, &lnl, &alpha), "initial optimisation");
            int moves = 0;
            check(ctx, pml_search(t, 5, rounds, a.eps, w, &lnl, &moves), "tree search");
            if (final_opt) check(ctx, pml_optimize(t, 1, a.eps, w, &lnl, &alpha), "final optimisation");
            if (lnl_out) *lnl_out = lnl;
            if (alpha_out) *alpha_out = alpha;
            return t;
        };
        if (a.parsimony_only) {  // `-y`: stop after the parsimony start tree (RAxMLRunner.runRaxmlParsimonyWithBranchLengths, 1st run)
            pml_tree* t = nullptr;
            check(ctx, pml_tree_start_parsimony(aln, a.pseed, nullptr, &t), "parsimony start tree");
            std::string s = tree_string(t);
            out_file(dir + "RAxML_parsimonyTree." + a.name) << s << "\n";
            return;
        }
        std::vector<std::string> boots;
        if (a.f == "a") {
            std::vector<int32_t> W((size_t)npat * a.reps);
            int64_t seed = a.bseed;
            check(ctx, pml_bootstrap_weights(aln, &seed, a.reps, W.data()), "bootstrap weights");
            std::ofstream bs = out_file(dir + "RAxML_bootstrap." + a.name);
            for (int r = 0; r < a.reps; ++r) {
                pml_tree* t = search_tree(W.data() + (size_t)r * npat, a.pseed + 1 + r, 1, false, nullptr, nullptr);
                boots.push_back(tree_string(t));
                bs << boots.back() << "\n";
                pml_tree_free(t);
            }
        }
        double lnl = 0.0, alpha = 1.0;
        pml_tree* t = search_tree(nullptr, a.pseed, 10, true, &lnl, &alpha);
        const std::string best = tree_string(t);
        out_file(dir + "RAxML_result." + a.name) << best << "\n";
        out_file(dir + "RAxML_bestTree." + a.name) << best << "\n";
        char buf[64];
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        std::snprintf(buf, sizeof buf, "%.6f", lnl);
        out_file(dir + "RAxML_log." + a.name) << secs << " " << buf << "\n";
        info << "Final GAMMA-based Score of best tree " << buf << "\n";
        info << "Final GAMMA  likelihood: " << buf << "\n";
        std::snprintf(buf, sizeof buf, "%.6f", alpha);
        info << "alpha: " << buf << "\n";
        std::snprintf(buf, sizeof buf, "%.6f", tree_length(t));
        info << "Tree-Length: " << buf << "\n";
        info << "Overall execution time: " << secs << " secs\n";
        if (a.f == "a") {
            std::vector<const char*> ptr;
            for (auto& s : boots) ptr.push_back(s.c_str());
            const int64_t n = pml_support_tree(best.c_str(), ptr.data(), (int)ptr.size(), 1, nullptr, 0);
            if (n < 0) die(std::string("supports: ") + pml_last_error(nullptr), 3);
            std::string out((size_t)n, '\0');
            pml_support_tree(best.c_str(), ptr.data(), (int)ptr.size(), 1, out.data(), (size_t)n);
            out.resize(std::strlen(out.c_str()));
            out_file(dir + "RAxML_bipartitions." + a.name) << out << "\n";
        }
        pml_tree_free(t);
        pml_aln_free(aln);
        return;
    }

    std::vector<std::string> trees;
    if (a.f == "e") {
        if (a.tree.empty()) die("-f e needs -t");
        trees = read_trees(a.tree);
        trees.resize(1);
    } else if (a.f == "g" || a.f == "n") {
        if (a.trees.empty()) die("-f " + a.f + " needs -z");
        trees = read_trees(a.trees);
    } else die("unsupported algorithm -f " + a.f, 2);

    std::ofstream result = out_file(dir + "RAxML_result." + a.name), logf = out_file(dir + "RAxML_log." + a.name);
    std::ofstream persite;
    if (a.f == "g") {
        if (lead) persite.open(dir + "RAxML_perSiteLLs." + a.name);
        persite << "  " << trees.size() << "  " << nsites << "\n";
    }
    char buf[64];
    for (size_t i = 0; i < trees.size(); ++i) {
        pml_tree* t = nullptr;
        check(ctx, pml_model_set(aln, a.model.c_str(), 1.0), "model");
        check(ctx, pml_tree_load(aln, trees[i].c_str(), &t), "tree");
        // raxmlHPC starts every branch at its default z = 0.9 regardless of the lengths in the file
        for (int e = 0; e < pml_tree_num_branches(t); ++e) pml_tree_set_branch(t, e, -std::log(0.9));
        double lnl = 0.0, alpha = 1.0;
        check(ctx, pml_optimize(t, 1, a.eps, nullptr, &lnl, &alpha), "optimisation");
        result << tree_string(t) << "\n";
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        std::snprintf(buf, sizeof buf, "%.6f", lnl);
        logf << secs << " " << buf << "\n";
        if (a.f == "g") {
            std::vector<double>& ps = per_site_of[(size_t)rank];
            ps.assign((size_t)nsites, 0.0);
            check(ctx, pml_evaluate(t, nullptr, &lnl, ps.data()), "per-site lnL");
            meeting.wait();  // a column's value is non-zero on exactly one rank: rank 0 adds the ranks' vectors
            if (lead) {
                char cell[64];
                persite << "tr" << (i + 1) << "\t";
                for (int64_t s = 0; s < nsites; ++s) {
                    double v = 0.0;
                    for (int r = 0; r < ngpu; ++r) v += per_site_of[(size_t)r][(size_t)s];
                    std::snprintf(cell, sizeof cell, "%.6f ", v);
                    persite << cell;
                }
                persite << "\n";
            }
            meeting.wait();
        }
        if (a.f == "n") info << "Tree " << i << " Likelihood " << buf << " Tree-Length " << tree_length(t) << "\n";
        if (i + 1 == trees.size()) {
            info << "\nOverall Time for Tree Evaluation " << secs << "\n";
            info << "Final GAMMA  likelihood: " << buf << "\n\n";
            info << "Model Parameters of Partition 0, Name: No Name Provided, Type of Data: AA\n";
            std::snprintf(buf, sizeof buf, "%.6f", alpha);
            info << "alpha: " << buf << "\n";
            std::snprintf(buf, sizeof buf, "%.6f", tree_length(t));
            info << "Tree-Length: " << buf << "\n";
        }
        pml_tree_free(t);
    }
    pml_aln_free(aln);
    };
    if (ngpu == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int r = 0; r < ngpu; ++r) pool.emplace_back(work, r);
        for (auto& th : pool) th.join();
    }
    for (pml_ctx* c : ctxs) pml_ctx_destroy(c);
    return 0;
}
