// Engine: contexts, resident alignments, trees with their CLV arenas, the likelihood / Newton-Raphson / optimisation
// drivers, and the extern "C" surface declared in include/peprml.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/peprml.h"
#include "host.h"
#include "kernels.h"
#include "model.h"
#include "pmatrix.cuh"

using namespace pml;

namespace {

thread_local std::string g_create_error;

// NCCL is bound at run time and only when a context joins a group of ranks: single-GPU use has no NCCL dependency.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) {
            err = std::string("cannot load libnccl.so.2: ") + dlerror();
            return false;
        }
        GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
            err = "libnccl.so.2 lacks required symbols";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;

}  // namespace

// Ranks of a site-sharded group that live in ONE process (pml_group_create): what they share.
struct PmlGroup {
    int n = 0;
    std::vector<int> devices;
    std::vector<double*> mail;  // one mailbox per rank, on that rank's device, written by all peers through NVLink peer access
    // pattern crunch of the alignment the ranks are loading: done once by the first rank that asks (see pml_aln_load)
    std::mutex mu;
    uint64_t crunch_key[5] = {0, 0, 0, 0, 0};
    int crunch_uses = 0;
    std::shared_ptr<const Patterns> crunch;
    ~PmlGroup() {
        for (int r = 0; r < (int)mail.size(); ++r)
            if (mail[r]) {
                cudaSetDevice(devices[r]);
                cudaFree(mail[r]);
            }
    }
};

struct pml_ctx {
    int device = 0, rank = 0, nranks = 1, sms = 148;
    std::shared_ptr<PmlGroup> group;  // set for ranks created by pml_group_create
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    std::string err;
    // pinned staging for small host->device parameter blocks and device->host results
    uint8_t* h_stage = nullptr;
    size_t stage_cap = 0, stage_used = 0;
    double* h_result = nullptr;
    // zero-copy result ring in mapped pinned memory: a branch pass publishes {lnL, d1, d2, length, NR status, sequence} in
    // slot (sequence mod kRing) and the host polls the sequence -- no D2H copy, no stream synchronisation per branch
    static constexpr int kRing = 8;
    volatile double* h_mapped = nullptr;
    double* d_mapped = nullptr;
    double sequence = 0.0;
    int* d_poison = nullptr;  // see Publish in kernels.h
    // in-kernel reduction over NVLink peer memory (see PeerReduce in kernels.h); falls back to NCCL when IPC is unavailable
    double* d_mail = nullptr;
    double** d_mail_ptrs = nullptr;
    std::vector<void*> peer_opened;
    bool peer_ok = false;
    bool fold = true;      // PEPRML_NO_FOLD=1: cherries are stored like every other inner node (reference mode for tests)
    bool fuse = true;      // PEPRML_NO_FUSE=1: CLV update and branch pass always as two launches (reference mode for tests)
    bool host_nr = false;  // PEPRML_HOST_NR=1: Newton-Raphson steps on the host, one wait per branch (reference mode for tests)
    volatile double* slot_host(double seq) const { return h_mapped + ((int64_t)seq % kRing) * kSlotDoubles; }
    double* slot_dev(double seq) const { return d_mapped + ((int64_t)seq % kRing) * kSlotDoubles; }
    // waits until the pass with this sequence number has published; out = {lnL, d1, d2, length, status}
    bool comm_lost = false;  // a pass reported kNrCommLost: the group is unusable, every later call fails with PML_ECOMM
    bool wait_slot(double seq, double out[5]) {
        const volatile unsigned long long* s = reinterpret_cast<const volatile unsigned long long*>(slot_host(seq));
        const uint32_t flag = ll_flag(seq);
        long spins = 0;
        for (int i = 0; i < 5; ++i) {
            while (!ll_try_read_host(s + 2 * i, flag, &out[i])) {
                if ((++spins & 0xFFFFF) == 0 && cudaStreamQuery(stream) != cudaErrorNotReady) {
                    if (ll_try_read_host(s + 2 * i, flag, &out[i])) break;
                    cuda(cudaStreamSynchronize(stream), "branch kernel");
                    if (err.empty()) err = "branch kernel finished without publishing its result";
                    return false;
                }
            }
        }
        if ((int)out[4] == kNrCommLost) {
            comm_lost = true;
            err = "peer reduction timed out: a rank of the group did not deliver its sums (crashed, failed before its launch, "
                  "or out of step); the group must be destroyed";
            return false;
        }
        return true;
    }
    int fail_code() const { return comm_lost || err.rfind("nccl", 0) == 0 ? PML_ECOMM : PML_ENODEVICE; }
    int* d_peer_lost = nullptr;            // sticky device flag of PeerReduce
    unsigned long long peer_timeout_ns = 10ull * 1000 * 1000 * 1000;  // PEPRML_PEER_TIMEOUT_MS
    // optional per-launch device timing (pml_profile_begin/end)
    struct Timed { int kind; int64_t rows; cudaEvent_t t0, t1; };
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    long long* d_trace = nullptr;  // pml_trace_enable (1: CLV kernel, 2: branch kernel)
    long long* d_trace_buf = nullptr;
    long long* d_trace_branch = nullptr;
    bool trace_branch_tip = false;
    bool trace_fused = false;
    int trace_newview_tips = -1;  // -1: every CLV kernel, 0 / 1: only those with that many tip children
    bool profiling = false;
    std::vector<Timed> timed;
    // pml_timeline_begin/read (profiling aid): six %globaltimer stamps per CLV / fused launch, in launch order
    static constexpr int kTimelineCap = 16384;
    unsigned long long* d_timeline = nullptr;
    std::vector<int> timeline_kinds;
    unsigned long long* timeline_slot(int kind) {
        if (!d_timeline || (int)timeline_kinds.size() >= kTimelineCap) return nullptr;
        timeline_kinds.push_back(kind);
        return d_timeline + 6 * (timeline_kinds.size() - 1);
    }
    std::vector<cudaEvent_t> spare_events;
    cudaEvent_t event() {
        cudaEvent_t e;
        if (!spare_events.empty()) { e = spare_events.back(); spare_events.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    int tick(int kind, int64_t rows) {  // call before a launch; returns index for tock()
        if (!profiling) return -1;
        Timed t{kind, rows, event(), event()};
        cudaEventRecord(t.t0, stream);
        timed.push_back(t);
        return (int)timed.size() - 1;
    }
    void tock(int idx) { if (idx >= 0) cudaEventRecord(timed[idx].t1, stream); }

    // Device blocks released by pml_aln_free / pml_tree_free stay cached in the context: PEPR calls the runner once per tree
    // (hundreds of times per refinement round), and a fresh cudaMalloc of a multi-GB CLV arena costs ~100 ms each time.
    std::vector<std::pair<size_t, void*>> cached_blocks;
    template <typename T>
    cudaError_t dev_alloc(T** out, size_t bytes) {
        bytes = (bytes + 511) & ~size_t(511);
        size_t best = cached_blocks.size();
        for (size_t i = 0; i < cached_blocks.size(); ++i)
            if (cached_blocks[i].first >= bytes && cached_blocks[i].first <= bytes + bytes / 4 + 4096 &&
                (best == cached_blocks.size() || cached_blocks[i].first < cached_blocks[best].first))
                best = i;
        if (best != cached_blocks.size()) {
            *out = (T*)cached_blocks[best].second;
            sizes[(void*)*out] = cached_blocks[best].first;
            cached_blocks.erase(cached_blocks.begin() + best);
            return cudaSuccess;
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess && !cached_blocks.empty()) {  // out of memory: drop the cache and retry once
            cudaGetLastError();
            for (auto& b : cached_blocks) cudaFree(b.second);
            cached_blocks.clear();
            e = cudaMalloc(&p, bytes);
        }
        if (e == cudaSuccess) {
            *out = (T*)p;
            sizes[p] = bytes;
        }
        return e;
    }
    void dev_free(void* p) {
        if (!p) return;
        auto it = sizes.find(p);
        if (it == sizes.end()) {
            cudaFree(p);
            return;
        }
        cached_blocks.push_back({it->second, p});
        sizes.erase(it);
    }
    std::unordered_map<void*, size_t> sizes;

    bool cuda(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    }
    bool bind() { return cuda(cudaSetDevice(device), "cudaSetDevice"); }
    bool sync() {
        if (!cuda(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
        stage_used = 0;
        return true;
    }
    // bump allocator over the pinned block; a wrap-around waits for the stream so in-flight copies are never overwritten
    void* stage(size_t bytes) {
        bytes = (bytes + 63) & ~size_t(63);
        if (stage_used + bytes > stage_cap) {
            if (!sync() || bytes > stage_cap) return nullptr;
        }
        void* p = h_stage + stage_used;
        stage_used += bytes;
        return p;
    }
    // sums n doubles at d_buf over all ranks (no-op for a single rank)
    bool allreduce(double* d_buf, int n) {
        if (nranks == 1) return true;
        ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, comm, stream);
        if (r != ncclSuccess) {
            err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
            return false;
        }
        return true;
    }
};

struct pml_aln {
    pml_ctx* ctx = nullptr;
    Patterns pat;
    int64_t p0 = 0, nloc = 0, npad = 0;  // this rank's pattern range [p0, p0+nloc), padded row count
    uint8_t* d_codes = nullptr;          // ntax x npad
    int32_t* d_weights = nullptr;        // npad (alignment's own)
    int32_t* d_wcustom = nullptr;        // npad (caller-supplied replicate)
    DeviceModel* d_model = nullptr;
    bool model_set = false;
    double alpha = 1.0, rates[kCats] = {1, 1, 1, 1};
    double alpha_step = 0.1;             // first probe distance (in log alpha) of the next alpha optimisation
    double* d_site_lnl = nullptr;
    double* d_partials = nullptr;
    double* d_result = nullptr;  // 16 doubles
    unsigned int* d_ticket = nullptr;
    double* d_sumtable = nullptr;
    int32_t* d_sumscale = nullptr;
    // The model constants and the product table are shared by every tree of this alignment, their validity is not:
    // model_epoch counts model uploads (a tree's views are only good for the epoch they were built under), and the
    // product table belongs to the tree that filled it last.
    uint64_t model_epoch = 0;
    const pml_tree* sumtable_owner = nullptr;
    Constraints constraints;  // pml_aln_set_constraints: splits every tree built or searched on this alignment must display
};

struct pml_tree {
    pml_aln* aln = nullptr;
    Topology topo;
    ViewState views;
    double* d_clv = nullptr;      // (ntax-2) x npad x 80
    int32_t* d_scale = nullptr;   // (ntax-2) x npad
    double* d_len = nullptr;      // branch lengths as the kernels read them (one double per branch id)
    std::vector<double> len_dev;  // what d_len holds (mirror), so that host-side edits of topo.len are uploaded lazily
    int64_t nr_retries = 0;       // Newton-Raphson passes that ended in the bad-curvature retry
    int last_swept = -1;          // branch the last smoothing sweep ended on
    int prepared_branch = -1;     // branch whose sumtable is resident (only while aln->sumtable_owner == this)
    uint64_t model_epoch = 0;     // aln->model_epoch the stored views were computed under
    int64_t site_updates[3] = {0, 0, 0};
    int64_t launches = 0;

    double* clv(int v) const { return d_clv + (size_t)(v - topo.ntax) * aln->npad * kRow; }
    int32_t* scale(int v) const { return d_scale + (size_t)(v - topo.ntax) * aln->npad; }
    Side side(int v) const {
        Side s{};
        if (topo.is_tip(v)) s.codes = aln->d_codes + (size_t)v * aln->npad;
        else {
            s.clv = clv(v);
            s.scale = scale(v);
        }
        return s;
    }
    // the view of v looking at its neighbour `toward`: a folded cherry is described by its two tips (never stored)
    Side side_toward(int v, int toward) const {
        if (!views.fold_cherries || !ViewState::is_cherry_view(topo, v, toward)) return side(v);
        Side s{};
        for (int k = 0; k < 3; ++k) {
            const int nb = topo.nbr[v][k];
            if (nb == toward) continue;
            const uint8_t* codes = aln->d_codes + (size_t)nb * aln->npad;
            if (!s.codes) {
                s.codes = codes;
                s.len1 = d_len + topo.edge[v][k];
            } else {
                s.codes2 = codes;
                s.len2 = d_len + topo.edge[v][k];
            }
        }
        return s;
    }
};

namespace {

// 64-bit content hash (8 bytes per step) -- the key under which ranks of one process share a pattern crunch
uint64_t hash_bytes(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 29;
    }
    for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

constexpr int kPad = 128;            // pattern rows are padded to a multiple of this (tile height of the CLV kernels)
constexpr double kDefaultLen = 0.1;

bool upload_model(pml_aln* a) {
    pml_ctx* c = a->ctx;
    const Eigensystem& es = wag_eigensystem();
    auto* h = (DeviceModel*)c->stage(sizeof(DeviceModel));
    if (!h) return false;
    std::memcpy(h->lambda, es.lambda, sizeof es.lambda);
    std::memcpy(h->V, es.V, sizeof es.V);
    std::memcpy(h->Vinv, es.Vinv, sizeof es.Vinv);
    std::memcpy(h->pi, es.pi, sizeof es.pi);
    std::memcpy(h->rates, a->rates, sizeof a->rates);
    for (int i = 0; i < kStates; ++i)
        for (int k = 0; k < kStates; ++k) h->piV[i][k] = es.pi[i] * es.V[i][k];
    for (int code = 0; code < kCodes; ++code)
        for (int k = 0; k < kStates; ++k) {
            double acc = 0.0;  // same summation order as the kernels used when they built the table themselves
            if (code < 20) acc = h->piV[code][k];
            else if (code == 20) acc = h->piV[2][k] + h->piV[3][k];
            else if (code == 21) acc = h->piV[5][k] + h->piV[6][k];
            else
                for (int i = 0; i < kStates; ++i) acc += h->piV[i][k];
            h->tipvec[code][k] = acc;
        }
    for (int k = 0; k < kStates; ++k) {
        double all = 0.0;
        for (int j = 0; j < kStates; ++j) all += es.Vinv[k][j];
        h->vinv_codes[k][0] = es.Vinv[k][2] + es.Vinv[k][3];
        h->vinv_codes[k][1] = es.Vinv[k][5] + es.Vinv[k][6];
        h->vinv_codes[k][2] = all;
        h->vinv_codes[k][3] = 0.0;
    }
    ++a->model_epoch;  // every tree of this alignment drops its views before it plans again (adopt_model)
    return c->cuda(cudaMemcpyAsync(a->d_model, h, sizeof(DeviceModel), cudaMemcpyHostToDevice, c->stream), "model upload");
}

// uploads a caller weight vector (global pattern order) for this rank's slice; returns the device pointer to use
const int32_t* device_weights(pml_aln* a, const int32_t* weights) {
    if (!weights) return a->d_weights;
    pml_ctx* c = a->ctx;
    if (a->nloc > 0) {
        // pageable source: cudaMemcpyAsync stages it before returning, so the caller's buffer may be reused at once
        if (!c->cuda(cudaMemcpyAsync(a->d_wcustom, weights + a->p0, sizeof(int32_t) * a->nloc, cudaMemcpyHostToDevice, c->stream),
                     "weights upload"))
            return nullptr;
    }
    return a->d_wcustom;
}

// A tree's stored views and its claim on the product table are void once the alignment's model changed under it (another
// tree's alpha optimisation, pml_model_set) or another tree filled the table.  Called by every entry point before planning.
void adopt_model(pml_tree* t) {
    pml_aln* a = t->aln;
    if (t->model_epoch != a->model_epoch) {
        t->views.reset(t->topo);
        t->prepared_branch = -1;
        t->model_epoch = a->model_epoch;
    }
    if (a->sumtable_owner != t) t->prepared_branch = -1;
}

// host-side edits of branch lengths (set_branch, SPR moves, a freshly loaded tree) reach the device before the next launch
bool sync_lengths(pml_tree* t) {
    pml_ctx* c = t->aln->ctx;
    const std::vector<double>& len = t->topo.len;
    if (t->len_dev.size() == len.size() && std::memcmp(t->len_dev.data(), len.data(), sizeof(double) * len.size()) == 0) return true;
    auto* h = (double*)c->stage(sizeof(double) * len.size());
    if (!h) return false;
    std::memcpy(h, len.data(), sizeof(double) * len.size());
    if (!c->cuda(cudaMemcpyAsync(t->d_len, h, sizeof(double) * len.size(), cudaMemcpyHostToDevice, c->stream), "lengths upload")) return false;
    t->len_dev = len;
    return true;
}

NewviewOp make_newview_op(pml_tree* t, const ViewOp& op) {
    NewviewOp nv{};
    nv.left = t->side_toward(op.child[0], op.node);
    nv.right = t->side_toward(op.child[1], op.node);
    nv.len_left = t->d_len + op.cedge[0];
    nv.len_right = t->d_len + op.cedge[1];
    nv.len_scale = 1.0;
    nv.dm = t->aln->d_model;
    nv.out = t->clv(op.node);
    nv.out_scale = t->scale(op.node);
    return nv;
}

// Kernel kinds of pml_profile_begin/end (PML_NKINDS in peprml.h; names, algorithmic bytes and DMMA counts: pml_kind_info).
// A side is an inner node (I), a tip (T) or a folded cherry (C).
//   0..5   CLV update by its children, ordered: TT, TI, II, TC, CC, CI
//   6..8   branch pass (one end inner), by the other end: I, T, C        9..11  the same as root evaluate (per-pattern lnL kept)
//   12     Newton-Raphson iteration on a stored product table
//   13..27 fused update + branch pass: 13 + 3 * (update's children: TI, II, TC, CC, CI) + (far end: I, T, C)
int side_rank(const Side& s) {
    const int k = side_kind(s);
    return k == kSideTip ? 0 : (k == kSideCherry ? 1 : 2);
}
int newview_kind(const Side& l, const Side& r) {
    static const int table[3][3] = {{0, 3, 1}, {3, 4, 5}, {1, 5, 2}};  // [rank][rank]: T, C, I
    return table[side_rank(l)][side_rank(r)];
}
int far_index(const Side& s) { return side_kind(s) == kSideInner ? 0 : (side_kind(s) == kSideTip ? 1 : 2); }

// one CLV kernel; counts the site-updates by the number of children that are not inner nodes (tips or folded cherries)
void launch_newview(pml_tree* t, const NewviewOp& nv, bool count = true) {
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    const int ntip = (nv.left.clv == nullptr) + (nv.right.clv == nullptr);
    const int tk = c->tick(newview_kind(nv.left, nv.right), a->nloc);
    NewviewOp timed_nv = nv;
    timed_nv.timeline = c->timeline_slot(newview_kind(nv.left, nv.right));
    launch_newview_mma(timed_nv, a->npad, c->sms, c->stream);
    c->tock(tk);
    ++t->launches;
    if (count) t->site_updates[2 - ntip] += a->nloc;
}

// executes a traversal descriptor: one CLV kernel per entry; each builds its own P matrices from the device lengths
bool run_ops(pml_tree* t, const std::vector<ViewOp>& ops) {
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    if (!sync_lengths(t)) return false;
    for (const ViewOp& op : ops) {
        NewviewOp nv = make_newview_op(t, op);
        const int ntip = (nv.left.clv == nullptr) + (nv.right.clv == nullptr);
        static const char* only = getenv("PEPRML_TRACE_KIND");  // profiling aid: trace CLV kernels of one kind only (launch_newview)
        nv.trace = (only ? atoi(only) == newview_kind(nv.left, nv.right) : (c->trace_newview_tips < 0 || c->trace_newview_tips == ntip)) ? c->d_trace : nullptr;
        launch_newview(t, nv);
    }
    return ops.empty() || c->cuda(cudaGetLastError(), "CLV kernels");
}

// Stores the folded view of cherry v (looking at `toward`) in v's own CLV slot after all -- the tip-tip kernel -- for the few
// consumers that cannot form it themselves; returns the inner-node side.
Side materialize_cherry(pml_tree* t, int v, int toward) {
    const Topology& T = t->topo;
    NewviewOp nv{};
    int k = 0;
    for (int q = 0; q < 3; ++q) {
        if (T.nbr[v][q] == toward) continue;
        (k == 0 ? nv.left : nv.right) = t->side(T.nbr[v][q]);
        (k == 0 ? nv.len_left : nv.len_right) = t->d_len + T.edge[v][q];
        ++k;
    }
    nv.len_scale = 1.0;
    nv.dm = t->aln->d_model;
    nv.out = t->clv(v);
    nv.out_scale = t->scale(v);
    launch_newview(t, nv, false);  // planning has already counted this update as a folded one
    t->views.orient[v - T.ntax] = T.slot_of(v, toward);
    return t->side(v);
}

// brings both ends of branch e up to date; (a, b) is returned with b inner and a the tip end if there is one
bool orient_branch(pml_tree* t, int e, int& a, int& b) {
    a = t->topo.ea[e];
    b = t->topo.eb[e];
    if (t->topo.is_tip(b)) std::swap(a, b);
    std::vector<ViewOp> ops;
    t->views.plan(t->topo, a, b, ops);
    t->views.plan(t->topo, b, a, ops);
    return run_ops(t, ops);
}


bool fetch_result(pml_ctx* c, const double* d_result, int n, double* out) {
    if (!c->cuda(cudaMemcpyAsync(c->h_result, d_result, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream), "result download"))
        return false;
    if (!c->sync()) return false;
    std::memcpy(out, c->h_result, sizeof(double) * n);
    return true;
}

bool ensure_sumtable(pml_aln* a) {
    pml_ctx* c = a->ctx;
    if (a->d_sumtable) return true;
    return c->cuda(c->dev_alloc(&a->d_sumtable, sizeof(double) * a->npad * kRow), "sumtable alloc") &&
           c->cuda(c->dev_alloc(&a->d_sumscale, sizeof(int32_t) * a->npad), "sumtable scale alloc");
}

// One pass over the two CLVs at the ends of branch e: {lnL, dlnL/dt, d2lnL/dt2} summed over ranks, published in the
// context's result ring under the returned sequence number (0 on failure).
//   device_nr = false: sums at length len; keep_table stores the eigen-space product table so that further lengths can be
//               tried without re-reading the CLVs; site_lnl fills the per-pattern lnL buffer (root evaluate)
//   device_nr = true:  sums at the branch's length in the tree's device array, followed on the device by the guarded
//               Newton-Raphson step that overwrites it (see Publish in kernels.h); len is ignored
enum : int { kWantLnl = 1, kWantDerivs = 2, kWantAll = 3 };
double branch_launch_sides(pml_tree* t, const Side& sa, const Side& sb, int e, const int32_t* dw, double len, bool fused,
                           bool keep_table, bool site_lnl, int want, bool device_nr, const NewviewOp* nv = nullptr);
// which consumers form a folded cherry themselves (the others get it stored by materialize_cherry)
constexpr bool kBranchTakesCherry = true;   // branch kernel, far end
constexpr bool kFusedTakesCherry = true;    // fused kernel: children of the update and far end

double branch_launch(pml_tree* t, int e, const int32_t* dw, double len, bool keep_table, bool site_lnl, int want, bool device_nr) {
    if (keep_table && !ensure_sumtable(t->aln)) return 0.0;
    pml_ctx* c = t->aln->ctx;
    const Topology& T = t->topo;
    int x = T.ea[e], y = T.eb[e];
    if (T.is_tip(y)) std::swap(x, y);
    std::vector<ViewOp> ops;
    t->views.plan(T, x, y, ops);
    t->views.plan(T, y, x, ops);
    // The last CLV update of the visit produces one end of this very branch: update and pass go out as ONE launch
    // (fused_mma.cu) unless the product table is wanted or both children of that node are tips.
    const bool fuse = c->fuse && !keep_table && !ops.empty() &&
                      !(T.is_tip(ops.back().child[0]) && T.is_tip(ops.back().child[1]));
    if (!fuse) {
        if (!run_ops(t, ops) || !sync_lengths(t)) return 0.0;
        // the pass wants an inner node at one end (b); the other end (a) may be an inner node, a tip or a folded cherry
        Side sa = t->side_toward(x, y), sb = t->side_toward(y, x);
        if (side_kind(sb) != kSideInner && side_kind(sa) == kSideInner) {
            std::swap(sa, sb);
            std::swap(x, y);
        }
        if (side_kind(sb) != kSideInner) {  // neither end is stored: a cherry against a tip or a cherry (3 / 4 taxa)
            if (side_kind(sb) != kSideCherry) {
                std::swap(sa, sb);
                std::swap(x, y);
            }
            sb = materialize_cherry(t, y, x);
        }
        if (side_kind(sa) == kSideCherry && !kBranchTakesCherry) sa = materialize_cherry(t, x, y);
        return branch_launch_sides(t, sa, sb, e, dw, len, false, keep_table, site_lnl, want, device_nr);
    }
    const ViewOp last = ops.back();
    ops.pop_back();
    if (!run_ops(t, ops) || !sync_lengths(t)) return 0.0;
    const int far = last.node == x ? y : x;
    Side sfar = t->side_toward(far, last.node);
    if (side_kind(sfar) == kSideCherry && !kFusedTakesCherry) sfar = materialize_cherry(t, far, last.node);
    NewviewOp nv = make_newview_op(t, last);
    if (!kFusedTakesCherry) {
        if (side_kind(nv.left) == kSideCherry) nv.left = materialize_cherry(t, last.child[0], last.node);
        if (side_kind(nv.right) == kSideCherry) nv.right = materialize_cherry(t, last.child[1], last.node);
    }
    const int ntip = (nv.left.clv == nullptr) + (nv.right.clv == nullptr);
    {
        static const char* only = getenv("PEPRML_TRACE_KIND");  // profiling aid: trace fused kernels of one kind only
        const int kind = 13 + 3 * (newview_kind(nv.left, nv.right) - 1) + far_index(sfar);
        if (c->trace_fused && (only ? atoi(only) == kind : kind == 16)) nv.trace = c->d_trace_buf;  // default: inner-inner update, inner far end
    }
    t->site_updates[2 - ntip] += t->aln->nloc;
    Side sx{};
    sx.clv = nv.out;
    sx.scale = nv.out_scale;
    return branch_launch_sides(t, sfar, sx, e, dw, len, true, false, site_lnl, want, device_nr, &nv);
}

// the pass itself between two given sides (sb inner; sa inner or tip); e names the branch for device NR and book-keeping;
// fused: sb is what the CLV update *nv is about to produce, both go out as one launch
double branch_launch_sides(pml_tree* t, const Side& sa, const Side& sb, int e, const int32_t* dw, double len, bool fused,
                           bool keep_table, bool site_lnl, int want, bool device_nr, const NewviewOp* nv) {
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    BranchArgs args{};
    args.a = sa;
    args.b = sb;
    args.dm = a->d_model;
    args.weights = dw;
    args.t = len;
    args.t_ptr = device_nr ? t->d_len + e : nullptr;
    args.site_lnl = site_lnl ? a->d_site_lnl : nullptr;
    args.want_lnl = want & 1;
    args.want_derivs = (want >> 1) & 1;
    args.sumtable = keep_table ? a->d_sumtable : nullptr;
    args.sum_scale = keep_table ? a->d_sumscale : nullptr;
    args.partials = a->d_partials;
    args.ticket = a->d_ticket;
    args.result = a->d_result;
    args.trace = c->trace_branch_tip == (args.a.clv == nullptr) ? c->d_trace_branch : nullptr;
    Publish pub{};
    pub.seq = (c->sequence += 1.0);
    pub.slot = c->slot_dev(pub.seq);
    if (device_nr) {
        pub.len = t->d_len + e;
        pub.poison = c->d_poison;
    }
    const bool in_kernel = c->nranks == 1 || c->peer_ok;
    if (in_kernel) args.pub = pub;
    if (c->peer_ok) args.peer = PeerReduce{c->d_mail_ptrs, c->rank, c->nranks, c->d_peer_lost, c->peer_timeout_ns};
    // kinds: see launch_newview
    const int fused_kind = fused ? 13 + 3 * (newview_kind(nv->left, nv->right) - 1) + far_index(args.a) : 0;
    const int tk = c->tick(fused ? fused_kind : (site_lnl ? 9 : 6) + far_index(args.a), a->nloc);
    if (fused) {
        NewviewOp timed_nv = *nv;
        timed_nv.timeline = c->timeline_slot(fused_kind);
        launch_fused(timed_nv, args, a->npad, c->sms, c->stream);
    } else launch_branch_mma(args, a->npad, c->sms, c->stream);
    c->tock(tk);
    t->launches += 1;
    t->prepared_branch = keep_table ? e : -1;
    if (keep_table) a->sumtable_owner = t;
    if (!c->cuda(cudaGetLastError(), "branch kernel")) return 0.0;
    if (!in_kernel) {
        if (!c->allreduce(a->d_result, 3)) return 0.0;
        launch_publish(a->d_result, pub, c->stream);
        t->launches += 1;
    }
    return pub.seq;
}

bool branch_pass(pml_tree* t, int e, const int32_t* dw, double len, bool keep_table, bool site_lnl, double out[3], int want = kWantAll) {
    pml_ctx* c = t->aln->ctx;
    const double seq = branch_launch(t, e, dw, len, keep_table, site_lnl, want, false);
    // everything queued before the pass has completed once its result is visible, so the staging area is free again
    double r[5];
    if (seq == 0.0 || !c->wait_slot(seq, r)) return false;
    c->stage_used = 0;
    out[0] = r[0];
    out[1] = r[1];
    out[2] = r[2];
    return true;
}

int evaluate_branch(pml_tree* t, int e, const int32_t* weights, double* lnl) {
    const int32_t* dw = device_weights(t->aln, weights);
    if (!dw) return PML_ENODEVICE;
    double r[3];
    if (!branch_pass(t, e, dw, t->topo.len[e], false, true, r, kWantLnl)) return t->aln->ctx->fail_code();
    *lnl = r[0];
    return PML_OK;
}

// lnL, dlnL/dt, d2lnL/dt2 of the branch whose product table is resident, at length len
bool core_at(pml_tree* t, const int32_t* dw, double len, double out[3]) {
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    const int tk = c->tick(12, a->nloc);
    launch_core(a->d_model, a->d_sumtable, a->d_sumscale, dw, a->npad, len, a->d_partials, a->d_result, c->stream);
    c->tock(tk);
    t->launches += 2;
    if (!c->cuda(cudaGetLastError(), "core kernel")) return false;
    if (!c->allreduce(a->d_result, 3)) return false;
    return fetch_result(c, a->d_result, 3, out);
}

void set_branch(pml_tree* t, int e, double len) {
    if (t->topo.len[e] == len) return;
    t->topo.len[e] = len;
    t->views.branch_changed(t->topo, e);
    t->prepared_branch = -1;
}

// guarded Newton-Raphson on z = exp(-t) in log z, constants and control flow of raxmlHPC topLevelMakenewz:
// bad curvature -> z = 0.37 z + 0.63; step z *= exp(-d1/d2) when that exponent is < 100; cap z <= 0.25 zprev + 0.75
bool newton_branch(pml_tree* t, int e, const int32_t* dw, int maxiter, double& z_out) {
    // a single NR step needs the CLVs once and no product table; longer iterations keep the table and re-use it
    const bool keep = maxiter > 1;
    bool first = true;
    double z = std::exp(-t->topo.len[e]);
    z = std::min(std::max(z, kZmin), kZmax);
    double zprev = z, zstep = 0.0;
    bool curvature_ok = true, done = false;
    while (!done) {
        if (curvature_ok) {
            curvature_ok = false;
            zprev = z;
            zstep = (1.0 - kZmax) * z + kZmin;
        }
        z = std::min(std::max(z, kZmin), kZmax);
        double r[3];
        if (first || !keep) {
            if (!branch_pass(t, e, dw, -std::log(z), keep, false, r, kWantDerivs)) return false;
        } else if (!core_at(t, dw, -std::log(z), r)) return false;
        first = false;
        const double d1 = -r[1], d2 = r[2];  // derivatives in lz = log z = -t
        if (d2 >= 0.0 && z < kZmax) {
            zprev = z = 0.37 * z + 0.63;
            ++t->nr_retries;
        } else curvature_ok = true;
        if (curvature_ok) {
            if (d2 < 0.0) {
                const double step = -d1 / d2;
                if (step < 100.0) {
                    z *= std::exp(step);
                    z = std::max(z, kZmin);
                    z = std::min(z, 0.25 * zprev + 0.75);
                } else {
                    z = 0.25 * zprev + 0.75;
                }
            }
            z = std::min(z, kZmax);
            --maxiter;
            done = !(maxiter > 0 && std::fabs(z - zprev) > zstep);
        }
    }
    z_out = z;
    return true;
}

// One sweep: depth-first over all branches starting at taxon 0's branch, one guarded NR step each (raxmlHPC
// smoothTree/update).  The step itself runs on the device (branch kernel tail / k_publish) and the host stays one branch
// ahead: while the GPU works on branch i the host has already queued branch i+1 and only then looks at the published
// outcome of branch i.  The speculation is that the step did not end in the bad-curvature retry (it never did on the bench
// workload); if it did, the device has refused the step queued behind it (poison flag), the host repeats the pass until
// the curvature is fine and queues the next branch again.
bool smooth_sweep(pml_tree* t, const int32_t* dw, bool& smoothed) {
    smoothed = true;
    pml_ctx* c = t->aln->ctx;
    Topology& T = t->topo;
    if (c->host_nr) {
        // reference mode (PEPRML_HOST_NR=1): the same sweep with the NR step on the host and a wait per branch; the tests
        // hold the pipelined path to it
        std::vector<std::pair<int, int>> stack;  // (branch, far node)
        stack.push_back({T.edge[0][0], T.nbr[0][0]});
        while (!stack.empty()) {
            const auto [e, far] = stack.back();
            stack.pop_back();
            const double z0 = std::min(std::max(std::exp(-T.len[e]), kZmin), kZmax);
            double z;
            if (!newton_branch(t, e, dw, 1, z)) return false;
            if (std::fabs(z - z0) > 1.0e-5) smoothed = false;
            set_branch(t, e, -std::log(z));
            t->views.branch_changed(T, e);
            t->last_swept = e;
            if (!T.is_tip(far)) {
                const int near = T.ea[e] == far ? T.eb[e] : T.ea[e];
                for (int s = 2; s >= 0; --s)
                    if (T.nbr[far][s] != near) stack.push_back({T.edge[far][s], T.nbr[far][s]});
            }
        }
        return true;
    }
    std::vector<int> order;
    {
        std::vector<std::pair<int, int>> stack;  // (branch, far node)
        stack.push_back({T.edge[0][0], T.nbr[0][0]});
        while (!stack.empty()) {
            const auto [e, far] = stack.back();
            stack.pop_back();
            order.push_back(e);
            if (!T.is_tip(far)) {
                const int near = T.ea[e] == far ? T.eb[e] : T.ea[e];
                for (int s = 2; s >= 0; --s)
                    if (T.nbr[far][s] != near) stack.push_back({T.edge[far][s], T.nbr[far][s]});
            }
        }
    }
    t->last_swept = order.empty() ? -1 : order.back();
    // z0: the branch's z before its FIRST pass of this visit -- raxmlHPC update() compares the final z with that one, also
    // when bad-curvature retries moved the starting point in between
    struct Pending { int e; double seq; double z0; };
    // adopts the outcome of a queued step: host mirror of the length, convergence flag; false = error
    auto adopt = [&](const Pending& p, double r[5]) {
        if (!c->wait_slot(p.seq, r)) return false;
        const int status = (int)r[4];
        if (status == kNrSkipped) return true;
        T.len[p.e] = t->len_dev[p.e] = r[3];
        if (status == kNrDone && std::fabs(std::exp(-r[3]) - p.z0) > 1.0e-5) smoothed = false;
        return true;
    };
    auto queue_step = [&](int e, double z0 = -1.0) {
        if (z0 < 0.0) z0 = std::min(std::max(std::exp(-T.len[e]), kZmin), kZmax);
        Pending p{e, branch_launch(t, e, dw, 0.0, false, false, kWantDerivs, true), z0};
        t->views.branch_changed(T, e);  // the length is about to change
        t->prepared_branch = -1;
        return p;
    };
    size_t i = 0;
    Pending prev{-1, 0.0, 0.0};
    while (i < order.size() || prev.e >= 0) {
        Pending cur{-1, 0.0, 0.0};
        if (i < order.size()) {
            cur = queue_step(order[i]);
            if (cur.seq == 0.0) return false;
        }
        if (prev.e >= 0) {
            double r[5];
            if (!adopt(prev, r)) return false;
            if ((int)r[4] == kNrRetry) {
                ++t->nr_retries;
                // the step queued behind (cur) was refused on the device: wait for it, lift the flag, redo prev until it is done
                if (cur.e >= 0 && !c->wait_slot(cur.seq, r)) return false;
                if (!c->cuda(cudaMemsetAsync(c->d_poison, 0, sizeof(int), c->stream), "flag clear")) return false;
                for (int guard = 0; guard < 64; ++guard) {
                    const Pending again = queue_step(prev.e, prev.z0);
                    if (again.seq == 0.0 || !adopt(again, r)) return false;
                    if ((int)r[4] != kNrRetry) break;
                    ++t->nr_retries;
                    if (!c->cuda(cudaMemsetAsync(c->d_poison, 0, sizeof(int), c->stream), "flag clear")) return false;
                }
                prev = Pending{-1, 0.0, 0.0};
                continue;  // cur is queued again on the next trip (i was not advanced)
            }
        }
        prev = cur;
        if (cur.e >= 0) ++i;
    }
    return true;
}

// raxmlHPC treeEvaluate: smoothing sweeps until no branch moves, then the tree's lnL.  The lnL is the same at every branch;
// where_last takes it at the branch the last sweep ended on -- both views are in place there, so it costs one branch pass
// instead of re-orienting the whole tree towards taxon 0 (worth it when the caller is about to change alpha, which
// invalidates every view anyway)
bool tree_evaluate(pml_tree* t, const int32_t* weights, const int32_t* dw, double factor, double* lnl, bool where_last = false) {
    int sweeps = (int)(32 * factor);
    while (--sweeps >= 0) {
        bool smoothed;
        if (!smooth_sweep(t, dw, smoothed)) return false;
        if (smoothed) break;
    }
    const int e = where_last && t->last_swept >= 0 ? t->last_swept : t->topo.edge[0][0];
    return evaluate_branch(t, e, weights, lnl) == PML_OK;
}

bool set_alpha(pml_tree* t, double alpha) {
    pml_aln* a = t->aln;
    a->alpha = alpha;
    gamma_mean_rates(alpha, kCats, a->rates);
    if (!upload_model(a)) return false;
    adopt_model(t);
    return true;
}

// Brent's minimiser on log(alpha) for -lnL with every other parameter fixed; each trial is a full traversal + evaluate
bool optimise_alpha(pml_tree* t, const int32_t* weights, double tol, double* best_lnl) {
    pml_aln* a = t->aln;
    const int root_branch = t->topo.edge[0][0];
    bool ok = true;
    static const bool debug = getenv("PEPRML_DEBUG_ALPHA") != nullptr;
    int nevals = 0;
    auto f = [&](double la) {
        double l = 0.0;
        if (!set_alpha(t, std::exp(la)) || evaluate_branch(t, root_branch, weights, &l) != PML_OK) ok = false;
        ++nevals;
        if (debug) fprintf(stderr, "  alpha trial %d: log alpha %.6f lnL %.4f\n", nevals, la, l);
        return -l;
    };
    const double lmin = std::log(kAlphaMin), lmax = std::log(kAlphaMax), gold = 1.6180339887498949, cgold = 0.3819660112501051;
    auto clamp = [&](double v) { return std::min(std::max(v, lmin), lmax); };
    // the caller has just evaluated the tree at the current alpha (*best_lnl): no trial is spent on it.  The first probe sits
    // as far away as alpha moved in the previous call (three times that, between 4 tol and 0.1): late modOpt rounds, where
    // alpha moves in the third decimal, bracket the optimum tightly at once
    const double x0 = std::log(a->alpha);
    double xa = x0, xb = clamp(xa + a->alpha_step), fa = -*best_lnl, fb = f(xb);
    if (xb == xa) xb = clamp(xa - a->alpha_step), fb = f(xb);
    if (fb > fa) {
        std::swap(xa, xb);
        std::swap(fa, fb);
    }
    double xc = clamp(xb + gold * (xb - xa)), fc = f(xc);
    for (int guard = 0; ok && fb > fc && guard < 64; ++guard) {
        xa = xb;
        fa = fb;
        xb = xc;
        fb = fc;
        xc = clamp(xb + gold * (xb - xa));
        if (xc == xb) break;
        fc = f(xc);
    }
    // Brent's memory starts from the bracket itself (best point, then the better and the worse end), so the very first
    // step is already a parabola through three evaluated points instead of a golden-section probe
    double lo = std::min(xa, xc), hi = std::max(xa, xc), x = xb, fx = fb;
    double w = fa <= fc ? xa : xc, fw = std::min(fa, fc), v = fa <= fc ? xc : xa, fv = std::max(fa, fc);
    double d = 0.5 * (hi - lo), e = hi - lo;
    for (int it = 0; ok && it < 100; ++it) {
        // absolute tolerance on log(alpha): a relative one degenerates when alpha is close to 1 (log alpha close to 0)
        const double mid = 0.5 * (lo + hi), tol1 = tol, tol2 = 2.0 * tol1;
        if (std::fabs(x - mid) <= tol2 - 0.5 * (hi - lo)) break;
        bool golden = true;
        if (std::fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), p = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = std::fabs(q);
            const double eprev = e;
            e = d;
            if (!(std::fabs(p) >= std::fabs(0.5 * q * eprev) || p <= q * (lo - x) || p >= q * (hi - x))) {
                d = p / q;
                const double u = x + d;
                if (u - lo < tol2 || hi - u < tol2) d = mid >= x ? tol1 : -tol1;
                golden = false;
            }
        }
        if (golden) {
            e = x >= mid ? lo - x : hi - x;
            d = cgold * e;
        }
        const double u = std::fabs(d) >= tol1 ? x + d : x + (d >= 0 ? tol1 : -tol1), fu = f(u);
        if (fu <= fx) {
            if (u >= x) lo = x; else hi = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) lo = u; else hi = u;
            if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
    }
    if (!ok) return false;
    if (!set_alpha(t, std::exp(x))) return false;
    a->alpha_step = std::min(0.1, std::max(4.0 * tol, 3.0 * std::fabs(x - x0)));
    *best_lnl = -fx;
    return true;
}

// lnL of the tree at branch e without touching parameters (weights already on the device)
bool lnl_at(pml_tree* t, int e, const int32_t* dw, double& lnl) {
    double r[3];
    if (!branch_pass(t, e, dw, t->topo.len[e], false, false, r, kWantLnl)) return false;
    lnl = r[0];
    return true;
}

// a few guarded NR steps on one branch
bool polish_branch(pml_tree* t, int e, const int32_t* dw, int iters) {
    double z;
    if (!newton_branch(t, e, dw, iters, z)) return false;
    set_branch(t, e, -std::log(z));
    return true;
}

// Lazy scores of a list of regraft targets for the pruned (p, s) (raxmlHPC rearrangeBIG / testInsertBIG without NR).
// The subtree is pruned ONCE; the remaining tree keeps its topology and its valid views while the regraft points are
// visited.  A candidate is a *virtual* insertion: orient the target branch (a, b) on the pruned tree (moving to a
// neighbouring target costs about one CLV update, as in a smoothing sweep), one CLV update for p from the two directed views
// of that branch with half its length on either side, and one branch pass between p and the subtree's root at the subtree
// branch's own length.  Nothing waits for a candidate's lnL: outcomes are collected from the result ring a few candidates
// later.  out[i] = lnL (or -inf when the move is illegal).
bool score_candidates(pml_tree* t, int p, int s, const std::vector<int>& targets, const int32_t* dw, std::vector<double>& out) {
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    Topology& T = t->topo;
    out.assign(targets.size(), -std::numeric_limits<double>::infinity());
    if (targets.empty() || T.is_tip(p)) return true;
    // the subtree's own view towards p is taken on the intact tree
    {
        std::vector<ViewOp> ops;
        t->views.plan(T, s, p, ops);
        if (!run_ops(t, ops)) return false;
    }
    SprMove mv;
    if (!spr_prune(T, t->views, p, s, mv)) return true;
    t->prepared_branch = -1;
    const double len_s = T.len[mv.e_s];
    Side side_p{};
    side_p.clv = t->clv(p);
    side_p.scale = t->scale(p);
    Side side_s = t->side_toward(s, p);  // the subtree's own view: an inner node, a tip or a folded cherry
    if (side_kind(side_s) == kSideCherry && !(kBranchTakesCherry && kFusedTakesCherry)) side_s = materialize_cherry(t, s, p);
    constexpr size_t kLag = pml_ctx::kRing - 2;
    std::vector<std::pair<double, size_t>> queued;  // (sequence number, candidate index)
    size_t head = 0;
    bool ok = true;
    auto collect = [&](size_t upto) {
        for (; head < upto; ++head) {
            double r[5];
            if (!c->wait_slot(queued[head].first, r)) return false;
            out[queued[head].second] = r[0];
        }
        return true;
    };
    for (size_t i = 0; ok && i < targets.size(); ++i) {
        const int e = targets[i];
        if (e == mv.e_q || e == mv.e_r || e == mv.e_s) continue;  // would put p back where it was
        int x, y;
        if (!orient_branch(t, e, x, y)) {  // both directed views of the target branch, on the pruned tree
            ok = false;
            break;
        }
        NewviewOp nv{};
        nv.left = t->side_toward(x, y);
        nv.right = t->side_toward(y, x);
        nv.len_left = nv.len_right = t->d_len + e;
        nv.len_scale = 0.5;
        nv.dm = a->d_model;
        nv.out = t->clv(p);
        nv.out_scale = t->scale(p);
        const bool both_tips = side_kind(nv.left) == kSideTip && side_kind(nv.right) == kSideTip;
        double seq;
        if (c->fuse && !both_tips) {  // the insertion update and the pass over the subtree's branch as one launch
            if (!kFusedTakesCherry) {
                if (side_kind(nv.left) == kSideCherry) nv.left = materialize_cherry(t, x, y);
                if (side_kind(nv.right) == kSideCherry) nv.right = materialize_cherry(t, y, x);
            }
            const int ntip = (nv.left.clv == nullptr) + (nv.right.clv == nullptr);
            t->site_updates[2 - ntip] += a->nloc;
            seq = branch_launch_sides(t, side_s, side_p, mv.e_s, dw, len_s, true, false, false, kWantLnl, false, &nv);
        } else {
            launch_newview(t, nv);
            seq = branch_launch_sides(t, side_s, side_p, mv.e_s, dw, len_s, false, false, false, kWantLnl, false);
        }
        if (seq == 0.0) {
            ok = false;
            break;
        }
        queued.push_back({seq, i});
        if (queued.size() - head > kLag && !collect(queued.size() - kLag)) ok = false;
    }
    ok = collect(queued.size()) && ok;
    spr_unprune(T, t->views, mv);
    t->prepared_branch = -1;
    return ok;
}

struct SearchStats {
    int64_t candidates = 0;
    int accepted = 0;
};

// one pass over every (node, subtree) pair: lazy scores for all targets, thorough check of the best one
bool spr_round(pml_tree* t, const int32_t* dw, int radius, double& best, SearchStats& st) {
    Topology& T = t->topo;
    const int root_branch = T.edge[0][0];
    for (int p = T.ntax; p < T.nnodes(); ++p) {
        for (int k = 0; k < 3; ++k) {
            const int s = T.nbr[p][k];
            std::vector<int> targets = spr_targets(T, p, s, radius);
            if (!t->aln->constraints.empty()) {  // topological constraints: only moves that keep every split (FastTree -constraints)
                std::vector<int> kept;
                for (int target : targets) {
                    Topology trial = T;
                    ViewState scratch;
                    scratch.reset(trial);
                    SprMove mv;
                    if (spr_apply(trial, scratch, p, s, target, mv) && satisfies(trial, t->aln->constraints)) kept.push_back(target);
                }
                targets.swap(kept);
            }
            if (targets.empty()) continue;
            int best_target = -1;
            double best_lazy = -1e300;
            std::vector<double> lazy;
            if (!score_candidates(t, p, s, targets, dw, lazy)) return false;
            for (size_t i = 0; i < targets.size(); ++i) {
                if (!std::isfinite(lazy[i])) continue;
                ++st.candidates;
                if (lazy[i] > best_lazy) {
                    best_lazy = lazy[i];
                    best_target = targets[i];
                }
            }
            // the lazy score leaves three branches unoptimised, so a candidate slightly below the current tree may still win
            if (best_target < 0 || best_lazy < best - 2.0) continue;
            SprMove mv;
            if (!spr_apply(T, t->views, p, s, best_target, mv)) continue;
            t->prepared_branch = -1;
            for (int pass = 0; pass < 2; ++pass)
                for (int e : {mv.e_t, mv.e_r, mv.e_s, mv.e_q})
                    if (!polish_branch(t, e, dw, 3)) return false;
            double cand;
            if (!lnl_at(t, root_branch, dw, cand)) return false;
            if (cand > best + 1e-3) {
                best = cand;
                ++st.accepted;
                break;  // p's surroundings changed: go on with the next node
            }
            spr_undo(T, t->views, mv);
            t->prepared_branch = -1;
        }
    }
    return true;
}

int fail(pml_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return code;
}

// Opens every rank's mailbox in every other rank (CUDA IPC over NVLink).  The 64-byte handles are gathered with the NCCL
// communicator that already exists (sum of zero-padded int32 arrays = gather).  Any failure leaves peer_ok = false and the
// engine on the NCCL path; set PEPRML_NO_PEER=1 to force that path.
void setup_peer_mail(pml_ctx* c) {
    if (c->nranks < 2 || c->nranks > kMaxPeers || getenv("PEPRML_NO_PEER")) return;
    const size_t bytes = sizeof(double) * kPeerRing * c->nranks * 6;
    constexpr int kWords = (int)(sizeof(cudaIpcMemHandle_t) / sizeof(int32_t));
    int32_t* d_gather = nullptr;
    std::vector<int32_t> h((size_t)kWords * c->nranks + c->nranks, 0);
    cudaIpcMemHandle_t mine;
    bool ok = cudaMalloc(&c->d_peer_lost, sizeof(int)) == cudaSuccess && cudaMemset(c->d_peer_lost, 0, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&c->d_mail, bytes) == cudaSuccess && cudaMemset(c->d_mail, 0, bytes) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine, c->d_mail) == cudaSuccess && cudaMalloc(&d_gather, sizeof(int32_t) * h.size()) == cudaSuccess;
    // every rank must take part in the collectives below, whatever happened locally: a flag per rank travels with the handles
    if (ok) {
        std::memcpy(h.data() + (size_t)kWords * c->rank, &mine, sizeof mine);
        h[(size_t)kWords * c->nranks + c->rank] = 1;
    }
    if (!d_gather && cudaMalloc(&d_gather, sizeof(int32_t) * h.size()) != cudaSuccess) {
        cudaGetLastError();
        return;  // cannot even take part; the other ranks will time out in NCCL -- as they would on any allocation failure
    }
    cudaMemcpy(d_gather, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice);
    const bool gathered = g_nccl.AllReduce(d_gather, d_gather, h.size(), ncclInt32, ncclSum, c->comm, c->stream) == ncclSuccess &&
                          cudaStreamSynchronize(c->stream) == cudaSuccess &&
                          cudaMemcpy(h.data(), d_gather, sizeof(int32_t) * h.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    int good = 0;
    for (int r = 0; r < c->nranks; ++r) good += h[(size_t)kWords * c->nranks + r];
    std::vector<double*> ptrs(c->nranks, nullptr);
    bool opened = gathered && good == c->nranks;
    for (int r = 0; opened && r < c->nranks; ++r) {
        if (r == c->rank) {
            ptrs[r] = c->d_mail;
            continue;
        }
        cudaIpcMemHandle_t hr;
        std::memcpy(&hr, h.data() + (size_t)kWords * r, sizeof hr);
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            opened = false;
            break;
        }
        c->peer_opened.push_back(p);
        ptrs[r] = (double*)p;
    }
    if (opened)
        opened = cudaMalloc(&c->d_mail_ptrs, sizeof(double*) * c->nranks) == cudaSuccess &&
                 cudaMemcpy(c->d_mail_ptrs, ptrs.data(), sizeof(double*) * c->nranks, cudaMemcpyHostToDevice) == cudaSuccess;
    // second round: the in-kernel path is used only if EVERY rank opened every mailbox (it doubles as the barrier that keeps
    // a fast rank from writing into a mailbox that is still being cleared)
    int32_t flag = opened ? 1 : 0;
    cudaMemcpy(d_gather, &flag, sizeof flag, cudaMemcpyHostToDevice);
    if (g_nccl.AllReduce(d_gather, d_gather, 1, ncclInt32, ncclSum, c->comm, c->stream) == ncclSuccess &&
        cudaStreamSynchronize(c->stream) == cudaSuccess && cudaMemcpy(&flag, d_gather, sizeof flag, cudaMemcpyDeviceToHost) == cudaSuccess)
        c->peer_ok = flag == c->nranks;
    cudaFree(d_gather);
    cudaGetLastError();
}

}  // namespace

// =============================================================================================== C ABI ======
#pragma GCC visibility push(default)
extern "C" {

const char* pml_version(void) { return "peprml-b200 0.1 (WAG+G4, FP64, sm_100a)"; }

int pml_comm_unique_id(unsigned char id[PML_UNIQUE_ID_BYTES]) {
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, PML_ECOMM, err);
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != ncclSuccess) return fail(nullptr, PML_ECOMM, "ncclGetUniqueId failed");
    static_assert(sizeof(u) == PML_UNIQUE_ID_BYTES, "unique id size");
    std::memcpy(id, &u, sizeof u);
    return PML_OK;
}

static int create_base(int gpu_id, int rank, int nranks, std::unique_ptr<pml_ctx>& c) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PML_ENODEVICE,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this engine has no CPU path");
    if (gpu_id < 0 || gpu_id >= ndev) return fail(nullptr, PML_EINVAL, "gpu_id out of range");
    c = std::make_unique<pml_ctx>();
    c->device = gpu_id;
    c->host_nr = getenv("PEPRML_HOST_NR") != nullptr;
    c->fuse = getenv("PEPRML_NO_FUSE") == nullptr;
    c->fold = getenv("PEPRML_NO_FOLD") == nullptr;
    c->rank = rank;
    c->nranks = nranks;
    if (const char* ms = getenv("PEPRML_PEER_TIMEOUT_MS")) c->peer_timeout_ns = (unsigned long long)std::max(1.0, atof(ms)) * 1000000ull;
    if (!c->bind() || !c->cuda(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream create"))
        return fail(nullptr, PML_ENODEVICE, c->err);
    cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, gpu_id);
    configure_mma_kernels();
    configure_branch_kernels();
    configure_fused_kernels();
    c->stage_cap = 1 << 20;
    if (!c->cuda(cudaMallocHost(&c->h_stage, c->stage_cap), "pinned alloc") ||
        !c->cuda(cudaMallocHost(&c->h_result, 4096), "pinned alloc") ||
        !c->cuda(cudaHostAlloc((void**)&c->h_mapped, sizeof(double) * pml_ctx::kRing * kSlotDoubles, cudaHostAllocMapped), "mapped alloc") ||
        !c->cuda(cudaHostGetDevicePointer((void**)&c->d_mapped, (void*)c->h_mapped, 0), "mapped pointer") ||
        !c->cuda(cudaMalloc(&c->d_poison, sizeof(int)), "flag alloc") || !c->cuda(cudaMemset(c->d_poison, 0, sizeof(int)), "flag init"))
        return fail(nullptr, PML_ENOMEM, c->err);
    for (int i = 0; i < pml_ctx::kRing * kSlotDoubles; ++i) c->h_mapped[i] = 0.0;
    return PML_OK;
}

int pml_ctx_create(int gpu_id, int rank, int nranks, const unsigned char* unique_id, pml_ctx** out) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return fail(nullptr, PML_EINVAL, "bad rank/nranks");
    *out = nullptr;
    std::unique_ptr<pml_ctx> c;
    const int rc = create_base(gpu_id, rank, nranks, c);
    if (rc != PML_OK) return rc;
    if (nranks > 1) {
        if (!unique_id) return fail(nullptr, PML_EINVAL, "unique_id required when nranks > 1");
        std::string err;
        if (!g_nccl.load(err)) return fail(nullptr, PML_ECOMM, err);
        ncclUniqueId u;
        std::memcpy(&u, unique_id, sizeof u);
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
        if (r != ncclSuccess) return fail(nullptr, PML_ECOMM, "ncclCommInitRank failed");
        setup_peer_mail(c.get());
    }
    *out = c.release();
    return PML_OK;
}

// All ranks of a group in THIS process (what `-T n` is for raxmlHPC-PTHREADS inside one JVM, RAxMLRunner.java:130-132):
// rank i runs on gpu_ids[i].  The mailboxes of the in-kernel reduction are plain device allocations reached through
// cudaDeviceEnablePeerAccess -- no CUDA IPC, which cannot map a handle in the process that exported it.  Every context must
// then be driven by its own host thread, all threads making the same calls (a branch pass waits for its peers' sums).
int pml_group_create(const int* gpu_ids, int ngpu, pml_ctx** out) {
    if (!gpu_ids || !out || ngpu < 1 || ngpu > kMaxPeers) return fail(nullptr, PML_EINVAL, "pml_group_create: bad arguments");
    for (int r = 0; r < ngpu; ++r) out[r] = nullptr;
    for (int r = 0; r < ngpu; ++r)
        for (int q = 0; q < r; ++q)
            if (gpu_ids[r] == gpu_ids[q]) return fail(nullptr, PML_EINVAL, "pml_group_create: a GPU may carry one rank only");
    std::vector<std::unique_ptr<pml_ctx>> cs(ngpu);
    auto destroy_all = [&]() {
        for (auto& c : cs)
            if (c) pml_ctx_destroy(c.release());
    };
    for (int r = 0; r < ngpu; ++r) {
        const int rc = create_base(gpu_ids[r], r, ngpu, cs[r]);
        if (rc != PML_OK) {
            cs[r].reset();
            destroy_all();
            return rc;
        }
    }
    if (ngpu > 1) {
        std::string err;
        if (!g_nccl.load(err) || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
            destroy_all();
            return fail(nullptr, PML_ECOMM, err.empty() ? "libnccl.so.2 lacks ncclGroupStart/End" : err);
        }
        ncclUniqueId u;
        bool ok = g_nccl.GetUniqueId(&u) == ncclSuccess && g_nccl.GroupStart() == ncclSuccess;
        for (int r = 0; ok && r < ngpu; ++r) ok = cudaSetDevice(gpu_ids[r]) == cudaSuccess && g_nccl.CommInitRank(&cs[r]->comm, ngpu, u, r) == ncclSuccess;
        ok = g_nccl.GroupEnd() == ncclSuccess && ok;
        if (!ok) {
            destroy_all();
            return fail(nullptr, PML_ECOMM, "pml_group_create: NCCL communicator setup failed");
        }
        auto g = std::make_shared<PmlGroup>();
        g->n = ngpu;
        g->devices.assign(gpu_ids, gpu_ids + ngpu);
        g->mail.assign(ngpu, nullptr);
        for (auto& c : cs) c->group = g;
        bool peers = getenv("PEPRML_NO_PEER") == nullptr;
        for (int r = 0; peers && r < ngpu; ++r)
            for (int q = 0; peers && q < ngpu; ++q) {
                int can = 0;
                if (q != r && (cudaDeviceCanAccessPeer(&can, gpu_ids[r], gpu_ids[q]) != cudaSuccess || !can)) peers = false;
            }
        const size_t bytes = sizeof(double) * kPeerRing * ngpu * 6;
        for (int r = 0; peers && r < ngpu; ++r) {
            cudaSetDevice(gpu_ids[r]);
            for (int q = 0; peers && q < ngpu; ++q) {
                if (q == r) continue;
                const cudaError_t pe = cudaDeviceEnablePeerAccess(gpu_ids[q], 0);
                if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (pe != cudaSuccess) peers = false;
            }
            peers = peers && cudaMalloc(&g->mail[r], bytes) == cudaSuccess && cudaMemset(g->mail[r], 0, bytes) == cudaSuccess &&
                    cudaMalloc(&cs[r]->d_peer_lost, sizeof(int)) == cudaSuccess && cudaMemset(cs[r]->d_peer_lost, 0, sizeof(int)) == cudaSuccess;
        }
        for (int r = 0; peers && r < ngpu; ++r) {
            cudaSetDevice(gpu_ids[r]);
            peers = cudaMalloc(&cs[r]->d_mail_ptrs, sizeof(double*) * ngpu) == cudaSuccess &&
                    cudaMemcpy(cs[r]->d_mail_ptrs, g->mail.data(), sizeof(double*) * ngpu, cudaMemcpyHostToDevice) == cudaSuccess &&
                    cudaDeviceSynchronize() == cudaSuccess;
        }
        cudaGetLastError();
        for (auto& c : cs) c->peer_ok = peers;
    }
    for (int r = 0; r < ngpu; ++r) out[r] = cs[r].release();
    return PML_OK;
}

void pml_ctx_destroy(pml_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    for (auto& b : c->cached_blocks) cudaFree(b.second);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->h_mapped) cudaFreeHost((void*)c->h_mapped);
    if (c->d_poison) cudaFree(c->d_poison);
    for (void* p : c->peer_opened) cudaIpcCloseMemHandle(p);
    if (c->d_mail_ptrs) cudaFree(c->d_mail_ptrs);
    if (c->d_mail) cudaFree(c->d_mail);
    if (c->d_peer_lost) cudaFree(c->d_peer_lost);
    if (c->d_trace_buf) cudaFree(c->d_trace_buf);
    if (c->d_timeline) cudaFree(c->d_timeline);
    for (auto& t : c->timed) { cudaEventDestroy(t.t0); cudaEventDestroy(t.t1); }
    for (auto e : c->spare_events) cudaEventDestroy(e);
    delete c;
}

const char* pml_last_error(const pml_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int pml_ctx_collective(const pml_ctx* c) { return !c ? PML_EINVAL : (c->nranks == 1 ? 0 : (c->peer_ok ? 2 : 1)); }

int pml_ctx_sync(pml_ctx* c) {
    if (!c) return PML_EINVAL;
    return c->bind() && c->sync() ? PML_OK : PML_ENODEVICE;
}

int pml_trace_enable(pml_ctx* c, int on) {
    if (!c || !c->bind() || !c->sync()) return PML_EINVAL;
    if (on && !c->d_trace_buf) {
        if (!c->cuda(cudaMalloc(&c->d_trace_buf, 96 * sizeof(long long)), "trace alloc")) return PML_ENOMEM;
    }
    if (c->d_trace_buf) cudaMemset(c->d_trace_buf, 0, 96 * sizeof(long long));
    c->d_trace = (on == 1 || on == 4 || on == 5) ? c->d_trace_buf : nullptr;
    c->trace_newview_tips = on == 4 ? 1 : (on == 5 ? 0 : -1);
    c->trace_fused = on == 6;
    c->d_trace_branch = on >= 2 ? c->d_trace_buf : nullptr;
    c->trace_branch_tip = on == 3;
    return PML_OK;
}

int pml_trace_read(pml_ctx* c, int64_t out[96]) {
    if (!c || !out || !c->d_trace_buf || !c->bind() || !c->sync()) return PML_EINVAL;
    return c->cuda(cudaMemcpy(out, c->d_trace_buf, 96 * sizeof(long long), cudaMemcpyDeviceToHost), "trace read") ? PML_OK : PML_ENODEVICE;
}

int pml_timeline_begin(pml_ctx* c) {
    if (!c || !c->bind() || !c->sync()) return PML_EINVAL;
    const size_t bytes = sizeof(unsigned long long) * 6 * pml_ctx::kTimelineCap;
    if (!c->d_timeline && !c->cuda(cudaMalloc(&c->d_timeline, bytes), "timeline alloc")) return PML_ENOMEM;
    cudaMemset(c->d_timeline, 0, bytes);
    c->timeline_kinds.clear();
    return PML_OK;
}

int pml_timeline_read(pml_ctx* c, uint64_t* stamps, int32_t* kinds, int cap) {
    if (!c || !c->d_timeline || !c->bind() || !c->sync()) return PML_EINVAL;
    const int n = std::min<int>(cap, (int)c->timeline_kinds.size());
    if (stamps && n > 0 && !c->cuda(cudaMemcpy(stamps, c->d_timeline, sizeof(uint64_t) * 6 * n, cudaMemcpyDeviceToHost), "timeline read")) return PML_ENODEVICE;
    if (kinds) std::copy(c->timeline_kinds.begin(), c->timeline_kinds.begin() + n, kinds);
    cudaFree(c->d_timeline);   // one capture per begin: launches stop being stamped
    c->d_timeline = nullptr;
    return n;
}

int pml_timer_start(pml_ctx* c) {
    if (!c || !c->bind()) return PML_EINVAL;
    if (!c->timer0) { cudaEventCreate(&c->timer0); cudaEventCreate(&c->timer1); }
    return c->cuda(cudaEventRecord(c->timer0, c->stream), "timer start") ? PML_OK : PML_ENODEVICE;
}

int pml_timer_stop(pml_ctx* c, double* ms) {
    if (!c || !ms || !c->timer0 || !c->bind()) return PML_EINVAL;
    float f = 0.f;
    if (!c->cuda(cudaEventRecord(c->timer1, c->stream), "timer stop") || !c->cuda(cudaEventSynchronize(c->timer1), "timer sync") ||
        !c->cuda(cudaEventElapsedTime(&f, c->timer0, c->timer1), "timer read"))
        return PML_ENODEVICE;
    *ms = f;
    return PML_OK;
}

// host builds of two pieces of device arithmetic, so that what DESIGN claims about them is a CPU test (tests/test_host.py)
double pml_debug_exp_neg(double x) { return pmat::exp_neg(x); }
int pml_debug_nr_step(double t, double d1, double d2, int in_t, double* t_new) {
    if (!t_new) return PML_EINVAL;
    const double tc = nr_clamp_length(t);
    return in_t ? nr_step_t(nr_early(tc), d1, d2, t_new) : nr_step_z(nr_z(tc), d1, d2, t_new);
}

int pml_kind_info(int kind, char* name, size_t cap, int* bytes_per_pattern, int* dmma_per_tile) {
    if (kind < 0 || kind >= PML_NKINDS) return PML_EINVAL;
    // per side: bytes read and DMMAs of its 20 x 20 product (a tip's is a table look-up)
    static const char* far_name[3] = {"inner", "tip", "cherry"};
    static const int side_bytes[3] = {640, 1, 2}, side_dmma[3] = {30, 0, 30};
    static const char* nv_name[6] = {"tip_tip", "tip_inner", "inner_inner", "tip_cherry", "cherry_cherry", "cherry_inner"};
    static const int nv_sides[6][2] = {{1, 1}, {1, 0}, {0, 0}, {1, 2}, {2, 2}, {2, 0}};
    std::string n;
    int bytes = 0, dmma = 0;
    if (kind < 6) {
        n = std::string("newview_") + nv_name[kind];
        bytes = side_bytes[nv_sides[kind][0]] + side_bytes[nv_sides[kind][1]] + 640;
        dmma = side_dmma[nv_sides[kind][0]] + side_dmma[nv_sides[kind][1]];
    } else if (kind < 12) {
        const int far = (kind - 6) % 3;
        const bool eval = kind >= 9;
        n = std::string(eval ? "evaluate_" : "branch_") + far_name[far];
        bytes = 640 + side_bytes[far] + 12 + (eval ? 8 : 0);   // + scaling counts and weight (+ per-pattern lnL written)
        dmma = 30 + side_dmma[far] + 12;
    } else if (kind == 12) {
        n = "core";
        bytes = 648;
    } else {
        const int nv = 1 + (kind - 13) / 3, far = (kind - 13) % 3;
        n = std::string("fused_") + nv_name[nv] + "_" + far_name[far];
        bytes = side_bytes[nv_sides[nv][0]] + side_bytes[nv_sides[nv][1]] + 640 + side_bytes[far];
        dmma = side_dmma[nv_sides[nv][0]] + side_dmma[nv_sides[nv][1]] + 30 + side_dmma[far] + 12;
    }
    if (name && cap > n.size()) std::memcpy(name, n.c_str(), n.size() + 1);
    if (bytes_per_pattern) *bytes_per_pattern = bytes;
    if (dmma_per_tile) *dmma_per_tile = dmma;
    return PML_OK;
}

int pml_profile_begin(pml_ctx* c) {
    if (!c) return PML_EINVAL;
    if (!c->bind() || !c->sync()) return PML_ENODEVICE;
    for (auto& t : c->timed) { c->spare_events.push_back(t.t0); c->spare_events.push_back(t.t1); }
    c->timed.clear();
    c->profiling = true;
    return PML_OK;
}

int pml_profile_end(pml_ctx* c, double ms[PML_NKINDS], int64_t launches[PML_NKINDS], int64_t rows[PML_NKINDS]) {
    if (!c) return PML_EINVAL;
    c->profiling = false;
    if (!c->bind() || !c->sync()) return PML_ENODEVICE;
    for (int k = 0; k < PML_NKINDS; ++k) {
        if (ms) ms[k] = 0.0;
        if (launches) launches[k] = 0;
        if (rows) rows[k] = 0;
    }
    for (auto& t : c->timed) {
        float f = 0.f;
        cudaEventElapsedTime(&f, t.t0, t.t1);
        if (ms) ms[t.kind] += f;
        if (launches) ++launches[t.kind];
        if (rows) rows[t.kind] += t.rows;
        c->spare_events.push_back(t.t0);
        c->spare_events.push_back(t.t1);
    }
    c->timed.clear();
    return PML_OK;
}

int pml_aln_load(pml_ctx* c, int ntax, int64_t nsites, const char* const* names, const uint8_t* chars,
                 const int32_t* site_weights, pml_aln** out) {
    if (!c || !out || ntax < 3 || nsites < 1 || !names || !chars) return fail(c, PML_EINVAL, "pml_aln_load: bad arguments");
    *out = nullptr;
    if (!c->bind()) return PML_ENODEVICE;
    auto a = std::make_unique<pml_aln>();
    a->ctx = c;
    // The column sort is global, the residue codes a rank keeps are those of its own pattern block.  Ranks of one process
    // (pml_group_create) crunch ONCE: the first rank to arrive does it, the others copy the result.  Ranks in separate
    // processes split the radix sort between them and exchange the sorted column order with one NCCL allreduce.
    if (c->group) {
        PmlGroup& g = *c->group;
        const uint64_t key[5] = {(uint64_t)(uintptr_t)chars, (uint64_t)ntax, (uint64_t)nsites,
                                 hash_bytes(chars, (size_t)ntax * (size_t)nsites),
                                 site_weights ? hash_bytes(site_weights, sizeof(int32_t) * (size_t)nsites) : 0};
        std::shared_ptr<const Patterns> shared;
        {
            std::lock_guard<std::mutex> lock(g.mu);
            if (!g.crunch || std::memcmp(g.crunch_key, key, sizeof key) != 0 || g.crunch_uses >= g.n) {
                auto fresh = std::make_shared<Patterns>();
                crunch_patterns(ntax, nsites, chars, site_weights, *fresh, 0, 1);
                fresh->codes.clear();
                fresh->codes.shrink_to_fit();
                g.crunch = fresh;
                std::memcpy(g.crunch_key, key, sizeof key);
                g.crunch_uses = 0;
            }
            ++g.crunch_uses;
            shared = g.crunch;
            if (g.crunch_uses >= g.n) g.crunch.reset();  // everybody has it
        }
        a->pat.ntax = shared->ntax;
        a->pat.nsites = shared->nsites;
        a->pat.npat = shared->npat;
        a->pat.weight = shared->weight;
        a->pat.site_to_pat = shared->site_to_pat;
        a->pat.codes_p0 = shared->npat * c->rank / c->nranks;
        a->pat.codes_n = shared->npat * (c->rank + 1) / c->nranks - a->pat.codes_p0;
        gather_codes(ntax, nsites, chars, shared->first, a->pat.codes_p0, a->pat.codes_n, a->pat.codes);
    } else if (c->nranks > 1 && !getenv("PEPRML_NO_SHARED_CRUNCH")) {
        const CrunchShare share = [c](int64_t* order, uint8_t* fresh, int64_t n) {
            int64_t* d_order = nullptr;
            uint8_t* d_fresh = nullptr;
            bool ok = c->cuda(c->dev_alloc(&d_order, sizeof(int64_t) * (size_t)n), "crunch exchange alloc") &&
                      c->cuda(c->dev_alloc(&d_fresh, (size_t)n), "crunch exchange alloc") &&
                      c->cuda(cudaMemcpyAsync(d_order, order, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream), "crunch upload") &&
                      c->cuda(cudaMemcpyAsync(d_fresh, fresh, (size_t)n, cudaMemcpyHostToDevice, c->stream), "crunch upload");
            // every rank must enter the collectives, whatever happened locally (a local failure shows up as a sum that is not a permutation)
            const bool sent = g_nccl.AllReduce(d_order, d_order, (size_t)n, ncclInt64, ncclSum, c->comm, c->stream) == ncclSuccess &&
                              g_nccl.AllReduce(d_fresh, d_fresh, (size_t)n, ncclUint8, ncclSum, c->comm, c->stream) == ncclSuccess;
            ok = ok && sent &&
                 c->cuda(cudaMemcpyAsync(order, d_order, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "crunch download") &&
                 c->cuda(cudaMemcpyAsync(fresh, d_fresh, (size_t)n, cudaMemcpyDeviceToHost, c->stream), "crunch download") && c->sync();
            c->dev_free(d_order);
            c->dev_free(d_fresh);
            return ok;
        };
        crunch_patterns(ntax, nsites, chars, site_weights, a->pat, c->rank, c->nranks, &share, true);  // own block now, the rest on demand
    } else {
        crunch_patterns(ntax, nsites, chars, site_weights, a->pat, c->rank, c->nranks);  // codes of this rank's block only
    }
    for (int i = 0; i < ntax; ++i) a->pat.names.emplace_back(names[i]);
    if (a->pat.npat == 0) return fail(c, PML_EINVAL, "alignment has no column with positive weight");
    const int64_t np = a->pat.npat;
    a->p0 = np * c->rank / c->nranks;
    a->nloc = np * (c->rank + 1) / c->nranks - a->p0;
    a->npad = std::max<int64_t>(kPad, (a->nloc + kPad - 1) / kPad * kPad);
    // device copies of this rank's slice; padded rows are "undetermined" with weight 0
    std::vector<uint8_t> hc((size_t)ntax * a->npad, 22);
    for (int t = 0; t < ntax; ++t)
        std::memcpy(hc.data() + (size_t)t * a->npad, a->pat.codes.data() + (size_t)t * a->pat.codes_n, a->nloc);
    std::vector<int32_t> hw(a->npad, 0);
    if (a->pat.full) std::copy(a->pat.weight.begin() + a->p0, a->pat.weight.begin() + a->p0 + a->nloc, hw.begin());
    else std::copy(a->pat.own_weight.begin(), a->pat.own_weight.end(), hw.begin());
    bool ok = c->cuda(c->dev_alloc(&a->d_codes, hc.size()), "codes alloc") &&
              c->cuda(c->dev_alloc(&a->d_weights, sizeof(int32_t) * a->npad), "weights alloc") &&
              c->cuda(c->dev_alloc(&a->d_wcustom, sizeof(int32_t) * a->npad), "weights alloc") &&
              c->cuda(c->dev_alloc(&a->d_model, sizeof(DeviceModel)), "model alloc") &&
              c->cuda(c->dev_alloc(&a->d_site_lnl, sizeof(double) * a->npad), "site lnl alloc") &&
              c->cuda(c->dev_alloc(&a->d_partials, sizeof(double) * reduce_partials_capacity(a->npad)), "partials alloc") &&
              c->cuda(c->dev_alloc(&a->d_result, sizeof(double) * 16), "result alloc") &&
              c->cuda(c->dev_alloc(&a->d_ticket, 64), "ticket alloc") &&
              c->cuda(cudaMemset(a->d_ticket, 0, 64), "ticket clear") &&
              c->cuda(cudaMemcpy(a->d_codes, hc.data(), hc.size(), cudaMemcpyHostToDevice), "codes upload") &&
              c->cuda(cudaMemcpy(a->d_weights, hw.data(), sizeof(int32_t) * a->npad, cudaMemcpyHostToDevice), "weights upload") &&
              c->cuda(cudaMemset(a->d_wcustom, 0, sizeof(int32_t) * a->npad), "weights clear");
    if (!ok) {
        pml_aln_free(a.release());
        return PML_ENOMEM;
    }
    *out = a.release();
    return PML_OK;
}

int pml_aln_load_phylip(pml_ctx* c, const char* path, const char* weights_path, pml_aln** out) {
    if (!c || !path || !out) return fail(c, PML_EINVAL, "pml_aln_load_phylip: bad arguments");
    std::vector<std::string> names;
    std::vector<uint8_t> chars;
    int64_t nsites = 0;
    std::string err;
    if (!read_phylip(path, names, chars, nsites, err)) return fail(c, PML_EINVAL, err);
    std::vector<int32_t> w;
    if (weights_path) {
        std::ifstream in(weights_path);
        long v;
        while (in >> v) w.push_back((int32_t)v);
        if ((int64_t)w.size() != nsites) return fail(c, PML_EINVAL, "weight file must hold one integer per alignment column");
    }
    std::vector<const char*> np;
    for (auto& s : names) np.push_back(s.c_str());
    return pml_aln_load(c, (int)names.size(), nsites, np.data(), chars.data(), w.empty() ? nullptr : w.data(), out);
}

void pml_aln_free(pml_aln* a) {
    if (!a) return;
    a->ctx->bind();
    cudaStreamSynchronize(a->ctx->stream);
    a->ctx->dev_free(a->d_codes);
    a->ctx->dev_free(a->d_weights);
    a->ctx->dev_free(a->d_wcustom);
    a->ctx->dev_free(a->d_model);
    a->ctx->dev_free(a->d_site_lnl);
    a->ctx->dev_free(a->d_partials);
    a->ctx->dev_free(a->d_result);
    a->ctx->dev_free(a->d_ticket);
    a->ctx->dev_free(a->d_sumtable);
    a->ctx->dev_free(a->d_sumscale);
    delete a;
}

int pml_aln_dims(const pml_aln* a, int* ntax, int64_t* nsites, int64_t* npat, int64_t* nloc) {
    if (!a) return PML_EINVAL;
    if (ntax) *ntax = a->pat.ntax;
    if (nsites) *nsites = a->pat.nsites;
    if (npat) *npat = a->pat.npat;
    if (nloc) *nloc = a->nloc;
    return PML_OK;
}

int pml_aln_patterns(const pml_aln* a, int32_t* weights, int64_t* site_to_pattern) {
    if (!a) return PML_EINVAL;
    finish_patterns(const_cast<pml_aln*>(a)->pat);
    if (weights) std::copy(a->pat.weight.begin(), a->pat.weight.end(), weights);
    if (site_to_pattern) std::copy(a->pat.site_to_pat.begin(), a->pat.site_to_pat.end(), site_to_pattern);
    return PML_OK;
}

const char* pml_aln_name(const pml_aln* a, int taxon) {
    return (a && taxon >= 0 && taxon < a->pat.ntax) ? a->pat.names[taxon].c_str() : nullptr;
}

int pml_model_set(pml_aln* a, const char* model, double alpha) {
    if (!a) return PML_EINVAL;
    pml_ctx* c = a->ctx;
    if (model && std::strcmp(model, "PROTGAMMAWAG") != 0)
        return fail(c, PML_EINVAL, std::string("unsupported model '") + model + "' (only PROTGAMMAWAG)");
    if (!(alpha >= kAlphaMin && alpha <= kAlphaMax)) return fail(c, PML_EINVAL, "alpha must lie in [0.02, 1000]");
    if (!c->bind()) return PML_ENODEVICE;
    a->alpha = alpha;
    gamma_mean_rates(alpha, kCats, a->rates);
    if (!upload_model(a)) return PML_ENODEVICE;
    a->model_set = true;
    return PML_OK;
}

int pml_model_get(const pml_aln* a, double* alpha, double rates4[4]) {
    if (!a) return PML_EINVAL;
    if (alpha) *alpha = a->alpha;
    if (rates4) std::memcpy(rates4, a->rates, sizeof a->rates);
    return PML_OK;
}

int pml_wag_pmatrix(double t, double rate, double P[400]) {
    wag_pmatrix(t, rate, P);
    return PML_OK;
}
int pml_wag_frequencies(double pi[20]) {
    std::memcpy(pi, wag_eigensystem().pi, sizeof(double) * 20);
    return PML_OK;
}
int pml_gamma_rates(double alpha, int ncat, double* rates) {
    if (!(alpha > 0) || ncat < 1 || !rates) return PML_EINVAL;
    gamma_mean_rates(alpha, ncat, rates);
    return PML_OK;
}

int pml_tree_load(pml_aln* a, const char* newick, pml_tree** out) {
    if (!a || !newick || !out) return PML_EINVAL;
    pml_ctx* c = a->ctx;
    *out = nullptr;
    if (!a->model_set) return fail(c, PML_ESTATE, "pml_model_set must be called before pml_tree_load");
    if (!c->bind()) return PML_ENODEVICE;
    auto t = std::make_unique<pml_tree>();
    t->aln = a;
    std::string err;
    if (!parse_newick(newick, a->pat.names, kDefaultLen, t->topo, err)) return fail(c, PML_EINVAL, err);
    t->views.fold_cherries = c->fold;
    t->views.reset(t->topo);
    const size_t inner = (size_t)a->pat.ntax - 2;
    bool ok = c->cuda(c->dev_alloc(&t->d_clv, sizeof(double) * inner * a->npad * kRow), "CLV arena alloc") &&
              c->cuda(c->dev_alloc(&t->d_scale, sizeof(int32_t) * inner * a->npad), "scaler alloc") &&
              c->cuda(c->dev_alloc(&t->d_len, sizeof(double) * std::max(1, t->topo.nedges())), "lengths alloc");
    if (!ok) {
        pml_tree_free(t.release());
        return PML_ENOMEM;
    }
    t->model_epoch = a->model_epoch;
    *out = t.release();
    return PML_OK;
}

void pml_tree_free(pml_tree* t) {
    if (!t) return;
    t->aln->ctx->bind();
    cudaStreamSynchronize(t->aln->ctx->stream);
    t->aln->ctx->dev_free(t->d_clv);
    t->aln->ctx->dev_free(t->d_scale);
    t->aln->ctx->dev_free(t->d_len);
    if (t->aln->sumtable_owner == t) t->aln->sumtable_owner = nullptr;
    delete t;
}

int pml_tree_num_branches(const pml_tree* t) { return t ? t->topo.nedges() : PML_EINVAL; }

int pml_tree_branch(const pml_tree* t, int e, int* a, int* b, double* len) {
    if (!t || e < 0 || e >= t->topo.nedges()) return PML_EINVAL;
    if (a) *a = t->topo.ea[e];
    if (b) *b = t->topo.eb[e];
    if (len) *len = t->topo.len[e];
    return PML_OK;
}

int pml_tree_set_branch(pml_tree* t, int e, double len) {
    if (!t || e < 0 || e >= t->topo.nedges() || !(len >= 0.0)) return PML_EINVAL;
    set_branch(t, e, len);
    return PML_OK;
}

int64_t pml_tree_newick(const pml_tree* t, char* buf, size_t cap) {
    if (!t) return PML_EINVAL;
    const std::string s = write_newick_result(t->topo, t->aln->pat.names);
    if (buf && cap > s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size() + 1;
}

int pml_tree_invalidate(pml_tree* t) {
    if (!t) return PML_EINVAL;
    t->views.reset(t->topo);
    t->prepared_branch = -1;
    return PML_OK;
}

int64_t pml_tree_nr_retries(const pml_tree* t) { return t ? t->nr_retries : PML_EINVAL; }

int pml_tree_stats(const pml_tree* t, int64_t site_updates[3], int64_t* launches) {
    if (!t) return PML_EINVAL;
    if (site_updates) {
        std::memcpy(site_updates, t->site_updates, sizeof t->site_updates);
        site_updates[0] += t->views.folded * t->aln->nloc;  // tip-tip updates formed inside their consumers (folded cherries)
    }
    if (launches) *launches = t->launches;
    return PML_OK;
}

int pml_evaluate(pml_tree* t, const int32_t* weights, double* lnl, double* per_site) {
    if (!t || !lnl) return PML_EINVAL;
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    const int rc = evaluate_branch(t, t->topo.edge[0][0], weights, lnl);
    if (rc != PML_OK || !per_site) return rc;
    finish_patterns(a->pat);
    std::vector<double> pp(a->npad);
    if (!c->cuda(cudaMemcpy(pp.data(), a->d_site_lnl, sizeof(double) * a->npad, cudaMemcpyDeviceToHost), "site lnL download"))
        return PML_ENODEVICE;
    for (int64_t s = 0; s < a->pat.nsites; ++s) {
        const int64_t p = a->pat.site_to_pat[s];
        per_site[s] = (p >= a->p0 && p < a->p0 + a->nloc) ? pp[p - a->p0] : 0.0;
    }
    return PML_OK;
}

int pml_branch_derivs(pml_tree* t, int branch, double len, const int32_t* weights, double* lnl, double* d1, double* d2) {
    if (!t || branch < 0 || branch >= t->topo.nedges() || !(len >= 0.0)) return PML_EINVAL;
    pml_ctx* c = t->aln->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    const int32_t* dw = device_weights(t->aln, weights);
    if (!dw) return PML_ENODEVICE;
    double r[3];
    if (t->prepared_branch != branch) {
        if (!branch_pass(t, branch, dw, len, true, false, r)) return c->fail_code();
    } else if (!core_at(t, dw, len, r)) return c->fail_code();
    if (lnl) *lnl = r[0];
    if (d1) *d1 = r[1];
    if (d2) *d2 = r[2];
    return PML_OK;
}

int pml_smooth_branches(pml_tree* t, int sweeps, const int32_t* weights, int* converged) {
    if (!t || sweeps < 1) return PML_EINVAL;
    pml_ctx* c = t->aln->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    const int32_t* dw = device_weights(t->aln, weights);
    if (!dw) return PML_ENODEVICE;
    bool smoothed = false;
    for (int s = 0; s < sweeps && !smoothed; ++s)
        if (!smooth_sweep(t, dw, smoothed)) return c->fail_code();
    if (converged) *converged = smoothed ? 1 : 0;
    return PML_OK;
}

int pml_optimize(pml_tree* t, int opt_alpha, double eps, const int32_t* weights, double* lnl, double* alpha) {
    if (!t) return PML_EINVAL;
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    if (!(eps > 0.0)) eps = 0.1;
    const int32_t* dw = device_weights(a, weights);
    if (!dw) return PML_ENODEVICE;
    // modOpt: { smooth (2 sweeps max), Brent on alpha, smooth (3 sweeps max) } until a round gains <= eps
    a->alpha_step = 0.1;  // every call brackets alpha from the same first probe: results do not depend on earlier calls
    double cur, best = 0.0;
    if (evaluate_branch(t, t->topo.edge[0][0], weights, &best) != PML_OK) return c->fail_code();
    int rounds = 0;
    do {
        cur = best;
        if (!tree_evaluate(t, weights, dw, 0.0625, &best, opt_alpha != 0)) return c->fail_code();
        if (opt_alpha) {
            if (!optimise_alpha(t, weights, 1.0e-3, &best)) return c->fail_code();
        }
        if (!tree_evaluate(t, weights, dw, 0.1, &best)) return c->fail_code();
    } while (std::fabs(cur - best) > eps && ++rounds < 200);
    if (lnl) *lnl = best;
    if (alpha) *alpha = a->alpha;
    return PML_OK;
}

int pml_tree_start_parsimony(pml_aln* a, int64_t seed, const int32_t* weights, pml_tree** out) {
    if (!a || !out) return PML_EINVAL;
    pml_ctx* c = a->ctx;
    *out = nullptr;
    if (!c->bind()) return PML_ENODEVICE;
    const int32_t* dw = device_weights(a, weights);
    if (!dw) return PML_ENODEVICE;
    // host control flow (host.cpp) + one device scan per added taxon (parsimony.cu)
    const int maxnodes = 2 * a->pat.ntax;
    uint32_t *d_down = nullptr, *d_up = nullptr;
    int4* d_nodes = nullptr;
    int* d_pre = nullptr;
    unsigned long long* d_out = nullptr;
    bool ok = c->cuda(c->dev_alloc(&d_down, sizeof(uint32_t) * (size_t)maxnodes * a->npad), "parsimony alloc") &&
              c->cuda(c->dev_alloc(&d_up, sizeof(uint32_t) * (size_t)maxnodes * a->npad), "parsimony alloc") &&
              c->cuda(c->dev_alloc(&d_nodes, sizeof(int4) * maxnodes), "parsimony alloc") &&
              c->cuda(c->dev_alloc(&d_pre, sizeof(int) * maxnodes), "parsimony alloc") &&
              c->cuda(c->dev_alloc(&d_out, sizeof(unsigned long long) * (maxnodes + 1)), "parsimony alloc");
    std::vector<unsigned long long> h_out(maxnodes + 1);
    ParsimonyScan scan = [&](const GrowTree& g, const std::vector<int>& pre, int next_taxon, int64_t& score, std::vector<int64_t>& cost) {
        const int N = (int)g.parent.size(), npre = (int)pre.size();
        auto* hn = (int4*)c->stage(sizeof(int4) * N + sizeof(int) * npre);
        if (!hn) return false;
        int* hp = (int*)(hn + N);
        for (int v = 0; v < N; ++v) hn[v] = make_int4(g.left[v], g.right[v], g.taxon[v], g.parent[v]);
        std::copy(pre.begin(), pre.end(), hp);
        ParsimonyArgs pa{a->d_codes, dw, a->npad, a->nloc, d_nodes, d_pre, npre, g.left[0], g.taxon[0], next_taxon, d_down, d_up, d_out};
        if (!c->cuda(cudaMemcpyAsync(d_nodes, hn, sizeof(int4) * N, cudaMemcpyHostToDevice, c->stream), "parsimony upload") ||
            !c->cuda(cudaMemcpyAsync(d_pre, hp, sizeof(int) * npre, cudaMemcpyHostToDevice, c->stream), "parsimony upload") ||
            !c->cuda(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long) * (npre + 1), c->stream), "parsimony clear"))
            return false;
        launch_parsimony_scan(pa, c->stream);
        if (c->nranks > 1 &&
            g_nccl.AllReduce(d_out, d_out, (size_t)npre + 1, ncclUint64, ncclSum, c->comm, c->stream) != ncclSuccess) {
            c->err = "ncclAllReduce: parsimony costs";
            return false;
        }
        if (!c->cuda(cudaMemcpyAsync(h_out.data(), d_out, sizeof(unsigned long long) * (npre + 1), cudaMemcpyDeviceToHost, c->stream),
                     "parsimony download") ||
            !c->sync())
            return false;
        score = (int64_t)h_out[0];
        cost.assign(npre, 0);
        for (int i = 0; i < npre; ++i) cost[i] = (int64_t)h_out[1 + i];
        return true;
    };
    Topology topo;
    ok = ok && parsimony_start_tree(a->pat, seed, kDefaultLen, topo, nullptr, &scan, &a->constraints);
    c->sync();
    c->dev_free(d_down);
    c->dev_free(d_up);
    c->dev_free(d_nodes);
    c->dev_free(d_pre);
    c->dev_free(d_out);
    if (!ok) return c->err.empty() ? fail(c, PML_EINVAL, "parsimony start tree failed") : PML_ENODEVICE;
    if (topo.ntax != a->pat.ntax) return fail(c, PML_EINVAL, "parsimony start tree failed");
    const std::string text = write_newick_result(topo, a->pat.names);
    return pml_tree_load(a, text.substr(0, text.size() - 5).append(";").c_str(), out);
}

int64_t pml_parsimony_tree(int ntax, int64_t nsites, const char* const* names, const uint8_t* chars, int64_t seed, char* buf,
                           size_t cap, int64_t* score) {
    if (ntax < 3 || nsites < 1 || !names || !chars) return PML_EINVAL;
    Patterns pat;
    crunch_patterns(ntax, nsites, chars, nullptr, pat);
    for (int i = 0; i < ntax; ++i) pat.names.emplace_back(names[i]);
    Topology topo;
    parsimony_start_tree(pat, seed, kDefaultLen, topo, score);
    if (topo.ntax != ntax) return fail(nullptr, PML_EINVAL, "parsimony start tree failed");
    const std::string s = write_newick_result(topo, pat.names);
    if (buf && cap > s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size() + 1;
}

int64_t pml_parsimony_tree_constrained(int ntax, int64_t nsites, const char* const* names, const uint8_t* chars, int64_t seed,
                                       const char* constraints, char* buf, size_t cap, int64_t* score) {
    if (ntax < 3 || nsites < 1 || !names || !chars) return PML_EINVAL;
    Patterns pat;
    crunch_patterns(ntax, nsites, chars, nullptr, pat);
    for (int i = 0; i < ntax; ++i) pat.names.emplace_back(names[i]);
    Constraints cons;
    std::string err;
    if (constraints && *constraints && !parse_constraints(constraints, pat.names, cons, err)) return fail(nullptr, PML_EINVAL, err);
    Topology topo;
    if (!parsimony_start_tree(pat, seed, kDefaultLen, topo, score, nullptr, &cons) || topo.ntax != ntax)
        return fail(nullptr, PML_EINVAL, "parsimony start tree failed (contradictory constraints?)");
    const std::string s = write_newick_result(topo, pat.names);
    if (buf && cap > s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size() + 1;
}

int pml_newick_satisfies_constraints(const char* newick, const char* const* names, int ntax, const char* constraints) {
    if (!newick || !names || ntax < 3 || !constraints) return PML_EINVAL;
    std::vector<std::string> nm(names, names + ntax);
    Topology topo;
    Constraints cons;
    std::string err;
    if (!parse_newick(newick, nm, kDefaultLen, topo, err) || !parse_constraints(constraints, nm, cons, err)) return fail(nullptr, PML_EINVAL, err);
    return satisfies(topo, cons) ? 1 : 0;
}

int pml_tree_spr(pml_tree* t, int node, int keep, int target) {
    if (!t) return PML_EINVAL;
    Topology& T = t->topo;
    if (node < T.ntax || node >= T.nnodes() || T.slot_of(node, keep) < 0 || target < 0 || target >= T.nedges())
        return fail(t->aln->ctx, PML_EINVAL, "bad pruning-regrafting move");
    const std::vector<int> ok = spr_targets(T, node, keep, T.nnodes());
    if (std::find(ok.begin(), ok.end(), target) == ok.end())
        return fail(t->aln->ctx, PML_EINVAL, "target branch lies inside the moved subtree or next to the pruning point");
    SprMove mv;
    if (!spr_apply(T, t->views, node, keep, target, mv)) return fail(t->aln->ctx, PML_EINVAL, "move rejected");
    t->prepared_branch = -1;
    return PML_OK;
}

int pml_tree_neighbors(const pml_tree* t, int node, int nbr[3]) {
    if (!t || !nbr || node < 0 || node >= t->topo.nnodes()) return PML_EINVAL;
    for (int k = 0; k < 3; ++k) nbr[k] = t->topo.nbr[node][k];
    return PML_OK;
}

int pml_score_spr_candidates(pml_tree* t, int node, int keep, int radius, const int32_t* weights, int* targets, double* lnl,
                             int* ncand) {
    if (!t || !ncand || !targets || !lnl || radius < 1) return PML_EINVAL;
    Topology& T = t->topo;
    if (node < T.ntax || node >= T.nnodes() || T.slot_of(node, keep) < 0) return fail(t->aln->ctx, PML_EINVAL, "bad (node, keep) pair");
    pml_ctx* c = t->aln->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    const int32_t* dw = device_weights(t->aln, weights);
    if (!dw) return PML_ENODEVICE;
    std::vector<int> cand = spr_targets(T, node, keep, radius);
    if ((int)cand.size() > *ncand) cand.resize(*ncand);
    std::vector<double> lazy;
    if (!score_candidates(t, node, keep, cand, dw, lazy)) return c->fail_code();
    int n = 0;
    for (size_t i = 0; i < cand.size(); ++i) {
        if (!std::isfinite(lazy[i])) continue;
        targets[n] = cand[i];
        lnl[n] = lazy[i];
        ++n;
    }
    *ncand = n;
    return PML_OK;
}

int pml_search(pml_tree* t, int radius, int max_rounds, double eps, const int32_t* weights, double* lnl, int* accepted_moves) {
    if (!t || radius < 1 || max_rounds < 1) return PML_EINVAL;
    pml_ctx* c = t->aln->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    if (!(eps > 0.0)) eps = 0.1;
    const int32_t* dw = device_weights(t->aln, weights);
    if (!dw) return PML_ENODEVICE;
    if (!satisfies(t->topo, t->aln->constraints))
        return fail(c, PML_ESTATE, "pml_search: the tree does not display the alignment's topological constraints (start from pml_tree_start_parsimony)");
    double best;
    if (!lnl_at(t, t->topo.edge[0][0], dw, best)) return c->fail_code();
    SearchStats st;
    for (int round = 0; round < max_rounds; ++round) {
        const double before = best;
        if (!spr_round(t, dw, radius, best, st)) return c->fail_code();
        // settle all branch lengths on the new topology before the next round
        bool smoothed = false;
        for (int sweep = 0; sweep < 2 && !smoothed; ++sweep)
            if (!smooth_sweep(t, dw, smoothed)) return c->fail_code();
        if (!lnl_at(t, t->topo.edge[0][0], dw, best)) return c->fail_code();
        if (best - before < eps) break;
    }
    if (lnl) *lnl = best;
    if (accepted_moves) *accepted_moves = st.accepted;
    return PML_OK;
}

int pml_aln_set_constraints(pml_aln* a, const char* text) {
    if (!a) return PML_EINVAL;
    if (!text || !*text) {
        a->constraints = Constraints();
        return PML_OK;
    }
    std::string err;
    Constraints parsed;
    if (!parse_constraints(text, a->pat.names, parsed, err)) return fail(a->ctx, PML_EINVAL, err);
    a->constraints = std::move(parsed);
    return PML_OK;
}

int pml_aln_num_constraints(const pml_aln* a) { return a ? (int)a->constraints.splits.size() : PML_EINVAL; }

int pml_tree_satisfies_constraints(const pml_tree* t) {
    if (!t) return PML_EINVAL;
    return satisfies(t->topo, t->aln->constraints) ? 1 : 0;
}

int64_t pml_constraints_from_tree(const char* newick, char* buf, size_t cap) {
    if (!newick) return PML_EINVAL;
    std::string err;
    const std::string s = constraints_from_tree(newick, err);
    if (s.empty()) return fail(nullptr, PML_EINVAL, err);
    if (buf && cap > s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size() + 1;
}

int pml_bootstrap_weights(const pml_aln* a, int64_t* seed, int nrep, int32_t* out) {
    if (!a || !seed || nrep < 0 || !out) return PML_EINVAL;
    finish_patterns(const_cast<pml_aln*>(a)->pat);
    bootstrap_replicates(seed, a->pat.weight, nrep, out);
    return PML_OK;
}

int pml_bootstrap_weights_host(const int32_t* pw, int64_t npat, int64_t* seed, int nrep, int32_t* out) {
    if (!pw || npat < 1 || !seed || nrep < 0 || !out) return PML_EINVAL;
    bootstrap_replicates(seed, std::vector<int32_t>(pw, pw + npat), nrep, out);
    return PML_OK;
}

int64_t pml_newick_capacity(const pml_aln* a) {
    if (!a) return PML_EINVAL;
    size_t longest = 0;
    for (const std::string& n : a->pat.names) longest = std::max(longest, n.size());
    return (int64_t)((size_t)a->pat.ntax * (2 * longest + 64) + 64);
}

int pml_bootstrap_trees(pml_aln* a, int64_t weight_seed, int64_t parsimony_seed, int nrep, int first, int stride, int radius,
                        int rounds, double eps, char* newicks, size_t cap, double* lnl, double* seconds) {
    if (!a || nrep < 0 || first < 0 || stride < 1 || radius < 1 || rounds < 0 || (!newicks && cap > 0)) return PML_EINVAL;
    pml_ctx* c = a->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    if (!(eps > 0.0)) eps = 0.1;
    finish_patterns(a->pat);
    const int64_t np = a->pat.npat;
    // the WHOLE weight stream is drawn (it is sequential by construction), so replicate r carries the same weights whatever
    // the sharding; only this share's vectors are kept
    std::vector<int32_t> w((size_t)np), mine;
    std::vector<int> ids;
    int64_t seed = weight_seed;
    for (int r = 0; r < nrep; ++r) {
        bootstrap_replicates(&seed, a->pat.weight, 1, w.data());
        if (r >= first && (r - first) % stride == 0) {
            mine.insert(mine.end(), w.begin(), w.end());
            ids.push_back(r);
        }
    }
    std::string text;
    const double alpha_saved = a->alpha;
    for (size_t k = 0; k < ids.size(); ++k) {
        const auto t0 = std::chrono::steady_clock::now();
        const int32_t* wk = mine.data() + k * (size_t)np;
        pml_tree* t = nullptr;
        int rc = pml_model_set(a, nullptr, 1.0);
        if (rc == PML_OK) rc = pml_tree_start_parsimony(a, parsimony_seed + 1 + ids[k], wk, &t);
        double l = 0.0, al = 1.0;
        if (rc == PML_OK) rc = pml_optimize(t, 1, 5.0, wk, &l, &al);
        int moves = 0;
        if (rc == PML_OK && rounds > 0) rc = pml_search(t, radius, rounds, eps, wk, &l, &moves);
        if (rc != PML_OK) {
            pml_tree_free(t);
            return rc;
        }
        const std::string nw = write_newick_result(t->topo, a->pat.names);
        text += nw.substr(0, nw.size() - 5) + ";\n";  // RAxML_bootstrap form: no ":0.0" after the last parenthesis
        if (lnl) lnl[ids[k]] = l;
        pml_tree_free(t);
        if (seconds) seconds[ids[k]] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    pml_model_set(a, nullptr, alpha_saved);
    if (text.size() + 1 > cap) return fail(c, PML_EINVAL, "pml_bootstrap_trees: buffer too small (see pml_newick_capacity)");
    if (newicks) std::memcpy(newicks, text.c_str(), text.size() + 1);
    return PML_OK;
}

int pml_crunch_patterns(int ntax, int64_t nsites, const uint8_t* chars, const int32_t* site_weights, uint8_t* codes_out,
                        int32_t* weights_out, int64_t* site_to_pattern, int64_t* npatterns) {
    if (ntax < 1 || nsites < 1 || !chars || !npatterns) return PML_EINVAL;
    Patterns p;
    crunch_patterns(ntax, nsites, chars, site_weights, p);
    *npatterns = p.npat;
    if (codes_out) std::copy(p.codes.begin(), p.codes.end(), codes_out);
    if (weights_out) std::copy(p.weight.begin(), p.weight.end(), weights_out);
    if (site_to_pattern) std::copy(p.site_to_pat.begin(), p.site_to_pat.end(), site_to_pattern);
    return PML_OK;
}

int pml_crunch_patterns_sharded(int nranks, int ntax, int64_t nsites, const uint8_t* chars, const int32_t* site_weights,
                                uint8_t* codes_out, int32_t* weights_out, int64_t* site_to_pattern, int64_t* npatterns) {
    if (nranks < 1 || nranks > 64 || ntax < 1 || nsites < 1 || !chars || !npatterns) return PML_EINVAL;
    // the ranks of a multi-process group, played by threads: the exchange is an element-wise sum behind a barrier
    struct Exchange {
        std::mutex mu;
        std::condition_variable cv;
        int arrived = 0, generation = 0;
        std::vector<int64_t> order;
        std::vector<uint8_t> fresh;
    } ex;
    std::vector<Patterns> pats(nranks);
    std::vector<std::thread> pool;
    for (int r = 0; r < nranks; ++r)
        pool.emplace_back([&, r] {
            const CrunchShare share = [&](int64_t* order, uint8_t* fresh, int64_t n) {
                std::unique_lock<std::mutex> lock(ex.mu);
                if (ex.arrived == 0) {
                    ex.order.assign((size_t)n, 0);
                    ex.fresh.assign((size_t)n, 0);
                }
                for (int64_t i = 0; i < n; ++i) {
                    ex.order[i] += order[i];
                    ex.fresh[i] = (uint8_t)(ex.fresh[i] + fresh[i]);
                }
                const int gen = ex.generation;
                if (++ex.arrived == nranks) {
                    ++ex.generation;
                    ex.cv.notify_all();
                } else ex.cv.wait(lock, [&] { return ex.generation != gen; });
                std::copy(ex.order.begin(), ex.order.end(), order);
                std::copy(ex.fresh.begin(), ex.fresh.end(), fresh);
                if (--ex.arrived == 0) ex.order.clear();
                return true;
            };
            crunch_patterns(ntax, nsites, chars, site_weights, pats[r], r, nranks, &share, true);
        });
    for (auto& th : pool) th.join();
    for (auto& p : pats) finish_patterns(p);
    const Patterns& p0 = pats[0];
    for (int r = 1; r < nranks; ++r)
        if (pats[r].npat != p0.npat || pats[r].weight != p0.weight || pats[r].site_to_pat != p0.site_to_pat || pats[r].first != p0.first)
            return fail(nullptr, PML_ESTATE, "sharded pattern crunch: ranks disagree");
    *npatterns = p0.npat;
    if (codes_out)  // every rank gathered the codes of its own block: put the blocks side by side
        for (int r = 0; r < nranks; ++r)
            for (int t = 0; t < ntax; ++t)
                std::copy(pats[r].codes.begin() + (size_t)t * pats[r].codes_n, pats[r].codes.begin() + (size_t)(t + 1) * pats[r].codes_n,
                          codes_out + (size_t)t * p0.npat + pats[r].codes_p0);
    if (weights_out) std::copy(p0.weight.begin(), p0.weight.end(), weights_out);
    if (site_to_pattern) std::copy(p0.site_to_pat.begin(), p0.site_to_pat.end(), site_to_pattern);
    return PML_OK;
}

int pml_evaluate_replicates(pml_tree* t, const int32_t* W, int nrep, double* lnl) {
    if (!t || !W || nrep < 1 || !lnl) return PML_EINVAL;
    pml_aln* a = t->aln;
    pml_ctx* c = a->ctx;
    if (!c->bind()) return PML_ENODEVICE;
    adopt_model(t);
    double base;
    const int rc = evaluate_branch(t, t->topo.edge[0][0], nullptr, &base);  // fills d_site_lnl
    if (rc != PML_OK) return rc;
    int32_t* dW = nullptr;
    double* dl = nullptr;
    const int64_t np = a->pat.npat;
    bool ok = c->cuda(c->dev_alloc(&dW, sizeof(int32_t) * (size_t)nrep * a->npad), "replicate weights alloc") &&
              c->cuda(c->dev_alloc(&dl, sizeof(double) * nrep), "replicate lnL alloc") &&
              c->cuda(cudaMemsetAsync(dW, 0, sizeof(int32_t) * (size_t)nrep * a->npad, c->stream), "replicate weights clear") &&
              c->cuda(cudaMemcpy2DAsync(dW, sizeof(int32_t) * a->npad, W + a->p0, sizeof(int32_t) * np, sizeof(int32_t) * a->nloc,
                                        nrep, cudaMemcpyHostToDevice, c->stream),
                      "replicate weights upload");
    if (ok) {
        launch_replicate_lnl(dW, nrep, a->npad, a->npad, a->d_site_lnl, dl, c->stream);
        ++t->launches;
        ok = c->cuda(cudaGetLastError(), "replicate kernel") && c->allreduce(dl, nrep) &&
             c->cuda(cudaMemcpyAsync(lnl, dl, sizeof(double) * nrep, cudaMemcpyDeviceToHost, c->stream), "replicate lnL download") &&
             c->sync();
    }
    c->dev_free(dW);
    c->dev_free(dl);
    return ok ? PML_OK : PML_ENODEVICE;
}

int64_t pml_support_tree(const char* main_newick, const char* const* trees, int ntrees, int as_percent, char* buf, size_t cap) {
    if (!main_newick || ntrees < 0 || (ntrees > 0 && !trees)) return PML_EINVAL;
    std::vector<std::string> ts;
    for (int i = 0; i < ntrees; ++i) ts.emplace_back(trees[i]);
    std::string err;
    const std::string s = support_tree(main_newick, ts, as_percent != 0, nullptr, err);
    if (s.empty()) return fail(nullptr, PML_EINVAL, err);
    if (buf && cap > s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int64_t)s.size() + 1;
}

int pml_support_counts(const char* main_newick, const char* const* trees, int ntrees, int32_t* counts, int* nsplits) {
    if (!main_newick || ntrees < 0 || (ntrees > 0 && !trees)) return PML_EINVAL;
    std::vector<std::string> ts;
    for (int i = 0; i < ntrees; ++i) ts.emplace_back(trees[i]);
    std::string err;
    std::vector<int32_t> cs;
    if (support_tree(main_newick, ts, false, &cs, err).empty()) return fail(nullptr, PML_EINVAL, err);
    if (nsplits) *nsplits = (int)cs.size();
    if (counts) std::copy(cs.begin(), cs.end(), counts);
    return PML_OK;
}

}  // extern "C"
#pragma GCC visibility pop
