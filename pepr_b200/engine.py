"""ctypes binding of include/peprml.h (the same binding a JNI / Panama stub would make)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


NKINDS = 28   # PML_NKINDS in include/peprml.h


def kinds():
    """kernel kinds of pml_profile_begin/end in index order: [(name, algorithmic bytes per pattern, DMMA per tile and warp)]"""
    out = []
    for k in range(NKINDS):
        name = C.create_string_buffer(64)
        b, d = C.c_int(0), C.c_int(0)
        lib().pml_kind_info(k, name, 64, C.byref(b), C.byref(d))
        out.append((name.value.decode(), b.value, d.value))
    return out


class EngineError(RuntimeError):
    pass


def lib():
    """loads pepr_b200/libpeprml.so; raises if it has not been built (python -m pepr_b200.build)"""
    global _LIB
    if _LIB is None:
        so = os.environ.get("PEPRML_LIB") or os.path.join(HERE, "libpeprml.so")  # PEPRML_LIB: an experiment build (build.build_variant)
        if not os.path.exists(so):
            raise EngineError("pepr_b200/libpeprml.so is missing: run `python -m pepr_b200.build` (nvcc, sm_100a); "
                              "there is no CPU fallback")
        L = C.CDLL(so)
        L.pml_version.restype = C.c_char_p
        L.pml_last_error.restype = C.c_char_p
        L.pml_last_error.argtypes = [C.c_void_p]
        L.pml_aln_name.restype = C.c_char_p
        L.pml_tree_newick.restype = C.c_int64
        L.pml_support_tree.restype = C.c_int64
        L.pml_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
        L.pml_ctx_destroy.argtypes = [C.c_void_p]
        L.pml_ctx_sync.argtypes = [C.c_void_p]
        L.pml_ctx_collective.argtypes = [C.c_void_p]
        L.pml_trace_enable.argtypes = [C.c_void_p, C.c_int]
        L.pml_trace_read.argtypes = [C.c_void_p, C.c_void_p]
        L.pml_aln_load.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_void_p)]
        L.pml_aln_load_phylip.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.pml_aln_free.argtypes = [C.c_void_p]
        L.pml_aln_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), c_i64p, c_i64p, c_i64p]
        L.pml_aln_patterns.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pml_aln_name.argtypes = [C.c_void_p, C.c_int]
        L.pml_model_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.pml_model_get.argtypes = [C.c_void_p, c_f64p, c_f64p]
        L.pml_wag_pmatrix.argtypes = [C.c_double, C.c_double, C.c_void_p]
        L.pml_wag_frequencies.argtypes = [C.c_void_p]
        L.pml_gamma_rates.argtypes = [C.c_double, C.c_int, C.c_void_p]
        L.pml_tree_load.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.pml_tree_free.argtypes = [C.c_void_p]
        L.pml_tree_num_branches.argtypes = [C.c_void_p]
        L.pml_tree_branch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), c_f64p]
        L.pml_tree_set_branch.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.pml_tree_newick.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.pml_tree_invalidate.argtypes = [C.c_void_p]
        L.pml_tree_stats.argtypes = [C.c_void_p, c_i64p, c_i64p]
        L.pml_tree_nr_retries.argtypes = [C.c_void_p]
        L.pml_tree_nr_retries.restype = C.c_int64
        L.pml_evaluate.argtypes = [C.c_void_p, C.c_void_p, c_f64p, C.c_void_p]
        L.pml_branch_derivs.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, c_f64p, c_f64p, c_f64p]
        L.pml_optimize.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, c_f64p, c_f64p]
        L.pml_smooth_branches.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.pml_bootstrap_weights.argtypes = [C.c_void_p, c_i64p, C.c_int, C.c_void_p]
        L.pml_evaluate_replicates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pml_support_tree.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.pml_support_counts.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.pml_comm_unique_id.argtypes = [C.c_char_p]
        L.pml_parsimony_tree.restype = C.c_int64
        L.pml_parsimony_tree.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_char_p), C.c_void_p, C.c_int64, C.c_char_p, C.c_size_t, c_i64p]
        L.pml_tree_spr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.pml_tree_neighbors.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pml_tree_start_parsimony.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]
        L.pml_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, c_f64p, C.POINTER(C.c_int)]
        L.pml_score_spr_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.pml_profile_begin.argtypes = [C.c_void_p]
        L.pml_timer_start.argtypes = [C.c_void_p]
        L.pml_timer_stop.argtypes = [C.c_void_p, c_f64p]
        L.pml_profile_end.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pml_aln_set_constraints.argtypes = [C.c_void_p, C.c_char_p]
        L.pml_aln_num_constraints.argtypes = [C.c_void_p]
        L.pml_tree_satisfies_constraints.argtypes = [C.c_void_p]
        L.pml_constraints_from_tree.restype = C.c_int64
        L.pml_constraints_from_tree.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.pml_parsimony_tree_constrained.restype = C.c_int64
        L.pml_parsimony_tree_constrained.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_char_p), C.c_void_p, C.c_int64, C.c_char_p, C.c_char_p,
                                                     C.c_size_t, c_i64p]
        L.pml_newick_satisfies_constraints.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p]
        L.pml_kind_info.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.pml_bootstrap_weights_host.argtypes = [C.c_void_p, C.c_int64, c_i64p, C.c_int, C.c_void_p]
        L.pml_crunch_patterns.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]
        L.pml_crunch_patterns_sharded.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64p]
        L.pml_group_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.pml_newick_capacity.restype = C.c_int64
        L.pml_newick_capacity.argtypes = [C.c_void_p]
        L.pml_bootstrap_trees.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                          C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _weights(w):
    return None if w is None else np.ascontiguousarray(w, np.int32)


def unique_id():
    buf = C.create_string_buffer(128)
    if lib().pml_comm_unique_id(buf) != 0:
        raise EngineError(lib().pml_last_error(None).decode())
    return buf.raw


class Context:
    """one GPU + one CUDA stream (+ one NCCL rank of a site-sharded group)"""

    def __init__(self, gpu=0, rank=0, nranks=1, uid=None, _handle=None):
        if _handle is not None:
            self.h, self.rank, self.nranks = _handle, rank, nranks
            return
        h = C.c_void_p()
        rc = lib().pml_ctx_create(gpu, rank, nranks, uid, C.byref(h))
        if rc != 0:
            raise EngineError("pml_ctx_create failed (%d): %s" % (rc, lib().pml_last_error(None).decode()))
        self.h, self.rank, self.nranks = h, rank, nranks

    def check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (%d): %s" % (what, rc, lib().pml_last_error(self.h).decode()))

    def sync(self):
        self.check(lib().pml_ctx_sync(self.h), "pml_ctx_sync")

    @property
    def collective(self):
        """how branch-pass sums cross the ranks: 'none' (one rank), 'nccl', or 'in-kernel nvlink' (peer mailboxes)"""
        return ("none", "nccl", "in-kernel nvlink")[lib().pml_ctx_collective(self.h)]

    def timer_start(self):
        self.check(lib().pml_timer_start(self.h), "pml_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self.check(lib().pml_timer_stop(self.h, C.byref(ms)), "pml_timer_stop")
        return ms.value

    def profile_begin(self):
        self.check(lib().pml_profile_begin(self.h), "pml_profile_begin")

    def profile_end(self):
        """-> dict kind -> (device ms, launches, pattern rows)"""
        ms = (C.c_double * NKINDS)()
        n = (C.c_int64 * NKINDS)()
        rows = (C.c_int64 * NKINDS)()
        self.check(lib().pml_profile_end(self.h, ms, n, rows), "pml_profile_end")
        return {k[0]: (ms[i], n[i], rows[i]) for i, k in enumerate(kinds())}

    def close(self):
        if self.h:
            lib().pml_ctx_destroy(self.h)
            self.h = None


class Group:
    """all ranks of a site-sharded group in THIS process (pml_group_create): one Context per GPU, each driven by its own
    host thread.  `run(fn)` calls fn(ctx) for every rank concurrently (ctypes releases the GIL inside the engine) and
    returns the results in rank order -- the shape PEPR's thread-per-runner code has (RAxMLRunner `-T n`)."""

    def __init__(self, gpus):
        gpus = list(gpus)
        ids = (C.c_int * len(gpus))(*gpus)
        hs = (C.c_void_p * len(gpus))()
        rc = lib().pml_group_create(ids, len(gpus), hs)
        if rc != 0:
            raise EngineError("pml_group_create failed (%d): %s" % (rc, lib().pml_last_error(None).decode()))
        self.contexts = [Context(rank=r, nranks=len(gpus), _handle=C.c_void_p(hs[r])) for r in range(len(gpus))]

    def __len__(self):
        return len(self.contexts)

    def run(self, fn):
        import threading
        out, errs = [None] * len(self.contexts), [None] * len(self.contexts)

        def work(r):
            try:
                out[r] = fn(self.contexts[r])
            except BaseException as ex:  # noqa: BLE001 -- re-raised below on the calling thread
                errs[r] = ex

        threads = [threading.Thread(target=work, args=(r,)) for r in range(len(self.contexts))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for ex in errs:
            if ex is not None:
                raise ex
        return out

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []


class Alignment:
    """resident, pattern-compressed alignment; `seqs` are equal-length strings (or a uint8 [ntax, nsites] array)"""

    def __init__(self, ctx, names=None, seqs=None, site_weights=None, phylip=None, weights_file=None, alpha=1.0,
                 model="PROTGAMMAWAG"):
        self.ctx = ctx
        h = C.c_void_p()
        if phylip is not None:
            ctx.check(lib().pml_aln_load_phylip(ctx.h, phylip.encode(), None if weights_file is None else weights_file.encode(),
                                                C.byref(h)), "pml_aln_load_phylip")
        else:
            if isinstance(seqs, np.ndarray):
                chars = np.ascontiguousarray(seqs, np.uint8)
            else:
                chars = np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
            arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
            sw = _weights(site_weights)
            ctx.check(lib().pml_aln_load(ctx.h, chars.shape[0], chars.shape[1], arr, _ptr(chars), _ptr(sw), C.byref(h)),
                      "pml_aln_load")
        self.h = h
        nt, ns, npat, nloc = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
        lib().pml_aln_dims(h, C.byref(nt), C.byref(ns), C.byref(npat), C.byref(nloc))
        self.ntax, self.nsites, self.npatterns, self.npatterns_local = nt.value, ns.value, npat.value, nloc.value
        self.names = [lib().pml_aln_name(h, i).decode() for i in range(self.ntax)]
        self.set_model(alpha, model)

    def set_model(self, alpha, model="PROTGAMMAWAG"):
        self.ctx.check(lib().pml_model_set(self.h, model.encode(), alpha), "pml_model_set")

    @property
    def alpha(self):
        a = C.c_double()
        lib().pml_model_get(self.h, C.byref(a), None)
        return a.value

    def patterns(self):
        w = np.zeros(self.npatterns, np.int32)
        s2p = np.zeros(self.nsites, np.int64)
        lib().pml_aln_patterns(self.h, _ptr(w), _ptr(s2p))
        return w, s2p

    def bootstrap_weights(self, seed, nrep):
        out = np.zeros((nrep, self.npatterns), np.int32)
        s = C.c_int64(seed)
        self.ctx.check(lib().pml_bootstrap_weights(self.h, C.byref(s), nrep, _ptr(out)), "pml_bootstrap_weights")
        return out, s.value

    def bootstrap_trees(self, nrep, weight_seed=12345, parsimony_seed=12345, first=0, stride=1, radius=5, rounds=1, eps=0.1):
        """replicate trees first, first + stride, ... of `nrep` on this context (see pml_bootstrap_trees);
        -> (list of (replicate index, newick)), lnL per replicate (nan elsewhere), wall seconds per replicate (nan elsewhere)"""
        ids = list(range(first, nrep, stride))
        cap = max(1, len(ids)) * int(lib().pml_newick_capacity(self.h)) + 16
        buf = C.create_string_buffer(cap)
        lnl = np.full(max(nrep, 1), np.nan)
        secs = np.full(max(nrep, 1), np.nan)
        self.ctx.check(lib().pml_bootstrap_trees(self.h, C.c_int64(weight_seed), C.c_int64(parsimony_seed), nrep, first, stride,
                                                 radius, rounds, eps, buf, cap, _ptr(lnl), _ptr(secs)), "pml_bootstrap_trees")
        trees = [t for t in buf.value.decode().split("\n") if t]
        return list(zip(ids, trees)), lnl[:nrep], secs[:nrep]

    def set_constraints(self, text):
        """FastTree constraint alignment (None clears): every tree grown or searched on this alignment must display its splits"""
        self.ctx.check(lib().pml_aln_set_constraints(self.h, text.encode() if text else None), "pml_aln_set_constraints")
        return lib().pml_aln_num_constraints(self.h)

    def close(self):
        if self.h:
            lib().pml_aln_free(self.h)
            self.h = None


class Tree:
    def __init__(self, aln, newick=None, parsimony_seed=None, weights=None):
        """newick: a given topology; parsimony_seed: randomised stepwise-addition parsimony start tree"""
        self.aln, self.ctx = aln, aln.ctx
        h = C.c_void_p()
        if newick is not None:
            self.ctx.check(lib().pml_tree_load(aln.h, newick.encode(), C.byref(h)), "pml_tree_load")
        else:
            w = _weights(weights)
            self.ctx.check(lib().pml_tree_start_parsimony(aln.h, C.c_int64(12345 if parsimony_seed is None else parsimony_seed),
                                                           _ptr(w), C.byref(h)), "pml_tree_start_parsimony")
        self.h = h

    @property
    def num_branches(self):
        return lib().pml_tree_num_branches(self.h)

    def branch(self, e):
        a, b, l = C.c_int(), C.c_int(), C.c_double()
        self.ctx.check(lib().pml_tree_branch(self.h, e, C.byref(a), C.byref(b), C.byref(l)), "pml_tree_branch")
        return a.value, b.value, l.value

    def set_branch(self, e, length):
        self.ctx.check(lib().pml_tree_set_branch(self.h, e, length), "pml_tree_set_branch")

    def invalidate(self):
        lib().pml_tree_invalidate(self.h)

    def stats(self):
        su = (C.c_int64 * 3)()
        ln = C.c_int64()
        lib().pml_tree_stats(self.h, su, C.byref(ln))
        return list(su), ln.value

    def evaluate(self, weights=None, per_site=False):
        w = _weights(weights)
        lnl = C.c_double()
        ps = np.zeros(self.aln.nsites) if per_site else None
        self.ctx.check(lib().pml_evaluate(self.h, _ptr(w), C.byref(lnl), _ptr(ps)), "pml_evaluate")
        return (lnl.value, ps) if per_site else lnl.value

    @property
    def nr_retries(self):
        return int(lib().pml_tree_nr_retries(self.h))

    def branch_derivs(self, branch, t, weights=None):
        w = _weights(weights)
        l, d1, d2 = C.c_double(), C.c_double(), C.c_double()
        self.ctx.check(lib().pml_branch_derivs(self.h, branch, t, _ptr(w), C.byref(l), C.byref(d1), C.byref(d2)),
                       "pml_branch_derivs")
        return l.value, d1.value, d2.value

    def smooth(self, sweeps=1, weights=None):
        w = _weights(weights)
        conv = C.c_int()
        self.ctx.check(lib().pml_smooth_branches(self.h, sweeps, _ptr(w), C.byref(conv)), "pml_smooth_branches")
        return bool(conv.value)

    def optimize(self, opt_alpha=True, eps=0.1, weights=None):
        w = _weights(weights)
        lnl, alpha = C.c_double(), C.c_double()
        self.ctx.check(lib().pml_optimize(self.h, int(opt_alpha), eps, _ptr(w), C.byref(lnl), C.byref(alpha)), "pml_optimize")
        return lnl.value, alpha.value

    def search(self, radius=5, max_rounds=10, eps=0.1, weights=None):
        """lazy-SPR hill climbing on this tree (topology changes in place); returns (lnL, accepted moves)"""
        w = _weights(weights)
        lnl, moves = C.c_double(), C.c_int()
        self.ctx.check(lib().pml_search(self.h, radius, max_rounds, eps, _ptr(w), C.byref(lnl), C.byref(moves)), "pml_search")
        return lnl.value, moves.value

    def score_spr_candidates(self, node, keep, radius=3, weights=None, cap=4096):
        w = _weights(weights)
        targets = np.zeros(cap, np.int32)
        lnl = np.zeros(cap)
        n = C.c_int(cap)
        self.ctx.check(lib().pml_score_spr_candidates(self.h, node, keep, radius, _ptr(w), _ptr(targets), _ptr(lnl), C.byref(n)),
                       "pml_score_spr_candidates")
        return targets[: n.value].copy(), lnl[: n.value].copy()

    def spr(self, node, keep, target):
        self.ctx.check(lib().pml_tree_spr(self.h, node, keep, target), "pml_tree_spr")

    def neighbors(self, node):
        nb = (C.c_int * 3)()
        self.ctx.check(lib().pml_tree_neighbors(self.h, node, nb), "pml_tree_neighbors")
        return list(nb)

    def evaluate_replicates(self, W):
        W = np.ascontiguousarray(W, np.int32)
        out = np.zeros(W.shape[0])
        self.ctx.check(lib().pml_evaluate_replicates(self.h, _ptr(W), W.shape[0], _ptr(out)), "pml_evaluate_replicates")
        return out

    def satisfies_constraints(self):
        return bool(lib().pml_tree_satisfies_constraints(self.h))

    def newick(self):
        n = lib().pml_tree_newick(self.h, None, 0)
        buf = C.create_string_buffer(n)
        lib().pml_tree_newick(self.h, buf, n)
        return buf.value.decode()

    def close(self):
        if self.h:
            lib().pml_tree_free(self.h)
            self.h = None


# ---- host-only entry points (no GPU needed) ----------------------------------------------------------------
def constraints_from_tree(newick):
    """FastTreeRunner.getFastTreeConstraintsForTree (FastTreeRunner.java:243-273): the constraint alignment of a tree"""
    n = lib().pml_constraints_from_tree(newick.encode(), None, 0)
    if n < 0:
        raise EngineError("pml_constraints_from_tree failed: %s" % lib().pml_last_error(None).decode())
    buf = C.create_string_buffer(int(n))
    lib().pml_constraints_from_tree(newick.encode(), buf, n)
    return buf.value.decode()


def wag_pmatrix(t, rate=1.0):
    P = np.zeros((20, 20))
    lib().pml_wag_pmatrix(t, rate, _ptr(P))
    return P


def wag_frequencies():
    pi = np.zeros(20)
    lib().pml_wag_frequencies(_ptr(pi))
    return pi


def gamma_rates(alpha, ncat=4):
    r = np.zeros(ncat)
    if lib().pml_gamma_rates(alpha, ncat, _ptr(r)) != 0:
        raise EngineError("pml_gamma_rates: bad arguments")
    return r


def bootstrap_weights(pattern_weights, seed, nrep):
    """host-only: replicate weight vectors from explicit pattern weights; returns (int32 [nrep, npat], next seed)"""
    pw = np.ascontiguousarray(pattern_weights, np.int32)
    out = np.zeros((nrep, len(pw)), np.int32)
    s = C.c_int64(seed)
    if lib().pml_bootstrap_weights_host(_ptr(pw), C.c_int64(len(pw)), C.byref(s), nrep, _ptr(out)) != 0:
        raise EngineError("pml_bootstrap_weights_host: bad arguments")
    return out, s.value


def crunch_patterns(seqs, site_weights=None, nranks=1):
    """host-only pattern crunch: (codes uint8 [ntax, npat], weights int32 [npat], site_to_pattern int64 [nsites]);
    nranks > 1: the way the ranks of a multi-process group share the sort (same result, bit for bit)"""
    chars = seqs if isinstance(seqs, np.ndarray) else np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
    chars = np.ascontiguousarray(chars, np.uint8)
    ntax, nsites = chars.shape
    codes = np.zeros(ntax * nsites, np.uint8)
    w = np.zeros(nsites, np.int32)
    s2p = np.zeros(nsites, np.int64)
    npat = C.c_int64()
    sw = _weights(site_weights)
    if nranks > 1:
        rc = lib().pml_crunch_patterns_sharded(nranks, ntax, C.c_int64(nsites), _ptr(chars), _ptr(sw), _ptr(codes), _ptr(w), _ptr(s2p), C.byref(npat))
    else:
        rc = lib().pml_crunch_patterns(ntax, C.c_int64(nsites), _ptr(chars), _ptr(sw), _ptr(codes), _ptr(w), _ptr(s2p), C.byref(npat))
    if rc != 0:
        raise EngineError("pml_crunch_patterns: " + (lib().pml_last_error(None).decode() or "bad arguments"))
    n = npat.value
    return codes[: ntax * n].reshape(ntax, n).copy(), w[:n].copy(), s2p


def pattern_range(npatterns, rank, nranks):
    """this rank's contiguous block of patterns: [p0, p1) -- the same split pml_aln_load applies"""
    return npatterns * rank // nranks, npatterns * (rank + 1) // nranks


def parsimony_tree(names, seqs, seed=12345):
    """host-only: randomised stepwise-addition parsimony tree (newick) and its score"""
    chars = seqs if isinstance(seqs, np.ndarray) else np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
    chars = np.ascontiguousarray(chars, np.uint8)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    score = C.c_int64()
    n = lib().pml_parsimony_tree(chars.shape[0], C.c_int64(chars.shape[1]), arr, _ptr(chars), C.c_int64(seed), None, 0, C.byref(score))
    if n < 0:
        raise EngineError("pml_parsimony_tree: " + lib().pml_last_error(None).decode())
    buf = C.create_string_buffer(n)
    lib().pml_parsimony_tree(chars.shape[0], C.c_int64(chars.shape[1]), arr, _ptr(chars), C.c_int64(seed), buf, n, C.byref(score))
    return buf.value.decode(), score.value


def parsimony_tree_constrained(names, seqs, constraints, seed=12345):
    """host-only: the stepwise-addition parsimony tree grown through constraint-preserving insertions only"""
    chars = seqs if isinstance(seqs, np.ndarray) else np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
    chars = np.ascontiguousarray(chars, np.uint8)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    score = C.c_int64()
    args = (chars.shape[0], C.c_int64(chars.shape[1]), arr, _ptr(chars), C.c_int64(seed), constraints.encode() if constraints else None)
    n = lib().pml_parsimony_tree_constrained(*args, None, 0, C.byref(score))
    if n < 0:
        raise EngineError("pml_parsimony_tree_constrained: " + lib().pml_last_error(None).decode())
    buf = C.create_string_buffer(n)
    lib().pml_parsimony_tree_constrained(*args, buf, n, C.byref(score))
    return buf.value.decode(), score.value


def newick_satisfies_constraints(newick, names, constraints):
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    rc = lib().pml_newick_satisfies_constraints(newick.encode(), arr, len(names), constraints.encode())
    if rc < 0:
        raise EngineError("pml_newick_satisfies_constraints: " + lib().pml_last_error(None).decode())
    return bool(rc)


def support_tree(main_newick, trees, as_percent=False):
    arr = (C.c_char_p * len(trees))(*[t.encode() for t in trees])
    n = lib().pml_support_tree(main_newick.encode(), arr, len(trees), int(as_percent), None, 0)
    if n < 0:
        raise EngineError("pml_support_tree: " + lib().pml_last_error(None).decode())
    buf = C.create_string_buffer(n)
    lib().pml_support_tree(main_newick.encode(), arr, len(trees), int(as_percent), buf, n)
    return buf.value.decode()


def support_counts(main_newick, trees):
    arr = (C.c_char_p * len(trees))(*[t.encode() for t in trees])
    n = C.c_int()
    if lib().pml_support_counts(main_newick.encode(), arr, len(trees), None, C.byref(n)) != 0:
        raise EngineError("pml_support_counts: " + lib().pml_last_error(None).decode())
    out = np.zeros(n.value, np.int32)
    lib().pml_support_counts(main_newick.encode(), arr, len(trees), _ptr(out), C.byref(n))
    return out
