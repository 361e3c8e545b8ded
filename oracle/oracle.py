"""ctypes front-end of the CPU checker (oracle/pml_oracle.c) plus small pure-Python restatements of the
integer paths (bipartition support counting).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under pepr_b200/ may import this module.

Parity status: PINNED against the reference binary's own outputs (tests/golden/, tests/golden/make_golden.py).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """compile libpml_oracle.so (and stage oracle/_ref when /root/reference is mounted)"""
    subprocess.run(["make", "-s", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libpml_oracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "pml_oracle.c")):
            build()
        L = C.CDLL(so)
        L.orc_model_new.restype = C.c_void_p
        L.orc_tree_parse.restype = C.c_void_p
        L.orc_evaluate.restype = C.c_double
        L.orc_optimize.restype = C.c_double
        L.orc_randum.restype = C.c_double
        L.orc_compress.restype = C.c_int64
        L.orc_tree_get_bl.restype = C.c_double
        _LIB = L
    return _LIB


def ref_binary(name="raxmlHPC"):
    """path of the staged reference executable (oracle/_ref), or None"""
    p = os.path.join(HERE, "_ref", name)
    return p if os.path.exists(p) else None


AA = "ARNDCQEGHILKMFPSTWYV"


def encode(seqs):
    """list of equal-length strings -> uint8 [ntax, nsites] codes 0..22"""
    lut = np.full(256, 22, np.uint8)
    for i, ch in enumerate(AA):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = i
    lut[ord("B")] = lut[ord("b")] = 20
    lut[ord("Z")] = lut[ord("z")] = 21
    return np.stack([lut[np.frombuffer(s.encode(), np.uint8)] for s in seqs])


def read_phylip(path):
    toks = open(path).read().split()
    ntax, nsites = int(toks[0]), int(toks[1])
    names, seqs = toks[2::2][:ntax], toks[3::2][:ntax]
    assert all(len(s) == nsites for s in seqs)
    return names, seqs


class Model:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_model_new())

    def arrays(self):
        pi, lam = np.zeros(20), np.zeros(20)
        V, Vi = np.zeros((20, 20)), np.zeros((20, 20))
        lib().orc_model_get(self.h, *[a.ctypes.data_as(C.c_void_p) for a in (pi, lam, V, Vi)])
        return pi, lam, V, Vi

    def pmatrix(self, t, rate=1.0):
        P = np.zeros((20, 20))
        lib().orc_pmatrix(self.h, C.c_double(t), C.c_double(rate), P.ctypes.data_as(C.c_void_p))
        return P


def gamma_rates(alpha, k=4):
    r = np.zeros(k)
    lib().orc_gamma_rates(C.c_double(alpha), C.c_int(k), r.ctypes.data_as(C.c_void_p))
    return r


def compress(codes, site_w=None):
    """pattern crunch: returns (pat_codes [ntax,npat] uint8, pat_w int32[npat], site2pat int64[nsites])"""
    codes = np.ascontiguousarray(codes, np.uint8)
    ntax, nsites = codes.shape
    pat = np.zeros((ntax, max(nsites, 1)), np.uint8).ravel()
    w = np.zeros(max(nsites, 1), np.int32)
    s2p = np.zeros(nsites, np.int64)
    sw = None if site_w is None else np.ascontiguousarray(site_w, np.int32)
    npat = lib().orc_compress(C.c_int(ntax), C.c_int64(nsites), codes.ctypes.data_as(C.c_void_p),
                              None if sw is None else sw.ctypes.data_as(C.c_void_p),
                              pat.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), s2p.ctypes.data_as(C.c_void_p))
    return pat[: ntax * npat].reshape(ntax, npat).copy(), w[:npat].copy(), s2p


class Tree:
    def __init__(self, newick, names, deflen=0.1):
        self.names = list(names)
        self._arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        h = lib().orc_tree_parse(newick.encode(), C.c_int(len(names)), self._arr, C.c_double(deflen))
        if not h:
            raise ValueError("oracle: cannot parse newick")
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            lib().orc_tree_free(self.h)
        except Exception:
            pass

    @property
    def nedge(self):
        return lib().orc_tree_nedge(self.h)

    def edge(self, e):
        a, b = C.c_int(), C.c_int()
        lib().orc_tree_edge(self.h, C.c_int(e), C.byref(a), C.byref(b))
        return a.value, b.value

    def get_bl(self, e):
        return lib().orc_tree_get_bl(self.h, C.c_int(e))

    def set_bl(self, e, v):
        lib().orc_tree_set_bl(self.h, C.c_int(e), C.c_double(v))

    def newick(self):
        buf = C.create_string_buffer(64 * len(self.names) * 4 + 1024)
        lib().orc_tree_newick(self.h, self._arr, buf)
        return buf.value.decode()


def evaluate(model, tree, pat_codes, weights, alpha, edge=0, per_pattern=False, scalers=False):
    pat_codes = np.ascontiguousarray(pat_codes, np.uint8)
    npat = pat_codes.shape[1]
    w = None if weights is None else np.ascontiguousarray(weights, np.int32)
    pp = np.zeros(npat) if per_pattern else None
    sc = np.zeros(npat, np.int32) if scalers else None
    v = lib().orc_evaluate(model.h, tree.h, C.c_int64(npat), pat_codes.ctypes.data_as(C.c_void_p),
                           None if w is None else w.ctypes.data_as(C.c_void_p), C.c_double(alpha), C.c_int(edge),
                           None if pp is None else pp.ctypes.data_as(C.c_void_p),
                           None if sc is None else sc.ctypes.data_as(C.c_void_p))
    out = [v]
    if per_pattern:
        out.append(pp)
    if scalers:
        out.append(sc)
    return out[0] if len(out) == 1 else tuple(out)


def clv(model, tree, pat_codes, alpha, node, frm):
    pat_codes = np.ascontiguousarray(pat_codes, np.uint8)
    npat = pat_codes.shape[1]
    x = np.zeros((npat, 80))
    sc = np.zeros(npat, np.int32)
    lib().orc_clv(model.h, tree.h, C.c_int64(npat), pat_codes.ctypes.data_as(C.c_void_p), C.c_double(alpha),
                  C.c_int(node), C.c_int(frm), x.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p))
    return x, sc


def branch_derivs(model, tree, pat_codes, weights, alpha, edge, t):
    pat_codes = np.ascontiguousarray(pat_codes, np.uint8)
    w = None if weights is None else np.ascontiguousarray(weights, np.int32)
    l, d1, d2 = C.c_double(), C.c_double(), C.c_double()
    lib().orc_branch_derivs(model.h, tree.h, C.c_int64(pat_codes.shape[1]), pat_codes.ctypes.data_as(C.c_void_p),
                            None if w is None else w.ctypes.data_as(C.c_void_p), C.c_double(alpha), C.c_int(edge),
                            C.c_double(t), C.byref(l), C.byref(d1), C.byref(d2))
    return l.value, d1.value, d2.value


def optimize(model, tree, pat_codes, weights, alpha, opt_alpha=True, eps=0.1):
    pat_codes = np.ascontiguousarray(pat_codes, np.uint8)
    w = None if weights is None else np.ascontiguousarray(weights, np.int32)
    a = C.c_double(alpha)
    lnl = lib().orc_optimize(model.h, tree.h, C.c_int64(pat_codes.shape[1]), pat_codes.ctypes.data_as(C.c_void_p),
                             None if w is None else w.ctypes.data_as(C.c_void_p), C.byref(a), C.c_int(int(opt_alpha)),
                             C.c_double(eps))
    return lnl, a.value


def randum(seed):
    s = C.c_int64(seed)
    r = lib().orc_randum(C.byref(s))
    return s.value, r


def bootstrap_weights(seed, pat_w, nrep):
    """returns (int32 [nrep, npat], seed after the last replicate); bit exact restatement of computeNextReplicate"""
    pat_w = np.ascontiguousarray(pat_w, np.int32)
    out = np.zeros((nrep, len(pat_w)), np.int32)
    s = C.c_int64(seed)
    lib().orc_bootstrap_weights(C.byref(s), C.c_int64(len(pat_w)), pat_w.ctypes.data_as(C.c_void_p), C.c_int(nrep),
                                out.ctypes.data_as(C.c_void_p))
    return out, s.value


# ------------------------------------------------------------------ bipartition support (integer path) ------
def _parse_topology(newick):
    """-> nested tuples of leaf names; labels/lengths dropped"""
    s = newick.strip().rstrip(";")
    pos = [0]

    def node():
        kids = []
        if s[pos[0]] == "(":
            pos[0] += 1
            while True:
                kids.append(node())
                if s[pos[0]] == ",":
                    pos[0] += 1
                    continue
                if s[pos[0]] == ")":
                    pos[0] += 1
                    break
                raise ValueError("bad newick at %d" % pos[0])
        m = re.match(r"([^:,()\[\]]*)(:[-+0-9.eE]+)?(\[[^\]]*\])?", s[pos[0]:])
        pos[0] += m.end()
        return tuple(kids) if kids else m.group(1).strip()

    return node()


def _leafsets(t, out):
    if isinstance(t, str):
        return frozenset([t])
    acc = frozenset()
    for k in t:
        acc |= _leafsets(k, out)
    out.append(acc)
    return acc


def _canon(side, taxa_sorted):
    """Bipartition canonical form of PEPR (Bipartition.java:41-64): the smaller side; on a tie the side that holds
    the lowest-index taxon (taxa sorted by name)."""
    n = len(taxa_sorted)
    other = frozenset(taxa_sorted) - side
    if len(side) < len(other):
        return side
    if len(other) < len(side):
        return other
    return side if taxa_sorted[0] in side else other


def support_counts(main_newick, support_newicks):
    """TreeSupportDecorator.addSupportValues semantics (TreeSupportDecorator.java:86-163): for every non-trivial
    split of the main tree, the number of support trees containing the same split.  Returns {frozenset(smaller side): count}."""
    main = _parse_topology(main_newick)
    sets = []
    taxa = sorted(_leafsets(main, sets))
    n = len(taxa)
    counts = {}
    sup_splits = []
    for nw in support_newicks:
        ss = []
        _leafsets(_parse_topology(nw), ss)
        sup_splits.append({_canon(x, taxa) for x in ss if 1 < len(x) < n - 1})
    for x in sets:
        if 1 < len(x) < n - 1:
            c = _canon(x, taxa)
            counts[c] = sum(1 for sp in sup_splits if c in sp)
    return counts
