"""Host-side logic of the engine (no GPU): model construction, pattern crunch, bootstrap stream, support counting,
newick handling, and the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import pepr_b200 as pb
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "peprml.h")).read()
    names = set(re.findall(r"\b(pml_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    L = ctypes.CDLL(os.path.join(ROOT, "pepr_b200", "libpeprml.so"))
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_context_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.EngineError, match="no CPU path"):
        pb.Context(0)


def test_transition_matrices_match_oracle():
    m = orc.Model()
    for t, r in [(1e-6, 1.0), (0.05, 0.137), (0.37, 1.3), (2.5, 2.386), (34.0, 4.0)]:
        P = pb.wag_pmatrix(t, r)
        assert np.abs(P - m.pmatrix(t, r)).max() < 1e-13
        assert np.abs(P.sum(1) - 1).max() < 1e-12
    pi = pb.wag_frequencies()
    assert abs(pi.sum() - 1.0) < 1e-12
    P = pb.wag_pmatrix(0.3, 1.0)
    assert np.abs(pi[:, None] * P - (pi[:, None] * P).T).max() < 1e-14   # detailed balance


def test_gamma_rates_match_oracle_and_have_mean_one():
    for a in (0.02, 0.155636, 0.5, 1.0, 3.705783, 142.7, 1000.0):
        r = pb.gamma_rates(a)
        assert np.abs(r - orc.gamma_rates(a)).max() < 1e-11
        assert abs(r.mean() - 1.0) < 1e-12
        assert (np.diff(r) > 0).all()


@pytest.mark.parametrize("case", ["small", "dup", "wide"])
def test_pattern_crunch_matches_oracle_and_reference_count(golden, case):
    g = golden(case)
    codes, w, s2p = pb.crunch_patterns(g.seqs)
    assert codes.shape[1] == g.meta["fe"]["patterns"]
    assert (codes == g.pat).all() and (w == g.w).all() and (s2p == g.s2p).all()
    assert w.sum() == len(g.seqs[0])


def test_pattern_crunch_with_column_weights(golden):
    g = golden("small")
    sw = np.array(g.meta["fw"]["weights"], np.int32)
    codes, w, s2p = pb.crunch_patterns(g.seqs, sw)
    c2, w2, s2 = orc.compress(g.codes, sw)
    assert codes.shape[1] == g.meta["fw"]["patterns"]
    assert (codes == c2).all() and (w == w2).all() and (s2p == s2).all()
    assert (s2p[sw == 0] == -1).all()


def test_pattern_crunch_threaded_radix_path_matches_oracle():
    """>= 20,000 columns take the multi-threaded MSD-radix path (level-0 buckets sorted by different threads, pattern
    boundaries from the radix pass, chunked weight accumulation): same patterns, weights and site map as the oracle's
    comparison sort, on a duplicate-rich alignment with gaps, B/Z codes, zero-weight columns and patterns that straddle
    the 32 k-column chunks of the weight pass"""
    rng = np.random.default_rng(11)
    ntax, nbase = 9, 900
    letters = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV-?XBZ", np.uint8)
    base = letters[rng.integers(0, len(letters), size=(ntax, nbase))]
    cols = rng.integers(0, nbase, size=70_000)
    cols[5_000:45_000] = 7                                        # one pattern 40,000 columns long: crosses chunk boundaries
    chars = np.ascontiguousarray(base[:, cols])
    seqs = [bytes(r).decode() for r in chars]
    sw = rng.integers(0, 3, size=chars.shape[1]).astype(np.int32)
    for weights in (None, sw):
        codes, w, s2p = pb.crunch_patterns(seqs, weights)
        oc, ow, os_ = orc.compress(orc.encode(seqs), weights)
        assert codes.shape == oc.shape and (codes == oc).all() and (w == ow).all() and (s2p == os_).all()
        assert w.sum() == (chars.shape[1] if weights is None else int(sw.sum()))


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_pattern_crunch_shared_between_ranks_equals_single_rank(golden, nranks):
    """N > 1: every rank sorts only its first-residue buckets and the sorted column order is exchanged by a sum
    (pml_aln_load does it with NCCL); patterns, weights, site map and the ranks' code blocks put side by side must be
    bit-identical to the single-rank crunch -- on a golden alignment and on the threaded large-input path"""
    g = golden("wide")
    codes, w, s2p = pb.crunch_patterns(g.seqs, nranks=nranks)
    assert (codes == g.pat).all() and (w == g.w).all() and (s2p == g.s2p).all()
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV-?XBZ", np.uint8)
    base = letters[rng.integers(0, len(letters), size=(7, 700))]
    chars = np.ascontiguousarray(base[:, rng.integers(0, 700, size=50_000)])
    sw = rng.integers(0, 3, size=chars.shape[1]).astype(np.int32)
    for weights in (None, sw):
        one = pb.crunch_patterns(chars, weights)
        many = pb.crunch_patterns(chars, weights, nranks=nranks)
        assert all((a == b).all() for a, b in zip(one, many))


def test_bootstrap_weights_bit_exact_with_reference(golden):
    g = golden("small")
    fj = g.meta["fj"]
    w, seed = pb.bootstrap_weights(g.w, fj["seed"], 3)
    assert w.tolist() == fj["replicate_weights"]
    # the stream continues across calls exactly like consecutive replicates of one run
    w1, s1 = pb.bootstrap_weights(g.w, fj["seed"], 1)
    w2, s2 = pb.bootstrap_weights(g.w, s1, 2)
    assert (np.vstack([w1, w2]) == w).all() and s2 == seed
    assert (w.sum(1) == g.w.sum()).all()


def test_bootstrap_stream_matches_oracle_on_weighted_patterns(golden):
    g = golden("dup")
    a, sa = pb.bootstrap_weights(g.w, 977, 5)
    b, sb = orc.bootstrap_weights(977, g.w, 5)
    assert (a == b).all() and sa == sb


def test_support_tree_matches_oracle_counts_and_raxml_percent(golden):
    from tests.test_oracle_golden import _labels
    g = golden("small")
    main, sup = g.meta["fe"]["tree"], g.meta["fb"]["support_trees"]
    counts = orc.support_counts(main, sup)
    got = _labels(pb.support_tree(main, sup, as_percent=False))
    assert got == {k: v for k, v in counts.items()}
    pct = _labels(pb.support_tree(main, sup, as_percent=True))
    assert pct == _labels(g.meta["fb"]["bipartitions"])
    assert sorted(pb.support_counts(main, sup).tolist()) == sorted(counts.values())


def test_support_tree_top_label_and_rooted_inputs():
    main = "((A:0.1,B:0.2):0.05,(C:0.3,D:0.4):0.05,E:0.1);"
    sup = ["((A,B),(C,D),E);", "(((A,B),E),(C,D));", "((A,C),(B,D),E);"]
    s = pb.support_tree(main, sup)
    # every inner node is labelled with a raw count; rooted support trees add their dissolved root to the tally of the
    # empty split, exactly as TreeSupportDecorator does (3 trees + 1 rooted one = 4 at the top)
    assert s == "((A:0.1,B:0.2)2:0.05,(C:0.3,D:0.4)2:0.05,E:0.1)4;"
    assert pb.support_counts(main, sup).tolist() == [2, 2]


def test_support_tie_break_uses_lowest_taxon_index():
    # 4 taxa: {A,B}|{C,D} has equal sides; both spellings must hit the same canonical split
    main = "((A:1.0,B:1.0):1.0,C:1.0,D:1.0);"
    sup = ["((C,D),A,B);", "((A,B),C,D);", "((A,C),B,D);"]
    assert pb.support_counts(main, sup).tolist() == [2]


def test_parsimony_start_tree_is_a_valid_good_tree(golden):
    from oracle import oracle as orc
    g = golden("wide")
    nw, score = pb.parsimony_tree(g.names, g.seqs, 12345)
    assert score > 0 and all(n in nw for n in g.names)
    t = orc.Tree(nw.replace("):0.0;", ");"), g.names)         # parses as a binary tree over exactly the taxa
    assert t.nedge == 2 * len(g.names) - 3
    found = orc.support_counts(g.meta["fe"]["tree"], [nw.replace("):0.0;", ");")])
    assert sum(found.values()) >= 0.9 * len(found)                # stepwise addition recovers most true splits
    nw2, score2 = pb.parsimony_tree(g.names, g.seqs, 12345)
    assert nw2 == nw and score2 == score                           # seeded, reproducible
    nw3, _ = pb.parsimony_tree(g.names, g.seqs, 777)
    assert nw3 != nw                                               # the addition order comes from the seed


# ---- topological constraints (FastTreeRunner.java:53-83, 243-273), tree-builder dispatch helpers, jackknife masks --------------
def _splits(newick, names):
    """non-trivial splits of a newick tree as frozensets of the side without names[0]"""
    toks = re.findall(r"[(),]|[^(),:;]+(?::[0-9.eE+-]+)?", newick.strip().rstrip(";"))
    stack, out = [], set()
    for tk in toks:
        if tk == "(":
            stack.append([])
        elif tk == ")":
            grp = stack.pop()
            s = frozenset().union(*grp)
            if stack:
                stack[-1].append(s)
            if 1 < len(s) < len(names) - 1:
                out.add(s if names[0] not in s else frozenset(names) - s)
        elif tk != "," and not tk.startswith(":"):
            nm = tk.split(":")[0]
            if nm and stack and nm in names:
                stack[-1].append(frozenset([nm]))
    return out


def test_constraint_alignment_of_a_tree_matches_the_reference_encoder():
    """getFastTreeConstraintsForTree: taxa sorted, one 0/1 column per node of the tree, 1 = the taxon descends from it"""
    text = pb.constraints_from_tree("((B:1,A:1):1,(C:1,D:1)0.9:1,E:2);")
    rows = dict(zip(text.split()[0::2], text.split()[1::2]))
    assert list(rows) == [">A", ">B", ">C", ">D", ">E"]
    cols = {"".join(rows[k][j] for k in rows) for j in range(len(rows[">A"]))}
    # nodes: top (all), (B,A), B, A, (C,D), C, D, E
    assert cols == {"11111", "11000", "01000", "10000", "00110", "00100", "00010", "00001"}


def test_constrained_parsimony_tree_displays_every_split(golden):
    g = golden("search")
    names = g.names
    free, _ = pb.parsimony_tree(names, g.seqs, seed=5)
    # force two splits the unconstrained tree does not have: taxa 0+5+11 together, and inside it 0+11
    a, b, c = names[0], names[5], names[11]
    assert not any(s == frozenset(names) - {a, b, c} or s == frozenset([a, b, c]) for s in _splits(free, names))
    rows = {n: "00" for n in names}
    rows[a], rows[b], rows[c] = "11", "10", "11"
    text = "".join(">%s\n%s\n" % (n, rows[n]) for n in names)
    for seed in (5, 6, 7):
        tree, score = pb.parsimony_tree_constrained(names, g.seqs, text, seed=seed)
        assert pb.newick_satisfies_constraints(tree, names, text)
        sp = _splits(tree, names)
        assert (frozenset(names) - {a, b, c}) in sp and (frozenset(names) - {a, c}) in sp
    assert not pb.newick_satisfies_constraints(free, names, text)
    # no constraints: the constrained entry point gives the plain tree; partial constraints (unlisted taxa are free) hold too
    assert pb.parsimony_tree_constrained(names, g.seqs, None, seed=5)[0] == free
    part = "".join(">%s\n%s\n" % (n, rows[n][0]) for n in names[:14])
    tree, _ = pb.parsimony_tree_constrained(names, g.seqs, part, seed=9)
    assert pb.newick_satisfies_constraints(tree, names, part)
    with pytest.raises(pb.EngineError, match="not in the alignment"):
        pb.parsimony_tree_constrained(names, g.seqs, ">nobody\n01\n", seed=1)
    # contradictory splits (AB|CD and AC|BD) cannot both be displayed: the entry point says so instead of returning a tree
    bad = {n: "--" for n in names}
    bad[names[0]], bad[names[1]], bad[names[2]], bad[names[3]] = "11", "10", "01", "00"
    with pytest.raises(pb.EngineError, match="contradictory"):
        pb.parsimony_tree_constrained(names, g.seqs, "".join(">%s\n%s\n" % (n, bad[n]) for n in names), seed=1)


def test_java_random_and_seeded_random_sets():
    """RandomSetUtils.getRandomSet draws from `new Random()`; the mirror takes a seed and reproduces java.util.Random"""
    from pepr_b200 import runner
    r = runner.JavaRandom(42)
    assert r.nextIntAll() == -1170105035                      # new Random(42).nextInt()
    r = runner.JavaRandom(42)
    assert [r.nextInt(10) for _ in range(10)] == [0, 3, 8, 4, 0, 5, 5, 8, 9, 3]
    s1 = runner.getRandomSet(20, 0, 39, False, seed=12345)
    assert s1 == runner.getRandomSet(20, 0, 39, False, seed=12345) and len(set(s1)) == 20 and all(0 <= v <= 39 for v in s1)
    assert runner.getRandomSet(5, 3, 6, False, seed=1) is None         # range too small without reuse: null in the reference
    assert len(runner.getRandomSet(9, 3, 6, True, seed=1)) == 9


def test_gene_block_weights_and_memory_throttle():
    from pepr_b200 import runner
    w = runner.gene_block_weights([3, 2, 4], [0, 2])
    assert w.dtype == np.int32 and w.tolist() == [1, 1, 1, 0, 0, 1, 1, 1, 1]
    assert runner.gene_block_weights([2, 2], [1, 1]).tolist() == [0, 0, 2, 2]             # a gene drawn twice (bootstrap over genes)
    # C4: a 500 x 250k replicate tree needs ~81 GB -> two fit in 180 GB, not three
    assert runner.support_threads_for_memory(500, 249975, 180e9, 8) == 2
    assert runner.support_threads_for_memory(2000, 1000000, 180e9, 8) == 1              # never below one, as in the reference
    assert runner.support_threads_for_memory(100, 95933, 180e9, 4) == 4


def test_branch_lengths_are_printed_like_java_double_tostring():
    """TreeSupportDecorator output goes through AdvancedTree.getTreeString -> Double.toString: shortest repr that round-trips,
    plain notation in [1e-3, 1e7), `d.dddE±n` outside it, always a fractional digit"""
    t = "((A:0.1,B:0.0001):1.0E7,(C:123456789,D:0.001):9999999,(E:2,F:1e-5):12345.678,G:0.30000000000000004);"
    out = pb.support_tree(t, [t], as_percent=False)
    assert out == "((A:0.1,B:1.0E-4)1:1.0E7,(C:1.23456789E8,D:0.001)1:9999999.0,(E:2.0,F:1.0E-5)1:12345.678,G:0.30000000000000004)1;"


def test_device_exponential_accuracy_on_host_build():
    """pmatrix.cuh exp_neg (the exponential of the P(t) and exponential-table prologues), host build of the same arithmetic:
    relative error <= 1.4 x 2^-53 on [-708, 0] against 60-digit arithmetic, exactly 1 at 0, exactly 0 below -708."""
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    L = ctypes.CDLL(os.path.join(ROOT, "pepr_b200", "libpeprml.so"))
    L.pml_debug_exp_neg.restype = ctypes.c_double
    L.pml_debug_exp_neg.argtypes = [ctypes.c_double]
    rng = np.random.default_rng(7)
    xs = np.concatenate([-rng.random(4000) * 708.0, -rng.random(3000) * 3.0, -rng.random(1000) * 1e-4, -10.0 ** rng.uniform(-300, -5, 500),
                         [0.0, -0.34657359027997264, -0.3465735902799727, -707.999, 3e-16]])
    worst = 0.0
    for x in xs:
        got, ref = Decimal(L.pml_debug_exp_neg(float(x))), Decimal(float(x)).exp()
        worst = max(worst, float(abs(got - ref) / ref))
    assert worst <= 1.4 * 2.0 ** -53, worst
    assert L.pml_debug_exp_neg(0.0) == 1.0
    assert L.pml_debug_exp_neg(-708.5) == 0.0 and L.pml_debug_exp_neg(-1e4) == 0.0


def test_guarded_newton_step_on_t_makes_the_decisions_of_the_rule_on_z():
    """kernels.h nr_step_t (what the device runs) against nr_step_z (raxmlHPC's update stated on z = exp(-t)): same status, same
    new length up to rounding -- ordinary steps, bad curvature, the step cap 0.25 z + 0.75, exponent >= 100, zmin and zmax."""
    L = ctypes.CDLL(os.path.join(ROOT, "pepr_b200", "libpeprml.so"))
    L.pml_debug_nr_step.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]

    def both(t, d1, d2):
        out = []
        for in_t in (0, 1):
            tn = ctypes.c_double()
            out.append((L.pml_debug_nr_step(t, d1, d2, in_t, ctypes.byref(tn)), tn.value))
        return out

    rng = np.random.default_rng(11)
    cases = [(0.1, -50.0, -800.0), (0.1, 50.0, -800.0), (0.1, 5.0, 3.0), (1e-7, 1.0, -1.0), (30.0, -1e5, -1.0), (0.5, 1e6, -1.0),
             (0.5, -150.0, -1.0), (1e-9, 0.0, 1.0), (40.0, 1.0, -1.0), (0.2, 0.0, -1.0)]
    for _ in range(5000):
        cases.append((float(10.0 ** rng.uniform(-7, 1.5)), float(rng.normal() * 10.0 ** rng.uniform(-3, 5)),
                      float(rng.normal() * 10.0 ** rng.uniform(-3, 5))))
    statuses = set()
    for t, d1, d2 in cases:
        (sz, tz), (st, tt) = both(t, d1, d2)
        assert sz == st, (t, d1, d2, sz, st)
        assert abs(tz - tt) <= 1e-12 * max(1.0, abs(tz)) + 1e-16, (t, d1, d2, tz, tt)
        if st == 1:  # a finished step ends inside the NR range (a retry point 0.37 z + 0.63 may lie above zmax: the next pass clamps it)
            assert 1.0e-6 <= tt <= 34.538776394910684 + 1e-12
        statuses.add(st)
    assert statuses == {1, 2}
