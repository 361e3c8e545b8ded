"""The checks every rank of a site-sharded group runs (tests/multirank_worker.py: one process per GPU under torchrun;
tests/test_gpu_group.py: all ranks as threads of one process).  Each rank makes the same calls; results must equal the
single-rank answers (oracle, raxmlHPC goldens, host-only integer code) and be bit-identical across ranks."""
import gzip
import json
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_case(case):
    from oracle import oracle as orc
    g = json.load(open(os.path.join(ROOT, "tests", "golden", case + ".json")))
    p = os.path.join(ROOT, "tests", "golden", case + ".phy")
    if os.path.exists(p):
        names, seqs = orc.read_phylip(p)
    else:
        toks = gzip.open(p + ".gz", "rt").read().split()
        n = int(toks[0])
        names, seqs = toks[2::2][:n], toks[3::2][:n]
    return g, names, seqs


def _strip(nw):
    return nw.replace("):0.0;", ");")


def _splits(newick):
    from oracle import oracle as orc
    sets = []
    taxa = sorted(orc._leafsets(orc._parse_topology(newick), sets))
    return {orc._canon(x, taxa) for x in sets if 1 < len(x) < len(taxa) - 1}


def likelihood_checks(ctx, case="wide"):
    """fixed-parameter lnL, derivatives, sharded pattern crunch, replicate lnL, `-f e`; returns floats to compare across ranks"""
    import pepr_b200 as pb
    from oracle import oracle as orc
    g, names, seqs = load_case(case)
    fe = g["fe"]
    aln = pb.Alignment(ctx, names, seqs, alpha=fe["alpha"])
    # the ranks shared the column sort: global pattern weights and the site -> pattern map must be the host-only ones
    codes, w_host, s2p_host = pb.crunch_patterns(seqs)
    w_eng, s2p_eng = aln.patterns()
    assert np.array_equal(w_eng, w_host) and np.array_equal(s2p_eng, s2p_host)
    tree = pb.Tree(aln, fe["tree"])
    lnl = tree.evaluate()
    assert abs(lnl - fe["lnl"]) <= 1e-6 * abs(fe["lnl"]), (lnl, fe["lnl"])
    pat, w, s2p = orc.compress(orc.encode(seqs))
    ot = orc.Tree(fe["tree"], names)
    m = orc.Model()
    want = orc.evaluate(m, ot, pat, w, fe["alpha"])
    assert abs(lnl - want) <= 1e-10 * abs(want), (lnl, want)
    olen = {round(ot.get_bl(e), 15): e for e in range(ot.nedge)}   # match branches of the two parsers by length
    out = [lnl]
    for e in (0, 3, tree.num_branches - 1):
        a, b, l = tree.branch(e)
        got = tree.branch_derivs(e, 0.5 * l + 0.01)
        ow = orc.branch_derivs(m, ot, pat, w, fe["alpha"], olen[round(l, 15)], 0.5 * l + 0.01)
        assert abs(got[0] - ow[0]) <= 1e-10 * abs(ow[0]), (e, got, ow)
        assert abs(got[1] - ow[1]) <= 1e-8 * max(1.0, abs(ow[1])) and abs(got[2] - ow[2]) <= 1e-8 * max(1.0, abs(ow[2])), (e, got, ow)
        out += list(got)
    # replicate weights are bit-exact integers from the global pattern weights; their lnL in one pass = direct evaluation
    W, _ = aln.bootstrap_weights(12345, 3)
    Wh, _ = pb.bootstrap_weights(w_host, 12345, 3)
    assert np.array_equal(W, Wh)
    rl = tree.evaluate_replicates(W)
    for r in range(3):
        direct = tree.evaluate(weights=W[r])
        assert abs(rl[r] - direct) <= 1e-10 * abs(direct), (r, rl[r], direct)
        out.append(float(rl[r]))
    # optimiser from default lengths: every rank must end on the same tree, within raxmlHPC's epsilon of its lnL
    t2 = pb.Tree(aln, re.sub(r":[0-9.eE+-]+", "", fe["tree"]))
    aln.set_model(1.0)
    l2, alpha = t2.optimize(True, 0.1)
    assert l2 >= fe["lnl"] - 0.1, (l2, fe["lnl"])
    out += [l2, alpha]
    t2.close(); tree.close(); aln.close()
    return out


def search_checks(ctx, case="search"):
    """integer parsimony scans, lazy-SPR scores against the oracle, a whole search against raxmlHPC's `-f d` tree, replicate trees"""
    import pepr_b200 as pb
    from oracle import oracle as orc
    g, names, seqs = load_case(case)
    fd = g["fd"]
    aln = pb.Alignment(ctx, names, seqs, alpha=fd["alpha"])
    pat, w, s2p = orc.compress(orc.encode(seqs))
    m = orc.Model()
    out = []
    # device Fitch scans summed over the ranks must pick the host-only tree
    for seed in (12345, 777):
        t = pb.Tree(aln, parsimony_seed=seed)
        want, _ = pb.parsimony_tree(names, seqs, seed)
        assert _splits(_strip(t.newick())) == _splits(_strip(want))
        t.close()
    tree = pb.Tree(aln, fd["tree"])
    base = tree.evaluate()
    checked = 0
    for node in (aln.ntax, aln.ntax + 5, 2 * aln.ntax - 3):
        keep = tree.neighbors(node)[1]
        targets, lnl = tree.score_spr_candidates(node, keep, radius=3)
        assert abs(tree.evaluate() - base) <= 1e-9 * abs(base)
        for tgt, l in list(zip(targets, lnl))[:2]:
            moved = pb.Tree(aln, fd["tree"])
            moved.spr(node, keep, int(tgt))
            want = orc.evaluate(m, orc.Tree(_strip(moved.newick()), names), pat, w, fd["alpha"])
            assert abs(l - want) <= 1e-9 * abs(want), (node, tgt, l, want)
            moved.close()
            out.append(float(l))
            checked += 1
    assert checked >= 4
    tree.close()
    t = pb.Tree(aln, parsimony_seed=12345)
    aln.set_model(1.0)
    t.optimize(True, 5.0)
    lnl, moves = t.search(radius=5, max_rounds=10, eps=0.1)
    lnl, alpha = t.optimize(True, 0.1)
    assert lnl >= fd["lnl"] - 0.5, (lnl, fd["lnl"])
    assert _splits(_strip(t.newick())) == _splits(fd["tree"])
    out += [lnl, alpha, float(moves)]
    t.close()
    trees, rl, _ = aln.bootstrap_trees(2, weight_seed=12345, parsimony_seed=12345)
    assert [i for i, _ in trees] == [0, 1]
    out += [float(x) for x in rl]
    out.append(float(sum(len(nw) for _, nw in trees)))
    aln.close()
    return out, [nw for _, nw in trees]
