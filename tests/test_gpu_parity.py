"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the reference binary's golden outputs.
Tolerances: FP64 mode, total lnL relative error <= 1e-6 against raxmlHPC (BASELINE.json north_star); against the oracle
(same arithmetic, different summation order) we hold 1e-10 relative; integer paths are bit exact."""
import numpy as np
import pytest

import pepr_b200 as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
REL_REF = 1e-6      # vs raxmlHPC (north_star tolerance)
REL_ORACLE = 1e-10  # vs the oracle restatement


def _load(gpu_ctx, g, alpha, newick, site_weights=None):
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, site_weights=site_weights, alpha=alpha)
    return aln, pb.Tree(aln, newick)


@pytest.mark.parametrize("case", ["small", "dup", "deep", "wide", "aquificales", "erysipelotrichales"])
def test_lnl_matches_reference_and_oracle(gpu_ctx, golden, case):
    g = golden(case)
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, fe["alpha"], fe["tree"])
    assert aln.npatterns == fe["patterns"]
    lnl = tree.evaluate()
    ref = orc.evaluate(orc.Model(), orc.Tree(fe["tree"], g.names), g.pat, g.w, fe["alpha"])
    assert abs(lnl - fe["lnl"]) / abs(fe["lnl"]) < REL_REF
    assert abs(lnl - ref) / abs(ref) < REL_ORACLE
    tree.close(); aln.close()


def test_per_site_lnl_in_original_column_order(gpu_ctx, golden):
    g = golden("small")
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, fe["alpha"], fe["tree"])
    lnl, ps = tree.evaluate(per_site=True)
    assert np.abs(ps - np.array(g.meta["fg"]["per_site"])).max() < 2e-6      # raxmlHPC prints 6 decimals
    _, pp = orc.evaluate(orc.Model(), orc.Tree(fe["tree"], g.names), g.pat, g.w, fe["alpha"], per_pattern=True)
    assert np.abs(ps - pp[g.s2p]).max() < 1e-10
    assert abs(ps.sum() - lnl) < 1e-8
    tree.close(); aln.close()


def test_column_weights(gpu_ctx, golden):
    g = golden("small")
    fw = g.meta["fw"]
    aln, tree = _load(gpu_ctx, g, fw["alpha"], fw["tree"], site_weights=np.array(fw["weights"], np.int32))
    assert aln.npatterns == fw["patterns"]
    assert abs(tree.evaluate() - fw["lnl"]) / abs(fw["lnl"]) < REL_REF
    tree.close(); aln.close()


def test_replicate_weight_vectors(gpu_ctx, golden):
    g = golden("small")
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, fe["alpha"], fe["tree"])
    W, seed = aln.bootstrap_weights(g.meta["fj"]["seed"], 3)
    assert W.tolist() == g.meta["fj"]["replicate_weights"]                  # bit exact with raxmlHPC -f j
    _, pp = orc.evaluate(orc.Model(), orc.Tree(fe["tree"], g.names), g.pat, g.w, fe["alpha"], per_pattern=True)
    want = W.astype(np.float64) @ pp
    got = tree.evaluate_replicates(W)
    assert np.abs(got - want).max() / np.abs(want).max() < REL_ORACLE
    for r in range(3):                                                      # the one-vector entry point agrees
        assert abs(tree.evaluate(weights=W[r]) - want[r]) / abs(want[r]) < REL_ORACLE
    tree.close(); aln.close()


def test_lnl_independent_of_traversal_state(gpu_ctx, golden):
    """partial traversals after branch edits must give what a full traversal gives"""
    g = golden("wide")
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, fe["alpha"], fe["tree"])
    base = tree.evaluate()
    rng = np.random.default_rng(0)
    for _ in range(5):
        e = int(rng.integers(0, tree.num_branches))
        _, _, l = tree.branch(e)
        tree.set_branch(e, l * 1.7 + 0.01)
        tree.branch_derivs(int(rng.integers(0, tree.num_branches)), 0.1)   # re-orients CLVs somewhere else
        partial = tree.evaluate()
        tree.invalidate()
        full = tree.evaluate()
        assert abs(partial - full) <= 1e-9 * abs(full)
        tree.set_branch(e, l)
    assert abs(tree.evaluate() - base) <= 1e-9 * abs(base)
    tree.close(); aln.close()


@pytest.mark.parametrize("case,edges", [("small", [0, 3, 7, 12]), ("deep", [0, 10, 100, 316])])
def test_branch_derivatives_match_oracle(gpu_ctx, golden, case, edges):
    g = golden(case)
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, fe["alpha"], fe["tree"])
    ot = orc.Tree(fe["tree"], g.names)
    m = orc.Model()
    # branch numbering differs between the two parsers: match branches by their end nodes' leaf sets via lengths
    olen = {round(ot.get_bl(e), 15): e for e in range(ot.nedge)}
    for e in edges:
        a, b, l = tree.branch(e)
        oe = olen[round(l, 15)]
        for t in (l, 0.5 * l + 0.01, 2.0 * l + 0.05):
            got = tree.branch_derivs(e, t)
            want = orc.branch_derivs(m, ot, g.pat, g.w, fe["alpha"], oe, t)
            assert abs(got[0] - want[0]) <= 1e-10 * abs(want[0])
            assert abs(got[1] - want[1]) <= 1e-8 * max(1.0, abs(want[1]))
            assert abs(got[2] - want[2]) <= 1e-8 * max(1.0, abs(want[2]))
    tree.close(); aln.close()


@pytest.mark.parametrize("case", ["small", "dup", "deep", "aquificales", "erysipelotrichales"])
def test_optimize_reaches_reference_optimum(gpu_ctx, golden, case):
    """`-f e` from the unoptimised input tree: raxmlHPC stops at dlnL <= 0.1, so that is the comparison tolerance."""
    g = golden(case)
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, 1.0, g.meta["tree_in"])
    lnl, alpha = tree.optimize(opt_alpha=True, eps=0.1)
    assert lnl > fe["lnl"] - 0.1, (lnl, fe["lnl"])
    assert abs(lnl - fe["lnl"]) < 0.5
    assert abs(alpha - fe["alpha"]) / fe["alpha"] < 0.05
    # the optimised tree scores the same on the oracle (engine result is self-consistent)
    ot = orc.Tree(tree.newick().replace("):0.0;", ");"), g.names)
    chk = orc.evaluate(orc.Model(), ot, g.pat, g.w, alpha)
    assert abs(chk - lnl) / abs(lnl) < 1e-9
    tree.close(); aln.close()


def test_optimize_wide_matches_reference(gpu_ctx, golden):
    g = golden("wide")
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, 1.0, g.meta["tree_in"])
    lnl, alpha = tree.optimize(opt_alpha=True, eps=0.1)
    assert lnl > fe["lnl"] - 0.1 and abs(lnl - fe["lnl"]) < 0.5, (lnl, fe["lnl"])
    assert abs(alpha - fe["alpha"]) / fe["alpha"] < 0.02
    tree.close(); aln.close()


def test_errors_are_reported_not_swallowed(gpu_ctx, golden):
    g = golden("small")
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs)
    with pytest.raises(pb.EngineError, match="not in the alignment"):
        pb.Tree(aln, "((TaxA,Nope),(TaxC,TaxD),((TaxE,TaxF),(TaxG,TaxH)));")
    with pytest.raises(pb.EngineError, match="missing from the tree|binary tree"):
        pb.Tree(aln, "((TaxA,TaxB),(TaxC,TaxD),(TaxE,TaxF));")
    with pytest.raises(pb.EngineError, match="PROTGAMMAWAG"):
        aln.set_model(1.0, "GTRGAMMA")
    aln.close()


def test_ragged_pattern_counts(gpu_ctx, golden):
    """pattern counts that are not multiples of the tile height, down to a single column"""
    g = golden("small")
    fe = g.meta["fe"]
    m = orc.Model()
    for ncol in (1, 31, 33, 129, 257):
        seqs = [s[:ncol] for s in g.seqs]
        pat, w, _ = orc.compress(orc.encode(seqs))
        aln = pb.Alignment(gpu_ctx, g.names, seqs, alpha=fe["alpha"])
        tree = pb.Tree(aln, fe["tree"])
        want = orc.evaluate(m, orc.Tree(fe["tree"], g.names), pat, w, fe["alpha"])
        assert abs(tree.evaluate() - want) <= 1e-10 * abs(want)
        tree.close(); aln.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["deep", "dup"])
def test_pipelined_smoothing_equals_host_reference_mode(golden, case):
    """The device-resident sweep (NR step in the kernel tail, host one branch ahead, poison-flag rollback on raxmlHPC's
    bad-curvature retry) against the same sweep with the step on the host and a wait per branch (PEPRML_HOST_NR=1).
    Bad starting lengths provoke the retry path."""
    import os
    g = golden(case)
    results, retries = {}, {}
    for mode in ("device", "host"):
        if mode == "host":
            os.environ["PEPRML_HOST_NR"] = "1"
        else:
            os.environ.pop("PEPRML_HOST_NR", None)
        ctx = pb.Context(0)
        os.environ.pop("PEPRML_HOST_NR", None)
        aln = pb.Alignment(ctx, g.names, g.seqs, alpha=0.7)
        out, nret = [], 0
        for start in (0.1, 2.5, 9.0, 30.0):
            tree = pb.Tree(aln, g.meta["tree_in"])
            for e in range(tree.num_branches):
                tree.set_branch(e, start if e % 3 else 0.001)
            tree.smooth(3)
            out.append([tree.branch(e)[2] for e in range(tree.num_branches)] + [tree.evaluate()])
            nret += tree.nr_retries
            tree.close()
        results[mode], retries[mode] = np.array(out), nret
        aln.close(); ctx.close()
    assert retries["device"] == retries["host"] and retries["device"] > 0, retries
    assert np.allclose(results["device"], results["host"], rtol=1e-8, atol=1e-12), np.abs(results["device"] - results["host"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["small", "deep", "wide"])
def test_fused_update_and_branch_pass_equals_two_launches(golden, case):
    """fused_mma.cu (last CLV update of a branch visit + the pass over that branch in one launch) against the same work as
    two launches (PEPRML_NO_FUSE=1): per-site lnL, smoothing sweeps, `-f e`; both against the reference lnL."""
    import os
    g = golden(case)
    fe = g.meta["fe"]
    res = {}
    for mode in ("fused", "two"):
        if mode == "two":
            os.environ["PEPRML_NO_FUSE"] = "1"
        else:
            os.environ.pop("PEPRML_NO_FUSE", None)
        ctx = pb.Context(0)
        os.environ.pop("PEPRML_NO_FUSE", None)
        aln = pb.Alignment(ctx, g.names, g.seqs, alpha=fe["alpha"])
        tree = pb.Tree(aln, fe["tree"])
        lnl, ps = tree.evaluate(per_site=True)
        assert abs(lnl - fe["lnl"]) <= REL_REF * abs(fe["lnl"])
        su0, ln0 = tree.stats()
        for e in range(tree.num_branches):
            tree.set_branch(e, 0.05 + 0.01 * (e % 7))
        tree.smooth(2)
        lens = [tree.branch(e)[2] for e in range(tree.num_branches)]
        l2 = tree.evaluate()
        su1, ln1 = tree.stats()
        tree.close()
        t2 = pb.Tree(aln, g.meta["tree_in"])
        l3, a3 = t2.optimize(True, 0.1)
        t2.close(); aln.close(); ctx.close()
        res[mode] = (lnl, ps, np.array(lens), l2, l3, a3, ln1 - ln0, sum(su1) - sum(su0))
    f, t = res["fused"], res["two"]
    assert abs(f[0] - t[0]) <= 1e-12 * abs(t[0]) and np.abs(f[1] - t[1]).max() <= 1e-9
    assert np.allclose(f[2], t[2], rtol=1e-8, atol=1e-12) and abs(f[3] - t[3]) <= 1e-10 * abs(t[3])
    assert abs(f[4] - t[4]) <= 1e-3 and abs(f[5] - t[5]) <= 1e-4 * t[5]
    assert f[7] == t[7] and f[6] < t[6]          # same site-updates, fewer launches


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["small", "dup", "deep", "wide"])
def test_folded_cherries_equal_stored_cherries(golden, case):
    """Cherry folding (an inner node with two tip children is never stored: its consumers form the product of the two tip
    look-ups themselves) against the engine with every CLV stored (PEPRML_NO_FOLD=1): per-site lnL, derivatives on every
    branch, smoothing sweeps, `-f e`, lazy SPR scores; same site-update count, fewer launches."""
    import os
    g = golden(case)
    fe = g.meta["fe"]
    res = {}
    for mode in ("fold", "stored"):
        if mode == "stored":
            os.environ["PEPRML_NO_FOLD"] = "1"
        else:
            os.environ.pop("PEPRML_NO_FOLD", None)
        ctx = pb.Context(0)
        os.environ.pop("PEPRML_NO_FOLD", None)
        aln = pb.Alignment(ctx, g.names, g.seqs, alpha=fe["alpha"])
        tree = pb.Tree(aln, fe["tree"])
        lnl, ps = tree.evaluate(per_site=True)
        assert abs(lnl - fe["lnl"]) <= REL_REF * abs(fe["lnl"])
        der = np.array([tree.branch_derivs(e, 0.07 + 0.01 * (e % 5)) for e in range(tree.num_branches)])
        su0, ln0 = tree.stats()
        for e in range(tree.num_branches):
            tree.set_branch(e, 0.05 + 0.01 * (e % 7))
        tree.smooth(2)
        lens = [tree.branch(e)[2] for e in range(tree.num_branches)]
        l2 = tree.evaluate()
        su1, ln1 = tree.stats()
        spr = []
        for node in range(len(g.names), min(len(g.names) + 4, 2 * len(g.names) - 2)):
            for keep in tree.neighbors(node):
                tg, sc = tree.score_spr_candidates(node, keep, 3)
                spr += sorted(zip(tg.tolist(), sc.tolist()))
        tree.close()
        t2 = pb.Tree(aln, g.meta["tree_in"])
        l3, a3 = t2.optimize(True, 0.1)
        t2.close(); aln.close(); ctx.close()
        res[mode] = (lnl, ps, np.array(lens), l2, l3, a3, ln1 - ln0, sum(su1) - sum(su0), der, np.array(spr))
    f, t = res["fold"], res["stored"]
    assert abs(f[0] - t[0]) <= 1e-12 * abs(t[0]) and np.abs(f[1] - t[1]).max() <= 1e-9
    assert np.allclose(f[8], t[8], rtol=1e-9, atol=1e-9), np.abs(f[8] - t[8]).max()
    assert np.allclose(f[2], t[2], rtol=1e-8, atol=1e-12) and abs(f[3] - t[3]) <= 1e-10 * abs(t[3])
    assert abs(f[4] - t[4]) <= 1e-3 and abs(f[5] - t[5]) <= 1e-4 * t[5]
    assert f[9].shape == t[9].shape and np.allclose(f[9], t[9], rtol=1e-10)
    assert f[7] <= t[7] and f[6] < t[6]          # no more site-updates (a folded view never evicts a stored one), fewer launches


def test_real_data_per_site_lnl_and_search(gpu_ctx, golden):
    """the reference's own example genomes (tests/golden/make_real.py): per-site lnL against raxmlHPC -f g, and the engine's
    search from a parsimony start tree ends at least as high as raxmlHPC -f d did (two strains are identical: branches at
    the zmax bound)"""
    import re
    g = golden("aquificales")
    fe = g.meta["fe"]
    aln, tree = _load(gpu_ctx, g, fe["alpha"], fe["tree"])
    lnl, ps = tree.evaluate(per_site=True)
    ref = np.array(g.meta["fg"]["per_site"])
    assert np.abs(ps - ref).max() < 2e-6 and abs(lnl - fe["lnl"]) <= REL_REF * abs(fe["lnl"])
    tree.close()
    t2 = pb.Tree(aln, parsimony_seed=12345)
    t2.optimize(True, 5.0)
    t2.search(radius=5, max_rounds=10, eps=0.1)
    l2, a2 = t2.optimize(True, 0.1)
    assert l2 >= g.meta["fd"]["lnl"] - 0.5, (l2, g.meta["fd"]["lnl"])
    t2.close(); aln.close()


def test_two_trees_on_one_alignment_do_not_see_each_others_state(gpu_ctx, golden):
    """The model constants and the product table live with the alignment, the views with each tree: changing alpha through
    tree A, or filling the table from tree B, must not leave the other tree on stale state (no manual invalidate here)."""
    g = golden("small")
    fe = g.meta["fe"]
    m = orc.Model()
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=fe["alpha"])
    ta = pb.Tree(aln, fe["tree"])
    tb = pb.Tree(aln, g.meta["fw"]["tree"])                      # same topology family, other branch lengths
    la, lb = ta.evaluate(), tb.evaluate()                        # both trees now hold views computed under fe's alpha
    ota, otb = orc.Tree(fe["tree"], g.names), orc.Tree(g.meta["fw"]["tree"], g.names)
    assert abs(lb - orc.evaluate(m, otb, g.pat, g.w, fe["alpha"])) <= REL_ORACLE * abs(lb)
    # alpha moves under tree B's feet
    aln.set_model(0.37)
    want_b = orc.evaluate(m, otb, g.pat, g.w, 0.37)
    assert abs(tb.evaluate() - want_b) <= REL_ORACLE * abs(want_b)
    want_a = orc.evaluate(m, ota, g.pat, g.w, 0.37)
    assert abs(ta.evaluate() - want_a) <= REL_ORACLE * abs(want_a)
    # an optimisation on A changes alpha again (through the engine this time); B must follow without being told
    _, alpha = ta.optimize(True, 0.1)
    want_b = orc.evaluate(m, otb, g.pat, g.w, alpha)
    assert abs(tb.evaluate() - want_b) <= REL_ORACLE * abs(want_b)
    # product table: A prepares branch 2, B overwrites the shared table with its own branch 2, A asks again at another length
    a0 = ta.branch_derivs(2, 0.11)
    b0 = tb.branch_derivs(2, 0.23)
    a1 = ta.branch_derivs(2, 0.17)
    fresh = pb.Tree(aln, ta.newick().replace("):0.0;", ");"))
    want = fresh.branch_derivs(2, 0.17)
    for x, y in zip(a1, want):
        assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (a1, want)
    assert a0 != a1 and b0 != a1
    fresh.close(); ta.close(); tb.close(); aln.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ntax", [3, 4, 5, 6])
def test_smallest_trees_where_cherries_meet(gpu_ctx, golden, ntax):
    """3 to 6 taxa: every inner node is a (folded) cherry or next to one -- cherry against tip, cherry against cherry, cherry as a
    child and as the far end -- including the cases where a folded cherry has to be stored after all.  lnL and derivatives on
    EVERY branch against the CPU oracle, the oracle's own `-f e` optimum within its stopping tolerance."""
    g = golden("small")
    names, seqs = g.names[:ntax], g.seqs[:ntax]
    shapes = {3: "(%s,%s,%s);", 4: "((%s,%s),%s,%s);", 5: "((%s,%s),(%s,%s),%s);", 6: "((%s,%s),(%s,%s),(%s,%s));"}
    nwk = shapes[ntax] % tuple(names)
    pat, w, _ = orc.compress(orc.encode(seqs))
    m = orc.Model()
    ot = orc.Tree(nwk, names)
    aln = pb.Alignment(gpu_ctx, names, seqs, alpha=0.7)
    tree = pb.Tree(aln, nwk)
    rng = np.random.default_rng(ntax)
    lens = {}
    for e in range(tree.num_branches):
        a, b, _ = tree.branch(e)
        lens[e] = float(rng.uniform(0.02, 0.6))
        tree.set_branch(e, lens[e])
    # the oracle numbers its branches differently: match them by their two end nodes' leaf sets through unique lengths
    for oe in range(ot.nedge):
        ot.set_bl(oe, 0.0)
    want_nwk = tree.newick()
    ot = orc.Tree(want_nwk.replace("):0.0;", ");"), names)
    want = orc.evaluate(m, ot, pat, w, 0.7)
    got = tree.evaluate()
    assert abs(got - want) <= 1e-10 * abs(want)
    olen = {round(ot.get_bl(oe), 12): oe for oe in range(ot.nedge)}
    for e in range(tree.num_branches):
        oe = olen[round(lens[e], 12)]
        for t in (lens[e], 0.3):
            gl, g1, g2 = tree.branch_derivs(e, t)
            wl, w1, w2 = orc.branch_derivs(m, ot, pat, w, 0.7, oe, t)
            assert abs(gl - wl) <= 1e-10 * abs(wl) and abs(g1 - w1) <= 1e-8 * max(1.0, abs(w1)) and abs(g2 - w2) <= 1e-8 * max(1.0, abs(w2))
    tree.invalidate()
    assert abs(tree.evaluate() - want) <= 1e-10 * abs(want)
    lnl, alpha = tree.optimize(True, 0.1)
    olnl, oalpha = orc.optimize(m, ot, pat, w, 0.7, True, 0.1)
    assert lnl >= olnl - 0.1 and lnl >= want - 1e-9
    if ntax >= 4:
        node = ntax                                                   # first inner node
        for keep in tree.neighbors(node):
            tg, sc = tree.score_spr_candidates(node, keep, 3)
            assert np.all(np.isfinite(sc))
    tree.close(); aln.close()
