"""Tree search on the GPU engine: lazy-SPR candidate scoring against the oracle, and whole searches against what
raxmlHPC `-f d` / `-f a` found on the same alignment (tests/golden/search.json).  Heuristic searches are compared by the
likelihood and the splits of the tree they end on, not move by move."""
import re

import numpy as np
import pytest

import pepr_b200 as pb
from oracle import oracle as orc
from pepr_b200 import runner as R

pytestmark = pytest.mark.gpu


def _splits(newick):
    sets = []
    taxa = sorted(orc._leafsets(orc._parse_topology(newick), sets))
    return {orc._canon(x, taxa) for x in sets if 1 < len(x) < len(taxa) - 1}


def _strip(nw):
    return nw.replace("):0.0;", ");")


def test_lazy_spr_scores_match_oracle(gpu_ctx, golden):
    g = golden("small")
    fe = g.meta["fe"]
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=fe["alpha"])
    tree = pb.Tree(aln, fe["tree"])
    base = tree.evaluate()
    m = orc.Model()
    checked = 0
    for node in range(aln.ntax, 2 * aln.ntax - 2):
        for keep in tree.neighbors(node):
            targets, lnl = tree.score_spr_candidates(node, keep, radius=3)
            assert abs(tree.evaluate() - base) <= 1e-9 * abs(base)          # scoring leaves the tree as it was
            for tgt, l in list(zip(targets, lnl))[:2]:
                moved = pb.Tree(aln, fe["tree"])                            # same parser -> same node / branch ids
                moved.spr(node, keep, int(tgt))
                want = orc.evaluate(m, orc.Tree(_strip(moved.newick()), g.names), g.pat, g.w, fe["alpha"])
                assert abs(l - want) <= 1e-9 * abs(want)
                assert abs(moved.evaluate() - want) <= 1e-9 * abs(want)
                moved.close()
                checked += 1
    assert checked >= 20
    tree.close(); aln.close()


@pytest.mark.parametrize("case", ["search", "deep", "small"])
def test_device_parsimony_scan_gives_the_host_tree(gpu_ctx, golden, case):
    """integer work: the GPU Fitch scans must pick exactly the branches the host implementation picks (same seed)"""
    g = golden(case)
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=1.0)
    for seed in (12345, 777):
        tree = pb.Tree(aln, parsimony_seed=seed)
        want, _ = pb.parsimony_tree(g.names, g.seqs, seed)
        assert _splits(_strip(tree.newick())) == _splits(_strip(want))
        tree.close()
    # a replicate's weights change the costs and (usually) the tree; zero-weight patterns must not count
    W, _ = aln.bootstrap_weights(4242, 1)
    t1 = pb.Tree(aln, parsimony_seed=12345, weights=W[0])
    assert t1.num_branches == 2 * len(g.names) - 3
    t1.close(); aln.close()


def test_search_finds_reference_tree(gpu_ctx, golden):
    g = golden("search")
    fd = g.meta["fd"]
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=1.0)
    tree = pb.Tree(aln, parsimony_seed=12345)
    tree.optimize(True, 5.0)
    lnl, moves = tree.search(radius=5, max_rounds=10, eps=0.1)
    lnl, alpha = tree.optimize(True, 0.1)
    # raxmlHPC -f d stops at dlnL <= 0.1 too; both searches are hill climbers from (different) parsimony trees
    assert lnl >= fd["lnl"] - 0.5, (lnl, fd["lnl"], moves)
    assert _splits(_strip(tree.newick())) == _splits(fd["tree"])
    # and the engine's lnL for that tree is what the oracle computes for it
    chk = orc.evaluate(orc.Model(), orc.Tree(_strip(tree.newick()), g.names), g.pat, g.w, alpha)
    assert abs(chk - lnl) <= 1e-9 * abs(lnl)
    tree.close(); aln.close()


def test_search_repairs_a_damaged_tree(gpu_ctx, golden):
    """start from the reference tree with a few random SPR moves applied: the search must climb back"""
    g = golden("search")
    fd = g.meta["fd"]
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=fd["alpha"])
    tree = pb.Tree(aln, fd["tree"])
    good = tree.evaluate()
    rng = np.random.default_rng(4)
    done = 0
    while done < 4:
        node = int(rng.integers(aln.ntax, 2 * aln.ntax - 2))
        keep = tree.neighbors(node)[int(rng.integers(0, 3))]
        targets, _ = tree.score_spr_candidates(node, keep, radius=4)
        if len(targets):
            tree.spr(node, keep, int(targets[int(rng.integers(0, len(targets)))]))
            done += 1
    bad = tree.evaluate()
    assert bad < good - 10
    lnl, moves = tree.search(radius=6, max_rounds=10, eps=0.1)
    assert moves >= 1 and lnl > good - 1.0
    assert _splits(_strip(tree.newick())) == _splits(fd["tree"])
    tree.close(); aln.close()


def test_runner_f_d_and_f_a(gpu_ctx, golden):
    g = golden("search")
    r = R.B200MLRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r.setBootstrapReps(6)
    r.run()                                                                  # `-f a -x 12345 -N 6`
    assert r.last_error is None
    assert r.getLikelihood() >= g.meta["fd"]["lnl"] - 0.5
    assert _splits(_strip(r.getBestTree())) == _splits(g.meta["fd"]["tree"])
    sup = r.getBestTreeWithSupports()
    labels = [int(x) for x in re.findall(r"\)(\d+)", sup)]
    assert len(labels) == len(g.names) - 3 and all(0 <= v <= 100 for v in labels)
    # the data carry strong signal: raxmlHPC's own rapid bootstrap gives (nearly) full support, so must the replicates here
    ref = [int(x) for x in re.findall(r"\)(\d+)", g.meta["fa"]["bipartitions"])]
    assert np.mean(labels) >= np.mean(ref) - 15


def test_fasttree_runner_mirror_gives_fasttrees_tree(gpu_ctx, golden):
    """FastTreeRunner.run (FastTreeRunner.java:38-135) served by the engine: same splits as the bundled FastTree_WAG found
    (golden made by running the reference binary), supports as integer percentages, unsupported options refused loudly"""
    g = golden("search")
    ft = g.meta["fasttree"]
    r = R.B200FastTreeRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(g.names, g.seqs))
    r.run()
    assert r.last_error is None and r.getResult()
    assert _splits(_strip(r.getResult())) == _splits(_strip(re.sub(r"\)[0-9.]+", ")", ft["tree"])))
    r.setBootstrapReps(4)
    r.setUseRaxmlBranchLengths(True)
    r.run()
    labels = [int(x) for x in re.findall(r"\)(\d+)", r.getResult())]
    assert len(labels) == len(g.names) - 3 and all(0 <= v <= 100 for v in labels)
    r.setNucleotide(True)
    r.run()
    assert r.getResult() is None and "not supported" in r.last_error


def test_fasttree_runner_honours_topological_constraints(gpu_ctx, golden):
    """`FastTree -constraints` (FastTreeRunner.java:53-83): a constraint tree that contradicts the data's own tree must come out
    displayed -- at a likelihood cost; the unconstrained part of the tree is still the search's; a constraint tree that agrees
    with the data changes nothing"""
    g = golden("search")
    names = g.names
    free = _splits(g.meta["fd"]["tree"])
    # a wrong clade: three taxa that do not form one in the ML tree
    a, b, c = names[0], names[5], names[11]
    taxa = sorted(names)
    assert orc._canon(frozenset([a, b, c]), taxa) not in free
    others = [n for n in names if n not in (a, b, c)]
    con_tree = "(((%s,%s),%s),%s);" % (a, c, b, ",".join(others))
    r = R.B200FastTreeRunner(ctx=gpu_ctx)
    r.setAlignment(R.SequenceAlignment(names, g.seqs))
    r.run()
    lnl_free = r.getLikelihood()
    r.setConstraintTree(con_tree)
    r.run()
    assert r.last_error is None and r.getResult()
    got = _splits(_strip(r.getResult()))
    assert orc._canon(frozenset([a, b, c]), taxa) in got and orc._canon(frozenset([a, c]), taxa) in got
    assert pb.newick_satisfies_constraints(_strip(r.getResult()), names, pb.constraints_from_tree(con_tree))
    assert r.getLikelihood() < lnl_free - 10                      # the forced clade costs likelihood
    assert len(got & free) >= len(free) - 6                        # ... and only the neighbourhood of the forced clade changes
    # through the tree-builder dispatch (PhylogeneticTreeBuilder.buildFastTree passes the constraint tree on)
    tb = R.B200TreeBuilder(ctx=gpu_ctx)
    tb.setAlignment(R.SequenceAlignment(names, g.seqs))
    tb.setTreeBuildingMethod(R.FAST_TREE)
    tb.setBootstrapReps(0)
    tb.setConstraintTree(_strip(g.meta["fd"]["tree"]))            # agrees with the data: the ML tree itself
    tb.run()
    assert tb.last_error is None and _splits(_strip(tb.getTreeString())) == free


def test_tree_builder_dispatch(gpu_ctx, golden):
    """PhylogeneticTreeBuilder.run (PhylogeneticTreeBuilder.java:97-129): ml / parsimony / parsimony_bl / FastTree reach the
    engine's runners and hand back what the reference's build* methods hand back; methods off the path are reported"""
    g = golden("search")
    aln = R.SequenceAlignment(g.names, g.seqs)
    out = {}
    for method in (R.MAXIMUM_LIKELIHOOD, R.PARSIMONY, R.PARSIMONY_BL, R.FAST_TREE):
        tb = R.B200TreeBuilder(ctx=gpu_ctx)
        tb.setAlignment(aln)
        tb.setTreeBuildingMethod(method)
        tb.setBootstrapReps(0)
        tb.setMLMatrix("PROTGAMMAWAG")
        tb.run()
        assert tb.last_error is None, (method, tb.last_error)
        out[method] = tb.getTreeString()
    ref = _splits(g.meta["fd"]["tree"])
    assert _splits(_strip(out[R.MAXIMUM_LIKELIHOOD])) == ref and _splits(_strip(out[R.FAST_TREE])) == ref
    assert ":" not in out[R.PARSIMONY] and all(n in out[R.PARSIMONY] for n in g.names)          # RAxML_parsimonyTree: topology only
    assert _splits(_strip(out[R.PARSIMONY_BL])) == _splits(out[R.PARSIMONY]) and ":" in out[R.PARSIMONY_BL]
    tb = R.B200TreeBuilder(ctx=gpu_ctx)
    tb.setAlignment(aln)
    tb.setTreeBuildingMethod(R.MAXIMUM_LIKELIHOOD)
    tb.setBootstrapReps(3)
    tb.run()
    assert len(re.findall(r"\)(\d+)", tb.getTreeString())) == len(g.names) - 3               # getBestTreeWithSupports
    tb.setTreeBuildingMethod(R.NEIGHBOR_JOINING)
    tb.run()
    assert tb.getTreeString() is None and "not on the accelerated path" in tb.last_error


def test_gene_wise_jackknife_as_weight_masks(gpu_ctx, golden):
    """a gene-wise jackknife replicate as a 0/1 column mask over the ONE resident supermatrix equals the
    alignment concatenated from the kept genes alone (what PhylogenomicPipeline2.java:1227-1275 builds per replicate): same
    lnL on a fixed tree to rounding, same `-f e` result, and the replicate tree search runs on the mask directly"""
    g = golden("wide")
    L = len(g.seqs[0])
    blocks = [300] * (L // 300)
    blocks[-1] += L - sum(blocks)
    keep = R.getRandomSet(len(blocks) // 2, 0, len(blocks) - 1, False, seed=2024)     # PEPR: half of the genes, no reuse
    w = R.gene_block_weights(blocks, keep)
    assert w.sum() == sum(blocks[k] for k in keep)
    starts = np.concatenate([[0], np.cumsum(blocks)])
    sub = ["".join(s[starts[k]:starts[k + 1]] for k in sorted(keep)) for s in g.seqs]
    tree_nwk = g.meta["fe"]["tree"]
    a_mask = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=0.8, site_weights=w)
    a_sub = pb.Alignment(gpu_ctx, g.names, sub, alpha=0.8)
    t1, t2 = pb.Tree(a_mask, tree_nwk), pb.Tree(a_sub, tree_nwk)
    l1, l2 = t1.evaluate(), t2.evaluate()
    assert abs(l1 - l2) <= 1e-11 * abs(l2)
    t1.close(); t2.close(); a_mask.close(); a_sub.close()
    res = []
    for seqs, weights in ((g.seqs, w), (sub, None)):
        r = R.B200MLRunner(ctx=gpu_ctx)
        r.setAlignment(R.SequenceAlignment(g.names, seqs))
        r.setSiteWeights(weights)
        r.setStartTree(g.meta["tree_in"])
        r.run()
        assert r.last_error is None
        res.append((r.getLikelihood(), r.getAlpha()))
    assert abs(res[0][0] - res[1][0]) < 1e-3 and abs(res[0][1] - res[1][1]) < 1e-4
