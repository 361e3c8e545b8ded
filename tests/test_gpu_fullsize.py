"""Parity at BASELINE.json's full single-GPU size (100 taxa x 100k sites, the bench workload) through size-independent
properties, anchored on the oracle for a slice it can finish in seconds:
  additivity over columns (lnL and both derivatives of the whole alignment = sum over disjoint column blocks, one of which is
  checked against the oracle), checksum of checksums (sum of per-site lnL = lnL), independence of the traversal state,
  replicate lnL = weighted sum of per-pattern lnL, replicate weights summing to the number of sites."""
import re

import numpy as np
import pytest

import pepr_b200 as pb
from oracle import oracle as orc
from pepr_b200 import synth

pytestmark = pytest.mark.gpu

NTAX, NSITES, SLICE = 100, 100_000, 3_000


@pytest.fixture(scope="module")
def workload():
    names, seqs, nwk = synth.simulate_wag(NTAX, NSITES, 3)
    return names, seqs, nwk


def test_additivity_checksums_and_oracle_anchor(gpu_ctx, workload):
    names, seqs, nwk = workload
    whole = pb.Alignment(gpu_ctx, names, seqs, alpha=0.8)
    tw = pb.Tree(whole, nwk)
    lnl, per_site = tw.evaluate(per_site=True)
    assert per_site.shape == (NSITES,)
    assert abs(per_site.sum() - lnl) <= 1e-11 * abs(lnl)                      # checksum of checksums
    e = 17
    _, _, t_e = tw.branch(e)
    dw = tw.branch_derivs(e, 0.5 * t_e + 0.01)
    # traversal state must not matter: derivative calls re-orient CLVs all over the tree
    for b in (3, 120, 60):
        tw.branch_derivs(b, 0.1)
    assert abs(tw.evaluate() - lnl) <= 1e-12 * abs(lnl)
    tw.invalidate()
    assert abs(tw.evaluate() - lnl) <= 1e-12 * abs(lnl)
    # disjoint column blocks: [0, SLICE) is small enough for the oracle
    cuts = [0, SLICE, 40_000, NSITES]
    tot, d1, d2 = 0.0, 0.0, 0.0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = pb.Alignment(gpu_ctx, names, [s[lo:hi] for s in seqs], alpha=0.8)
        tp = pb.Tree(part, nwk)
        l = tp.evaluate()
        assert abs(l - per_site[lo:hi].sum()) <= 1e-10 * abs(l)
        d = tp.branch_derivs(e, 0.5 * t_e + 0.01)
        tot, d1, d2 = tot + l, d1 + d[1], d2 + d[2]
        if lo == 0:
            pat, w, _ = orc.compress(orc.encode([s[lo:hi] for s in seqs]))
            want = orc.evaluate(orc.Model(), orc.Tree(nwk, names), pat, w, 0.8)
            assert abs(l - want) <= 1e-10 * abs(want), (l, want)
        tp.close(); part.close()
    assert abs(tot - lnl) <= 1e-11 * abs(lnl)
    assert abs(d1 - dw[1]) <= 1e-8 * max(1.0, abs(dw[1])) and abs(d2 - dw[2]) <= 1e-8 * max(1.0, abs(dw[2]))
    tw.close(); whole.close()


def test_replicates_at_full_size(gpu_ctx, workload):
    names, seqs, nwk = workload
    aln = pb.Alignment(gpu_ctx, names, seqs, alpha=1.0)
    tree = pb.Tree(aln, nwk)
    W, _ = aln.bootstrap_weights(12345, 3)
    assert (W.sum(axis=1) == NSITES).all() and (W >= 0).all()
    # bit-exact against the host restatement of raxmlHPC's stream (oracle), on the real pattern weights
    _, w, _ = pb.crunch_patterns(seqs)
    want = orc.bootstrap_weights(12345, w, 3)
    want = want[0] if isinstance(want, tuple) else want
    assert np.array_equal(W, np.asarray(want, dtype=np.int32))
    rl = tree.evaluate_replicates(W)
    for r in range(3):
        direct = tree.evaluate(weights=W[r])
        assert abs(direct - rl[r]) <= 1e-11 * abs(direct)
    tree.close(); aln.close()


def test_fe_at_full_size_is_a_fixed_point(gpu_ctx, workload):
    """`-f e` from default lengths on the true topology: the optimum beats the generating parameters, a second run changes
    nothing measurable, and the branch lengths land near the simulated ones (100k sites carry a lot of signal)."""
    names, seqs, nwk = workload
    topo = re.sub(r":[0-9.eE+-]+", "", nwk)
    aln = pb.Alignment(gpu_ctx, names, seqs, alpha=1.0)
    truth = pb.Tree(aln, nwk)
    l_true = truth.evaluate()
    true_len = {frozenset(truth.branch(e)[:2]): truth.branch(e)[2] for e in range(truth.num_branches)}
    truth.close()
    tree = pb.Tree(aln, topo)
    lnl, alpha = tree.optimize(True, 0.1)
    assert lnl >= l_true and abs(alpha - 1.0) < 0.05
    again, alpha2 = tree.optimize(True, 0.1)
    assert 0.0 <= again - lnl <= 0.1 + 1e-6
    got = np.array([tree.branch(e)[2] for e in range(tree.num_branches)])
    ref = np.array([true_len[frozenset(tree.branch(e)[:2])] for e in range(tree.num_branches)])
    assert np.abs(got - ref).max() < 0.02 and np.corrcoef(got, ref)[0, 1] > 0.999
    tree.close(); aln.close()
