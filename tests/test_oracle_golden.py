"""The CPU oracle against the reference binary's recorded outputs (tests/golden, made by make_golden.py running
/root/reference/pepr-bin_linux/raxmlHPC).  raxmlHPC prints lnL with 6 decimals, so 'all printed digits' = 1e-6 absolute
(plus the second-order effect of alpha being printed with 6 decimals)."""
import numpy as np
import pytest

from oracle import oracle as orc


@pytest.mark.parametrize("case", ["small", "dup", "deep", "wide", "aquificales", "erysipelotrichales"])
def test_fixed_parameter_lnl_matches_reference(golden, case):
    g = golden(case)
    fe = g.meta["fe"]
    assert g.pat.shape[1] == fe["patterns"]
    t = orc.Tree(fe["tree"], g.names)
    lnl = orc.evaluate(orc.Model(), t, g.pat, g.w, fe["alpha"])
    assert abs(lnl - fe["lnl"]) <= 2e-6 * max(1.0, abs(fe["lnl"]) / 1e3), (lnl, fe["lnl"])
    assert abs(lnl - fe["lnl"]) / abs(fe["lnl"]) < 1e-8


def test_lnl_is_the_same_on_every_branch(golden):
    g = golden("small")
    t = orc.Tree(g.meta["fe"]["tree"], g.names)
    m = orc.Model()
    vals = [orc.evaluate(m, t, g.pat, g.w, g.meta["fe"]["alpha"], edge=e) for e in range(t.nedge)]
    assert max(vals) - min(vals) < 1e-9


@pytest.mark.parametrize("case", ["small", "aquificales", "erysipelotrichales"])
def test_per_site_lnl_matches_reference(golden, case):
    g = golden(case)
    fg = g.meta["fg"]
    t = orc.Tree(g.meta["fe"]["tree"], g.names)
    lnl, pp = orc.evaluate(orc.Model(), t, g.pat, g.w, g.meta["fe"]["alpha"], per_pattern=True)
    ref = np.array(fg["per_site"])
    assert len(ref) == len(g.s2p)
    assert np.abs(pp[g.s2p] - ref).max() < 2e-6   # 6 printed decimals


def test_integer_column_weights_match_reference(golden):
    g = golden("small")
    fw = g.meta["fw"]
    pat, w, _ = orc.compress(g.codes, np.array(fw["weights"], np.int32))
    assert pat.shape[1] == fw["patterns"]
    t = orc.Tree(fw["tree"], g.names)
    lnl = orc.evaluate(orc.Model(), t, pat, w, fw["alpha"])
    assert abs(lnl - fw["lnl"]) < 2e-6


def test_bootstrap_weights_bit_exact(golden):
    g = golden("small")
    fj = g.meta["fj"]
    assert g.w.tolist() == fj["pattern_weights"]
    w, _ = orc.bootstrap_weights(fj["seed"], g.w, 3)
    assert w.tolist() == fj["replicate_weights"]


def test_scaling_fires_on_deep_tree(golden):
    g = golden("deep")
    t = orc.Tree(g.meta["fe"]["tree"], g.names)
    _, sc = orc.evaluate(orc.Model(), t, g.pat, g.w, g.meta["fe"]["alpha"], scalers=True)
    assert sc.max() >= 1


def test_optimiser_reaches_reference_optimum(golden):
    g = golden("small")
    t = orc.Tree(g.meta["tree_in"], g.names)
    lnl, alpha = orc.optimize(orc.Model(), t, g.pat, g.w, 1.0)
    assert abs(lnl - g.meta["fe"]["lnl"]) < 0.1       # raxml's own stopping tolerance (-e 0.1)
    assert abs(alpha - g.meta["fe"]["alpha"]) / g.meta["fe"]["alpha"] < 0.05


def test_derivatives_agree_with_finite_differences(golden):
    g = golden("small")
    t = orc.Tree(g.meta["fe"]["tree"], g.names)
    m, a = orc.Model(), g.meta["fe"]["alpha"]
    for e in (0, 3, 7):
        x = max(t.get_bl(e), 0.05)
        h = 1e-5
        l0, d1, d2 = orc.branch_derivs(m, t, g.pat, g.w, a, e, x)
        lp = orc.branch_derivs(m, t, g.pat, g.w, a, e, x + h)[0]
        lm = orc.branch_derivs(m, t, g.pat, g.w, a, e, x - h)[0]
        assert abs(d1 - (lp - lm) / (2 * h)) < 1e-4 * max(1, abs(d1))
        assert abs(d2 - (lp - 2 * l0 + lm) / h ** 2) < 1e-2 * max(1, abs(d2))


def _labels(newick):
    """split (frozenset of the side not containing the first taxon alphabetically) -> integer label"""
    import re
    t = orc._parse_topology(newick)
    lab = {}
    s = newick.strip().rstrip(";")
    pos = [0]

    def node():
        leaves = frozenset()
        if s[pos[0]] == "(":
            pos[0] += 1
            while True:
                leaves |= node()
                if s[pos[0]] == ",":
                    pos[0] += 1
                    continue
                pos[0] += 1
                break
            m = re.match(r"([^:,()]*)(:[-+0-9.eE]+)?", s[pos[0]:])
            pos[0] += m.end()
            if m.group(1):
                lab[leaves] = int(m.group(1))
            return leaves
        m = re.match(r"([^:,()]*)(:[-+0-9.eE]+)?", s[pos[0]:])
        pos[0] += m.end()
        return frozenset([m.group(1)])

    allt = node()
    out = {}
    for k, v in lab.items():
        if 1 < len(k) < len(allt) - 1:
            out[orc._canon(k, sorted(allt))] = v
    return out


def test_support_counts_match_raxml_f_b(golden):
    g = golden("small")
    fb = g.meta["fb"]
    counts = orc.support_counts(g.meta["fe"]["tree"], fb["support_trees"])
    ref = _labels(fb["bipartitions"])
    n = len(fb["support_trees"])
    assert set(counts) == set(ref)
    for k, c in counts.items():
        assert int(np.floor(100.0 * c / n + 0.5)) == ref[k]
