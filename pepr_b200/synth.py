"""Synthetic WAG+Gamma4 supermatrices of the BASELINE.json shapes.

Random unrooted binary topology by sequential random joining, branch lengths 0.005 + Exp(mean 0.08), root states ~ pi,
per-site Gamma category uniform over the 4 mean-rate categories, evolved with P_c(t).  Optional PEPR-like missing-gene
blocks ('?' runs, the padding MSAConcatenator.concatenate writes for absent genes, MSAConcatenator.java:164-170).

The caller supplies `pmatrix(t, rate) -> (20,20) row-stochastic array` and `rates` so that this module has no
dependency on either the CUDA engine or the test oracle.
"""
import numpy as np

AA = "ARNDCQEGHILKMFPSTWYV"


def random_tree(ntax, rng, mean_bl=0.08, min_bl=0.005):
    """returns (newick string with lengths, nested structure) ; taxa are named T0000.."""
    items = [("T%04d" % i,) for i in range(ntax)]
    while len(items) > 3:
        i, j = sorted(rng.choice(len(items), 2, replace=False))
        a, b = items[i], items[j]
        items = [x for k, x in enumerate(items) if k not in (i, j)] + [((a, b),)]
    return items


def simulate(ntax, nsites, seed, pmatrix, rates, pi, missing_frac=0.0, block=(200, 400), mean_bl=0.08):
    """-> names (sorted), list of sequence strings, newick of the true tree"""
    rng = np.random.default_rng(seed)
    rates = np.asarray(rates)
    cat = rng.integers(0, len(rates), nsites)
    items = random_tree(ntax, rng)
    seqs = {}

    def evolve(node, parent):
        t = float(rng.exponential(mean_bl)) + 0.005
        child = np.empty(nsites, dtype=np.int64)
        for k, r in enumerate(rates):
            m = cat == k
            P = np.clip(pmatrix(t, r), 0, None)
            P /= P.sum(1, keepdims=True)
            cdf = P.cumsum(1)[parent[m]]
            u = rng.random(int(m.sum()))[:, None]
            child[m] = np.minimum((u > cdf).sum(1), 19)
        nm = node[0]
        if isinstance(nm, str):
            seqs[nm] = child
            return "%s:%.6f" % (nm, t)
        return "(%s,%s):%.6f" % (evolve(nm[0], child), evolve(nm[1], child), t)

    root = rng.choice(20, nsites, p=np.asarray(pi) / np.sum(pi))
    nwk = "(" + ",".join(evolve(x, root) for x in items) + ");"
    names = sorted(seqs)
    A = np.frombuffer(AA.encode(), np.uint8)
    mat = np.stack([A[seqs[k]] for k in names])
    if missing_frac > 0:
        target = int(missing_frac * ntax * nsites)
        done = 0
        while done < target:
            tx = int(rng.integers(0, ntax))
            ln = int(rng.integers(block[0], block[1] + 1))
            st = int(rng.integers(0, max(1, nsites - ln)))
            mat[tx, st:st + ln] = ord("?")
            done += ln
    return names, [bytes(r).decode() for r in mat], nwk


def write_phylip(path, names, seqs):
    """relaxed phylip as SequenceAlignment.getAlignmentAsExtendedPhylipUsingTaxonNames writes it
    (SequenceAlignment.java:489-522): 'ntax len', then name padded to longest+1 followed by the whole sequence."""
    w = max(len(n) for n in names) + 1
    with open(path, "w") as f:
        f.write("%d %d\n" % (len(names), len(seqs[0])))
        for n, s in zip(names, seqs):
            f.write(n.ljust(w) + s + "\n")
