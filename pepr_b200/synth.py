"""Synthetic WAG+Gamma4 supermatrices of the BASELINE.json shapes.

Random unrooted binary topology by sequential random joining, branch lengths 0.005 + Exp(mean 0.08), root states ~ pi,
per-site Gamma category uniform over the 4 mean-rate categories, evolved with P_c(t).  Optional PEPR-like missing-gene
blocks ('?' runs, the padding MSAConcatenator.concatenate writes for absent genes, MSAConcatenator.java:164-170).

The caller supplies `pmatrix(t, rate) -> (20,20) row-stochastic array` and `rates` so that this module has no
dependency on either the CUDA engine or the test oracle.
"""
import numpy as np

AA = "ARNDCQEGHILKMFPSTWYV"

# WAG exchangeabilities (PAML wag.dat x100, lower triangle) and the fixed 3-decimal frequencies; the generator carries its
# own copy so that neither the engine nor the oracle is needed to produce benchmark input.
_WAG = """55.1571 50.9848 63.5346 73.8998 14.7304 542.942 102.704 52.8191 26.5256 3.02949 90.8598 303.55 154.364 61.6783
9.88179 158.285 43.9157 94.7198 617.416 2.1352 546.947 141.672 58.4665 112.556 86.5584 30.6674 33.0052 56.7717 31.6954
213.715 395.629 93.0676 24.8972 429.411 57.0025 24.941 19.3335 18.6979 55.4236 3.9437 17.0135 11.3917 12.7395 3.04501
13.819 39.7915 49.7671 13.1528 8.48047 38.4287 86.9489 15.4263 6.13037 49.9462 317.097 90.6265 535.142 301.201 47.9855
7.40339 389.49 258.443 37.3558 89.0432 32.3832 25.7555 89.3496 68.3162 19.8221 10.3754 39.0482 154.526 31.5124 17.41
40.4141 425.746 485.402 93.4276 21.0494 10.2711 9.61621 4.67304 39.802 9.99208 8.11339 4.9931 67.9371 105.947 211.517
8.8836 119.063 143.855 67.9489 19.5081 42.3984 10.9404 93.3372 68.2355 24.357 69.6198 9.99288 41.5844 55.6896 17.1329
16.1444 337.079 122.419 397.423 107.176 140.766 102.887 70.4939 134.182 74.0169 31.944 34.4739 96.713 49.3905 54.5931
161.328 212.111 55.4413 203.006 37.4866 51.2984 85.7928 82.2765 22.5833 47.3307 145.816 32.6622 138.698 151.612 17.1903
79.5384 437.802 11.3133 116.392 7.19167 12.9767 71.707 21.5737 15.6557 33.6983 26.2569 21.2483 66.5309 13.7505 51.5706
152.964 13.9405 52.3742 11.0864 24.0735 38.1533 108.6 32.5711 54.3833 22.771 19.6303 10.3604 387.344 42.017 39.8618
13.3264 42.8437 645.428 21.6046 78.6993 29.1148 248.539 200.601 25.1849 19.6246 15.2335 100.214 30.1281 58.8731 18.7247
11.8358 782.13 180.034 30.5434 205.845 64.9892 31.4887 23.2739 138.823 36.5369 31.473"""
WAG_PI = np.array([0.087, 0.044, 0.039, 0.057, 0.019, 0.037, 0.058, 0.083, 0.024, 0.049,
                   0.086, 0.062, 0.020, 0.038, 0.046, 0.070, 0.061, 0.014, 0.035, 0.071])


class WagModel:
    """numpy-only WAG+Gamma4 used to SIMULATE data (not to score it)"""

    def __init__(self, alpha=1.0, ncat=4):
        from scipy.special import gammainc
        from scipy.stats import gamma as G
        S = np.zeros((20, 20))
        S[np.tril_indices(20, -1)] = np.array(_WAG.split(), float)
        S = S + S.T
        Q = S * WAG_PI[None, :]
        np.fill_diagonal(Q, 0.0)
        np.fill_diagonal(Q, -Q.sum(1))
        Q /= -(WAG_PI * np.diag(Q)).sum()
        sq = np.sqrt(WAG_PI)
        A = (sq[:, None] * Q) / sq[None, :]
        self.lam, U = np.linalg.eigh(0.5 * (A + A.T))
        self.V, self.Vi = U / sq[:, None], U.T * sq[None, :]
        self.pi = WAG_PI
        b = G.ppf(np.arange(1, ncat) / ncat, alpha, scale=1.0 / alpha)
        c = np.concatenate([[0.0], gammainc(alpha + 1.0, b * alpha), [1.0]])
        self.rates = np.diff(c) * ncat

    def pmatrix(self, t, rate=1.0):
        return (self.V * np.exp(self.lam * rate * t)) @ self.Vi


def simulate_wag(ntax, nsites, seed, alpha=1.0, missing_frac=0.0):
    m = WagModel(alpha)
    return simulate(ntax, nsites, seed, m.pmatrix, m.rates, m.pi, missing_frac=missing_frac)


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

    cat_sites = [np.flatnonzero(cat == k) for k in range(len(rates))]

    def evolve(node, parent):
        t = float(rng.exponential(mean_bl)) + 0.005
        child = np.empty(nsites, dtype=np.int64)
        for k, r in enumerate(rates):
            P = np.clip(pmatrix(t, r), 0, None)
            P /= P.sum(1, keepdims=True)
            cdf = P.cumsum(1)
            # inverse-CDF draw: the new state is the number of cdf entries of the parent's row below u.  Sites are grouped by
            # parent state so that each group is one searchsorted on one row (20 x faster than comparing against all 20
            # entries per site, same draws and the same result bit for bit)
            sites = cat_sites[k]
            u = rng.random(len(sites))
            par = parent[sites]
            order = np.argsort(par, kind="stable")
            bounds = np.searchsorted(par[order], np.arange(21))
            res = np.empty(len(sites), dtype=np.int64)
            for s in range(20):
                grp = order[bounds[s]:bounds[s + 1]]
                res[grp] = np.searchsorted(cdf[s], u[grp], side="left")
            child[sites] = np.minimum(res, 19)
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


def simulate_wag_device(ntax, nsites, seed, device="cuda", alpha=1.0, mean_bl=0.08):
    """The same model and the same kind of tree as simulate_wag, drawn on a CUDA device with torch (plumbing: 2000 taxa x 1 M
    sites take seconds instead of a quarter of an hour on the host).  The random stream is torch's, so the alignment is NOT
    the one simulate_wag gives for the same seed.  -> names, uint8 array [ntax, nsites] of residue letters, newick"""
    import torch
    m = WagModel(alpha)
    rng = np.random.default_rng(seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    ncat = len(m.rates)
    # sites are laid out category by category while the states evolve; a fixed column permutation mixes them at the end
    bounds = [nsites * k // ncat for k in range(ncat + 1)]
    items = random_tree(ntax, rng)
    rows = {}

    def evolve(node, parent):
        t = float(rng.exponential(mean_bl)) + 0.005
        child = torch.empty(nsites, dtype=torch.uint8, device=device)
        for k in range(ncat):
            P = np.clip(m.pmatrix(t, m.rates[k]), 0, None)
            P /= P.sum(1, keepdims=True)
            cdf = torch.from_numpy(P.cumsum(1)).to(device)
            lo, hi = bounds[k], bounds[k + 1]
            u = torch.rand(hi - lo, generator=gen, dtype=torch.float64, device=device)
            child[lo:hi] = (u[:, None] > cdf[parent[lo:hi].long()]).sum(1).clamp_(max=19).to(torch.uint8)
        nm = node[0]
        if isinstance(nm, str):
            rows[nm] = child
            return "%s:%.6f" % (nm, t)
        return "(%s,%s):%.6f" % (evolve(nm[0], child), evolve(nm[1], child), t)

    pi = torch.from_numpy(np.asarray(m.pi) / np.sum(m.pi)).to(device)
    root = torch.multinomial(pi, nsites, replacement=True, generator=gen).to(torch.uint8)
    nwk = "(" + ",".join(evolve(x, root) for x in items) + ");"
    names = sorted(rows)
    perm = torch.randperm(nsites, generator=gen, device=device)
    letters = torch.from_numpy(np.frombuffer(AA.encode(), np.uint8).copy()).to(device)
    out = np.empty((ntax, nsites), np.uint8)
    for i, nme in enumerate(names):
        out[i] = letters[rows.pop(nme)[perm].long()].cpu().numpy()
    return names, out, nwk


def write_phylip(path, names, seqs):
    """relaxed phylip as SequenceAlignment.getAlignmentAsExtendedPhylipUsingTaxonNames writes it
    (SequenceAlignment.java:489-522): 'ntax len', then name padded to longest+1 followed by the whole sequence."""
    w = max(len(n) for n in names) + 1
    with open(path, "w") as f:
        f.write("%d %d\n" % (len(names), len(seqs[0])))
        for n, s in zip(names, seqs):
            f.write(n.ljust(w) + s + "\n")
