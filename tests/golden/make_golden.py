#!/usr/bin/env python
"""Regenerates tests/golden/* by RUNNING THE REFERENCE'S OWN IMPLEMENTATION of the path in this container:
/root/reference/pepr-bin_linux/raxmlHPC (RAxML 7.2.5) -- the executable PEPR's RAxMLRunner/FastTreeRunner exec
(RAxMLRunner.java:87-147,164-208,216-272,490-503).  The reference has no test fixtures of its own,
so these outputs are the parity pins.  Inputs are produced here from fixed seeds and committed next to the outputs.

usage: python tests/golden/make_golden.py [case ...]     (needs /root/reference; not runnable on the GPU box)
"""
import gzip
import json
import os
import random
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from pepr_b200 import synth  # noqa: E402

RAX = "/root/reference/pepr-bin_linux/raxmlHPC"
RAXP = "/root/reference/pepr-bin_linux/raxmlHPC-PTHREADS"


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("%s failed:\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def info(cwd, n):
    txt = open(os.path.join(cwd, "RAxML_info." + n)).read()
    out = {}
    m = re.search(r"Final GAMMA\s+likelihood: (\S+)", txt)
    if m:
        out["lnl"] = float(m.group(1))
    m = re.search(r"Final GAMMA-based Score of best tree (\S+)", txt)
    if m:
        out["lnl"] = float(m.group(1))
    m = re.search(r"alpha: (\S+)", txt) or re.search(r"alpha\[0\]: (\S+)", txt)
    if m:
        out["alpha"] = float(m.group(1))
    m = re.search(r"Tree-Length: (\S+)", txt)
    if m:
        out["tree_length"] = float(m.group(1))
    m = re.search(r"Alignment has (\d+) distinct alignment patterns", txt)
    if m:
        out["patterns"] = int(m.group(1))
    m = re.search(r"Overall Time for Tree Evaluation (\S+)", txt)
    if m:
        out["seconds"] = float(m.group(1))
    return out


def small_alignment(seed, ntax, nsites, tree_fn, mut=0.12, qblocks=((2, 40, 20),), dashblocks=((5, 100, 10),), extra=()):
    """star-ish evolution by copying with mutation along a given pairing so columns carry signal"""
    rnd = random.Random(seed)
    aa = orc.AA
    pick = lambda: aa[int(rnd.random() * 20)]
    anc = [pick() for _ in range(nsites)]
    names = ["Tax" + chr(ord("A") + i) for i in range(ntax)]
    seqs = []
    prev = anc
    for i in range(ntax):
        base = prev if i % 2 else anc
        s = [c if rnd.random() > mut * (1 + i % 3) else pick() for c in base]
        prev = s
        seqs.append(s)
    for (t, st, ln) in qblocks:
        seqs[t][st:st + ln] = "?" * ln
    for (t, st, ln) in dashblocks:
        seqs[t][st:st + ln] = "-" * ln
    for (t, pos, ch) in extra:
        seqs[t][pos] = ch
    return names, ["".join(s) for s in seqs]


def case_small(d):
    """8 x 300 with '?'/'-' blocks and X/B/Z/* characters: -f e, -f g, -a weights, -f j bootstrap, -f b supports"""
    names, seqs = small_alignment(7, 8, 300, None, extra=((0, 3, "X"), (1, 7, "B"), (3, 11, "Z"), (4, 13, "*"), (6, 200, "B"), (7, 250, "Z")))
    tree = "((TaxA,TaxB),(TaxC,TaxD),((TaxE,TaxF),(TaxG,TaxH)));"
    tmp = tempfile.mkdtemp()
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    g = {"names": names, "tree_in": tree}
    run([RAX, "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fe", "-t", "t.nwk"], tmp)
    g["fe"] = info(tmp, "fe")
    g["fe"]["tree"] = open(os.path.join(tmp, "RAxML_result.fe")).read().strip()
    # per-site lnL (-f g); RAxMLRunner.runRaxmlPerSiteLL + PhylogenomicPipeline2.getTreeScore sum line 2
    run([RAX, "-f", "g", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fg", "-z", "t.nwk"], tmp)
    ps = open(os.path.join(tmp, "RAxML_perSiteLLs.fg")).read().split("\n")[1].split()[1:]
    g["fg"] = {"per_site": [float(x) for x in ps], "tree": open(os.path.join(tmp, "RAxML_result.fg")).read().strip() if os.path.exists(os.path.join(tmp, "RAxML_result.fg")) else None}
    g["fg"].update(info(tmp, "fg"))
    # integer column weights (-a)
    rnd = random.Random(3)
    w = [0] * 300
    for _ in range(300):
        w[int(rnd.random() * 300)] += 1
    open(os.path.join(tmp, "w.txt"), "w").write(" ".join(map(str, w)) + "\n")
    run([RAX, "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fw", "-t", "t.nwk", "-a", "w.txt"], tmp)
    g["fw"] = info(tmp, "fw")
    g["fw"]["tree"] = open(os.path.join(tmp, "RAxML_result.fw")).read().strip()
    g["fw"]["weights"] = w
    # bootstrap replicate alignments (-f j -b seed -# n): columns come out in sorted-pattern order
    run([RAX, "-f", "j", "-b", "12345", "-#", "3", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fj"], tmp)
    codes = orc.encode(seqs)
    pat, pw, s2p = orc.compress(codes)
    reps = []
    for r in range(3):
        _, bs = orc.read_phylip(os.path.join(tmp, "t.phy.BS%d" % r))
        bc = orc.encode(bs)
        cols = {}
        for j in range(bc.shape[1]):
            k = bytes(bc[:, j])
            cols[k] = cols.get(k, 0) + 1
        reps.append([cols.get(bytes(pat[:, p]), 0) for p in range(pat.shape[1])])
        assert sum(reps[-1]) == 300
    g["fj"] = {"seed": 12345, "pattern_weights": pw.tolist(), "replicate_weights": reps}
    # -f b: draw supports of 7 trees on the -f e tree (RAxMLRunner.getSupportDecoratedTree)
    sup = ["((TaxA,TaxB),(TaxC,TaxD),((TaxE,TaxF),(TaxG,TaxH)));", "((TaxA,TaxB),(TaxC,TaxD),((TaxE,TaxF),(TaxG,TaxH)));",
           "((TaxA,TaxC),(TaxB,TaxD),((TaxE,TaxF),(TaxG,TaxH)));", "((TaxA,TaxB),(TaxC,TaxD),((TaxE,TaxG),(TaxF,TaxH)));",
           "(((TaxA,TaxB),TaxC),TaxD,((TaxE,TaxF),(TaxG,TaxH)));", "((TaxA,TaxB),(TaxC,TaxD),(TaxE,(TaxF,(TaxG,TaxH))));",
           "((TaxA,TaxD),(TaxC,TaxB),((TaxE,TaxH),(TaxG,TaxF)));"]
    open(os.path.join(tmp, "sup.trees"), "w").write("\n".join(sup) + "\n")
    open(os.path.join(tmp, "main.nwk"), "w").write(g["fe"]["tree"] + "\n")
    run([RAX, "-f", "b", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fb", "-z", "sup.trees", "-t", "main.nwk"], tmp)
    g["fb"] = {"support_trees": sup, "bipartitions": open(os.path.join(tmp, "RAxML_bipartitions.fb")).read().strip()}
    shutil.copy(os.path.join(tmp, "t.phy"), os.path.join(d, "small.phy"))
    json.dump(g, open(os.path.join(d, "small.json"), "w"), indent=1)
    shutil.rmtree(tmp)


def case_dup(d):
    """10 x 500, low divergence -> many duplicate columns (pattern weights > 1) + a '?' block"""
    names, seqs = small_alignment(11, 10, 500, None, mut=0.01, qblocks=((4, 100, 60),), dashblocks=((9, 300, 40),))
    tree = "(((TaxA,TaxB),(TaxC,TaxD)),TaxE,((TaxF,TaxG),(TaxH,(TaxI,TaxJ))));"
    tmp = tempfile.mkdtemp()
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(tree + "\n")
    run([RAX, "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fe", "-t", "t.nwk"], tmp)
    g = {"names": names, "tree_in": tree, "fe": info(tmp, "fe")}
    g["fe"]["tree"] = open(os.path.join(tmp, "RAxML_result.fe")).read().strip()
    shutil.copy(os.path.join(tmp, "t.phy"), os.path.join(d, "dup.phy"))
    json.dump(g, open(os.path.join(d, "dup.json"), "w"), indent=1)
    shutil.rmtree(tmp)


def _sim(ntax, nsites, seed, missing=0.0):
    m = orc.Model()
    pi = m.arrays()[0]
    return synth.simulate(ntax, nsites, seed, m.pmatrix, orc.gamma_rates(1.0), pi, missing_frac=missing)


def case_deep(d):
    """160 taxa x 150 sites WAG-simulated: deep enough that x2^256 scaling fires"""
    names, seqs, nwk = _sim(160, 150, 5)
    tmp = tempfile.mkdtemp()
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(nwk + "\n")
    run([RAX, "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fe", "-t", "t.nwk"], tmp)
    g = {"names": names, "tree_in": nwk, "fe": info(tmp, "fe")}
    g["fe"]["tree"] = open(os.path.join(tmp, "RAxML_result.fe")).read().strip()
    with open(os.path.join(tmp, "t.phy"), "rb") as f, gzip.GzipFile(os.path.join(d, "deep.phy.gz"), "wb", mtime=0) as o:
        o.write(f.read())
    json.dump(g, open(os.path.join(d, "deep.json"), "w"), indent=1)
    shutil.rmtree(tmp)


def case_wide(d):
    """100 taxa x 3000 sites WAG+G4 simulated with PEPR-like '?' gene blocks; -f e with the PTHREADS binary (8 threads)"""
    names, seqs, nwk = _sim(100, 3000, 1, missing=0.15)
    tmp = tempfile.mkdtemp()
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    open(os.path.join(tmp, "t.nwk"), "w").write(nwk + "\n")
    run([RAXP, "-T", "8", "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fe", "-t", "t.nwk"], tmp)
    g = {"names": names, "tree_in": nwk, "fe": info(tmp, "fe")}
    g["fe"]["tree"] = open(os.path.join(tmp, "RAxML_result.fe")).read().strip()
    with open(os.path.join(tmp, "t.phy"), "rb") as f, gzip.GzipFile(os.path.join(d, "wide.phy.gz"), "wb", mtime=0) as o:
        o.write(f.read())
    json.dump(g, open(os.path.join(d, "wide.json"), "w"), indent=1)
    shutil.rmtree(tmp)


CASES = {"small": case_small, "dup": case_dup, "deep": case_deep, "wide": case_wide}



def case_search(d):
    """24 taxa x 1200 sites WAG+G4 simulated: raxmlHPC -f d (PEPR's default full-tree call, RAxMLRunner.java:112-132) and
    -f a -x 12345 -N 10 (rapid bootstrap + ML search, the bootstrapReps > 0 branch)"""
    names, seqs, nwk = _sim(24, 1200, 11)
    tmp = tempfile.mkdtemp()
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    run([RAX, "-f", "d", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fd", "-p", "12345"], tmp)
    g = {"names": names, "true_tree": nwk, "fd": info(tmp, "fd")}
    g["fd"]["tree"] = open(os.path.join(tmp, "RAxML_result.fd")).read().strip()
    run([RAX, "-f", "a", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fa", "-x", "12345", "-N", "10", "-p", "12345"], tmp)
    g["fa"] = info(tmp, "fa")
    g["fa"]["bipartitions"] = open(os.path.join(tmp, "RAxML_bipartitions.fa")).read().strip()
    with open(os.path.join(tmp, "t.phy"), "rb") as f, gzip.GzipFile(os.path.join(d, "search.phy.gz"), "wb", mtime=0) as o:
        o.write(f.read())
    json.dump(g, open(os.path.join(d, "search.json"), "w"), indent=1)
    shutil.rmtree(tmp)


CASES["search"] = case_search

FASTTREE = "/root/reference/pepr-bin_linux/FastTree_WAG"


def case_search_fasttree(d):
    """adds FastTree_WAG's trees for the `search` alignment (FastTreeRunner.run: `-gamma -nosupport`, and `-gamma` when
    bootstrapReps > 0; FastTreeRunner.java:66-94) to search.json without touching the raxmlHPC entries"""
    toks = gzip.open(os.path.join(d, "search.phy.gz"), "rt").read().split()
    n = int(toks[0])
    names, seqs = toks[2::2][:n], toks[3::2][:n]
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "s.faa"), "w") as f:
        for a, b in zip(names, seqs):
            f.write(">%s\n%s\n" % (a, b))
    g = json.load(open(os.path.join(d, "search.json")))
    out = {}
    for key, extra in (("tree", ["-nosupport"]), ("tree_support", [])):
        r = subprocess.run([FASTTREE, "-gamma"] + extra + ["s.faa"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        out[key] = r.stdout.strip().splitlines()[0]
        m = re.search(r"Gamma\(20\) LogLk = (\S+) alpha = (\S+)", r.stderr)
        out["gamma20_loglk"], out["alpha"] = float(m.group(1)), float(m.group(2))
    g["fasttree"] = out
    json.dump(g, open(os.path.join(d, "search.json"), "w"), indent=1)
    shutil.rmtree(tmp)


CASES["search_fasttree"] = case_search_fasttree


def case_search_parsimony(d):
    """adds to search.json what RAxMLRunner.runRaxmlParsimonyWithBranchLengths produces (RAxMLRunner.java:215-280):
    `-f d -y` -> RAxML_parsimonyTree, then `-f e -t` on it -> RAxML_result.<run>BL; and FastTree_WAG under a topological
    constraint in the 0/1 alignment form FastTreeRunner writes (FastTreeRunner.java:243-273)"""
    toks = gzip.open(os.path.join(d, "search.phy.gz"), "rt").read().split()
    n = int(toks[0])
    names, seqs = toks[2::2][:n], toks[3::2][:n]
    tmp = tempfile.mkdtemp()
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    run([RAX, "-f", "d", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "py", "-y", "-p", "12345"], tmp)
    ptree = open(os.path.join(tmp, "RAxML_parsimonyTree.py")).read().strip()
    run([RAX, "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "pyBL", "-t", "RAxML_parsimonyTree.py"], tmp)
    g = json.load(open(os.path.join(d, "search.json")))
    g["fy"] = info(tmp, "pyBL")
    g["fy"]["parsimony_tree"] = ptree
    g["fy"]["tree"] = open(os.path.join(tmp, "RAxML_result.pyBL")).read().strip()
    # a constraint that CONTRADICTS the data: the first taxa of two different clades of the -f d tree forced together
    sets = []
    taxa = sorted(orc._leafsets(orc._parse_topology(g["fd"]["tree"]), sets))
    cherries = [sorted(x) for x in sets if len(x) == 2]
    forced = sorted([cherries[0][0], cherries[1][0], cherries[2][0]])
    with open(os.path.join(tmp, "s.faa"), "w") as f:
        for a, b in zip(names, seqs):
            f.write(">%s\n%s\n" % (a, b))
    with open(os.path.join(tmp, "c.txt"), "w") as f:
        for a in names:
            f.write(">%s\n%s\n" % (a, "1" if a in forced else "0"))
    r = subprocess.run([FASTTREE, "-gamma", "-nosupport", "-constraints", "c.txt", "s.faa"], cwd=tmp, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    g["fasttree_constrained"] = {"forced_clade": forced, "tree": r.stdout.strip().splitlines()[0]}
    json.dump(g, open(os.path.join(d, "search.json"), "w"), indent=1)
    shutil.rmtree(tmp)


CASES["search_parsimony"] = case_search_parsimony


if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    for c in which:
        print("golden case", c, flush=True)
        CASES[c](HERE)
