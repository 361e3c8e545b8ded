"""Real-data golden: a supermatrix of the reference's own example genomes (examples/Aquificales, 11 ingroup genomes + the
outgroup) scored by the bundled raxmlHPC (-f d / -f e / -f g, PROTGAMMAWAG).
PEPR builds its supermatrix with blat + mcl + muscle + Gblocks; the bundled muscle and Gblocks are 32-bit executables that
do not start in this image (no 32-bit loader) and the Java pipeline needs a JDK, so the matrix is built from what needs
neither: single-copy families by PATRIC product annotation (exactly one protein of that name in every genome); every
protein is aligned globally (Needleman-Wunsch, the reference's BLOSUM62 file, linear gap cost) to the family's protein of
the first genome, columns are that protein's positions (insertions relative to it are dropped, deletions become '-'), and
columns with more than half gaps are trimmed (the role Gblocks plays, MSATrimmer.java:94-102).  A star alignment is cruder
than muscle, but the point here is real data for the likelihood engine: unequal residue composition, indel-derived gaps,
strong among-site rate variation, invariant columns, duplicate patterns.
Writes tests/golden/<example>.phy.gz and <example>.json (aquificales, erysipelotrichales).  Needs /root/reference (run once; the outputs are committed)."""
import glob
import gzip
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import RAX, info, run  # noqa: E402

REF = "/root/reference"
NFAM = 40


def read_fasta(path):
    out, name, seq = [], None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(seq)))
            name, seq = line[1:], []
        elif line:
            seq.append(line)
    if name is not None:
        out.append((name, "".join(seq)))
    return out


def load_blosum(path):
    rows = [l.split() for l in open(path) if l.strip() and not l.startswith("#")]
    cols = rows[0]
    M = {}
    for r in rows[1:]:
        for c, v in zip(cols, r[1:]):
            M[(r[0], c)] = int(v)
    return M


def star_align(center, seq, M, gap=6):
    """global alignment of seq onto center's coordinates: returns a string of len(center) (residue or '-')"""
    import numpy as np
    n, m = len(center), len(seq)
    S = np.array([[M.get((a, b), -4) for b in seq] for a in center], dtype=np.int32)
    H = np.zeros((n + 1, m + 1), np.int32)
    H[0, :] = -gap * np.arange(m + 1)
    H[:, 0] = -gap * np.arange(n + 1)
    ar = gap * np.arange(m + 1)
    for i in range(1, n + 1):
        T = np.empty(m + 1, np.int64)
        T[0] = H[i, 0]
        T[1:] = np.maximum(H[i - 1, :-1] + S[i - 1], H[i - 1, 1:] - gap)
        H[i] = np.maximum.accumulate(T + ar) - ar          # left moves as a max-plus prefix scan
    out, i, j = [], n, m
    while i > 0:
        if j > 0 and H[i, j] == H[i - 1, j - 1] + S[i - 1, j - 1]:
            out.append(seq[j - 1]); i -= 1; j -= 1
        elif H[i, j] == H[i - 1, j] - gap:
            out.append("-"); i -= 1
        else:
            j -= 1                                           # insertion relative to the centre: dropped
    return "".join(reversed(out))


def main(example="Aquificales"):
    files = sorted(glob.glob(os.path.join(REF, "examples", example, "*.faa"))) + \
        sorted(glob.glob(os.path.join(REF, "examples", example, "outgroup", "*.faa")))
    genomes = []
    for f in files:
        taxon = re.sub(r"\W+", "_", os.path.basename(f).replace(".PATRIC.faa", ""))[:40]
        fam = {}
        for title, seq in read_fasta(f):
            parts = title.split("|")
            product = parts[4].strip().split("[")[0].strip() if len(parts) > 4 else ""
            if not product or "hypothetical" in product.lower() or len(seq) < 80:
                continue
            seq = seq.replace("*", "")
            if seq not in fam.setdefault(product, []):   # one example file lists every protein twice
                fam[product].append(seq)
        genomes.append((taxon, fam))
    shared = set(genomes[0][1])
    for _, fam in genomes:
        shared &= {k for k, v in fam.items() if len(v) == 1}
    print("genomes %d, single-copy annotated families %d" % (len(genomes), len(shared)))
    tmp = tempfile.mkdtemp()
    M = load_blosum(os.path.join(REF, "BLOSUM62"))
    blocks = [[] for _ in genomes]
    used = []
    for famname in sorted(shared):
        rows = [fam[famname][0].upper() for _, fam in genomes]
        if not (100 <= len(rows[0]) <= 600) or max(len(r) for r in rows) > 1.3 * min(len(r) for r in rows):
            continue
        aligned = [rows[0]] + [star_align(rows[0], r, M) for r in rows[1:]]
        ident = min(sum(a == b for a, b in zip(rows[0], r)) / len(rows[0]) for r in aligned)
        if ident < 0.30:
            continue
        keep = [k for k in range(len(rows[0])) if sum(r[k] == "-" for r in aligned) * 2 <= len(aligned)]
        for i, r in enumerate(aligned):
            blocks[i].append("".join(r[k] for k in keep))
        used.append(famname)
        if len(used) == NFAM:
            break
    names = [g[0] for g in genomes]
    seqs = ["".join(b) for b in blocks]
    print("families kept %d, supermatrix %d x %d" % (len(used), len(names), len(seqs[0])))
    from pepr_b200 import synth
    synth.write_phylip(os.path.join(tmp, "t.phy"), names, seqs)
    run([RAX, "-f", "d", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fd", "-p", "12345"], tmp)
    g = {"names": names, "families": used, "fd": info(tmp, "fd")}
    g["fd"]["tree"] = open(os.path.join(tmp, "RAxML_result.fd")).read().strip()
    shutil.copy(os.path.join(tmp, "RAxML_result.fd"), os.path.join(tmp, "best.nwk"))
    run([RAX, "-f", "e", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fe", "-t", "best.nwk"], tmp)
    g["fe"] = info(tmp, "fe")
    g["fe"]["tree"] = open(os.path.join(tmp, "RAxML_result.fe")).read().strip()
    shutil.copy(os.path.join(tmp, "RAxML_result.fe"), os.path.join(tmp, "fe.nwk"))
    run([RAX, "-f", "g", "-m", "PROTGAMMAWAG", "-s", "t.phy", "-n", "fg", "-z", "fe.nwk"], tmp)
    lines = open(os.path.join(tmp, "RAxML_perSiteLLs.fg")).read().split("\n")
    g["fg"] = {"per_site": [float(x) for x in lines[1].split("\t")[1].split()]}
    g["tree_in"] = re.sub(r":[0-9.eE+-]+", "", g["fd"]["tree"])
    with open(os.path.join(tmp, "t.phy"), "rb") as f, gzip.GzipFile(os.path.join(HERE, example.lower() + ".phy.gz"), "wb", mtime=0) as o:
        o.write(f.read())
    json.dump(g, open(os.path.join(HERE, example.lower() + ".json"), "w"), indent=1)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    for ex in sys.argv[1:] or ["Aquificales", "Erysipelotrichales"]:
        main(ex)
