import os, re, sys
sys.path.insert(0, "/root/repo")
import pepr_b200 as pb
from pepr_b200 import synth
names, seqs, nwk = synth.simulate_wag(100, 100000, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
ctx = pb.Context(0); aln = pb.Alignment(ctx, names, seqs, alpha=1.0); tree = pb.Tree(aln, topo)
print(tree.optimize(True, 0.1))
