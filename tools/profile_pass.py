"""Short driver for ncu: a few full likelihood passes (98 newview launches + evaluate) and a few NR branch updates on
the bench workload.  usage: python tools/profile_pass.py [sites] [passes]"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pepr_b200 as pb
from pepr_b200 import synth

sites = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
names, seqs, nwk = synth.simulate_wag(100, sites, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
ctx = pb.Context(0)
aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
tree = pb.Tree(aln, topo)
for _ in range(passes):
    tree.invalidate()
    lnl = tree.evaluate()
for e in range(0, 40, 5):
    tree.branch_derivs(e, 0.1)
tree.smooth(1)
print("lnL", lnl, tree.stats())
