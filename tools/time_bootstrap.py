"""Where a bootstrap-replicate tree spends its time (bench workload): parsimony start tree, coarse optimisation, SPR round."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pepr_b200 as pb
from pepr_b200 import synth

sites = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
names, seqs, nwk = synth.simulate_wag(100, sites, 3)
ctx = pb.Context(0)
aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
W, _ = aln.bootstrap_weights(12345, 2)
for r in range(2):
    t0 = time.perf_counter()
    bt = pb.Tree(aln, parsimony_seed=12346 + r, weights=W[r])
    ctx.sync()
    t1 = time.perf_counter()
    l0, a0 = bt.optimize(False, 5.0, weights=W[r])
    t2 = time.perf_counter()
    bl, bm = bt.search(radius=5, max_rounds=1, eps=0.1, weights=W[r])
    t3 = time.perf_counter()
    su, ln = bt.stats()
    print("replicate %d: parsimony %.3f s, coarse branch lengths %.3f s (lnL %.1f), SPR round %.3f s (lnL %.1f, %d moves); %d launches, %.2e site-updates"
          % (r, t1 - t0, t2 - t1, l0, t3 - t2, bl, bm, ln, sum(su)))
    bt.close()
