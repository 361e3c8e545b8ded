"""One GPU's share of BASELINE configs[4] (2000 taxa x 1M sites over 8 GPUs = 125k sites per GPU; default here 2000 x 100k so
that the 128 GB CLV arena leaves headroom on a 180 GB B200): likelihood pass, one smoothing sweep, lazy-SPR candidate
scoring sweep.  usage: python tools/run_c5_shard.py [ntax] [sites] [prune nodes]"""
import json
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import synth

ntax = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sites = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
nprune = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = {"ntax": ntax, "sites": sites}
t0 = time.perf_counter()
names, seqs, nwk = synth.simulate_wag(ntax, sites, 1)
out["simulate_s"] = time.perf_counter() - t0
chars = np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
ctx = pb.Context(0)
t0 = time.perf_counter()
aln = pb.Alignment(ctx, names, chars, alpha=1.0)
tree = pb.Tree(aln, nwk)
out["load_s"] = time.perf_counter() - t0
out["patterns"] = aln.npatterns
out["clv_GB"] = (ntax - 2) * aln.npatterns * 640 / 1e9
t0 = time.perf_counter()
lnl = tree.evaluate()
out["first_pass_s"] = time.perf_counter() - t0
out["lnl_true_tree"] = lnl
ctx.timer_start()
tree.invalidate()
l2 = tree.evaluate()
ms = ctx.timer_stop()
assert l2 == lnl
out["likelihood_pass_ms"] = ms
out["likelihood_pass_site_updates_per_s"] = (ntax - 2) * aln.npatterns / (ms * 1e-3)
su0, ln0 = tree.stats()
t0 = time.perf_counter()
tree.smooth(1)
out["smoothing_sweep_s"] = time.perf_counter() - t0
su1, ln1 = tree.stats()
out["sweep_site_updates_per_s"] = (sum(su1) - sum(su0)) / out["smoothing_sweep_s"]
out["lnl_after_sweep"] = tree.evaluate()
assert out["lnl_after_sweep"] >= lnl - 1e-6 * abs(lnl)
# lazy-SPR candidate scoring sweep: every candidate = 1-3 CLV updates + one branch pass over all patterns
rng = np.random.default_rng(5)
ncand, t0 = 0, time.perf_counter()
su1, ln1 = tree.stats()
best_gain = -1e300
for _ in range(nprune):
    node = int(rng.integers(ntax, 2 * ntax - 2))
    keep = tree.neighbors(node)[int(rng.integers(0, 3))]
    targets, scores = tree.score_spr_candidates(node, keep, radius=5)
    ncand += len(targets)
    if len(scores):
        best_gain = max(best_gain, float(scores.max()) - out["lnl_after_sweep"])
out["spr_sweep_s"] = time.perf_counter() - t0
su2, ln2 = tree.stats()
out["spr_candidates"] = ncand
out["spr_candidates_per_s"] = ncand / out["spr_sweep_s"]
out["spr_site_updates_per_s"] = (sum(su2) - sum(su1)) / out["spr_sweep_s"]
out["spr_launches"] = int(ln2 - ln1)
out["best_candidate_minus_current_lnl"] = best_gain      # the true topology: no candidate should beat it by much
print(json.dumps(out, indent=1))
