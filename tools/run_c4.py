"""BASELINE configs[3] at full size on ONE B200: synthetic 500 taxa x 250k sites WAG+G4 (79.7 GB of CLVs), likelihood pass,
`-f e`, 100 bootstrap replicates as integer weight vectors (lnL of every replicate in one pass), one replicate tree.
usage: python tools/run_c4.py [ntax] [sites]"""
import json
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import synth

ntax = int(sys.argv[1]) if len(sys.argv) > 1 else 500
sites = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
out = {"ntax": ntax, "sites": sites}
t0 = time.perf_counter()
names, seqs, nwk = synth.simulate_wag(ntax, sites, 2)
out["simulate_s"] = time.perf_counter() - t0
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
chars = np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
ctx = pb.Context(0)
t0 = time.perf_counter()
aln = pb.Alignment(ctx, names, chars, alpha=1.0)
tree = pb.Tree(aln, nwk)
out["load_s"] = time.perf_counter() - t0
out["patterns"] = aln.npatterns
out["clv_GB"] = (ntax - 2) * aln.npatterns * 640 / 1e9
t0 = time.perf_counter()
out["lnl_true_lengths"] = tree.evaluate()
out["first_pass_s"] = time.perf_counter() - t0
for e in range(tree.num_branches):   # `-f e` starts from default lengths
    tree.set_branch(e, 0.1)
ctx.timer_start()
for _ in range(3):
    tree.invalidate()
    tree.evaluate()
ms = ctx.timer_stop() / 3
out["likelihood_pass_ms"] = ms
out["likelihood_pass_site_updates_per_s"] = (ntax - 2) * aln.npatterns / (ms * 1e-3)
su0, ln0 = tree.stats()
ctx.timer_start()
t0 = time.perf_counter()
lnl, alpha = tree.optimize(True, 0.1)
out["fe_device_ms"] = ctx.timer_stop()
out["fe_wall_s"] = time.perf_counter() - t0
su1, ln1 = tree.stats()
out["fe_lnl"], out["fe_alpha"] = lnl, alpha
out["fe_site_updates"] = int(sum(su1) - sum(su0))
out["fe_site_updates_per_s"] = out["fe_site_updates"] / (out["fe_device_ms"] * 1e-3)
out["fe_launches"] = int(ln1 - ln0)
assert lnl >= out["lnl_true_lengths"] - 0.1   # the optimum cannot be worse than the generating parameters
t0 = time.perf_counter()
W, _ = aln.bootstrap_weights(12345, 100)
out["bootstrap_weights_100_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
rl = tree.evaluate_replicates(W)
out["replicate_lnl_100_s"] = time.perf_counter() - t0
# size-independent properties: weights of every replicate sum to the number of sites; a replicate's lnL is the weighted sum
assert (W.sum(axis=1) == sites).all()
l_w0 = tree.evaluate(weights=W[0]) if "weights" in tree.evaluate.__code__.co_varnames else None
if l_w0 is not None:
    assert abs(l_w0 - rl[0]) <= 1e-9 * abs(rl[0]), (l_w0, rl[0])
out["replicate_lnl_mean"] = float(rl.mean())
t0 = time.perf_counter()
bt = pb.Tree(aln, parsimony_seed=12346, weights=W[0])
ctx.sync()
out["replicate_parsimony_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
bt.optimize(False, 5.0, weights=W[0])
out["replicate_branch_lengths_s"] = time.perf_counter() - t0
if len(sys.argv) <= 3:
    t0 = time.perf_counter()
    bl, bm = bt.search(radius=5, max_rounds=1, eps=0.1, weights=W[0])
    out["replicate_spr_round_s"] = time.perf_counter() - t0
    out["replicate_lnl"], out["replicate_moves"] = bl, bm
bt.close()
print(json.dumps(out, indent=1))
