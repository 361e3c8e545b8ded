"""Event-timed smoothing sweeps with and without the fused kernel, and with the finishing arithmetic switched off
(want flags) to see what the non-MMA FP64 work costs."""
import os, re, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pepr_b200 as pb
from pepr_b200 import synth
names, seqs, nwk = synth.simulate_wag(100, 100000, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
for mode in ("fused", "two"):
    if mode == "two":
        os.environ["PEPRML_NO_FUSE"] = "1"
    ctx = pb.Context(0)
    aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
    tree = pb.Tree(aln, topo)
    tree.smooth(2)
    ctx.timer_start()
    tree.smooth(5)
    ms = ctx.timer_stop()
    print("%s: %.2f ms per sweep" % (mode, ms / 5))
    ctx.profile_begin()
    tree.smooth(3)
    prof = ctx.profile_end()
    print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in prof.items() if v[1]})
    tree.close(); aln.close(); ctx.close()
