"""Phase breakdown of the fused update + branch kernel (inner children, inner far end): cycles per tile for CTA 0."""
import ctypes as C, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import synth
names, seqs, nwk = synth.simulate_wag(100, 100000, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
ctx = pb.Context(0)
aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
tree = pb.Tree(aln, topo)
tree.smooth(1)
L = pb.lib()
L.pml_trace_enable(ctx.h, 6)
tree.smooth(2)
out = np.zeros(96, np.int64)
L.pml_trace_read(ctx.h, out.ctypes.data_as(C.c_void_p))
out = out.reshape(12, 8)
print("NV warps: wait_data frags+turn mma products wait_slot store | BR warps: wait_prod wait_far frags wait_turn mma post")
for w in range(8):
    n = max(out[w, 6], 1)
    print("warp %d:" % w, " ".join("%7.0f" % (out[w, k] / n) for k in range(6)), " tiles", out[w, 6], " sum %.0f" % (out[w, :6].sum() / n))
