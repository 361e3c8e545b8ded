"""Phase breakdown of the warp-specialised branch kernel: cycles per tile and phase for the warps of CTA 0.
usage: python tools/trace_branch.py [tip|inner]"""
import ctypes as C
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import synth

names, seqs, nwk = synth.simulate_wag(100, 100000, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
ctx = pb.Context(0)
aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
tree = pb.Tree(aln, topo)
tree.evaluate()
L = pb.lib()
L.pml_trace_enable.argtypes = [C.c_void_p, C.c_int]
L.pml_trace_read.argtypes = [C.c_void_p, C.c_void_p]
for want_tip in (True, False):
    L.pml_trace_enable(ctx.h, 3 if want_tip else 2)
    tree.smooth(1)
    out = np.zeros(96, np.int64)
    L.pml_trace_read(ctx.h, out.ctypes.data_as(C.c_void_p))
    L.pml_trace_enable(ctx.h, 0)
    last = out[88:90].copy()
    out = out.reshape(12, 8)
    print("=== branch kernel, %s end; launches %d" % ("tip" if want_tip else "inner", out[8, 6]))
    print("phase: wait_data frags+turn mma contraction wait_slot store (cycles per tile), tiles")
    nl = max(out[8, 6], 1)
    for w in range(8):
        n = max(out[w, 6], 1)
        print("warp %d:" % w, " ".join("%7.0f" % (out[w, k] / n) for k in range(6)), " tiles/launch %.1f" % (out[w, 6] / nl),
              " sum %.0f" % (out[w, :6].sum() / n), " prologue/launch %.0f" % (out[w, 7] / nl), " loop/launch %.0f" % (out[w, :6].sum() / nl))
    for w in (8, 9):
        nt = max(out[w, 4], 1)
        print("finishing warp %d: kernel cycles/launch %.0f; per pair: wait sums %.0f, finish %.0f; pairs/launch %.1f; entry->loop %.0f; partial+ticket %.0f"
              % (w, out[w, 0] / nl, out[w, 1] / nt, out[w, 2] / nt, out[w, 4] / nl, out[w, 7] / nl, out[w, 3] / nl))
    print("last CTA: final sum + publication %.0f cycles per launch (%d launches)" % (last[0] / max(last[1], 1), last[1]))
