"""Phase breakdown of the warp-specialised CLV kernel: cycles per tile and phase for the MMA warps of CTA 0."""
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
L.pml_trace_enable(ctx.h, int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for _ in range(3):
    tree.invalidate()
    tree.evaluate()
out = np.zeros(96, np.int64)
L.pml_trace_read(ctx.h, out.ctypes.data_as(C.c_void_p))
print("prologue probes (cycles after pdl_wait, per launch): init sync %.0f, model in smem %.0f, P built %.0f" % tuple(out[90:93] / max(out[70], 1)))
out = out.reshape(12, 8)
print("phase: wait_data frags+turn mma products wait_slot store  (cycles per tile), tiles")
for w in range(8):
    n = max(out[w, 6], 1)
    print("warp %d:" % w, " ".join("%7.0f" % (out[w, k] / n) for k in range(6)), " tiles", out[w, 6], " sum %.0f" % (out[w, :6].sum() / n),
          " prologue/launch %.0f" % (out[w, 7] / max(out[8, 6], 1)), " loop/launch %.0f" % (out[w, :6].sum() / max(out[8, 6], 1)))
for w in (8, 9):
    nt = max(out[w, 4], 1)
    print("epilogue warp %d: kernel cycles per launch %.0f over %d launches; per tile: wait products %.0f, test+store issue %.0f, wait prev store %.0f; final drain per launch %.0f; entry->loop %.0f; tiles/launch %.1f"
          % (w, out[w, 0] / max(out[w, 6], 1), out[w, 6], out[w, 1] / nt, out[w, 2] / nt, out[w, 3] / nt, out[w, 5] / max(out[w, 6], 1),
             out[w, 7] / max(out[w, 6], 1), out[w, 4] / max(out[w, 6], 1)))
