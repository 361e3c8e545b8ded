"""Profiling aid: where a CLV launch's prologue goes, from two extra %globaltimer stamps (PML_TL_PROBE_A / _B builds of the engine,
see newview_mma.cu).  usage: PEPRML_LIB=... python tools/prologue_stamps.py A B   (A, B = the probe points of that build)"""
import ctypes as C, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import engine, synth
names, seqs, nwk = synth.simulate_wag(100, 100000, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
ctx = pb.Context(0)
aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
tree = pb.Tree(aln, topo)
tree.smooth(2)
L = pb.lib()
L.pml_timeline_begin.argtypes = [C.c_void_p]
L.pml_timeline_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
kind_names = [k[0] for k in engine.kinds()]
L.pml_timeline_begin(ctx.h)
for _ in range(3):
    tree.invalidate()
    tree.evaluate()
tree.smooth(1)
cap = 16384
st = np.zeros((cap, 6), np.uint64)
kd = np.zeros(cap, np.int32)
n = L.pml_timeline_read(ctx.h, st.ctypes.data_as(C.c_void_p), kd.ctypes.data_as(C.c_void_p), cap)
st, kd = st[:n].astype(np.int64), kd[:n]
a, b = sys.argv[1], sys.argv[2]
for k in sorted(set(kd.tolist())):
    if not kind_names[k].startswith("newview"):
        continue
    sel = (kd == k) & (st[:, 4] > 0)
    if sel.any():
        print("%-24s %4d launches: wait returns -> point %s %.2f us, -> point %s %.2f us, -> first turn %.2f us" % (
            kind_names[k], int(sel.sum()), a, ((st[sel, 4] - st[sel, 1]) / 1e3).mean(), b, ((st[sel, 5] - st[sel, 1]) / 1e3).mean(),
            ((st[sel, 2] - st[sel, 1]) / 1e3).mean()))
