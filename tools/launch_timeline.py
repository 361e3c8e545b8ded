"""Where the time between the steady states of consecutive launches goes: %globaltimer stamps of every CLV / fused launch of two
smoothing sweeps and of three likelihood passes on the bench workload (pml_timeline_begin / pml_timeline_read).
usage: python tools/launch_timeline.py > profiles/r02_launch_timeline.txt"""
import ctypes as C, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import engine, synth

names, seqs, nwk = synth.simulate_wag(100, int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 3)
topo = re.sub(r":[0-9.eE+-]+", "", nwk)
ctx = pb.Context(0)
aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
tree = pb.Tree(aln, topo)
tree.smooth(2)
L = pb.lib()
L.pml_timeline_begin.argtypes = [C.c_void_p]
L.pml_timeline_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
kind_names = [k[0] for k in engine.kinds()]


def capture(fn, title):
    L.pml_timeline_begin(ctx.h)
    fn()
    cap = 16384
    st = np.zeros((cap, 6), np.uint64)
    kd = np.zeros(cap, np.int32)
    n = L.pml_timeline_read(ctx.h, st.ctypes.data_as(C.c_void_p), kd.ctypes.data_as(C.c_void_p), cap)
    st, kd = st[:n].astype(np.int64), kd[:n]
    fused = st[:, 5] > 0
    print("# %s: %d launches (%d fused); patterns %d" % (title, n, int(fused.sum()), aln.npatterns))
    end = np.where(fused, st[:, 5], st[:, 3])              # what the launch ends with as seen by the stamps
    gap_prev = st[1:, 1] - end[:-1]                         # predecessor's last stamp -> this launch's dependency wait returns
    rows = [("entry -> dependency wait returns (waits for the predecessor)", st[:, 1] - st[:, 0]),
            ("wait returns -> first MMA turn (prologue)", st[:, 2] - st[:, 1]),
            ("first MMA turn -> CTA 0's last tile (steady state)", st[:, 3] - st[:, 2])]
    print("%-78s %9s %9s %9s" % ("all launches, microseconds", "mean", "median", "p90"))
    for name, v in rows:
        v = v / 1e3
        print("%-78s %9.2f %9.2f %9.2f" % (name, v.mean(), np.median(v), np.percentile(v, 90)))
    if fused.any():
        for name, v in (("fused: CTA 0's last tile -> last CTA's ticket (drain + imbalance)", st[fused, 4] - st[fused, 3]),
                        ("fused: ticket -> published (sum over CTAs, NR step, publication)", st[fused, 5] - st[fused, 4])):
            v = v / 1e3
            print("%-78s %9.2f %9.2f %9.2f" % (name, v.mean(), np.median(v), np.percentile(v, 90)))
    for label, sel in (("after a fused launch", fused[:-1]), ("after a CLV launch", ~fused[:-1])):
        if sel.any():
            v = gap_prev[sel] / 1e3
            print("%-78s %9.2f %9.2f %9.2f" % ("last stamp of the predecessor -> wait returns, " + label, v.mean(), np.median(v), np.percentile(v, 90)))
    total = (end[-1] - st[0, 1]) / 1e3
    steady = ((st[:, 3] - st[:, 2]) / 1e3).sum()
    print("capture: %.1f us from the first wait return to the last stamp; steady states add up to %.1f us (%.0f %%)" % (total, steady, 100 * steady / total))
    for k in sorted(set(kd.tolist())):
        sel = kd == k
        print("  %-28s %5d launches: prologue %.2f us, steady %.2f us" % (kind_names[k], int(sel.sum()), ((st[sel, 2] - st[sel, 1]) / 1e3).mean(),
                                                                        ((st[sel, 3] - st[sel, 2]) / 1e3).mean()))
    print()


def passes():
    for _ in range(3):
        tree.invalidate()
        tree.evaluate()


capture(lambda: tree.smooth(2), "two smoothing sweeps")
capture(passes, "three likelihood passes")
