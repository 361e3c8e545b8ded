"""Where the warps of a streaming kernel spend their time, from the warp-state samples of an `ncu --set full --import-source on`
capture: samples between consecutive named-barrier instructions (the MMA turns), the share of DMMA + their issue slots, the
stall reasons per region.  usage: python tools/ncu_stalls.py capture.ncu-rep [kernel-substring] > profiles/...txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "k_fused<(int)1, (int)0, (int)1>"
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1])
hdr = rows[start + 1]
body = []
for r in rows[start + 2:]:
    if r and r[0] == "Kernel Name":
        break
    body.append(r)
si, src = hdr.index("# Samples"), hdr.index("Source")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("# %s: %s, %d SASS instructions, %d samples" % (rep, rows[start][1], len(body), sum(int(r[si]) for r in body)))
marks = [i for i, r in enumerate(body) if "BAR." in r[src] or "EXIT" in r[src]]
print("# regions between barrier / exit instructions (index: instruction): samples, DMMA count, samples on DMMA + following NOP, top stall reasons")
prev = 0
for m in marks + [len(body)]:
    seg = body[prev:m]
    n = sum(int(r[si]) for r in seg)
    if n >= 20:
        dm = sum(1 for r in seg if "DMMA" in r[src])
        dms = sum(int(r[si]) for j, r in enumerate(seg) if "DMMA" in r[src] or ("NOP" in r[src] and j and "DMMA" in seg[j - 1][src]))
        agg = collections.Counter()
        for r in seg:
            for c in stalls:
                agg[hdr[c][6:]] += int(r[c])
        print("%5d .. %5d  %-44s samples %5d  DMMA %3d  on DMMA+NOP %5d  %s" % (prev, m, body[prev][src].strip()[:44], n, dm, dms,
                                                                                 ", ".join("%s %d" % kv for kv in agg.most_common(4))))
    prev = m
