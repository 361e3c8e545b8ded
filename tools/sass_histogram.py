"""Per-kernel SASS opcode histogram of pepr_b200/libpeprml.so (evidence for DESIGN.md section 4: which pipes and copy engines the
kernels use).  usage: python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pepr_b200", "libpeprml.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
WATCH = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ACQBULK", "UTCMMA", "UTCHMMA", "LDTM", "HMMA", "LDS", "STS",
         "LDG", "STG", "BAR", "SHFL", "MUFU", "ATOM", "RED", "ELECT", "FENCE", "MEMBAR"]
kern, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|pml::|void ", "", kern)
        kern = re.sub(r"\((?:int|bool)\)", "", kern)
        kern = re.sub(r"\(.*$", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", line)
    if m and kern:
        op = m.group(1)
        key = op + (m.group(2) if op in ("DMMA", "SYNCS", "UBLKCP", "FENCE") else "")
        counts[kern][key] += 1
print("# cuobjdump -sass %s ; architectures: %s" % (os.path.relpath(so, ROOT), ", ".join(sorted(arch))))
print("# opcode counts per kernel (static instructions); watched families: " + " ".join(WATCH))
for k, c in counts.items():
    total = sum(c.values())
    picks = [(op, n) for op, n in sorted(c.items()) if any(op.startswith(w) for w in WATCH)]
    print("%-40s %6d instr | %s" % (k, total, "  ".join("%s %d" % p for p in picks)))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# library totals: " + "  ".join("%s %d" % (op, n) for op, n in sorted(tot.items()) if any(op.startswith(w) for w in
      ("DMMA", "UBLKCP", "SYNCS", "UTC", "LDTM", "UTMA", "HMMA", "ACQBULK"))))
