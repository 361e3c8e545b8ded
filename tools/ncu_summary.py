"""Summaries of ncu output for profiles/: a launch list (CSV of gpu__time_duration.sum) or a --set full report.
usage: python tools/ncu_summary.py launches <csv> ["header comment"]
       python tools/ncu_summary.py full <file.ncu-rep> ["header comment"]"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"void |pml::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", name)
    name = re.sub(r"\((?:bool|int)\)", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def launches(path, note):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) <= vi or "gpu__time_duration" not in ",".join(r):
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        k = tot[short(r[ki])]
        k[0] += 1
        k[1] += v
    total = sum(v[1] for v in tot.values())
    print("# " + note)
    print("%-34s %9s %12s %10s %8s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-34s %9d %12.1f %10.2f %7.1f%%" % (k, n, us, us / n, 100 * us / total))


FULL = [("time", "gpu__time_duration.sum"), ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
        ("dram_%peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dmma_pipe_%", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fp64_pipe_%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm_active_cyc", "sm__cycles_active.avg"), ("sm_elapsed_cyc", "sm__cycles_elapsed.avg"),
        ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("grid", "launch__grid_size"), ("smem_KB", "launch__shared_mem_per_block_dynamic"),
        ("bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu.sum")]


def full(path, note):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# " + note)
    print("%-30s" % "kernel" + "".join("%15s" % n for n, _ in FULL))
    seen = collections.Counter()
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        seen[name] += 1
        if seen[name] > 2:
            continue
        cells = []
        for n, m in FULL:
            if m in hdr:
                i = hdr.index(m)
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    v = "%.2f" % f if f < 1e6 else "%.4g" % f
                except ValueError:
                    pass
                cells.append("%15s" % (v + (units[i] if units[i] in ("us", "Mbyte", "Kbyte", "byte", "ns") else "")))
            else:
                cells.append("%15s" % "-")
        print("%-30s" % name + "".join(cells))


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else path
    (launches if mode == "launches" else full)(path, note)
