"""Experiment aid: the same short measurement on several builds of the engine (pepr_b200.build.build_variant), one process per
build.  Per build: ms per smoothing sweep, ms per likelihood pass, lnL after the sweeps (must be the same bits on every build
that only changes scheduling), event-timed us per launch and kind.
usage: python tools/variant_bench.py [sites] name=path/to/lib.so ...      (name "base" = the product library)"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(sites):
    import pepr_b200 as pb
    from pepr_b200 import synth
    names, seqs, nwk = synth.simulate_wag(100, sites, 3)
    topo = re.sub(r":[0-9.eE+-]+", "", nwk)
    ctx = pb.Context(0)
    aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
    tree = pb.Tree(aln, topo)
    tree.smooth(2)
    best = {"sweep_ms": 1e30, "sweep_us_per_launch": 1e30, "pass_ms": 1e30}
    for _ in range(3):
        n0 = tree.stats()[1]
        ctx.timer_start()
        tree.smooth(4)
        ms = ctx.timer_stop()
        best["sweep_ms"] = min(best["sweep_ms"], ms / 4)
        best["sweep_us_per_launch"] = min(best["sweep_us_per_launch"], ms * 1e3 / max(tree.stats()[1] - n0, 1))
    lnl = tree.evaluate()
    for _ in range(3):
        ctx.timer_start()
        for _ in range(5):
            tree.invalidate()
            tree.evaluate()
        best["pass_ms"] = min(best["pass_ms"], ctx.timer_stop() / 5)
    ctx.profile_begin()
    tree.smooth(2)
    for _ in range(2):
        tree.invalidate()
        tree.evaluate()
    prof = ctx.profile_end()
    best["lnl"] = repr(lnl)
    best["us_per_launch"] = {k: round(v[0] / v[1] * 1e3, 2) for k, v in prof.items() if v[1]}
    print("RESULT " + json.dumps(best))


if __name__ == "__main__":
    args = sys.argv[1:]
    sites = 100000
    if args and args[0].isdigit():
        sites = int(args.pop(0))
    if not args:
        one(sites)
        sys.exit(0)
    results = {}
    for spec in args:
        name, _, path = spec.partition("=")
        env = dict(os.environ)
        if path:
            env["PEPRML_LIB"] = os.path.join(ROOT, path)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), str(sites)], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(name, "FAILED", out.stdout[-2000:], out.stderr[-2000:])
            continue
        results[name] = json.loads(line[0][7:])
    names = list(results)
    print("%-28s" % "" + "".join("%12s" % n for n in names))
    for key in ("sweep_ms", "sweep_us_per_launch", "pass_ms"):
        print("%-28s" % key + "".join("%12.3f" % results[n][key] for n in names))
    print("%-28s" % "lnl" + "".join("%12s" % ("same" if results[n]["lnl"] == results[names[0]]["lnl"] else results[n]["lnl"]) for n in names))
    for k in sorted({k for n in names for k in results[n]["us_per_launch"]}):
        print("%-28s" % k + "".join("%12.2f" % results[n]["us_per_launch"].get(k, float("nan")) for n in names))
    print("JSON " + json.dumps(results))
