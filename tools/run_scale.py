"""BASELINE configs[4] and configs[3] at FULL size on the N GPUs of one box, driven from ONE process (pml_group_create / one
context per GPU, a host thread each -- the shape PEPR's JVM has).

  c5  synthetic 2000 taxa x 1 M sites, site-sharded over the N GPUs (125 k sites, ~150 GB of CLVs per GPU at N = 8):
      likelihood pass, one smoothing sweep (3,997 guarded NR steps), lazy-SPR candidate scoring sweep; checks: lnL bit-identical
      on all ranks, sum over ranks of the per-site lnL of the rank's own patterns = the group's lnL, lnL does not drop in the
      sweep, and (after the group is gone) the first columns alone against the CPU oracle when it is available.
  c4  synthetic 500 taxa x 250 k sites, bootstrap replicate trees sharded by REPLICATE (replicate r on GPU r mod N, the whole
      pattern set and an 80 GB arena on every GPU, no collective): one replicate per GPU on 1, 2, 4, .. N GPUs -> wall seconds per
      step; 100 replicates take ceil(100 / N) such steps.

usage: python tools/run_scale.py --gpus 8 [--skip-c5] [--skip-c4] [--c5-taxa 2000 --c5-sites 1000000] [--c4-taxa 500 --c4-sites 250000]
The alignments are drawn on cuda:0 (synth.simulate_wag_device)."""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pepr_b200 as pb
from pepr_b200 import synth


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def run_c5(args):
    n = args.gpus
    out = {"config": "synthetic %d taxa x %d sites WAG+G4, site-sharded over %d GPU(s), one process" % (args.c5_taxa, args.c5_sites, n),
           "gpus": n}
    t0 = time.perf_counter()
    names, chars, nwk = synth.simulate_wag_device(args.c5_taxa, args.c5_sites, 1)
    out["simulate_s"] = time.perf_counter() - t0
    log("c5: alignment drawn in %.1f s" % out["simulate_s"])
    import torch
    torch.cuda.empty_cache()
    grp = pb.Group(list(range(n)))
    out["collective"] = grp.contexts[0].collective
    ntax = args.c5_taxa

    def work(ctx):
        r = {}
        t0 = time.perf_counter()
        aln = pb.Alignment(ctx, names, chars, alpha=1.0)
        tree = pb.Tree(aln, nwk)
        r["load_s"] = time.perf_counter() - t0
        r["patterns"], r["patterns_local"] = aln.npatterns, aln.npatterns_local
        t0 = time.perf_counter()
        lnl, ps = tree.evaluate(per_site=True)
        r["first_pass_s"] = time.perf_counter() - t0
        r["lnl"], r["own_per_site_sum"] = lnl, float(np.sum(ps))
        del ps
        ctx.timer_start()
        for _ in range(3):
            tree.invalidate()
            l2 = tree.evaluate()
        r["pass_ms"] = ctx.timer_stop() / 3
        assert l2 == lnl
        su0, ln0 = tree.stats()
        ctx.timer_start()
        t0 = time.perf_counter()
        tree.smooth(1)
        r["sweep_ms"] = ctx.timer_stop()
        r["sweep_wall_s"] = time.perf_counter() - t0
        su1, ln1 = tree.stats()
        r["sweep_site_updates"], r["sweep_launches"] = int(sum(su1) - sum(su0)), int(ln1 - ln0)
        r["lnl_after_sweep"] = tree.evaluate()
        rng = np.random.default_rng(5)
        ncand, best = 0, -1e300
        su1, ln1 = tree.stats()
        ctx.timer_start()
        t0 = time.perf_counter()
        for _ in range(args.c5_prune):
            node = int(rng.integers(ntax, 2 * ntax - 2))
            keep = tree.neighbors(node)[int(rng.integers(0, 3))]
            targets, scores = tree.score_spr_candidates(node, keep, radius=5)
            ncand += len(targets)
            if len(scores):
                best = max(best, float(scores.max()) - r["lnl_after_sweep"])
        r["spr_ms"] = ctx.timer_stop()
        r["spr_wall_s"] = time.perf_counter() - t0
        su2, ln2 = tree.stats()
        r["spr_candidates"], r["spr_site_updates"], r["spr_launches"] = ncand, int(sum(su2) - sum(su1)), int(ln2 - ln1)
        r["best_candidate_minus_current_lnl"] = best
        tree.close()
        aln.close()
        return r

    try:
        res = grp.run(work)
    finally:
        grp.close()
    r0 = res[0]
    npat = r0["patterns"]
    out.update({
        "patterns": npat, "patterns_per_gpu": [r["patterns_local"] for r in res],
        "clv_arena_gb_per_gpu": (ntax - 2) * max(r["patterns_local"] for r in res) * 640 / 1e9,
        "load_s": max(r["load_s"] for r in res), "first_pass_s": max(r["first_pass_s"] for r in res),
        "lnl_true_tree": r0["lnl"],
        "likelihood_pass_ms": max(r["pass_ms"] for r in res),
        "sweep_ms": max(r["sweep_ms"] for r in res), "sweep_wall_s": max(r["sweep_wall_s"] for r in res),
        "sweep_launches": r0["sweep_launches"], "lnl_after_sweep": r0["lnl_after_sweep"],
        "spr_ms": max(r["spr_ms"] for r in res), "spr_wall_s": max(r["spr_wall_s"] for r in res),
        "spr_candidates": r0["spr_candidates"], "spr_launches": r0["spr_launches"],
        "best_candidate_minus_current_lnl": r0["best_candidate_minus_current_lnl"]})
    out["likelihood_pass_site_updates_per_s"] = (ntax - 2) * npat / (out["likelihood_pass_ms"] * 1e-3)
    out["sweep_site_updates_per_s"] = sum(r["sweep_site_updates"] for r in res) / (out["sweep_ms"] * 1e-3)
    out["spr_candidates_per_s"] = r0["spr_candidates"] / (out["spr_ms"] * 1e-3)
    out["spr_site_updates_per_s"] = sum(r["spr_site_updates"] for r in res) / (out["spr_ms"] * 1e-3)
    parts = sum(r["own_per_site_sum"] for r in res)
    out["parity"] = {
        "lnl_bit_identical_on_all_ranks": all(r["lnl"] == r0["lnl"] and r["lnl_after_sweep"] == r0["lnl_after_sweep"] for r in res),
        "sum_over_ranks_of_own_per_site_lnl": parts, "rel_err_vs_group_lnl": abs(parts - r0["lnl"]) / abs(r0["lnl"]),
        "sweep_does_not_lower_lnl": r0["lnl_after_sweep"] >= r0["lnl"] - 1e-9 * abs(r0["lnl"])}
    # the first columns alone, single rank, against the CPU oracle (checker only; absent oracle = said so)
    try:
        from oracle import oracle as orc
        nchk = args.c5_check_columns
        cseqs = [bytes(row[:nchk]).decode() for row in chars]
        ctx = pb.Context(0)
        a = pb.Alignment(ctx, names, cseqs, alpha=1.0)
        t = pb.Tree(a, nwk)
        got = t.evaluate()
        t.close(); a.close(); ctx.close()
        pat, w, _ = orc.compress(orc.encode(cseqs))
        want = orc.evaluate(orc.Model(), orc.Tree(nwk, names), pat, w, 1.0)
        out["parity"]["oracle_slice"] = {"columns": nchk, "engine_lnl": got, "oracle_lnl": want, "rel_err": abs(got - want) / abs(want)}
    except Exception as ex:  # noqa: BLE001
        out["parity"]["oracle_slice"] = {"unavailable": str(ex)}
    out["parity"]["ok"] = bool(out["parity"]["lnl_bit_identical_on_all_ranks"] and out["parity"]["rel_err_vs_group_lnl"] <= 1e-10 and
                               out["parity"]["sweep_does_not_lower_lnl"] and out["parity"]["oracle_slice"].get("rel_err", 0.0) <= 1e-10)
    return out


def run_c4(args):
    out = {"config": "synthetic %d taxa x %d sites WAG+G4: bootstrap replicate trees sharded by replicate, one replicate per GPU" % (
        args.c4_taxa, args.c4_sites)}
    t0 = time.perf_counter()
    names, chars, nwk = synth.simulate_wag_device(args.c4_taxa, args.c4_sites, 2)
    out["simulate_s"] = time.perf_counter() - t0
    import torch
    torch.cuda.empty_cache()
    ctxs = [pb.Context(g) for g in range(args.gpus)]
    alns = [None] * args.gpus

    def load(g):
        alns[g] = pb.Alignment(ctxs[g], names, chars, alpha=1.0)

    def in_threads(fn, gpus):
        errs = []

        def wrap(g):
            try:
                fn(g)
            except BaseException as ex:  # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=wrap, args=(g,)) for g in gpus]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    t0 = time.perf_counter()
    in_threads(load, range(args.gpus))
    out["load_s"] = time.perf_counter() - t0
    out["patterns"] = alns[0].npatterns
    out["clv_arena_gb_per_gpu"] = (args.c4_taxa - 2) * alns[0].npatterns * 640 / 1e9
    curve = []
    sizes = [k for k in (1, 2, 4, 8, 16) if k <= args.gpus]
    trees_seen = {}
    for k in sizes:
        results = {}

        def work(g, k=k, results=results):
            results[g] = alns[g].bootstrap_trees(k, weight_seed=12345, parsimony_seed=12345, first=g, stride=k)
            ctxs[g].sync()

        t0 = time.perf_counter()
        in_threads(work, range(k))
        wall = time.perf_counter() - t0
        lnls = {}
        for g in range(k):
            mine, rl, secs = results[g]
            for rid, nw in mine:
                lnls[rid] = float(rl[rid])
                # the same replicate gives the same tree whatever the sharding (same weights, same seeds, no collective)
                assert trees_seen.setdefault(rid, nw) == nw, "replicate %d differs between shardings" % rid
        curve.append({"gpus": k, "replicates": k, "wall_s": wall, "replicate_lnl": [lnls[r] for r in sorted(lnls)],
                      "time_for_100_replicates_s": -(-100 // k) * wall})
        log("c4: %d GPU(s), %d replicate(s): %.1f s" % (k, k, wall))
    out["curve"] = curve
    out["speedup_vs_1_gpu_for_100_replicates"] = {str(c["gpus"]): curve[0]["time_for_100_replicates_s"] / c["time_for_100_replicates_s"] for c in curve}
    out["replicate_trees_identical_across_shardings"] = True
    for a in alns:
        a.close()
    for c in ctxs:
        c.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--c5-taxa", type=int, default=2000)
    ap.add_argument("--c5-sites", type=int, default=1_000_000)
    ap.add_argument("--c5-prune", type=int, default=40)
    ap.add_argument("--c5-check-columns", type=int, default=192)
    ap.add_argument("--c4-taxa", type=int, default=500)
    ap.add_argument("--c4-sites", type=int, default=250_000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    out = {}
    if not args.skip_c5:
        out["c5"] = run_c5(args)
        log(json.dumps(out["c5"])[:600])
    if not args.skip_c4:
        out["c4"] = run_c4(args)
    text = json.dumps(out, indent=1)
    if args.out:
        open(args.out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
