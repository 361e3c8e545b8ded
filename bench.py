#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 maximum-likelihood engine (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[2], the 1-GPU configuration the metric is quoted on): synthetic WAG+Gamma4 supermatrix,
100 taxa x 100,000 sites per GPU, task = raxmlHPC `-f e` (alpha + all branch lengths on the fixed true topology, from
default branch lengths).  One STEP = one complete `-f e` optimisation.
  value      CLV site-updates/s with the alignment resident in HBM (timed: reset parameters -> pml_optimize)
  e2e        the same through the C ABI from HOST buffers (timed: pml_aln_load [pattern crunch + H2D] -> pml_tree_load
             -> pml_optimize -> newick + lnL back on the host)
  likelihood_pass  one full newview traversal + root evaluate (the kernel-bound figure the roofline explains)
N > 1 (torchrun): weak scaling -- every rank holds its own block of 100k sites of an N x 100k-site alignment; the only
exchange is the NCCL allreduce of lnL / (lnL, d1, d2) scalars inside the engine.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/raxmlHPC-PTHREADS, all host threads) on a
bounded column sample of the same task; its site-updates are "effective": the engine's site-update count per pattern for
this task (bench_workmodel.json) x the sample's patterns / wall time.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "CLV site-updates/s (WAG+G4, 20-state)"
UNIT = "site-updates/s"
NTAX = 100
SITES_PER_GPU = 100_000
SEED = 3
BYTES = {"newview_tip_tip": 642, "newview_tip_inner": 1281, "newview_inner_inner": 1920, "evaluate": 660,
         "branch_inner_inner": 1292, "branch_tip_inner": 652, "core": 648}   # algorithmic bytes per pattern (SURVEY 8d / DESIGN.md)
# DMMA.8x8x4 issued per 16-pattern tile and MMA warp by the branch kernel (its binding roofline is the FP64 tensor pipe)
FLOP_PER_PATTERN = {"branch_inner_inner": 4 * 72 * 512 / 16.0, "branch_tip_inner": 4 * 42 * 512 / 16.0}
WORKMODEL = os.path.join(ROOT, "bench_workmodel.json")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(ntax, nsites):
    from pepr_b200 import synth
    names, seqs, nwk = synth.simulate_wag(ntax, nsites, SEED)
    topo = re.sub(r":[0-9.eE+-]+", "", nwk)   # `-f e` ignores input lengths: start from defaults
    return names, seqs, topo


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = max((float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ reference arm
def run_raxml(names, seqs, topo, threads, workdir, tag):
    from pepr_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "raxmlHPC-PTHREADS" if threads > 1 else "raxmlHPC")
    if not os.path.exists(exe):
        return None
    synth.write_phylip(os.path.join(workdir, tag + ".phy"), names, seqs)
    open(os.path.join(workdir, tag + ".nwk"), "w").write(topo + "\n")
    cmd = [exe, "-f", "e", "-m", "PROTGAMMAWAG", "-s", tag + ".phy", "-t", tag + ".nwk", "-n", tag]
    if threads > 1:
        cmd += ["-T", str(threads)]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.perf_counter() - t0
    info = os.path.join(workdir, "RAxML_info." + tag)
    txt = open(info).read() if os.path.exists(info) else r.stdout
    m = re.search(r"Final GAMMA\s+likelihood: (\S+)", txt)
    pat = re.search(r"Alignment has (\d+) distinct alignment patterns", txt)
    for f in os.listdir(workdir):
        if f.startswith("RAxML_") and f.endswith("." + tag):
            os.remove(os.path.join(workdir, f))
    if r.returncode != 0 or not m:
        return None
    return {"wall_s": wall, "lnl": float(m.group(1)), "patterns": int(pat.group(1)) if pat else len(seqs[0])}


def workmodel_updates_per_pattern(ntax):
    if os.path.exists(WORKMODEL):
        wm = json.load(open(WORKMODEL))
        if wm.get("ntax") == ntax:
            return wm["fe_site_updates_per_pattern"], "bench_workmodel.json (engine count on this task)"
    return 12.0 * (ntax - 2), "nominal 12 full traversals (no bench_workmodel.json)"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    total = max(1, args.steps + args.warmup)
    # about 150 s for the whole run; raxmlHPC-PTHREADS `-f e` does roughly 35 columns/s/thread on 100 taxa (8-thread probe)
    budget = 150.0 / total
    sample = int(min(SITES_PER_GPU, max(500, budget * 30.0 * cores)))
    sample = min(sample, args.ref_sites) if args.ref_sites else sample
    names, seqs, topo = make_workload(NTAX, SITES_PER_GPU)
    sseqs = [s[:sample] for s in seqs]
    tmp = tempfile.mkdtemp(prefix="pepr_ref_")
    times, last = [], None
    try:
        for i in range(total):
            r = run_raxml(names, sseqs, topo, cores, tmp, "r%d" % i)
            if r is None:
                emit({"impl": "reference", "unavailable": "oracle/_ref/raxmlHPC-PTHREADS missing or failed"})
                return 0
            last = r
            if i >= args.warmup:
                times.append(r["wall_s"])
            log("reference step %d: %.2f s lnL %.3f" % (i, r["wall_s"], r["lnl"]))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    per_pat, how = workmodel_updates_per_pattern(NTAX)
    t = sum(times) / len(times)
    value = per_pat * last["patterns"] / t
    desc = "first %d of %d columns of the 100-taxon workload, raxmlHPC-PTHREADS -T %d -f e; effective site-updates = %s" % (
        sample, SITES_PER_GPU, cores, how)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 100 taxa x 100k sites WAG+G4: -f e (alpha + branch lengths, fixed topology)",
                   "sample_sites": sample, "patterns": last["patterns"], "effective": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lnl": last["lnl"]})
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def emit(obj):
    """the ONE JSON line goes to the real stdout; everything libraries print (e.g. NCCL's version banner) was sent to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", type=int, default=SITES_PER_GPU, help="sites per GPU (default: the named workload)")
    ap.add_argument("--ref-sites", type=int, default=0, help="cap of the CPU sample (columns)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bootstrap-reps", type=int, default=2,
                    help="also time this many bootstrap-replicate trees (second half of BASELINE.json's metric); 0 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import pepr_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("warning: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [pb.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sites = args.sites
    names, seqs, topo = make_workload(NTAX, sites * world)
    chars = np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
    host = torch.empty(chars.shape, dtype=torch.uint8).pin_memory()
    host.numpy()[:] = chars
    hchars = host.numpy()

    ctx = pb.Context(local, rank, world, uid)
    aln = pb.Alignment(ctx, names, hchars, alpha=1.0)
    tree = pb.Tree(aln, topo)
    init = [tree.branch(e)[2] for e in range(tree.num_branches)]
    npat_local, npat = aln.npatterns_local, aln.npatterns

    def reset():
        for e, l in enumerate(init):
            tree.set_branch(e, l)
        aln.set_model(1.0)
        tree.invalidate()

    def counts():
        su, ln = tree.stats()
        return sum(su), ln

    # ---- value: resident alignment, one `-f e` per step --------------------------------------------------
    for _ in range(args.warmup):
        reset()
        lnl, alpha = tree.optimize(True, 0.1)
    barrier()
    su0, ln0 = counts()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reset()
        lnl, alpha = tree.optimize(True, 0.1)
    ms = ctx.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    su1, ln1 = counts()
    # the same K steps once more with a pair of CUDA events around EVERY launch (per-kernel times for the roofline); kept
    # out of the region above because an event between two kernels forbids their programmatic overlap
    ctx.profile_begin()
    ctx.timer_start()
    for _ in range(args.steps):
        reset()
        tree.optimize(True, 0.1)
    ms_profiled = ctx.timer_stop()
    prof = ctx.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_all = torch.tensor([ms], dtype=torch.float64, device="cuda")
    su_all = torch.tensor([float(su1 - su0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(su_all, op=dist.ReduceOp.SUM)
    ms_step = ms_all.item() / args.steps
    value = su_all.item() / (ms_all.item() * 1e-3)
    launches = ln1 - ln0

    # ---- likelihood pass: full traversal + evaluate (kernel-bound) -------------------------------------------
    for _ in range(3):
        tree.invalidate()
        tree.evaluate()
    barrier()
    reps = 10
    ctx.timer_start()
    for _ in range(reps):
        tree.invalidate()
        tree.evaluate()
    pass_ms = ctx.timer_stop() / reps
    pass_all = torch.tensor([pass_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(pass_all, op=dist.ReduceOp.MAX)
    pass_value = (NTAX - 2) * npat / (pass_all.item() * 1e-3)

    # ---- optional: bootstrap-replicate trees (replicate weights -> parsimony start tree -> lazy SPR -> branch lengths) ------
    boot = None
    if args.bootstrap_reps > 0:
        W, _ = aln.bootstrap_weights(12345, args.bootstrap_reps)
        barrier()
        t0 = time.perf_counter()
        for r in range(args.bootstrap_reps):
            bt = pb.Tree(aln, parsimony_seed=12346 + r, weights=W[r])
            bt.optimize(False, 5.0, weights=W[r])
            bl, bm = bt.search(radius=5, max_rounds=1, eps=0.1, weights=W[r])
            bt.close()
        barrier()
        boot = {"bootstrap_tree_wall_s": (time.perf_counter() - t0) / args.bootstrap_reps, "replicates": args.bootstrap_reps,
                "what": "replicate site weights (raxmlHPC stream, seed 12345) -> parsimony start tree on the replicate (GPU Fitch scans) -> "
                        "branch lengths (eps 5) -> one lazy-SPR round (radius 5) with branch smoothing; wall seconds per replicate tree, "
                        "replicates run one after the other on every rank's pattern shard"}

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------------
    tree.close()
    aln.close()

    def e2e_step():
        a = pb.Alignment(ctx, names, hchars, alpha=1.0)      # pattern crunch on the host + H2D of codes/weights
        t = pb.Tree(a, topo)
        l, al = t.optimize(True, 0.1)
        nw = t.newick()                                       # result tree string back on the host
        su, _ = t.stats()
        h2d = a.ntax * a.npatterns_local + 8 * a.npatterns_local
        t.close()
        a.close()
        return l, sum(su), h2d, len(nw)

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_su = 0
    for _ in range(args.steps):
        l2, su, h2d, nwlen = e2e_step()
        e_su += su
    barrier()
    e_wall = time.perf_counter() - t0
    e_t = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
    e_s = torch.tensor([float(e_su)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e_s, op=dist.ReduceOp.SUM)
    e2e_value = e_s.item() / e_t.item()
    # device -> host: every branch pass publishes five (value, seq) pairs (80 B) through mapped memory; + the result tree
    npass = sum(prof[k][1] for k in ("evaluate", "branch_inner_inner", "branch_tip_inner", "core", "fused_update_branch_inner",
                                     "fused_update_branch_tip")) / args.steps
    d2h = int(80 * npass + nwlen)

    if rank == 0:
        peak, peak_src = peaks()
        k = "newview_inner_inner"
        kms, kn, krows = prof[k]
        achieved = BYTES[k] * krows / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(k)
        kernels = {}
        for name, (kms_, kn_, krows_) in prof.items():
            if kn_:
                kernels[name] = {"launches_per_step": kn_ / args.steps, "ms_per_step": kms_ / args.steps, "avg_launch_us": 1e3 * kms_ / kn_}
                if name in BYTES:
                    kernels[name]["GBps"] = BYTES[name] * krows_ / (kms_ * 1e-3) / 1e9
                if name in FLOP_PER_PATTERN:   # issued FP64 tensor flops (24-wide padding included) against the measured 37.0 TFLOP/s
                    kernels[name]["dmma_TFLOPs"] = FLOP_PER_PATTERN[name] * krows_ / (kms_ * 1e-3) / 1e12
                    kernels[name]["dmma_frac_of_37.0"] = kernels[name]["dmma_TFLOPs"] / 37.0
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "synthetic 100 taxa x 100k sites/GPU WAG+G4 (seed 3): -f e (alpha + branch lengths, fixed topology)",
                       "taxa": NTAX, "sites_per_gpu": sites, "patterns": npat, "partition": "sites (pattern blocks)",
                       "collective": ctx.collective,
                       "cache": "CLV working set %.1f GB per GPU >> 126 MB L2 (inputs larger than L2)" % ((NTAX - 2) * npat_local * 640 / 1e9),
                       "final_lnl": lnl, "final_alpha": alpha},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e_t.item() / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": k, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES[k] * npat_local, "avg_launch_us": 1e3 * kms / max(kn, 1)},
            "likelihood_pass": {"value": pass_value, "unit": UNIT, "ms": pass_all.item()},
            "kernels": kernels,
            "wall_s_timed_region": wall,
            "ms_per_step_with_per_launch_events": ms_profiled / args.steps,
            "site_updates_per_pattern": (su1 - su0) / args.steps / max(npat_local, 1),
        }
        if boot:
            out["bootstrap"] = boot
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.ref_sites or min(sites, max(500, int(20.0 * 30.0 * cores)))
            sseqs = [s[:sample] for s in seqs]
            tmp = tempfile.mkdtemp(prefix="pepr_cpu_")
            try:
                r = run_raxml(names, sseqs, topo, cores, tmp, "cpu")
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
            if r:
                # the engine's own site-update count for the SAME sampled task makes the CPU figure "effective"
                a = pb.Alignment(ctx, names, sseqs, alpha=1.0)
                t = pb.Tree(a, topo)
                l3, _ = t.optimize(True, 0.1)
                su, _ = t.stats()
                out["cpu_baseline"] = {"value": sum(su) / r["wall_s"], "unit": UNIT, "cores": cores, "kind": "reference",
                                       "sample": "first %d columns (%d patterns), oracle/_ref/raxmlHPC-PTHREADS -T %d -f e, %.1f s; "
                                                 "effective = engine site-update count for the same task / CPU wall" % (sample, r["patterns"], cores, r["wall_s"]),
                                       "lnl_cpu": r["lnl"], "lnl_engine": l3}
                t.close()
                a.close()
            else:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"}
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
