#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 maximum-likelihood engine (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[2], the 1-GPU configuration the metric is quoted on): synthetic WAG+Gamma4 supermatrix,
100 taxa x 100,000 sites per GPU, task = raxmlHPC `-f e` (alpha + all branch lengths on the fixed true topology, from
default branch lengths).  One STEP = one complete `-f e` optimisation.
  value      CLV site-updates/s with the alignment resident in HBM (timed: reset parameters -> pml_optimize)
  e2e        the same through the C ABI from HOST buffers (timed: pml_aln_load [pattern crunch + H2D] -> pml_tree_load
             -> pml_optimize -> newick + lnL back on the host)
  likelihood_pass  one full newview traversal + root evaluate (the kernel-bound figure the roofline explains)
  parity     at every N: lnL of the true tree on the N site shards together against the sum of the shards evaluated one
             by one on single-rank contexts, and bit-identity of the group result over the ranks
  bootstrap  the second half of BASELINE's metric: replicate trees sharded by REPLICATE (replicate r on rank r mod N, the
             full 100k-site pattern set on every rank, no collective), wall seconds per replicate tree for the whole job
  strong     100 taxa x `--strong-sites` (1 M) sites IN TOTAL, site-sharded over the N ranks: `-f e` step and likelihood pass
N > 1 (torchrun): weak scaling -- every rank holds its own block of 100k sites of an N x 100k-site alignment; the only
exchange is the sum of lnL / (lnL, d1, d2) scalars inside the branch kernels.

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
# Per pattern and kernel kind: algorithmic bytes (CLV traffic + tip codes + weights) and DMMA.8x8x4 issued per
# 16-pattern tile and category warp come from the engine (pml_kind_info).  One DMMA = 512 flop; four
# category warps per tile.
KIND = {}
FP64_PEAK_TFLOPS = 37.0
FP64_PEAK_SOURCE = ("FP64 DMMA.8x8x4, measured on a pool B200 with tools/fp64_peak.cu (profiles/r01_fp64_pipe_peaks.log: 37.0 TFLOP/s "
                    "sustained, DFMA 33.5, one shared pipe); MEASURED_PEAKS.json carries no FP64 figure")
WORKMODEL = os.path.join(ROOT, "bench_workmodel.json")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(ntax, nsites):
    from pepr_b200 import synth
    names, seqs, nwk = synth.simulate_wag(ntax, nsites, SEED)
    topo = re.sub(r":[0-9.eE+-]+", "", nwk)   # `-f e` ignores input lengths: start from defaults
    return names, seqs, topo, nwk


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


def run_raxml_rapid_bootstrap(names, seqs, threads, workdir, reps, limit_s=240.0):
    """`-f a -x 12345 -N reps` (what RAxMLRunner.run issues when bootstrapReps > 0): the replicate searches are timed by
    raxmlHPC itself ("Average Time per Rapid Bootstrap"); the ML search that follows them is not part of the metric and
    the process is stopped once that line has been written"""
    from pepr_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "raxmlHPC-PTHREADS" if threads > 1 else "raxmlHPC")
    if not os.path.exists(exe):
        return None
    synth.write_phylip(os.path.join(workdir, "rb.phy"), names, seqs)
    cmd = [exe, "-f", "a", "-x", "12345", "-N", str(reps), "-m", "PROTGAMMAWAG", "-s", "rb.phy", "-n", "rb"]
    if threads > 1:
        cmd += ["-T", str(threads)]
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, cwd=workdir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    info, m = os.path.join(workdir, "RAxML_info.rb"), None
    while time.perf_counter() - t0 < limit_s:
        time.sleep(0.2)
        txt = open(info).read() if os.path.exists(info) else ""
        m = re.search(r"Average Time per Rapid Bootstrap (\S+)", txt)
        if m or p.poll() is not None:
            break
    p.kill()
    p.wait()
    return float(m.group(1)) if m else None


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
    names, seqs, topo, _ = make_workload(NTAX, SITES_PER_GPU)
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
                   "sample_sites": sample, "patterns": last["patterns"], "effective": True,
                   "extrapolation": "the CPU arm runs %.1f %% of the columns; its site-updates/s assume cost linear in patterns" % (100.0 * sample / SITES_PER_GPU)},
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


def kernel_table(prof, steps, hbm_peak):
    """per kind: launches, device ms, achieved GB/s and FP64-tensor TFLOP/s, and the fraction of its own floor
    max(HBM time, FP64 pipe time) -- the roofline that actually binds that kernel"""
    kernels, floor_ms_total, ms_total = {}, 0.0, 0.0
    for name, (kms, kn, krows) in prof.items():
        if not kn:
            continue
        b, d = KIND.get(name, (0, 0))
        k = {"launches_per_step": kn / steps, "ms_per_step": kms / steps, "avg_launch_us": 1e3 * kms / kn}
        hbm_ms = b * krows / (hbm_peak * 1e9) * 1e3
        pipe_ms = 4 * d * 512 / 16.0 * krows / (FP64_PEAK_TFLOPS * 1e12) * 1e3
        if kms > 0:
            k["GBps"] = b * krows / (kms * 1e-3) / 1e9
            k["hbm_frac"] = hbm_ms / kms
            if d:
                k["dmma_TFLOPs"] = 4 * d * 512 / 16.0 * krows / (kms * 1e-3) / 1e12
                k["fp64_frac"] = pipe_ms / kms
            k["bound"] = "tensor" if pipe_ms > hbm_ms else "hbm"
            k["frac_of_floor"] = max(hbm_ms, pipe_ms) / kms
        kernels[name] = k
        floor_ms_total += max(hbm_ms, pipe_ms)
        ms_total += kms
    return kernels, floor_ms_total, ms_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", type=int, default=SITES_PER_GPU, help="sites per GPU (default: the named workload)")
    ap.add_argument("--ref-sites", type=int, default=0, help="cap of the CPU sample (columns)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bootstrap-reps", type=int, default=8,
                    help="replicate trees in the replicate-sharded section, IN TOTAL at every N (second half of BASELINE.json's metric); 0 = skip")
    ap.add_argument("--strong-sites", type=int, default=1_000_000,
                    help="sites IN TOTAL of the strong-scaling section (100 taxa, site-sharded over the ranks); 0 = skip")
    ap.add_argument("--strong-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import pepr_b200 as pb
    from pepr_b200 import engine as _eng
    KIND.update({name: (b, d) for name, b, d in _eng.kinds()})

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

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    def pinned(chars):
        host = torch.empty(chars.shape, dtype=torch.uint8).pin_memory()
        host.numpy()[:] = chars
        return host

    sites = args.sites
    names, seqs, topo, true_nwk = make_workload(NTAX, sites * world)
    chars = np.stack([np.frombuffer(s.encode(), np.uint8) for s in seqs])
    host = pinned(chars)
    hchars = host.numpy()

    ctx = pb.Context(local, rank, world, uid)
    aln = pb.Alignment(ctx, names, hchars, alpha=1.0)
    tree = pb.Tree(aln, topo)
    init = [tree.branch(e)[2] for e in range(tree.num_branches)]
    npat_local, npat = aln.npatterns_local, aln.npatterns

    def reset(t=None, a=None, lens=None):
        t, a, lens = t or tree, a or aln, lens or init
        for e, l in enumerate(lens):
            t.set_branch(e, l)
        a.set_model(1.0)
        t.invalidate()

    def counts(t=None):
        su, ln = (t or tree).stats()
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
    ms_max = reduce_max(ms)
    su_all = reduce_sum(float(su1 - su0))
    ms_step = ms_max / args.steps
    value = su_all / (ms_max * 1e-3)
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
    pass_ms = reduce_max(ctx.timer_stop() / reps)
    pass_value = (NTAX - 2) * npat / (pass_ms * 1e-3)

    # ---- parity at this N: the group's lnL of the TRUE tree against the sum of its shards evaluated alone ----------------
    # every rank rebuilds ITS pattern block as a stand-alone weighted alignment (a representative column per pattern, the
    # pattern weight as column weight) on a single-rank context of the same GPU: same kernels, no collective; the sum of
    # those over the ranks must be the group's number, and every rank must hold the same bits for the group's number
    solo = pb.Context(local) if world > 1 else ctx
    ptree = pb.Tree(aln, true_nwk)
    aln.set_model(1.0)
    group_lnl = ptree.evaluate()
    ptree.close()
    w_glob, s2p = aln.patterns()
    rep_col = np.full(npat, -1, np.int64)
    seen = s2p >= 0
    rep_col[s2p[seen][::-1]] = np.flatnonzero(seen)[::-1]      # first column of every pattern
    p0, p1 = pb.pattern_range(npat, rank, world)
    halves = [(p0, (p0 + p1) // 2), ((p0 + p1) // 2, p1)] if world == 1 else [(p0, p1)]
    block_sum = 0.0
    for lo, hi in halves:
        if hi <= lo:
            continue
        ba = pb.Alignment(solo, names, np.ascontiguousarray(hchars[:, rep_col[lo:hi]]), site_weights=w_glob[lo:hi], alpha=1.0)
        bt = pb.Tree(ba, true_nwk)
        block_sum += bt.evaluate()
        bt.close()
        ba.close()
    blocks_lnl = reduce_sum(block_sum)
    g_all = torch.tensor([group_lnl], dtype=torch.float64, device="cuda")
    if world > 1:
        gl = [torch.zeros_like(g_all) for _ in range(world)]
        dist.all_gather(gl, g_all)
        identical = all(torch.equal(gl[0], x) for x in gl)
    else:
        identical = True
    parity = {"what": "lnL of the generating tree (alpha 1) on %d site shard(s) together vs the %s" % (
                  world, "shards evaluated alone on single-rank contexts, summed" if world > 1 else "two halves of the patterns evaluated alone, summed"),
              "group_lnl": group_lnl, "sum_of_blocks_lnl": blocks_lnl, "rel_err": abs(group_lnl - blocks_lnl) / abs(blocks_lnl),
              "bit_identical_on_all_ranks": bool(identical), "tolerance": 1e-10}
    parity["ok"] = bool(parity["rel_err"] <= parity["tolerance"] and identical)

    # ---- bootstrap-replicate trees, sharded by replicate (no collective): 100 taxa x 100k sites on EVERY rank -----------------
    boot = None
    if args.bootstrap_reps > 0:
        R = args.bootstrap_reps
        bchars = np.ascontiguousarray(hchars[:, :sites])
        ba = pb.Alignment(solo, names, bchars, alpha=1.0)
        ba.bootstrap_trees(min(R, world), first=rank % max(1, min(R, world)), stride=max(1, min(R, world)))   # warm-up: one replicate per rank
        barrier()
        t0 = time.perf_counter()
        mine, rl, secs = ba.bootstrap_trees(R, weight_seed=12345, parsimony_seed=12345, first=rank, stride=world)
        solo.sync()
        own_wall = time.perf_counter() - t0
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)                   # the newick texts travel to the host that draws the supports
        else:
            gathered = [mine]
        boot_wall = reduce_max(own_wall)
        ba.close()
        boot = {"bootstrap_tree_wall_s": boot_wall / R, "replicates": R, "wall_s": boot_wall,
                "replicates_per_rank": [len(range(r, R, world)) for r in range(world)],
                "sharding": "by replicate: replicate r on rank r mod N, full 100 taxa x %d-site pattern set on every rank, no collective; newick texts gathered on rank 0" % sites,
                "what": "replicate site weights (raxmlHPC stream, seed 12345) -> parsimony start tree on the replicate (GPU Fitch scans) -> "
                        "alpha + branch lengths (eps 5) -> one lazy-SPR round (radius 5) with branch smoothing; whole-job wall seconds / replicates",
                "seconds_per_replicate_on_one_gpu": float(np.nanmean(secs)) if len(mine) else None}
        if rank == 0:
            trees = [nw for part in gathered for _, nw in part]
            ids = sorted(i for part in gathered for i, _ in part)
            boot["trees_gathered"] = len(trees)
            boot["all_replicates_present"] = ids == list(range(R))
            boot["support_tree_chars"] = len(pb.support_tree(trees[0], trees, as_percent=True)) if trees else 0
    if world > 1:
        solo.close()

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------------
    tree.close()
    aln.close()

    def e2e_step(c, hc, tp):
        a = pb.Alignment(c, names, hc, alpha=1.0)      # pattern crunch on the host + H2D of codes/weights
        t = pb.Tree(a, tp)
        l, al = t.optimize(True, 0.1)
        nw = t.newick()                                 # result tree string back on the host
        su, _ = t.stats()
        h2d = a.ntax * a.npatterns_local + 8 * a.npatterns_local
        t.close()
        a.close()
        return l, sum(su), h2d, len(nw)

    for _ in range(min(args.warmup, 2)):
        e2e_step(ctx, hchars, topo)
    barrier()
    t0 = time.perf_counter()
    e_su = 0
    for _ in range(args.steps):
        l2, su, h2d, nwlen = e2e_step(ctx, hchars, topo)
        e_su += su
    barrier()
    e_wall = reduce_max(time.perf_counter() - t0)
    e2e_value = reduce_sum(float(e_su)) / e_wall
    # device -> host: every branch pass publishes five flagged doubles (80 B) through mapped memory; + the result tree
    npass = sum(v[1] for k, v in prof.items() if k.startswith(("evaluate", "branch", "core", "fused"))) / args.steps
    d2h = int(80 * npass + nwlen)

    # ---- strong scaling: 100 taxa x strong_sites IN TOTAL over the N ranks -------------------------------------------------
    strong = None
    if args.strong_sites > 0:
        del host, hchars, chars
        s_names, s_seqs, s_topo, _ = make_workload(NTAX, args.strong_sites)
        s_host = pinned(np.stack([np.frombuffer(s.encode(), np.uint8) for s in s_seqs]))
        del s_seqs
        sc = s_host.numpy()
        barrier()
        t0 = time.perf_counter()
        s_aln = pb.Alignment(ctx, s_names, sc, alpha=1.0)
        load_s = reduce_max(time.perf_counter() - t0)
        s_tree = pb.Tree(s_aln, s_topo)
        s_init = [s_tree.branch(e)[2] for e in range(s_tree.num_branches)]
        reset(s_tree, s_aln, s_init)
        s_lnl, s_alpha = s_tree.optimize(True, 0.1)        # warm-up step (also brings the arena in)
        barrier()
        a0, _ = counts(s_tree)
        ctx.timer_start()
        for _ in range(args.strong_steps):
            reset(s_tree, s_aln, s_init)
            s_lnl, s_alpha = s_tree.optimize(True, 0.1)
        s_ms = reduce_max(ctx.timer_stop())
        barrier()
        a1, _ = counts(s_tree)
        s_su = reduce_sum(float(a1 - a0))
        for _ in range(2):
            s_tree.invalidate()
            s_tree.evaluate()
        barrier()
        ctx.timer_start()
        for _ in range(5):
            s_tree.invalidate()
            s_tree.evaluate()
        s_pass = reduce_max(ctx.timer_stop() / 5)
        s_npat, s_nloc = s_aln.npatterns, s_aln.npatterns_local
        s_tree.close()
        s_aln.close()
        barrier()
        t0 = time.perf_counter()
        _, e_su1, _, _ = e2e_step(ctx, sc, s_topo)
        s_e2e = reduce_max(time.perf_counter() - t0)
        strong = {"workload": "synthetic 100 taxa x %d sites IN TOTAL (seed 3), site-sharded over %d rank(s): -f e" % (args.strong_sites, world),
                  "scaling": "strong", "sites_total": args.strong_sites, "patterns": s_npat, "patterns_per_gpu": s_nloc,
                  "clv_arena_gb_per_gpu": (NTAX - 2) * s_nloc * 640 / 1e9,
                  "value": s_su / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / args.strong_steps, "steps": args.strong_steps,
                  "likelihood_pass_ms": s_pass, "likelihood_pass_value": (NTAX - 2) * s_npat / (s_pass * 1e-3),
                  "e2e_ms_per_step": 1e3 * s_e2e, "e2e_value": reduce_sum(float(e_su1)) / s_e2e, "aln_load_s": load_s,
                  "final_lnl": s_lnl, "final_alpha": s_alpha}

    if rank == 0:
        peak, peak_src = peaks()
        kernels, floor_ms, prof_ms = kernel_table(prof, args.steps, peak)
        # the roofline line describes the kernel kind that takes the largest share of the step
        k = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
        kms, kn, krows = prof[k]
        kb, kd = KIND[k]
        kk = kernels[k]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(k)
        tensor_bound = kk["bound"] == "tensor"
        roofline = {"kernel": k, "share_of_step": kk["ms_per_step"] / (prof_ms / args.steps),
                    "bound": kk["bound"],
                    "achieved": kk["dmma_TFLOPs"] if tensor_bound else kk["GBps"],
                    "peak": FP64_PEAK_TFLOPS if tensor_bound else peak,
                    "unit": "TFLOP/s" if tensor_bound else "GB/s",
                    "frac": kk["frac_of_floor"], "traffic": traffic,
                    "peak_source": FP64_PEAK_SOURCE if tensor_bound else peak_src,
                    "hbm": {"achieved_GBps": kk["GBps"], "peak_GBps": peak, "frac": kk["hbm_frac"], "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": kb * npat_local},
                    "fp64_tensor": {"achieved_TFLOPs": kk.get("dmma_TFLOPs"), "peak_TFLOPs": FP64_PEAK_TFLOPS, "frac": kk.get("fp64_frac"),
                                    "fp64_peak_source": FP64_PEAK_SOURCE, "dmma_per_tile_and_warp": kd,
                                    "flop_per_launch": 4 * kd * 512 / 16.0 * npat_local},
                    "avg_launch_us": 1e3 * kms / max(kn, 1),
                    "step": {"sum_of_kernel_floors_ms": floor_ms / args.steps, "ms_per_step_with_per_launch_events": ms_profiled / args.steps,
                             "ms_per_step": ms_step, "frac_of_floor": floor_ms / args.steps / ms_step,
                             "note": "floor of a launch = max(algorithmic bytes / HBM peak, issued DMMA flop / FP64 tensor peak)"},
                    "newview_inner_inner_hbm_frac": kernels.get("newview_inner_inner", {}).get("hbm_frac")}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "synthetic 100 taxa x 100k sites/GPU WAG+G4 (seed 3): -f e (alpha + branch lengths, fixed topology)",
                       "taxa": NTAX, "sites_per_gpu": sites, "patterns": npat, "partition": "sites (pattern blocks)",
                       "collective": ctx.collective,
                       "cache": "CLV working set %.1f GB per GPU >> 126 MB L2 (inputs larger than L2)" % ((NTAX - 2) * npat_local * 640 / 1e9),
                       "cherries": "inner nodes with two tip children are never stored: their consumers form them from two tip look-ups "
                                   "(a folded update counts as a site-update whenever a stored one would have been recomputed)",
                       "launches_per_step": launches / args.steps,
                       "final_lnl": lnl, "final_alpha": alpha},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e_wall / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "parity": parity,
            "likelihood_pass": {"value": pass_value, "unit": UNIT, "ms": pass_ms},
            "kernels": kernels,
            "wall_s_timed_region": wall,
            "ms_per_step_with_per_launch_events": ms_profiled / args.steps,
            "site_updates_per_pattern": (su1 - su0) / args.steps / max(npat_local, 1),
        }
        if boot:
            out["bootstrap"] = boot
        if strong:
            out["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg(pb, ctx, args, names, seqs, topo, true_nwk, out)
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline_leg(pb, ctx, args, names, seqs, topo, true_nwk, out):
    """rank 0, N = 1 only: the reference's CPU implementation on the box's host cores on BOUNDED samples of the same workload
    (`-f e`, and rapid-bootstrap replicates for the second half of the metric), and -- the one place bench.py uses the test
    oracle, as a checker -- the engine's lnL of a slice of the workload against the CPU restatement"""
    import numpy as np
    cores = os.cpu_count() or 1
    sites = len(seqs[0])
    sample = args.ref_sites or min(sites, max(500, int(20.0 * 30.0 * cores)))
    sseqs = [s[:sample] for s in seqs]
    tmp = tempfile.mkdtemp(prefix="pepr_cpu_")
    try:
        r = run_raxml(names, sseqs, topo, cores, tmp, "cpu")
        bsample = 600
        rb = run_raxml_rapid_bootstrap(names, [s[:bsample] for s in seqs], cores, tmp, 1)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not r:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref missing"}
    # the engine's own site-update count for the SAME sampled task makes the CPU figure "effective"
    a = pb.Alignment(ctx, names, sseqs, alpha=1.0)
    t = pb.Tree(a, topo)
    l3, _ = t.optimize(True, 0.1)
    su, _ = t.stats()
    t.close()
    a.close()
    cb = {"value": sum(su) / r["wall_s"], "unit": UNIT, "cores": cores, "kind": "reference",
          "sample": "first %d columns (%d patterns), oracle/_ref/raxmlHPC-PTHREADS -T %d -f e, %.1f s; "
                    "effective = engine site-update count for the same task / CPU wall" % (sample, r["patterns"], cores, r["wall_s"]),
          "lnl_cpu": r["lnl"], "lnl_engine": l3}
    if rb:
        mine = out.get("bootstrap", {}).get("seconds_per_replicate_on_one_gpu")
        cb["bootstrap"] = {"average_time_per_rapid_bootstrap_s": rb, "sample": "first %d columns, raxmlHPC-PTHREADS -T %d -f a -x 12345 -N 1 "
                           "(stopped after the replicate search)" % (bsample, cores),
                           "extrapolated_to_100k_columns_s": rb * sites / bsample,
                           "note": "linear extrapolation in columns; raxmlHPC's rapid bootstrap is a different (CAT-based) search than the engine's replicate search",
                           "engine_seconds_per_replicate_tree_100k_columns": mine}
    # oracle check (checker only): fixed-parameter lnL of the first 1,500 columns, engine vs the CPU restatement
    try:
        from oracle import oracle as orc
        nchk = 1500
        cseqs = [s[:nchk] for s in seqs]
        a = pb.Alignment(ctx, names, cseqs, alpha=1.0)
        t = pb.Tree(a, true_nwk)
        got = t.evaluate()
        t.close()
        a.close()
        pat, w, _ = orc.compress(orc.encode(cseqs))
        want = orc.evaluate(orc.Model(), orc.Tree(true_nwk, names), pat, w, 1.0)
        out["parity"]["oracle_slice"] = {"columns": nchk, "engine_lnl": got, "oracle_lnl": want, "rel_err": abs(got - want) / abs(want)}
        out["parity"]["ok"] = bool(out["parity"]["ok"] and out["parity"]["oracle_slice"]["rel_err"] <= 1e-10)
    except Exception as ex:  # the checker is absent: say so, do not fail the measurement
        out["parity"]["oracle_slice"] = {"unavailable": str(ex)}
    return cb


if __name__ == "__main__":
    sys.exit(main())
