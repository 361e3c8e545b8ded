"""Host-side logic of the site-sharded (N > 1) path on CPU: two gloo ranks split the patterns exactly as pml_aln_load
does, compute their shard's weighted lnL / derivative sums with the oracle, and the allreduce of those scalars equals the
single-rank answer -- the contract the engine's NCCL allreduce of 1-3 doubles relies on."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json
    import pepr_b200 as pb
    from oracle import oracle as orc
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "dup.json")))
    names, seqs = orc.read_phylip(os.path.join(ROOT, "tests", "golden", "dup.phy"))
    codes, w, s2p = pb.crunch_patterns(seqs)                       # every rank crunches the full alignment
    p0, p1 = pb.pattern_range(codes.shape[1], rank, world)         # ... and keeps one contiguous block
    tree = orc.Tree(g["fe"]["tree"], names)
    m = orc.Model()
    shard = np.ascontiguousarray(codes[:, p0:p1])
    lnl = orc.evaluate(m, tree, shard, w[p0:p1], g["fe"]["alpha"])
    d = orc.branch_derivs(m, tree, shard, w[p0:p1], g["fe"]["alpha"], 2, 0.05)
    buf = torch.tensor([lnl, d[0], d[1], d[2], float(p1 - p0)], dtype=torch.float64)
    dist.all_reduce(buf)
    # replicate weights are generated from the GLOBAL pattern weights on every rank, then sliced
    W, _ = pb.bootstrap_weights(w, 4242, 2)
    part = torch.tensor([float(W[:, p0:p1].sum())], dtype=torch.float64)
    dist.all_reduce(part)
    if rank == 0:
        full = orc.evaluate(m, tree, codes, w, g["fe"]["alpha"])
        dfull = orc.branch_derivs(m, tree, codes, w, g["fe"]["alpha"], 2, 0.05)
        q.put((buf.tolist(), [full, dfull[0], dfull[1], dfull[2], float(codes.shape[1])], part.item(), float(W.sum())))
    dist.destroy_process_group()


def test_two_rank_pattern_sharding_reduces_to_single_rank_answer():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want, wsum_parts, wsum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[4] == want[4]                                        # the blocks tile the pattern range exactly
    assert abs(got[0] - want[0]) < 1e-9 * abs(want[0]) and abs(got[1] - want[1]) < 1e-9 * abs(want[1])
    assert abs(got[2] - want[2]) < 1e-7 * max(1, abs(want[2])) and abs(got[3] - want[3]) < 1e-7 * max(1, abs(want[3]))
    assert wsum_parts == wsum


def test_pattern_range_tiles_without_gaps():
    import pepr_b200 as pb
    for npat in (1, 7, 128, 95933):
        for world in (1, 2, 3, 8):
            edges = [pb.pattern_range(npat, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == npat
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
