"""Worker of tests/test_gpu_multirank.py (run under torch.distributed.run, one rank per GPU): the site-sharded engine must
give the single-rank answers -- fixed-parameter lnL and per-branch derivatives against the oracle, and `-f e` against the
raxmlHPC golden -- on both collective paths (in-kernel NVLink reduction, NCCL fallback selected with PEPRML_NO_PEER=1)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import pepr_b200 as pb
from oracle import oracle as orc


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [pb.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = pb.Context(local, rank, world, box[0])
    case = sys.argv[1] if len(sys.argv) > 1 else "wide"
    g = json.load(open(os.path.join(ROOT, "tests", "golden", case + ".json")))
    p = os.path.join(ROOT, "tests", "golden", case + ".phy")
    if os.path.exists(p):
        names, seqs = orc.read_phylip(p)
    else:
        import gzip
        toks = gzip.open(p + ".gz", "rt").read().split()
        n = int(toks[0])
        names, seqs = toks[2::2][:n], toks[3::2][:n]
    fe = g["fe"]
    aln = pb.Alignment(ctx, names, seqs, alpha=fe["alpha"])
    tree = pb.Tree(aln, fe["tree"])
    lnl = tree.evaluate()
    assert abs(lnl - fe["lnl"]) <= 1e-6 * abs(fe["lnl"]), (lnl, fe["lnl"])
    pat, w, s2p = orc.compress(orc.encode(seqs))
    ot = orc.Tree(fe["tree"], names)
    want = orc.evaluate(orc.Model(), ot, pat, w, fe["alpha"])
    assert abs(lnl - want) <= 1e-10 * abs(want), (lnl, want)
    m = orc.Model()
    olen = {round(ot.get_bl(e), 15): e for e in range(ot.nedge)}   # match branches of the two parsers by length
    for e in (0, 3, tree.num_branches - 1):
        a, b, l = tree.branch(e)
        got = tree.branch_derivs(e, 0.5 * l + 0.01)
        ow = orc.branch_derivs(m, ot, pat, w, fe["alpha"], olen[round(l, 15)], 0.5 * l + 0.01)
        assert abs(got[0] - ow[0]) <= 1e-10 * abs(ow[0]), (e, got, ow)
        assert abs(got[1] - ow[1]) <= 1e-8 * max(1.0, abs(ow[1])) and abs(got[2] - ow[2]) <= 1e-8 * max(1.0, abs(ow[2])), (e, got, ow)
    # optimiser from default lengths: every rank must end on the same tree, within raxmlHPC's epsilon of its lnL
    topo_only = __import__("re").sub(r":[0-9.eE+-]+", "", fe["tree"])
    t2 = pb.Tree(aln, topo_only)
    aln.set_model(1.0)
    l2, alpha = t2.optimize(True, 0.1)
    assert l2 >= fe["lnl"] - 0.1, (l2, fe["lnl"])
    out = torch.tensor([lnl, l2, alpha], dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    assert all(torch.equal(gathered[0], x) for x in gathered), gathered   # bit-identical control flow on every rank
    if rank == 0:
        print("MULTIRANK_OK lnl %.6f fe %.6f alpha %.6f peer %s" % (lnl, l2, alpha, os.environ.get("PEPRML_NO_PEER", "0") != "1"))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
