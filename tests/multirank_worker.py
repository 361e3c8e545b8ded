"""Worker of tests/test_gpu_multirank.py (run under torch.distributed.run, one rank per GPU): the site-sharded engine must
give the single-rank answers on both collective paths (in-kernel NVLink reduction, NCCL fallback selected with
PEPRML_NO_PEER=1).  The checks themselves live in tests/multirank_checks.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import pepr_b200 as pb
from tests import multirank_checks as mc


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [pb.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = pb.Context(local, rank, world, box[0])
    mode = sys.argv[1] if len(sys.argv) > 1 else "likelihood"
    if mode == "lost":
        # rank 1 never makes the call: rank 0 must come back with PML_ECOMM after the timeout instead of hanging
        g, names, seqs = mc.load_case("small")
        aln = pb.Alignment(ctx, names, seqs, alpha=1.0)
        tree = pb.Tree(aln, g["fe"]["tree"])
        tree.evaluate()
        dist.barrier()
        failed = False
        if rank == 0:
            try:
                tree.smooth(1)
            except pb.EngineError as ex:
                failed = "(-4)" in str(ex)
        dist.barrier()
        if rank == 0:
            print("MULTIRANK_OK lost-peer handling: %s" % ("PML_ECOMM" if failed else "NO ERROR"))
            assert failed
        ctx.close()
        dist.destroy_process_group()
        return
    vals = mc.likelihood_checks(ctx, "wide")
    if mode == "all":
        v2, trees = mc.search_checks(ctx, "search")
        vals += v2
    out = torch.tensor(vals, dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    assert all(torch.equal(gathered[0], x) for x in gathered), gathered   # bit-identical control flow on every rank
    if rank == 0:
        print("MULTIRANK_OK %s: %d values bit-identical on %d ranks, collective %s" % (mode, len(vals), world, ctx.collective))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
