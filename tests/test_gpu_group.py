"""All ranks of a site-sharded group as threads of ONE process (pml_group_create): the in-kernel NVLink reduction must be
in use (collective == 2, no CUDA IPC involved) and every check of tests/multirank_checks.py must hold with bit-identical
results on all ranks.  Needs two GPUs; the replicate-sharded path below it runs on one."""
import numpy as np
import pytest

import pepr_b200 as pb
from tests import multirank_checks as mc

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_threads_of_one_process_reach_the_in_kernel_reduction():
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    grp = pb.Group([0, 1])
    try:
        assert [c.collective for c in grp.contexts] == ["in-kernel nvlink"] * 2
        res = grp.run(lambda ctx: mc.likelihood_checks(ctx, "wide"))
        assert res[0] == res[1]                      # floats compared exactly: identical control flow on both ranks
        res = grp.run(lambda ctx: mc.search_checks(ctx, "search"))
        assert res[0] == res[1]
    finally:
        grp.close()


def test_replicates_sharded_by_replicate_equal_the_unsharded_run(gpu_ctx, golden):
    """replicate r on share r mod N with the full pattern set and no collective.  Shares run here one after
    the other on one GPU (the sharding logic does not care where a share runs): same weights, same trees, same lnL."""
    g = golden("search")
    aln = pb.Alignment(gpu_ctx, g.names, g.seqs, alpha=1.0)
    whole, lnl_all, _ = aln.bootstrap_trees(4, weight_seed=12345, parsimony_seed=777)
    assert [i for i, _ in whole] == [0, 1, 2, 3]
    got = {}
    for share in range(2):
        part, lnl, secs = aln.bootstrap_trees(4, weight_seed=12345, parsimony_seed=777, first=share, stride=2)
        assert [i for i, _ in part] == [share, share + 2]
        for i, nw in part:
            got[i] = nw
            assert lnl[i] == lnl_all[i] and secs[i] > 0
    assert [got[i] for i in range(4)] == [nw for _, nw in whole]
    # the replicate's weights are raxmlHPC's stream (bit-exact vs the `-f j -b` golden in test_host / test_gpu_parity)
    W, _ = aln.bootstrap_weights(12345, 4)
    assert W.sum(axis=1).tolist() == [len(g.seqs[0])] * 4
    aln.close()
