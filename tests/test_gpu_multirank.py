"""Site-sharded engine on two GPUs of one box (one process per GPU under torch.distributed.run): same answers as one rank,
on the in-kernel NVLink reduction and on the NCCL fallback.  Skipped on boxes with a single GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("no_peer", ["0", "1"])
def test_two_ranks_match_single_rank(no_peer):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, PEPRML_NO_PEER=no_peer)
    if no_peer == "0":
        env.pop("PEPRML_NO_PEER")
    port = 29700 + os.getpid() % 200 + int(no_peer)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py"), "wide"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, r.stdout[-3000:]
