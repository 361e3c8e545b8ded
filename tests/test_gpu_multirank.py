"""Site-sharded engine on two GPUs of one box (one process per GPU under torch.distributed.run): same answers as one rank,
on the in-kernel NVLink reduction and on the NCCL fallback; a rank that never shows up is reported, not waited for.
Skipped on boxes with a single GPU (tests/test_gpu_group.py covers ranks as threads of one process the same way)."""
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


def _run(mode, env_extra, port_off):
    env = dict(os.environ, **env_extra)
    port = 29700 + os.getpid() % 200 + port_off
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py"), mode],
                          env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)


@pytest.mark.gpu
@pytest.mark.parametrize("no_peer", ["0", "1"])
def test_two_ranks_match_single_rank(no_peer):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    r = _run("all" if no_peer == "0" else "likelihood", {"PEPRML_NO_PEER": "1"} if no_peer == "1" else {}, int(no_peer))
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
def test_lost_peer_is_reported_not_waited_for():
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    r = _run("lost", {"PEPRML_PEER_TIMEOUT_MS": "500"}, 2)
    assert r.returncode == 0 and "PML_ECOMM" in r.stdout, r.stdout[-3000:]
