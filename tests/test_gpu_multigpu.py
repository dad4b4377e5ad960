"""Cross-GPU ghost-layer exchange inside the library -- over NVLink peer memory (default) and
with ncclSend / ncclRecv (AITHER_B200_HALO_P2P=0) --, -m gpu; needs >= 2 GPUs on the box, otherwise skipped. One process per GPU under torch.distributed.run; the check itself
is in tests/multigpu_worker.py (multi-process NCCL == single-process same-GPU halo == CPU oracle).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("solver,exchange", [("dplur", "peer"), ("lusgs", "peer"), ("viscous", "peer"),
                                             ("dplur", "nccl"), ("viscous", "nccl")])
def test_two_ranks_match_one_rank_and_oracle(solver, exchange):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, AITHER_B200_HALO_P2P="1" if exchange == "peer" else "0")
    port = 29600 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multigpu_worker.py"), solver]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=600, env=env)
    assert res.returncode == 0 and "MULTIGPU_OK" in res.stdout, res.stdout[-4000:]
    assert "exchange=%s" % exchange in res.stdout, res.stdout[-2000:]
