"""Host-side multi-rank logic on CPU (gloo, world_size 2): rank assignment of a connected lattice,
the byte broadcast that carries the NCCL id, the byte all-gather that carries the peer-memory
handles of the direct ghost exchange, and the reduction of residual norms. The ghost-layer
exchange itself is device code (tests/test_gpu_multigpu.py)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from aither_b200 import distributed as adist
    from aither_b200 import synthetic
    adist.init_process_group("gloo")
    # the 128-byte id travels from rank 0 to everyone
    ident = bytes(range(128)) if rank == 0 else bytes(128)
    got = adist.broadcast_bytes(ident, 128, src=0)
    assert got == bytes(range(128))
    # the 64-byte peer-memory handles of all ranks, in rank order, on every rank
    mine = bytes([rank + 1] * 64)
    everyone = adist.all_gather_bytes(mine, 64)
    assert everyone == b"".join(bytes([r + 1] * 64) for r in range(world))
    # every rank derives the same placement from the same connection list
    prob = synthetic.lattice_problem(4, (1, 2, 2), only=[])
    per = synthetic.assign_ranks(prob, world)
    assert per == [[0, 1], [2, 3]]
    for cn in prob.conns:
        for s in range(2):
            assert cn.rank[s] == cn.block[s] // 2 and cn.localBlock[s] == cn.block[s] % 2
    cross = sum(1 for cn in prob.conns if cn.rank[0] != cn.rank[1])
    assert cross == 2 and len(prob.conns) == 4
    # norms: sum over ranks; matrix residual recombined over the global size
    l2, mr = adist.reduce_norms(np.arange(5.0) * (rank + 1), 2.0 * (rank + 1), 10 * (rank + 1))
    assert np.allclose(l2, np.arange(5.0) * 3)
    assert abs(mr - (2.0 * 10 + 4.0 * 20) / 30) < 1e-15
    assert adist.max_over_ranks(float(rank)) == world - 1
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
