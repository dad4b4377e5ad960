"""Worker of tests/test_gpu_multigpu.py (run under torch.distributed.run, one rank per GPU).

Every rank owns some blocks of a connected lattice and exchanges ghost layers over NCCL inside the
library; rank 0 also runs the whole lattice on its own GPU in one process (same-GPU halo path) and
checks that the two give IDENTICAL residual histories and states -- the kernels and the data are
the same, only the transport differs -- and that both agree with the CPU oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import aither_b200
    from aither_b200 import ctypes_abi as abi
    from aither_b200 import distributed as adist
    from aither_b200 import synthetic

    solver = sys.argv[1] if len(sys.argv) > 1 else "dplur"
    # "viscous": BASELINE configs[3]'s scheme -- WENO5 + 4th-order central viscous fluxes, DPLUR;
    # the state exchange takes the multi-level, reference-order plan (edge ghost cells) over NCCL
    kw = dict(solver=solver, sweeps=2, limiter="vanAlbada", amplitude=0.02)
    n = 16
    if solver == "viscous":
        n = 12
        kw = dict(solver="dplur", sweeps=2, amplitude=0.02, viscous=True, recon="weno",
                  visc_recon="centralFourth", size=12 * 2e-6)
    rank, world, local = adist.env_rank()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = adist.make_comm(local)
    splits = {2: (1, 1, 2), 4: (2, 1, 2), 8: (2, 2, 2)}[world]
    nblk = splits[0] * splits[1] * splits[2]
    iters, cfl = 6, 40.0
    if solver == "viscous" and world == 2:
        # four blocks around one edge, two per rank: connections on the same GPU and across GPUs,
        # and a state exchange whose levels depend on each other through the edge ghost cells
        splits = (2, 2, 1)
    nblk = splits[0] * splits[1] * splits[2]
    prob = synthetic.lattice_problem(n, splits, **kw)
    per = synthetic.assign_ranks(prob, world)
    lvl = aither_b200.GridLevel(prob, device=local, rank=rank, n_ranks=world, block_ids=per[rank],
                                nccl_comm=comm)
    # direct exchange over peer memory unless AITHER_B200_HALO_P2P=0 (then ncclSend / ncclRecv)
    peer = lvl.enable_peer_exchange()
    hist = np.zeros((iters, prob.neq))
    for it in range(iters):
        lvl.store_old_solution(it)
        l2, _, _ = lvl.iterate(cfl)
        hist[it] = adist.reduce_norms(l2)
    g = prob.cfg.numGhosts
    mine = {b: lvl.field(n_, abi.FIELD_STATE) for n_, b in enumerate(per[rank])}
    minex = {b: lvl.field(n_, abi.FIELD_UPDATE) for n_, b in enumerate(per[rank])}
    lvl.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, minex))
    ok = True
    if rank == 0:
        import goldencheck as gc
        import oracle
        single = synthetic.lattice_problem(n, splits, **kw)
        one = aither_b200.GridLevel(single, device=local)
        ref = oracle.OracleLevel(single)
        for it in range(iters):
            one.store_old_solution(it)
            ref.store_old_solution(it)
            l2, _, _ = one.iterate(cfl)
            l2r, _, _ = ref.iterate(cfl)
            # sums over ranks are associated differently from sums over blocks: 1e-14
            assert np.all(np.abs(l2 - hist[it]) <= 1e-13 * np.abs(l2)), (it, l2, hist[it])
            assert np.all(np.abs(l2r - hist[it]) <= (1e-9 if solver == "viscous" else 1e-10) *
                          np.abs(l2r)), (it, l2r, hist[it])
        m = (gc.non_corner_mask if solver == "viscous" else gc.non_edge_mask)((n + 2 * g,) * 3, g)
        for states, xs in gathered:
            for b, st in states.items():
                assert np.array_equal(st[m], one.field(b, abi.FIELD_STATE)[m]), "state of block %d" % b
                assert np.array_equal(xs[b][m], one.field(b, abi.FIELD_UPDATE)[m]), "update of block %d" % b
                sr = ref.field(b, abi.FIELD_STATE)
                assert np.abs(st[m] - sr[m]).max() <= (1e-11 if solver == "viscous" else 1e-12) * \
                    np.abs(sr).max()
        one.close()
        ref.close()
        print("MULTIGPU_OK world=%d solver=%s blocks=%d exchange=%s" % (world, solver, nblk, "peer" if peer else "nccl"), flush=True)
    adist.destroy_comm(comm)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
