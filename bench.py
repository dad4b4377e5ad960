#!/usr/bin/env python
"""Benchmark of the B200 hot path: Mcell-iter/s of one full nonlinear iteration
(BC fill + residual + time step + implicit DPLUR solve + update + norms; the body of the
reference's mgSolution::Iterate, src/mgSolution.cpp:246-269).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference on host cores

Workload (BASELINE.json configs[1]): synthetic 256^3 single block per GPU, Euler,
Roe + MUSCL (kappa = 1/3, no limiter), implicit Euler, DPLUR x4, CFL 50, characteristic i-faces,
slip-wall j/k faces, seed-fixed +-1 % perturbed state. fp64 throughout.

A "step" is one nonlinear iteration over the whole block. `value` times K steps back to back
with everything resident in HBM (CUDA events on the library's stream). `e2e` times the same K
steps through the C-ABI entry points a host solver would call when it owns the state: every step
uploads the state from pinned host memory (aither_gpu_upload_state), runs aither_gpu_iterate and
reads the residual norms back. Prints ONE JSON line on rank 0.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SWEEPS = 4
CFL = 50.0
# block lattice of the multi-GPU runs: one n^3 block per GPU, joined by interblock connections
LATTICE = {2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}
NEQ = 5
# algorithmic (compulsory) doubles per cell, per kernel family of one iteration
ALG_DOUBLES = {"residual": 27, "dt_diag_init": 13, "dplur_sweep": 34, "matrix_residual": 30,
               "update_norms": 20, "store_time_n": 10}


def bytes_per_cell_iter(sweeps):
    return 8 * (100 + 34 * sweeps)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons, sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([t.strip() for t in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2.0)
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = max(smax, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        # samples under load = the upper half (the sampler also sees idle gaps)
        sm.sort()
        med = float(np.median(sm[len(sm) // 2:])) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
AIR_DAT = """n: 2.5
molarMass: 28.97
vibrationalTemperature: [3056.0]
heatOfFormation: 0
referencePressure: 101325
referenceTemperature: 298.15
referenceEntropy: 0
sutherlandViscosityC1: 1.458e-6
sutherlandViscosityS: 110.4
sutherlandConductivityC1: 2.495e-3
sutherlandConductivityS: 194.0
"""
HARNESS = os.path.join(ROOT, "oracle", "_ref", "aither_dump")


def _run_reference_sample(n, iters, procs):
    """Time the UNMODIFIED reference (oracle/_ref/aither_dump, built from /root/reference against
    a single-rank MPI stub) on the bench workload at n^3 cells: `procs` independent single-rank
    processes, one per host core, each on its own n^3 block (no halo cost). Returns the list of
    per-iteration wall times of the slowest process (first iteration dropped by the caller)."""
    from aither_b200 import synthetic
    tmp = tempfile.mkdtemp(prefix="aither_ref_")
    try:
        dirs = []
        for p in range(procs):
            d = os.path.join(tmp, "p%d" % p)
            synthetic.write_case(d, "box", n, n, n, iterations=iters, solver="dplur",
                                 sweeps=SWEEPS, cfl=CFL, perturb=(p, 0.01))
            open(os.path.join(d, "air.dat"), "w").write(AIR_DAT)
            dirs.append(d)
        running = []
        for d in dirs:
            env = dict(os.environ, AITHER_INSTALL_DIRECTORY=d)
            running.append(subprocess.Popen(
                [HARNESS, "box.inp", os.path.join(d, "dump.bin"), "--iters", str(iters), "--time"],
                cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        times = []
        for pr in running:
            out, _ = pr.communicate()
            if pr.returncode != 0:
                raise RuntimeError("reference harness failed:\n" + out[-2000:])
            times.append([float(m) for m in re.findall(r"time_s ([0-9.eE+-]+)", out)])
        return [max(t[i] for t in times) for i in range(iters)]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _oracle_port_sample(n, iters):
    """Fallback when oracle/_ref is absent: the plain-C restatement, 1 core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from aither_b200 import synthetic
    prob = synthetic.box_problem(n, n, n, solver="dplur", sweeps=SWEEPS)
    lvl = oracle.OracleLevel(prob)
    out = []
    for it in range(iters):
        t0 = time.perf_counter()
        lvl.store_old_solution(it)
        lvl.iterate(CFL)
        out.append(time.perf_counter() - t0)
    lvl.close()
    return out


def cpu_baseline(n=40, iters=4):
    """Bounded CPU sample for the GPU arm's JSON line: 1 core, n^3 cells, `iters` iterations
    (first dropped). ~10-20 s of CPU work."""
    if os.path.exists(HARNESS):
        t = _run_reference_sample(n, iters, 1)
        kind = "reference"
    else:
        t = _oracle_port_sample(n, iters)
        kind = "port"
    per = float(np.mean(t[1:])) if len(t) > 1 else float(t[0])
    return {"value": n ** 3 / per / 1e6, "unit": "Mcell-iter/s", "cores": 1, "kind": kind,
            "sample": "same workload at %d^3 cells, %d iterations (first dropped), 1 process on "
                      "1 host core, %.2f s/iter" % (n, iters, per)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU code on all host cores of the box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    procs = max(1, min(cores, 64))
    n = 32
    iters = args.warmup + args.steps
    if os.path.exists(HARNESS):
        t = _run_reference_sample(n, iters, procs)
        kind = "reference"
    else:
        t = _oracle_port_sample(n, iters)
        kind, procs = "port", 1
    timed = t[args.warmup:]
    per = float(np.mean(timed))
    value = procs * n ** 3 / per / 1e6
    line = {
        "impl": "reference", "metric": "Mcell-iter/s (residual+implicit)", "value": value,
        "unit": "Mcell-iter/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_cells_note="%d independent %d^3 blocks, one per host core" %
                                  (procs, n)),
        "cpu_baseline": {"value": value, "unit": "Mcell-iter/s", "cores": procs, "kind": kind,
                         "sample": "each step = one iteration of %d single-rank reference "
                                   "processes (one per core, no MPI on the box so no halo cost), "
                                   "each on the bench workload at %d^3 cells" % (procs, n)},
        "e2e": {"value": value, "unit": "Mcell-iter/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n=256, n_cells_note=None, gpus=1, recon="thirdOrder", viscous=False,
                    turb=None, solver="dplur"):
    scheme = {"thirdOrder": "Roe+MUSCL(kappa=1/3)", "weno": "Roe+WENO5"}.get(recon, "Roe+" + recon)
    physics = ("RANS %s" % turb if turb else
               ("laminar viscous (centralFourth)" if viscous else "inviscid"))
    cfg = {"workload": "synthetic %d^3 single block per GPU, %s %s, "
                       "implicit Euler, %s x%d, CFL %g" %
                       (n, physics, scheme, solver.upper(), SWEEPS, CFL),
           "cells_per_gpu": n ** 3, "matrix_sweeps": SWEEPS,
           "l2": "inputs larger than L2 (each field %.0f MB, ~40 fields)" % (n ** 3 * 8 / 1e6),
           "parallelism": "blocks%d" % gpus}
    if gpus > 1 and gpus in LATTICE:
        cfg["parallelism"] = ("%dx%dx%d lattice of connected blocks, one per GPU; ghost layers of "
                              "state and implicit update written straight into the neighbour's "
                              "memory over NVLink (pack kernel -> peer buffer -> flag -> unpack; "
                              "AITHER_B200_HALO_P2P=0: ncclSend/ncclRecv), %d exchanges per "
                              "iteration" % (*LATTICE[gpus], SWEEPS + 2))
    if n_cells_note:
        cfg["sample"] = n_cells_note
    return cfg


def measure_configs3(args, rank, world, local, comm, dist):
    """BASELINE configs[3]'s block and scheme on every GPU of the run: one 512 x 256 x 256 block per
    GPU (the config's 1024 x 512 x 512 grid is eight of them), WENO5 + 4th-order central viscous
    fluxes, DPLUR, ghost layers exchanged over NCCL with the reference-order (multi-level) plan for
    the state; and the same block alone on this GPU (its joined faces turned into slip walls) as
    the 1-GPU time of the same run. Device time, max over ranks. AITHER_BENCH_CONFIGS3_DIMS=ni,nj,nk
    overrides the block size (tests)."""
    import aither_b200
    from aither_b200 import ctypes_abi as abi
    from aither_b200 import synthetic
    from aither_b200.problem import Block, Problem
    dims = tuple(int(v) for v in os.environ.get("AITHER_BENCH_CONFIGS3_DIMS", "512,256,256").split(","))
    splits = LATTICE[world]
    t0 = time.perf_counter()
    prob = synthetic.lattice_problem(dims, splits, only=[rank], solver="dplur", sweeps=SWEEPS,
                                     recon="weno", viscous=True, visc_recon="centralFourth",
                                     size=dims[0] * 2e-6)
    synthetic.assign_ranks(prob, world)
    setup_s = time.perf_counter() - t0
    cells = dims[0] * dims[1] * dims[2]

    def timed(lvl):
        def barrier():
            lvl.synchronize()
            dist.barrier()
            lvl.synchronize()
        lvl.run(args.warmup, CFL)
        lvl.profile_enable(True)
        barrier()
        lvl.timer_start()
        hist = lvl.run(args.steps, CFL)
        ms = lvl.timer_stop()
        barrier()
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = lvl.profile()
        lvl.profile_enable(False)
        assert np.isfinite(hist).all(), "non-finite residual history (configs3)"
        return float(t.item()) / args.steps, prof

    lvl = aither_b200.GridLevel(prob, device=local, rank=rank, n_ranks=world, block_ids=[rank],
                                nccl_comm=comm)
    lvl.enable_peer_exchange()
    ms_n, prof_n = timed(lvl)
    lvl.close()
    mine = prob.blocks[rank]
    alone = [(abi.BC_SLIP_WALL, *sf[1:7], 0) if sf[0] == abi.BC_INTERBLOCK else sf
             for sf in mine.surfaces]
    prob1 = Problem(prob.cfg, [Block(mine.ni, mine.nj, mine.nk, alone, mine.arrays,
                                     parent_block=0, global_pos=0)], [])
    lvl1 = aither_b200.GridLevel(prob1, device=local)
    ms_1, prof_1 = timed(lvl1)
    lvl1.close()
    return {
        "workload": "one %d x %d x %d block per GPU (BASELINE configs[3]: 1024 x 512 x 512 = 8 of "
                    "them), laminar viscous (centralFourth) Roe+WENO5, implicit Euler, DPLUR x%d, "
                    "CFL %g; %dx%dx%d lattice, state exchanged with the reference-order multi-level "
                    "plan" % (*dims, SWEEPS, CFL, *splits),
        "cells_per_gpu": cells, "n_gpus": world,
        "ms_per_step": ms_n, "value": cells * world / (ms_n * 1e-3) / 1e6, "unit": "Mcell-iter/s",
        "ms_per_step_1gpu_same_run": ms_1, "value_1gpu_same_run": cells / (ms_1 * 1e-3) / 1e6,
        "exchange_ms_per_step": prof_n.get("halo_pack_unpack", (0.0, 0))[0] / args.steps,
        "kernel_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in prof_n.items() if v[1]},
        "kernel_ms_per_step_1gpu": {k: round(v[0] / args.steps, 4) for k, v in prof_1.items() if v[1]},
        "setup_s": round(setup_s, 1),
    }


# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import aither_b200
    from aither_b200 import ctypes_abi as abi
    from aither_b200 import synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    comm = None
    n = args.n
    t_setup = time.perf_counter()
    if world > 1:
        # weak scaling WITH halo exchange: a lattice of `world` connected n^3 blocks, one per GPU
        # (reference `manual` decomposition, one block per rank); ghost layers of the state and of
        # the implicit update travel with ncclSend/ncclRecv inside the library every iteration
        import torch
        import torch.distributed as dist
        from aither_b200 import distributed as adist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = adist.make_comm(local)
        splits = LATTICE[world]
        if args.turb or args.solver != "dplur":
            raise SystemExit("--turb / --solver are single-GPU workloads in this round")
        extra = (dict(viscous=True, visc_recon="centralFourth", size=n * 2e-6) if args.viscous
                 else {})
        prob = synthetic.lattice_problem(n, splits, only=[rank], solver="dplur", sweeps=SWEEPS,
                                         recon=args.recon, **extra)
        synthetic.assign_ranks(prob, world)
        mine = rank
        lvl = aither_b200.GridLevel(prob, device=local, rank=rank, n_ranks=world,
                                    block_ids=[mine], nccl_comm=comm)
        # ghost exchange over NVLink peer memory (AITHER_B200_HALO_P2P=0: ncclSend / ncclRecv)
        peer_exchange = lvl.enable_peer_exchange()
    else:
        extra = dict(viscous=True, visc_recon="centralFourth", size=n * 2e-6) if args.viscous else {}
        if args.turb:
            extra = dict(turb=args.turb, limiter="vanAlbada", size=n * 1e-4)
        prob = synthetic.box_problem(n, n, n, solver=args.solver, sweeps=SWEEPS, seed=rank,
                                     recon=args.recon, **extra)
        mine = 0
        lvl = aither_b200.GridLevel(prob, device=local, rank=rank, n_ranks=world)
    cells = n ** 3
    g = prob.cfg.numGhosts
    neq = prob.neq
    state_shape = prob.blocks[mine].padded_shape(g) + (neq,)
    host_state = aither_b200.pinned_array(state_shape)
    host_state[...] = prob.blocks[mine].arrays["state"]
    # the physical cells alone (what a host really has to hand over: the ghost shell is filled at
    # the start of every iteration), for the overlapped upload
    host_interior = aither_b200.pinned_array((n, n, n, neq))
    host_interior[...] = host_state[g:g + n, g:g + n, g:g + n]
    for b in prob.blocks:          # host copies of the metrics are no longer needed
        b.arrays = {"state": None}
    t_setup = time.perf_counter() - t_setup

    def barrier():
        lvl.synchronize()
        if dist is not None:
            dist.barrier()
        lvl.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident: K iterations back to back, no host sync in between ----------------------
    lvl.run(args.warmup, CFL)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    lvl.profile_enable(True)
    barrier()
    launches0 = lvl.launch_count
    lvl.timer_start()
    hist = lvl.run(args.steps, CFL)
    ms = lvl.timer_stop()
    barrier()
    launches = lvl.launch_count - launches0
    ms = max_over_ranks(ms)
    prof = lvl.profile()
    lvl.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    assert np.isfinite(hist).all(), "non-finite residual history"

    # ---- e2e: host-owned state, H2D every step, norms D2H every step -------------------------
    # The host hands a fresh copy of the state over every step (703 MB at 256^3) and reads the
    # norms back. The copy of step n+1 is started before step n is iterated
    # (aither_gpu_upload_state_async) so that PCIe and the kernels overlap; every copy, the layout
    # conversion in front of each iteration and every read-back are inside the timed region.
    h2d = host_state.nbytes
    d2h = (neq + 1) * 8 + 32
    for it in range(2):
        lvl.upload_state(0, host_state)
        lvl.store_old_solution(it)
        lvl.iterate(CFL)
    barrier()
    t0 = time.perf_counter()
    lvl.timer_start()
    lvl.upload_interior_async(0, host_interior)
    for it in range(args.steps):
        lvl.upload_state_commit()
        if it + 1 < args.steps:
            lvl.upload_interior_async(0, host_interior)
        lvl.store_old_solution(it)
        l2, linf, mr = lvl.iterate(CFL)
    ms_e2e = lvl.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))
    # the same without overlap (synchronous aither_gpu_upload_state), reported beside it
    barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        lvl.upload_state(0, host_state)
        lvl.store_old_solution(it)
        lvl.iterate(CFL)
    lvl.synchronize()
    ms_e2e_sync = max_over_ranks((time.perf_counter() - t0) * 1e3)
    # a round trip that also brings the state back to the host (output / restart iterations)
    lvl.download_state_into(0, host_state)

    lvl.close()
    configs3 = None
    if world > 1 and args.recon == "thirdOrder" and not args.viscous and not args.no_configs3:
        configs3 = measure_configs3(args, rank, world, local, comm, dist)

    def shutdown():
        if comm is not None:
            from aither_b200 import distributed as adist
            adist.destroy_comm(comm)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        shutdown()
        return
    total_cells = cells * world
    value = total_cells * args.steps / (ms * 1e-3) / 1e6
    e2e = total_cells * args.steps / (ms_e2e * 1e-3) / 1e6
    e2e_sync = total_cells * args.steps / (ms_e2e_sync * 1e-3) / 1e6
    peak, peak_src = peaks()
    # dominant kernel family of the timed region
    alg = dict(ALG_DOUBLES)
    if prof.get("dt_diag_init", (0.0, 0))[1] == 0:
        # time step / diagonal / rhs / x0 ride in the residual kernel's epilogue: its compulsory
        # traffic is phase A + phase B of SURVEY 8d
        alg["residual"] += alg["dt_diag_init"]
    if neq != NEQ or args.solver != "dplur":
        # secondary workloads: SURVEY 8d's per-family table is for the headline (5 equations,
        # scalar diagonal); scale the state-sized entries by neq / 5 as a first-order figure
        alg = {k: int(round(v * neq / NEQ)) for k, v in alg.items()}
    if args.solver == "lusgs":
        # one launch = one half sweep of the pencil wavefront: per cell the packed
        # record (2 neq + 4 doubles; + 2 viscous slots), the behind-side faces (14; 18 viscous),
        # the ahead-sum read and the carried sum written (neq each), the update written (neq)
        visc = bool(args.viscous or args.turb)
        alg["lusgs_plane"] = (2 * neq + (6 if visc else 4)) + (18 if visc else 14) + 3 * neq
    fam = {k: v for k, v in prof.items() if v[1] > 0 and k in alg}
    top = max(fam, key=lambda k: fam[k][0])
    top_ms, top_n = fam[top]
    alg_bytes = alg[top] * 8 * cells
    achieved = alg_bytes / (top_ms / top_n * 1e-3) / 1e9
    total_fam_ms = sum(v[0] for v in prof.values())
    bpc = bytes_per_cell_iter(SWEEPS)
    iter_gbs = cells * bpc / (ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": "Mcell-iter/s (residual+implicit)", "value": value, "unit": "Mcell-iter/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, gpus=world, recon=args.recon, viscous=args.viscous,
                                  turb=args.turb, solver=args.solver),
        "e2e": {"value": max(e2e, e2e_sync), "unit": "Mcell-iter/s",
                "h2d_bytes_per_step": host_interior.nbytes if e2e >= e2e_sync else h2d,
                "d2h_bytes_per_step": d2h,
                "ms_per_step": min(ms_e2e, ms_e2e_sync) / args.steps,
                "what": "per step: the host hands over a fresh copy of the state from pinned memory, "
                        "store_old_solution, aither_gpu_iterate, norms copied "
                        "back. Two ways through the C ABI, both timed in full, the faster one is "
                        "`value`: `overlapped` = aither_gpu_upload_interior_async / _commit (the "
                        "physical cells only, 671 MB at 256^3 -- the ghost shell is filled at the "
                        "start of every iteration; the copy of the next step runs on the copy "
                        "stream during this step's kernels), `synchronous` = aither_gpu_upload_state "
                        "(ghost-padded state, 703 MB). Which one wins depends on the "
                        "box: the DMA copy slows down when the kernels keep HBM busy.",
                "mode": "overlapped" if e2e >= e2e_sync else "synchronous",
                "value_overlapped": e2e, "value_synchronous": e2e_sync},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "avg_launch_ms": top_ms / top_n,
                     "share_of_step": top_ms / total_fam_ms},
        "roofline_iteration": {"bytes_per_cell_iter": bpc, "achieved": iter_gbs,
                               "frac": iter_gbs / peak, "unit": "GB/s"},
        "kernel_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]},
        # ghost-layer exchanges of one iteration (pack + ncclSend/ncclRecv + unpack, device time)
        "exchange_ms_per_step": round(prof.get("halo_pack_unpack", (0.0, 0))[0] / args.steps, 4),
        "setup_s": round(t_setup, 1),
    }
    if configs3 is not None:
        line["configs3"] = configs3
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    ncu_traffic = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(ncu_traffic):
        tr = json.load(open(ncu_traffic))
        if top in tr and tr[top].get("cells") == cells:
            line["roofline"]["traffic"] = tr[top]["dram_bytes_per_launch"]
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--cells", dest="n", type=int, default=256,
                    help="cells per side of the block (256); use --cells under torchrun, whose own "
                         "parser treats --n as an abbreviation")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--no-configs3", action="store_true",
                    help="N > 1: skip the BASELINE configs[3] block (512 x 256 x 256 WENO5 + viscous "
                         "per GPU) that is measured after the headline workload")
    ap.add_argument("--recon", default="thirdOrder",
                    help="face reconstruction of the workload (default thirdOrder = the headline "
                         "config; weno = the inviscid half of BASELINE configs[3])")
    ap.add_argument("--viscous", action="store_true",
                    help="laminar Navier-Stokes with 4th-order central viscous reconstruction "
                         "(with --recon weno: BASELINE configs[3]'s scheme); not the headline")
    ap.add_argument("--turb", default=None, choices=["kOmegaWilcox2006", "sst2003"],
                    help="RANS workload (viscous wall on j-lo, vanAlbada limiter); not the headline")
    ap.add_argument("--solver", default="dplur", choices=["dplur", "lusgs", "bdplur", "blusgs"],
                    help="implicit solver of the workload (default dplur = the headline config)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
