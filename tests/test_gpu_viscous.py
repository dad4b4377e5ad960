"""Laminar Navier-Stokes on the GPU (-m gpu) against the CPU oracle on product-side synthetic
cases (the oracle itself is pinned to the reference on viscousFlatPlate / box_visc4 / box_visc_iso,
tests/test_oracle_pinned.py; the GPU path is compared with those reference dumps directly in
tests/test_gpu_golden.py). Covers what the goldens do not: LU-SGS + 4th-order viscous
reconstruction, MUSCL + DPLUR, and viscous flow across block connections.
"""
import numpy as np
import pytest

import goldencheck as gc
import oracle
from aither_b200 import ctypes_abi as abi
from aither_b200 import synthetic

pytestmark = pytest.mark.gpu

CASES = [
    dict(solver="dplur", sweeps=3, recon="thirdOrder", limiter="none", visc_recon="central"),
    dict(solver="lusgs", sweeps=2, recon="weno", limiter="none", visc_recon="centralFourth"),
    dict(solver="dplur", sweeps=2, recon="thirdOrder", limiter="vanAlbada",
         visc_recon="centralFourth", wall=("isothermal", 310.0)),
]


def rel(a, b):
    return gc.rel(a, b)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(str(v) for v in c.values()))
def test_viscous_phases_match_oracle(case):
    check_viscous_phases(case, (18, 11, 9), 2e-5)


def test_viscous_phases_match_oracle_multi_tile():
    """BASELINE configs[3]'s scheme (WENO5 + 4th-order central viscous fluxes, DPLUR) on a
    70 x 40 x 40 block: three tiles in i, five in j, several k-chunks of every kernel."""
    check_viscous_phases(dict(solver="dplur", sweeps=2, recon="weno", limiter="none",
                              visc_recon="centralFourth"), (70, 40, 40), 70 * 2e-6)


def check_viscous_phases(case, dims, size):
    import aither_b200
    prob = synthetic.box_problem(*dims, seed=21, amplitude=0.02, viscous=True, size=size,
                                 **case)
    g = prob.cfg.numGhosts
    gpu, ref = aither_b200.GridLevel(prob), oracle.OracleLevel(prob)
    for lvl in (gpu, ref):
        lvl.store_old_solution(0)
        lvl.get_boundary_conditions()
    m = gc.non_corner_mask(gpu.field(0, abi.FIELD_STATE).shape[:3], g)
    assert rel(gpu.field(0, abi.FIELD_STATE)[m], ref.field(0, abi.FIELD_STATE)[m]) <= 1e-13
    for lvl in (gpu, ref):
        lvl.calc_residual()
    # ghost cells as the viscous fluxes saw them, viscosity, residual, spectral radius
    assert rel(gpu.field(0, abi.FIELD_STATE)[m], ref.field(0, abi.FIELD_STATE)[m]) <= 1e-13
    assert rel(gpu.field(0, abi.FIELD_VISCOSITY)[m], ref.field(0, abi.FIELD_VISCOSITY)[m]) <= 1e-13
    assert rel(gpu.field(0, abi.FIELD_RESIDUAL), ref.field(0, abi.FIELD_RESIDUAL)) <= 1e-12
    assert rel(gpu.field(0, abi.FIELD_SPEC_RADIUS)[..., :1],
               ref.field(0, abi.FIELD_SPEC_RADIUS)[..., :1]) <= 1e-13
    for lvl in (gpu, ref):
        lvl.calc_time_step(30.0)
        lvl.invert_diagonal()
        lvl.initialize_matrix_update()
    for f in (abi.FIELD_DT, abi.FIELD_DIAG, abi.FIELD_DIAG_INV):
        assert rel(gpu.field(0, f), ref.field(0, f)) <= 1e-13
    mg, mr = gpu.relax(), ref.relax()
    cut = lambda a: a[g:-g, g:-g, g:-g]
    assert rel(cut(gpu.field(0, abi.FIELD_UPDATE)), cut(ref.field(0, abi.FIELD_UPDATE))) <= 1e-11
    assert abs(mg - mr) <= 1e-9 * abs(mr)
    (l2g, _), (l2r, _) = gpu.update_blocks(), ref.update_blocks()
    assert np.all(np.abs(l2g - l2r) <= 1e-12 * np.abs(l2r))
    assert rel(cut(gpu.field(0, abi.FIELD_STATE)), cut(ref.field(0, abi.FIELD_STATE))) <= 1e-12
    gpu.close()
    ref.close()


@pytest.mark.parametrize("solver,sweeps", [("dplur", 3), ("lusgs", 1)])
def test_viscous_multiblock_history_matches_oracle(solver, sweeps):
    """2x2x2 connected blocks, viscous wall on the j-lo side: ghost exchange + edge cells +
    viscous terms together; the oracle runs the same decomposition."""
    import aither_b200
    prob = synthetic.box_problem(16, 12, 10, seed=23, amplitude=0.02, viscous=True, size=2e-5,
                                 solver=solver, sweeps=sweeps, recon="weno",
                                 visc_recon="centralFourth")
    sp = synthetic.split_problem(prob, (2, 2, 2))
    gpu, ref = aither_b200.GridLevel(sp), oracle.OracleLevel(sp)
    for it in range(6):
        gpu.store_old_solution(it)
        ref.store_old_solution(it)
        l2g, _, mrg = gpu.iterate(40.0)
        l2r, _, mrr = ref.iterate(40.0)
        assert np.all(np.abs(l2g - l2r) <= 1e-9 * np.abs(l2r)), (it, l2g, l2r)
        assert abs(mrg - mrr) <= 1e-9 * abs(mrr)
    g = sp.cfg.numGhosts
    for b in range(len(sp.blocks)):
        sg = gpu.field(b, abi.FIELD_STATE)[g:-g, g:-g, g:-g]
        sr = ref.field(b, abi.FIELD_STATE)[g:-g, g:-g, g:-g]
        assert np.abs(sg - sr).max() <= 1e-12 * np.abs(sr).max()
    gpu.close()
    ref.close()


def test_viscous_lattice_matches_oracle():
    """The multi-GPU workload of BASELINE configs[3]'s scheme (WENO5 + 4th-order central viscous
    fluxes, DPLUR) as a 2x2x1 lattice of connected blocks in one process: the state exchange takes
    the full, reference-order plan (edge ghost cells are read by the viscous stencils), the
    implicit update the single-level one."""
    import aither_b200
    prob = synthetic.lattice_problem(10, (2, 2, 1), viscous=True, visc_recon="centralFourth",
                                     recon="weno", size=10 * 2e-6, sweeps=3)
    gpu, ref = aither_b200.GridLevel(prob), oracle.OracleLevel(prob)
    for it in range(5):
        gpu.store_old_solution(it)
        ref.store_old_solution(it)
        l2g, _, mrg = gpu.iterate(30.0)
        l2r, _, mrr = ref.iterate(30.0)
        assert np.all(np.abs(l2g - l2r) <= 1e-9 * np.abs(l2r)), (it, l2g, l2r)
        assert abs(mrg - mrr) <= 1e-9 * abs(mrr)
    gpu.close()
    ref.close()
