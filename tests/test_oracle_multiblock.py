"""CPU: the oracle's connection swaps against the UNMODIFIED reference (multiblockCylinder golden)
and its decomposition invariance on a split synthetic box."""
import numpy as np

import goldencheck as gc
import refcase
import oracle
from aither_b200 import ctypes_abi as abi
from aither_b200 import synthetic

TOL = dict(ghosts=1e-12, residual=1e-12, specRadius=1e-14, dt=1e-14, diag=1e-14, x0=1e-13,
           x=1e-12, matrixResid=1e-10, state=1e-13, l2=1e-13)


def test_oracle_multiblock_phases_match_reference():
    d = gc.load("multiblockCylinder")
    for it in gc.full_iterations(d):
        gc.check_phases(oracle.OracleLevel, d, it, TOL)


def test_oracle_multiblock_history_matches_reference():
    d = gc.load("multiblockCylinder")
    assert gc.check_history(oracle.OracleLevel, d, 100, 1e-9, name="multiblockCylinder") <= 1e-9


def test_oracle_split_box_equals_whole_box():
    prob = synthetic.box_problem(12, 10, 8, seed=3, amplitude=0.02, sweeps=3)
    sp = synthetic.split_problem(prob, (2, 2, 2))
    a, b = oracle.OracleLevel(prob), oracle.OracleLevel(sp)
    for it in range(3):
        a.store_old_solution(it)
        b.store_old_solution(it)
        la, _, _ = a.iterate(30.0)
        lb, _, _ = b.iterate(30.0)
        assert np.all(np.abs(la - lb) <= 1e-13 * np.abs(la))
    g = 2
    sa = a.field(0, abi.FIELD_STATE)[g:-g, g:-g, g:-g]
    sb = synthetic.reassemble(sp, (2, 2, 2), [b.field(i, abi.FIELD_STATE)[g:-g, g:-g, g:-g]
                                              for i in range(8)])
    assert np.abs(sa - sb).max() <= 1e-14 * np.abs(sa).max()
    a.close()
    b.close()


def test_oracle_uniform_flow_all_orientations():
    """testCases/uniformFlow joins its 10 blocks through all 8 patch orientations (as Euler, from
    a perturbed state): ghost cells after the swap are bit-identical to the reference's."""
    d = gc.load("uniformFlow_euler")
    assert sorted(set(int(c[26]) for c in d["connections"])) == [1, 2, 3, 4, 5, 6, 7, 8]
    for it in gc.full_iterations(d):
        out = gc.check_phases(oracle.OracleLevel, d, it, TOL)
        assert out["ghosts"] == 0.0
    assert gc.check_history(oracle.OracleLevel, d, 20, 1e-9) <= 1e-9


def test_oracle_shock_tube_bdf2_dual_time():
    """testCases/shockTube (regressionTests.py:271-287): two blocks, WENO, BDF2 with dual time
    stepping, 5 nonlinear iterations per step -- 40 steps = 200 history records."""
    d = gc.load("shockTube")
    gc.check_phases(oracle.OracleLevel, d, 0, TOL)
    assert gc.check_history(oracle.OracleLevel, d, 200, 1e-9) <= 1e-9


def test_oracle_uniform_flow_rans():
    """The shipped testCases/uniformFlow (SST 2003, LU-SGS x2, 10 blocks through all 8 patch
    orientations) from a perturbed state: eddy viscosity / blending functions swapped across
    connections and used by the implicit off-diagonals of the neighbouring block."""
    d = gc.load("uniformFlow_rans")
    gc.check_phases(oracle.OracleLevel, d, 0, dict(TOL, turb=1e-12))
    assert gc.check_history(oracle.OracleLevel, d, 20, 1e-9) <= 1e-9


def test_oracle_wall_law():
    """testCases/wallLaw (regressionTests.py:430-446): SST 2003 + BLU-SGS, two blocks, adiabatic
    viscous wall with the wall law -- y+ by Ridder's method per wall face, prescribed wall shear
    stress, k and omega ghost cells from the law. The shipped uniform start for 20 iterations
    (with the reference's own regression golden), and the same case from a perturbed state phase
    by phase."""
    d = gc.load("wallLaw_cloud")
    gc.check_phases(oracle.OracleLevel, d, 0, dict(TOL, turb=1e-12, diag=1e-13))
    assert gc.check_history(oracle.OracleLevel, d, 6, 1e-9) <= 1e-9
    d = gc.load("wallLaw")
    assert gc.check_history(oracle.OracleLevel, d, 20, 1e-9, name="wallLaw") <= 1e-9


def test_oracle_wall_data():
    """The oracle's wall variables (y+, wall shear stress, heat flux, wall temperature /
    viscosities / density, friction velocity, k, omega per wall face) after the first residual
    evaluation of testCases/wallLaw from a perturbed state, against the reference's own wallData
    (include/wallData.hpp:40-57) -- the CPU pin of what aither_gpu_download_wall_data returns
    (tests/test_gpu_multiblock.py::test_wall_data_matches_reference)."""
    d = gc.load("wallLaw_cloud")
    prob = refcase.problem_from_dump(d, state_key="state0")
    lvl = oracle.OracleLevel(prob)
    lvl.store_old_solution(0)
    lvl.get_boundary_conditions()
    lvl.calc_residual()
    seen = 0
    for bb, blk in enumerate(prob.blocks):
        ww = 0
        while "b%d/wall%d/surface" % (bb, ww) in d:
            sf = [int(v) for v in d["b%d/wall%d/surface" % (bb, ww)]]
            ref = d["b%d/wall%d/vars@it0" % (bb, ww)]
            idx = [n for n, s in enumerate(blk.surfaces) if list(s[1:7]) == sf[:6]]
            assert len(idx) == 1, (sf, blk.surfaces)
            mine = lvl.wall_data(bb, idx[0])
            assert mine is not None and mine.shape == ref.shape, (None if mine is None else mine.shape, ref.shape)
            # the three shear-stress components are one vector
            err = gc.rel(mine, ref, groups=([1, 2, 3],))
            assert err <= 1e-11, (bb, ww, err)
            seen += 1
            ww += 1
    assert seen >= 1
    lvl.close()


def test_oracle_convecting_vortex_nonreflecting():
    """testCases/convectingVortex (regressionTests.py:498-514): laminar, BDF2 dual time stepping
    (10 nonlinear iterations per step), LU-SGS, periodic pair, non-reflecting `inlet` and
    `pressureOutlet`: LODI relaxation towards the boundary state with the state at time n, the
    time step and the pressure / velocity gradients of the boundary-adjacent cell from the
    previous evaluation, and the average / maximum normal Mach number of the patch
    (src/ghostStates.cpp:435-466, :614-643; src/procBlock.cpp:6235-6261). 4 time steps."""
    d = gc.load("convectingVortex")
    assert any(int(d["cfg/bc%d/isNonreflecting" % q][0]) for q in range(3))
    gc.check_phases(oracle.OracleLevel, d, 0, TOL)
    assert gc.check_history(oracle.OracleLevel, d, 40, 1e-9) <= 1e-9


def test_oracle_couette():
    """The shipped testCases/couette: laminar, LU-SGS at CFL 1e5, periodic pair in j, isothermal
    walls with the upper one moving (viscousWall with a velocity). CFL 1e5 from a uniform start:
    a nearly singular implicit system, as viscousFlatPlate."""
    d = gc.load("couette")
    gc.check_phases(oracle.OracleLevel, d, 0, dict(TOL, x=1e-10, matrixResid=1e-9))
    assert gc.check_history(oracle.OracleLevel, d, 30, 1e-9) <= 1e-9


def test_oracle_rae2822():
    """The shipped testCases/rae2822: SST 2003, LU-SGS, C-mesh (368 x 64 cells) whose wake cut is
    an interblock connection of the block with itself; referenceLength 0.3048. The fixture
    (14 MB) is not committed: `python tests/golden/make_golden.py rae2822` regenerates it."""
    import os
    import pytest
    if not os.path.exists(os.path.join(gc.GOLDEN_DIR, "rae2822.npz")):
        pytest.skip("tests/golden/rae2822.npz has not been generated")
    d = gc.load("rae2822")
    assert gc.check_history(oracle.OracleLevel, d, 10, 1e-9) <= 1e-9


def test_oracle_periodic_connection():
    """A periodic pair (the block's own i-lo and i-hi faces, `periodic(startTag; endTag;
    translation)`, reference src/boundaryConditions.cpp:2224-2300): the block exchanges ghost layers
    with itself. Euler + DPLUR and laminar + LU-SGS, against the reference's dumps."""
    for name in ("box_periodic", "box_periodic_visc"):
        d = gc.load(name)
        assert int(d["connections"][0][27]) == 0  # isInterblock = 0: periodic
        for it in gc.full_iterations(d):
            out = gc.check_phases(oracle.OracleLevel, d, it, dict(TOL, diag=1e-13))
            assert out["ghosts"] == 0.0
        assert gc.check_history(oracle.OracleLevel, d, 12, 1e-9) <= 1e-9
