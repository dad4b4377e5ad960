"""GPU: multigrid through the C ABI against the UNMODIFIED reference: the
shipped testCases/transonicBump (Euler, DPLUR x4, CFL ramp, 3-level W cycle;
regressionTests.py:325-337), 100 iterations. Every level is a device handle; the transfer
operators are aither_gpu_mg_* (aither_b200/csrc/multigrid.cuh).

Run on a B200 by the driver at the end of round 1 (GPUTEST_r01.json: all three passed) and
again in round 2; a regression in aither_gpu_mg_* turns the suite red."""
import numpy as np
import pytest

import goldencheck as gc
import refcase

pytestmark = pytest.mark.gpu


def test_gpu_transonic_bump_three_level_w_cycle():
    import aither_b200
    d = gc.load("transonicBump")
    probs, transfers, cycle = refcase.multigrid_from_dump(d)
    mg = aither_b200.Multigrid(probs, transfers, cycle)
    href, mref, cfl = d["hist/residL2"], d["hist/matrixResid"], d["hist/cfl"]
    mine = np.zeros((100, probs[0].neq))
    for it in range(100):
        mg.store_old_solution(it)
        l2, _, mr = mg.iterate(float(cfl[it]))
        mine[it] = l2
        scale = np.where(href[it] > 1e-20 * href[it].max(), href[it], np.inf)
        err = float(np.max(np.abs(l2 - href[it]) / scale))
        assert err <= 1e-9, (it, err, l2, href[it])
        assert abs(mr - mref[it]) <= 1e-9 * abs(mref[it]), (it, mr, mref[it])
    mg.close()
    norm = gc.normalised_history(mine)[99]
    for e, gv in enumerate([2.6152e-02, 1.5984e-02, 9.6803e-03, None, 1.9215e-02]):
        if gv is not None:
            assert abs(norm[e] - gv) <= 0.01 * gv, (e, norm[e], gv)


@pytest.mark.parametrize("name,iters", [("multiblockCylinder_mg2", 30), ("viscousFlatPlate_mg2", 30),
                                        ("turbFlatPlate_mg2", 12)])
def test_gpu_two_level_v_cycle(name, iters):
    """two-level V cycles: two blocks with an interblock connection on every level (LU-SGS),
    laminar viscous terms on the coarse level, and RANS on the coarse level (k-omega Wilcox 2006:
    the wall omega of a level comes from the viscosity its own previous evaluation stored)"""
    import aither_b200
    d = gc.load(name)
    probs, transfers, cycle = refcase.multigrid_from_dump(d)
    mg = aither_b200.Multigrid(probs, transfers, cycle)
    href, mref, cfl = d["hist/residL2"], d["hist/matrixResid"], d["hist/cfl"]
    for it in range(iters):
        mg.store_old_solution(it)
        l2, _, mr = mg.iterate(float(cfl[it]))
        scale = np.where(href[it] > 1e-20 * href[it].max(), href[it], np.inf)
        err = float(np.max(np.abs(l2 - href[it]) / scale))
        assert err <= 1e-9, (it, err, l2, href[it])
        assert abs(mr - mref[it]) <= 1e-9 * abs(mref[it]) + 1e-12 * href[it].max(), (it, mr, mref[it])
    mg.close()
