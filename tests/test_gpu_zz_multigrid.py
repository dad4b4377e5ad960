"""GPU: multigrid through the C ABI against the UNMODIFIED reference: the
shipped testCases/transonicBump (Euler, DPLUR x4, CFL ramp, 3-level W cycle;
regressionTests.py:325-337), 100 iterations. Every level is a device handle; the transfer
operators are aither_gpu_mg_* (aither_b200/csrc/multigrid.cuh).

(The file name sorts last on purpose: these tests have not run on hardware in their final form,
and whatever they do must not stand in front of the verified GPU tests.)"""
import numpy as np
import pytest

import goldencheck as gc
import refcase

pytestmark = pytest.mark.gpu


# State of the evidence (profiles/r01r_multigrid_gpu.md): the one B200 run of this round, made
# BEFORE the coarse levels accumulated their diagonal over the restrictions of one W cycle
# (aither_gpu_mg_restrict, the reference's quirk), returned a first-iteration matrix residual of
# 4.097511605788737e-06; the CPU oracle with that accumulation switched off returns
# 4.097511605788731e-06 (the reference, and the oracle as committed: 4.0939659881952944e-06). The
# accumulation was added after the GPU budget of the round was spent, so this test has not run
# on hardware in its final form: non-strict xfail keeps the suite going either way.
@pytest.mark.xfail(strict=False, reason="diagonal accumulation of the coarse levels not yet "
                                        "re-run on a B200 (see profiles/r01r_multigrid_gpu.md)")
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


@pytest.mark.xfail(strict=False, reason="multigrid on the GPU has not been re-run on a B200 in its "
                                        "final form (see profiles/r01r_multigrid_gpu.md)")
@pytest.mark.parametrize("name", ["multiblockCylinder_mg2", "viscousFlatPlate_mg2"])
def test_gpu_two_level_v_cycle(name):
    """two-level V cycles: two blocks with an interblock connection on every level (LU-SGS),
    and laminar viscous terms on the coarse level"""
    import aither_b200
    d = gc.load(name)
    probs, transfers, cycle = refcase.multigrid_from_dump(d)
    mg = aither_b200.Multigrid(probs, transfers, cycle)
    href, mref, cfl = d["hist/residL2"], d["hist/matrixResid"], d["hist/cfl"]
    for it in range(30):
        mg.store_old_solution(it)
        l2, _, mr = mg.iterate(float(cfl[it]))
        scale = np.where(href[it] > 1e-20 * href[it].max(), href[it], np.inf)
        err = float(np.max(np.abs(l2 - href[it]) / scale))
        assert err <= 1e-9, (it, err, l2, href[it])
        assert abs(mr - mref[it]) <= 1e-9 * abs(mref[it]) + 1e-12 * href[it].max(), (it, mr, mref[it])
    mg.close()
