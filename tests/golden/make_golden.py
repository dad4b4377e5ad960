"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/aither_dump).

TEST INFRASTRUCTURE ONLY. Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py [case ...]

Each fixture is one compressed .npz holding, for a small case, the reference's set-up (config,
grid metrics, initial state, connections), its arrays after every phase of selected iterations
(`--full`), and its full-precision residual history. The tests (tests/test_oracle_pinned.py on
CPU, tests/test_gpu_golden.py on the GPU box) read only these files: /root/reference does not
exist on the GPU box.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refcase  # noqa: E402
from aither_b200 import synthetic  # noqa: E402

REF_CASES = "/root/reference/testCases"

# name -> (how to stage, iterations, iterations dumped phase by phase)
CASES = {
    # reference regression case (testCases/regressionTests.py:233-249): Euler, Roe, MUSCL k=1/3,
    # LU-SGS x1, slipWall + stagnationInlet + pressureOutlet, 33x2x41 nodes
    "subsonicCylinder": dict(src="subsonicCylinder", iters=100, full=(0, 50), edits={}),
    # supersonic wedge: supersonicInflow / supersonicOutflow BCs, vanAlbada; the shipped .inp is
    # explicit Euler, which is not on the hot path -> edited to implicit Euler + DPLUR x3
    "supersonicWedge": dict(src="supersonicWedge", iters=30, full=(0, 10),
                            edits={"timeIntegration": "implicitEuler", "cflStart": "5.0",
                                   "cflMax": "5.0", "matrixSolver": "dplur",
                                   "matrixSweeps": "3", "matrixRelaxation": "1.0"}),
    # synthetic box of SURVEY 8d (the bench workload, small): DPLUR x4
    "box_dplur": dict(synthetic=dict(ni=14, nj=10, nk=8, solver="dplur", sweeps=4), iters=30,
                      full=(0, 5)),
    "box_lusgs_va": dict(synthetic=dict(ni=12, nj=9, nk=7, solver="lusgs", sweeps=2,
                                        limiter="vanAlbada", flux="ausm"), iters=20, full=(0, 5)),
    "box_weno": dict(synthetic=dict(ni=12, nj=8, nk=8, solver="dplur", sweeps=2, recon="weno"),
                     iters=12, full=(0, 4)),
    # laminar Navier-Stokes (regressionTests.py:340-358): viscous wall (adiabatic), central
    # viscous reconstruction, LU-SGS, CFL 1e4; 65x65x2 nodes
    "viscousFlatPlate": dict(src="viscousFlatPlate", iters=100, full=(0, 1), edits={},
                             drop=("velocityGrad@", "diagRaw@", "temperature@")),
    # synthetic viscous boxes, 20 micron wide so that the viscous fluxes are ~10 % of the
    # inviscid ones: WENO + 4th-order central viscous reconstruction, adiabatic wall, DPLUR
    "box_visc4": dict(synthetic=dict(ni=12, nj=10, nk=8, solver="dplur", sweeps=3, recon="weno",
                                     viscous=True, visc_recon="centralFourth", size=2e-5),
                      iters=12, full=(0, 4)),
    # MUSCL + 2nd-order central, isothermal wall, LU-SGS
    "box_visc_iso": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="lusgs", sweeps=2,
                                        limiter="vanAlbada", viscous=True, size=2e-5,
                                        wall=("isothermal", 300.0)), iters=12, full=(0, 4)),
    # reference regression case (regressionTests.py:271-287): two blocks, WENO, BDF2 with dual
    # time stepping (5 nonlinear iterations per time step), LU-SGS
    "shockTube": dict(src="shockTube", iters=40, full=(0,), edits={}),
    # testCases/uniformFlow: 10 blocks joined through all 8 patch orientations. The shipped case is
    # RANS (run-only check, regressionTests.py:482-496); here it runs as Euler from a perturbed
    # state so that every orientation of the ghost exchange carries non-trivial data
    "uniformFlow_euler": dict(src="uniformFlow", iters=20, full=(0, 5), cloud=(17, 0.01),
                              edits={"equationSet": "euler", "turbulenceModel": "none",
                                     "iterations": "20",
                                     "initialConditions": "<icState(tag=-1; file=ic.dat)>"}),
    # block-matrix solvers (full flux Jacobians on the diagonal, Gauss-Jordan inverse, block
    # off-diagonals): BDPLUR on the Euler box, BLU-SGS on a laminar box (thin-shear-layer viscous
    # Jacobian) and on the SST box (2x2 turbulence blocks, source Jacobian)
    "box_bdplur": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="bdplur", sweeps=3), iters=12,
                       full=(0, 4)),
    "box_blusgs_visc": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="blusgs", sweeps=2,
                                           limiter="vanAlbada", viscous=True, size=2e-5),
                            iters=12, full=(0, 4)),
    "box_sst_blusgs": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="blusgs", sweeps=2,
                                          turb="sst2003", limiter="vanAlbada", size=1e-3),
                           iters=12, full=(0,)),
    # approximateRoe flux Jacobian (off-diagonals as Roe-flux changes), Euler, LU-SGS x2
    "box_roe_jac": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="lusgs", sweeps=2,
                                       jac="approximateRoe"), iters=12, full=(0, 4)),
    # the shipped uniformFlow case itself: SST 2003, LU-SGS x2, 10 blocks / 8 orientations, from
    # a perturbed state (turbulence included; CFL 20 instead of 1000, which diverges from such a
    # state): pins the exchange of eddy viscosity and blending
    # functions across connections (gridLevel::SwapEddyViscAndGradients, SwapTurbVars)
    "uniformFlow_rans": dict(src="uniformFlow", iters=20, full=(0,), cloud=(19, 0.01),
                             turb="sst2003",
                             edits={"iterations": "20", "cflStart": "20", "cflMax": "20",
                                    "initialConditions": "<icState(tag=-1; file=ic.dat)>"},
                             drop=("diagRaw@", "temperature@", "state@it0.start", "x0@",
                                   "velocityGrad@")),
    # synthetic three-species boxes (H2O / H2 / N2 with perturbed mass fractions, Schmidt-number
    # diffusion, Wilke mixing): laminar + AUSMPW+ + minmod + 4th-order viscous reconstruction +
    # LU-SGS, SST + DPLUR, and inviscid + DPLUR (AUSMPW+ as the shipped multi-species case)
    "box_mix3_visc": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="lusgs", sweeps=2, flux="ausm",
                                         limiter="minmod", viscous=True,
                                         visc_recon="centralFourth", size=2e-5,
                                         species={"H2O": 0.233, "H2": 0.001, "N2": 0.766}),
                          iters=12, full=(0, 4)),
    "box_mix3_sst": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="dplur", sweeps=3, flux="ausm",
                                        limiter="vanAlbada", turb="sst2003", size=1e-3, cfl=5.0,
                                        species={"H2O": 0.233, "H2": 0.001, "N2": 0.766}),
                         iters=12, full=(0,)),
    "box_mix3_euler": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="dplur", sweeps=3, cfl=5.0,
                                          flux="ausm", limiter="minmod",
                                          species={"H2O": 0.233, "H2": 0.001, "N2": 0.766}),
                           iters=12, full=(0, 4)),
    # options no other fixture holds: WENO-Z + Crank-Nicolson with a global time step and
    # matrixRelaxation 1.1; first-order (constant) reconstruction; constant-heat-flux wall
    "box_wenoz_cn": dict(synthetic=dict(ni=12, nj=8, nk=8, solver="dplur", sweeps=3, recon="wenoZ",
                                        overrides={"timeIntegration": "crankNicholson",
                                                   "timeStep": "2.0e-5",
                                                   "matrixRelaxation": "1.1"}),
                         iters=10, full=(0, 4)),
    "box_first_order": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="lusgs", sweeps=2,
                                           recon="constant"), iters=10, full=(0, 4)),
    "box_visc_heatflux": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="dplur", sweeps=3,
                                             limiter="minmod", viscous=True, size=2e-5,
                                             wall=("heatFlux", 2.0e5)), iters=10, full=(0, 4)),
    # subsonic `inlet` / `pressureOutlet` pair (reflecting forms) with WENO and LU-SGS
    "box_inlet_outlet": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="lusgs", sweeps=2,
                                            recon="weno", inlet_outlet=True), iters=10,
                             full=(0, 4)),
    # periodic connection (the block's i-lo and i-hi faces, translation one box length): the ghost
    # exchange of a block with itself; Euler + DPLUR and laminar + LU-SGS
    "box_periodic": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="dplur", sweeps=3,
                                        periodic=True), iters=12, full=(0, 4)),
    "box_periodic_visc": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="lusgs", sweeps=2,
                                             limiter="vanAlbada", viscous=True, size=2e-5,
                                             periodic=True), iters=12, full=(0,)),
    # two species (H2 / N2), laminar, AUSMPW+, DPLUR
    "box_mix2_visc": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="dplur", sweeps=3, flux="ausm",
                                         limiter="vanAlbada", viscous=True, size=2e-5, cfl=5.0,
                                         species={"H2": 0.05, "N2": 0.95}),
                          iters=12, full=(0,)),
    # the Roe flux with three species (the reference's Roe scheme is not stable for this mixture
    # beyond a few iterations, the heats of formation being large: 3 iterations of the laminar box)
    "box_mix3_roe": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="lusgs", sweeps=2, cfl=5.0,
                                        flux="roe", limiter="minmod", viscous=True, size=2e-5,
                                        species={"H2O": 0.233, "H2": 0.001, "N2": 0.766}),
                         iters=3, full=(0,)),
    # reference regression case (regressionTests.py:515-540), BASELINE configs[4]'s base: three
    # species (H2O, H2, N2) with Schmidt-number diffusion and Wilke mixing, SST 2003, AUSMPW+,
    # minmod, 4th-order central viscous reconstruction, LU-SGS x2, 5 blocks
    "supersonicMixing": dict(src="supersonicMixing", iters=20, full=(0,), edits={},
                             fluids=("H2O", "H2", "N2"),
                             drop=("diagRaw@", "temperature@", "state@it0.start", "x0@",
                                   "velocityGrad@", "tkeGrad@", "omegaGrad@", "f2@", "diagInv@",
                                   "matrixResid@", "dt@", "state@it0.bc")),
    # BASELINE configs[2] on the shipped grid: turbFlatPlate with `turbulenceModel: sst2003` and
    # `matrixSolver: blusgs` (the config string differs from the shipped .inp)
    "turbFlatPlate_sst_blusgs": dict(src="turbFlatPlate", iters=20, full=(), 
                                     edits={"turbulenceModel": "sst2003", "matrixSolver": "blusgs"},
                                     drop=("state@",)),
    # BASELINE configs[4] on the shipped grid: supersonicMixing with BDF2 dual time stepping
    # (3 nonlinear iterations per step)
    "supersonicMixing_bdf2": dict(src="supersonicMixing", iters=8, full=(),
                                  fluids=("H2O", "H2", "N2"),
                                  edits={"timeIntegration": "bdf2", "timeStep": "1.0e-7",
                                         "dualTimeCFL": "100", "nonlinearIterations": "3"},
                                  drop=("state@",)),
    # reference regression case transonicBump (regressionTests.py:325-337) on a single grid level:
    # first-order-upwind-biased MUSCL (kappa = -1) + vanAlbada, DPLUR x4, CFL ramp 1000 -> 10000
    # (cflStep), patched slip walls; the shipped case adds a 3-level W-cycle multigrid, which is a
    # SURVEY 8(f) row, so `multigridLevels: 1` here
    "transonicBump_sg": dict(src="transonicBump", iters=100, full=(0, 50),
                             edits={"multigridLevels": "1"}),
    # the shipped transonicBump itself: 3-level W-cycle multigrid (full-approximation-storage
    # coarse grid correction: volume-weighted restriction of state and update, summed matrix
    # residual as forcing, trilinear prolongation); regressionTests.py:325-337
    "transonicBump": dict(src="transonicBump", iters=100, full=(), edits={}, drop=("state@",)),
    # multigrid beyond the shipped case: two blocks with an interblock connection on every level
    # (AUSMPW+, LU-SGS, 2-level V cycle) and laminar viscous terms on the coarse level
    # (viscousFlatPlate, LU-SGS, 2 levels)
    "multiblockCylinder_mg2": dict(src="multiblockCylinder", iters=30, full=(),
                                   edits={"multigridLevels": "2", "multigridCycle": "V"},
                                   drop=("state@",)),
    # RANS on the coarse level (k-omega Wilcox 2006, wall omega from the stored viscosity of that
    # level's previous evaluation), 2-level V cycle
    "turbFlatPlate_mg2": dict(src="turbFlatPlate", iters=12, full=(),
                              edits={"multigridLevels": "2", "multigridCycle": "V"},
                              drop=("state@",)),
    "viscousFlatPlate_mg2": dict(src="viscousFlatPlate", iters=30, full=(),
                                 edits={"multigridLevels": "2", "multigridCycle": "V"},
                                 drop=("state@",)),
    # two-block cylinder with interblock halo, AUSMPW+ (regressionTests.py:252-268)
    "multiblockCylinder": dict(src="multiblockCylinder", iters=100, full=(0, 1), edits={}),
    # RANS, reference regression case (regressionTests.py:364-381): k-omega Wilcox 2006, LU-SGS,
    # CFL 1e5, viscous wall with the omega wall BC, stagnation inlet / pressure outlet with
    # farfield turbulence; 136x96 cells. Phases only at iteration 0: the wall omega BC reads the
    # viscosity stored by the previous evaluation, which a restart cannot reproduce
    "turbFlatPlate": dict(src="turbFlatPlate", iters=20, full=(0,), edits={},
                          drop=("diagRaw@", "temperature@", "state@it0.start", "x0@")),
    # reference regression case wallLaw (regressionTests.py:430-446): SST 2003 + BLU-SGS, two
    # blocks, adiabatic viscous wall with the wall law (y+ root by Ridder's method per wall face,
    # prescribed wall stress, k and omega from the law); phases at iteration 0 only
    "wallLaw": dict(src="wallLaw", iters=20, full=(), edits={}, drop=("state@",)),
    # the same case from a perturbed state (1 % noise about the shipped initial condition), so
    # that the phase-by-phase dump of iteration 0 is not a uniform flow
    "wallLaw_cloud": dict(src="wallLaw", iters=6, full=(0,), cloud=(23, 0.01), turb="sst2003",
                          ic=dict(density=1.2256, velocity=[0.0, 0.0, 75.0], pressure=101325.0),
                          edits={"initialConditions": "<icState(tag=-1; file=ic.dat)>"},
                          drop=("diagRaw@", "temperature@", "state@it0.start", "x0@",
                                "velocityGrad@", "f2@")),
    # shipped case couette: laminar, LU-SGS at CFL 1e5, periodic pair in j, isothermal walls, the
    # upper one moving (viscousWall with a velocity)
    "couette": dict(src="couette", iters=30, full=(0,), edits={},
                    drop=("diagRaw@", "temperature@", "state@it0.start", "x0@")),
    # shipped case rae2822: SST 2003, LU-SGS, C-mesh whose wake cut is an interblock connection of
    # the block with itself, referenceLength 0.3048
    "rae2822": dict(src="rae2822", iters=10, full=(), edits={}, drop=("state@",)),
    # reference regression case convectingVortex (regressionTests.py:498-514): laminar, BDF2 dual
    # time stepping (10 nonlinear iterations per step), LU-SGS, periodic pair, non-reflecting
    # inlet and pressure outlet (LODI relaxation with the state at time n, the time step, the
    # pressure / velocity gradients of the previous evaluation and the patch Mach numbers)
    "convectingVortex": dict(src="convectingVortex", iters=40, full=(0,), edits={}, fluids=("N2",),
                             drop=("diagRaw@", "temperature@", "state@it0.start", "x0@")),
    # Euler box with non-reflecting inlet / outlet: the gradients come from the reference's
    # gradient-only pass (CalcGradsI/J/K); and a laminar box with the wall law (1 cm).
    # (vanAlbada: at iteration 0 the time step is 0 and the outlet ghost cells equal the interior
    # cell to the last bit, where the unlimited eps-regularised MUSCL ratio is discontinuous)
    "box_nonrefl_euler": dict(synthetic=dict(ni=12, nj=9, nk=8, solver="dplur", sweeps=3,
                                             limiter="vanAlbada", inlet_outlet=True,
                                             nonreflecting=1.0), iters=10,
                              full=(0,)),
    "box_walllaw_laminar": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="lusgs", sweeps=2,
                                               limiter="vanAlbada", viscous=True, size=1e-2,
                                               wall_law=True), iters=10, full=(0,)),
    # synthetic SST boxes (1 cm: y+ of the wall cells ~ 50) with the wall law on an isothermal
    # and on a constant-heat-flux wall (wallLaw::IsothermalBCs / HeatFluxBCs; the shipped case is
    # adiabatic), DPLUR and BLU-SGS
    "box_walllaw_isothermal": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="dplur", sweeps=3,
                                                  turb="sst2003", limiter="vanAlbada", size=1e-2,
                                                  cfl=5.0, wall=("isothermal", 320.0),
                                                  wall_law=True), iters=10, full=(0,)),
    "box_walllaw_heatflux": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="blusgs", sweeps=2,
                                                turb="sst2003", limiter="vanAlbada", size=1e-2,
                                                cfl=5.0, wall=("heatFlux", 2.0e4),
                                                wall_law=True), iters=10, full=(0,)),
    # synthetic RANS boxes (1 mm, viscous wall on j-lo): SST 2003 + DPLUR + 4th-order
    # viscous reconstruction, and k-omega Wilcox 2006 + LU-SGS + AUSMPW+ + minmod at CFL 5
    "box_sst": dict(synthetic=dict(ni=12, nj=10, nk=8, solver="dplur", sweeps=3, turb="sst2003",
                                   limiter="vanAlbada", visc_recon="centralFourth", size=1e-3),
                    iters=12, full=(0,)),
    "box_kw": dict(synthetic=dict(ni=10, nj=9, nk=8, solver="lusgs", sweeps=2, flux="ausm",
                                  limiter="minmod", cfl=5.0, turb="kOmegaWilcox2006",
                                  size=1e-3), iters=12, full=(0,)),
}

DROP = ("nodes", "fCenterI", "fCenterJ", "fCenterK")


def generate(name):
    spec = CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        if "synthetic" in spec:
            kw = dict(spec["synthetic"])
            ni, nj, nk = kw.pop("ni"), kw.pop("nj"), kw.pop("nk")
            # perturbed initial state: a uniform start leaves cells that are equal to the last
            # bit, where the reference's eps-regularised MUSCL ratio (reconstruction.hpp:128-153)
            # is discontinuous and a 1-ulp difference moves the residual by 1e-6
            inp = synthetic.write_case(tmp, name, ni, nj, nk, iterations=spec["iters"],
                                       perturb=(11, 0.01), **kw)
            for fl in (kw.get("species") or ()):  # species data of the reference's database
                shutil.copy("/root/reference/fluidDatabase/%s.dat" % fl, os.path.join(tmp, fl + ".dat"))
                os.chmod(os.path.join(tmp, fl + ".dat"), 0o644)
        else:
            inp = refcase.stage_case(os.path.join(REF_CASES, spec["src"]), tmp, spec["edits"],
                                     iterations=spec["iters"])
            for fl in spec.get("fluids", ()):  # species data of the reference's fluid database
                shutil.copy("/root/reference/fluidDatabase/%s.dat" % fl, os.path.join(tmp, fl + ".dat"))
                os.chmod(os.path.join(tmp, fl + ".dat"), 0o644)
            if "cloud" in spec:  # perturbed initial state at every cell centroid of the grid
                blocks = synthetic.read_plot3d(os.path.join(tmp, inp[:-4] + ".xyz"))
                nodes = np.concatenate([synthetic.centroids(b).reshape(-1, 3) for b in blocks])
                synthetic.write_cloud_points(os.path.join(tmp, "ic.dat"), nodes, *spec["cloud"],
                                             turb=spec.get("turb"), ic=spec.get("ic"))
        d = refcase.run_harness(tmp, inp, spec["iters"], full=spec["full"], geom=True)
    out = {k: np.asarray(v) for k, v in d.items()
           if not k.startswith("__") and k.split("/")[-1] not in DROP and k != "hist/time" and
           not any(k.split("/")[-1].startswith(p) for p in spec.get("drop", ()))}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %d records, %.1f kB" % (name, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    if not refcase.have_harness():
        sys.exit("oracle/_ref/aither_dump is not built (make -C oracle ref)")
    for nm in (sys.argv[1:] or list(CASES)):
        generate(nm)
