"""Deterministic synthetic cases for the benchmark and parity tests.

Grid: an n_i x n_j x n_k-cell unit box with nodes x = X + 0.02 sin(2 pi Y) sin(2 pi Z) (non-trivial
metrics, positive volumes). Euler, implicit Euler, Roe + MUSCL(kappa = 1/3), DPLUR; i-faces
`characteristic`, j/k faces `slipWall`. `write_case` emits the raw-binary Plot3D grid and the
`.inp` the reference itself reads (reference src/plot3d.cpp:363-444, src/input.cpp:162-598), so the
very same case runs through the reference harness (tests) and the GPU path (bench).
"""
import os

import numpy as np

from . import ctypes_abi as abi
from . import nondim
from .geometry import block_metrics
from .problem import Block, Problem

REF_T = 288.0
REF_RHO = 1.2256
IC = dict(pressure=101300.0, density=1.2256, velocity=(100.0, 20.0, 10.0))


def box_nodes(ni, nj, nk, lengths=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), warp=0.02):
    """Node coordinates (nk+1, nj+1, ni+1, 3) of a warped box."""
    X = origin[0] + lengths[0] * np.arange(ni + 1) / ni
    Y = origin[1] + lengths[1] * np.arange(nj + 1) / nj
    Z = origin[2] + lengths[2] * np.arange(nk + 1) / nk
    zz, yy, xx = np.meshgrid(Z, Y, X, indexing="ij")
    x = xx + warp * np.sin(2 * np.pi * yy) * np.sin(2 * np.pi * zz)
    return np.stack([x, yy, zz], axis=-1)


def write_plot3d(path, blocks_nodes):
    """Raw-binary multi-block Plot3D: int32 nBlocks; 3 x int32 node dims per block; then per block
    all x, all y, all z as little-endian float64, i fastest."""
    with open(path, "wb") as f:
        np.array([len(blocks_nodes)], dtype="<i4").tofile(f)
        for nd in blocks_nodes:
            nk1, nj1, ni1 = nd.shape[:3]
            np.array([ni1, nj1, nk1], dtype="<i4").tofile(f)
        for nd in blocks_nodes:
            for c in range(3):
                np.ascontiguousarray(nd[..., c], dtype="<f8").tofile(f)


def inp_text(name, ni, nj, nk, *, solver="dplur", sweeps=4, cfl=50.0, limiter="none",
             recon="thirdOrder", flux="roe", iterations=10, ic_file=None):
    vel = "[%g, %g, %g]" % IC["velocity"]
    state = "pressure=%g; density=%g; velocity=%s" % (IC["pressure"], IC["density"], vel)
    ic = "icState(tag=-1; %s)" % state if ic_file is None else "icState(tag=-1; file=%s)" % ic_file
    return "\n".join([
        "gridName: %s" % name,
        "equationSet: euler",
        "timeIntegration: implicitEuler",
        "cflStart: %g" % cfl, "cflMax: %g" % cfl,
        "faceReconstruction: %s" % recon,
        "limiter: %s" % limiter,
        "inviscidFlux: %s" % flux,
        "iterations: %d" % iterations,
        "outputFrequency: 1000000",
        "outputVariables: <density, vel_x, vel_y, vel_z, pressure>",
        "referenceTemperature: %g" % REF_T,
        "referenceDensity: %g" % REF_RHO,
        "initialConditions: <%s>" % ic,
        "matrixSolver: %s" % solver,
        "matrixSweeps: %d" % sweeps,
        "matrixRelaxation: 1.0",
        "boundaryStates: <characteristic(tag=1; %s)>" % state,
        "boundaryConditions: 1",
        "2 2 2",
        "characteristic %d %d %d %d %d %d 1" % (0, 0, 0, nj, 0, nk),
        "characteristic %d %d %d %d %d %d 1" % (ni, ni, 0, nj, 0, nk),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, 0, 0, 0, nk),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, nj, nj, 0, nk),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, 0, nj, 0, 0),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, 0, nj, nk, nk),
        ""])


def write_cloud(path, nodes, seed=0, amplitude=0.01, species="air"):
    """Initial-condition cloud file (reference src/utility.cpp:513-520: `numberOfPoints`, species
    line, then `x y z rho u v w p tke omega mf...` per point), one point per cell centroid, with
    seed-fixed +-amplitude noise on rho, u, v, w, p. The reference assigns each cell the state of
    its nearest cloud point."""
    x = np.asarray(nodes)
    cen = 0.125 * (x[:-1, :-1, :-1] + x[:-1, :-1, 1:] + x[:-1, 1:, :-1] + x[:-1, 1:, 1:] +
                   x[1:, :-1, :-1] + x[1:, :-1, 1:] + x[1:, 1:, :-1] + x[1:, 1:, 1:]).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    base = np.array([IC["density"], *IC["velocity"], IC["pressure"]])
    vals = base[None, :] * (1.0 + amplitude * (2.0 * rng.random((cen.shape[0], 5)) - 1.0))
    with open(path, "w") as f:
        f.write("%d\n%s\n" % (cen.shape[0], species))
        for c, v in zip(cen, vals):
            f.write(" ".join("%.17g" % t for t in (*c, *v, 0.0, 0.0, 1.0)) + "\n")


def write_case(case_dir, name, ni, nj, nk, perturb=None, **kw):
    """Write `<name>.xyz` + `<name>.inp` (+ `ic.dat` when perturb=(seed, amplitude))."""
    os.makedirs(case_dir, exist_ok=True)
    write_plot3d(os.path.join(case_dir, name + ".xyz"), [box_nodes(ni, nj, nk)])
    if perturb is not None:
        write_cloud(os.path.join(case_dir, "ic.dat"), box_nodes(ni, nj, nk), *perturb)
        kw["ic_file"] = "ic.dat"
    with open(os.path.join(case_dir, name + ".inp"), "w") as f:
        f.write(inp_text(name, ni, nj, nk, **kw))
    return name + ".inp"


def perturbed_state(shape_kji, neq_state, seed=0, amplitude=0.01):
    """Seed-fixed +-1 % noise on the nondimensional primitive IC (rho, u, v, w, p)."""
    rng = np.random.default_rng(seed)
    base = nondim.nondim_primitive(IC["density"], IC["velocity"], IC["pressure"], REF_RHO, REF_T)
    noise = 1.0 + amplitude * (2.0 * rng.random(shape_kji + (neq_state,)) - 1.0)
    return base[None, None, None, :] * noise


def box_problem(ni, nj, nk, *, solver="dplur", sweeps=4, limiter="none", flux="roe",
                recon="thirdOrder", seed=0, amplitude=0.01):
    """The synthetic single-block Euler case as a `Problem` (product-side set-up, no reference)."""
    g = {"constant": 1, "weno": 3, "wenoZ": 3}.get(recon, 2)  # input.cpp:1127-1144
    m = block_metrics(box_nodes(ni, nj, nk), g)
    fluid = nondim.air(REF_RHO, REF_T)
    free = nondim.nondim_primitive(IC["density"], IC["velocity"], IC["pressure"], REF_RHO, REF_T)
    cfg = nondim.euler_cfg(fluid, g=g, solver=solver, sweeps=sweeps, limiter=limiter, flux=flux,
                           recon=recon,
                           bc_states=[dict(tag=1, type=abi.BC_CHARACTERISTIC, density=free[0],
                                           velocity=list(free[1:4]), pressure=free[4],
                                           massFractions=[1.0])])
    state = perturbed_state((nk + 2 * g, nj + 2 * g, ni + 2 * g), 5, seed, amplitude)
    surfaces = [
        (abi.BC_CHARACTERISTIC, 0, 0, 0, nj, 0, nk, 1),
        (abi.BC_CHARACTERISTIC, ni, ni, 0, nj, 0, nk, 1),
        (abi.BC_SLIP_WALL, 0, ni, 0, 0, 0, nk, 0),
        (abi.BC_SLIP_WALL, 0, ni, nj, nj, 0, nk, 0),
        (abi.BC_SLIP_WALL, 0, ni, 0, nj, 0, 0, 0),
        (abi.BC_SLIP_WALL, 0, ni, 0, nj, nk, nk, 0),
    ]
    arrays = {k: m[k] for k in ("vol", "fAreaI", "fAreaJ", "fAreaK", "center", "cellWidthI",
                                "cellWidthJ", "cellWidthK")}
    arrays["state"] = state
    arrays["wallDist"] = None
    return Problem(cfg, [Block(ni, nj, nk, surfaces, arrays)])
