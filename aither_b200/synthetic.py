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


def box_nodes(ni, nj, nk, lengths=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), warp=0.02, period=1.0):
    """Node coordinates (nk+1, nj+1, ni+1, 3) of a warped box."""
    X = origin[0] + lengths[0] * np.arange(ni + 1) / ni
    Y = origin[1] + lengths[1] * np.arange(nj + 1) / nj
    Z = origin[2] + lengths[2] * np.arange(nk + 1) / nk
    zz, yy, xx = np.meshgrid(Z, Y, X, indexing="ij")
    x = xx + warp * np.sin(2 * np.pi * yy / period) * np.sin(2 * np.pi * zz / period)
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
             recon="thirdOrder", flux="roe", iterations=10, ic_file=None, viscous=False,
             visc_recon="central", wall=None, turb=None, jac="rusanov", species=None,
             periodic=None, overrides=None, inlet_outlet=False, wall_law=False,
             nonreflecting=None):
    """`wall_law`: the viscous wall uses the wall law (`wallTreatment=wallLaw`).
    `nonreflecting`: None, or the length scale [m] of non-reflecting `inlet` / `pressureOutlet`
    states (with `inlet_outlet`).
    `periodic`: None, or the box length: the two i-faces become a periodic pair (translation
    [length, 0, 0]) instead of characteristic boundaries.
    `inlet_outlet`: the i-lo face becomes an `inlet` and the i-hi face a `pressureOutlet`
    (reflecting forms) instead of characteristic boundaries.
    `overrides`: dict of `.inp` keys replacing (or adding to) the lines below, e.g.
    {"timeIntegration": "crankNicholson", "timeStep": "1e-6", "matrixRelaxation": "1.1"}.
    `species`: None (air) or a dict name -> reference mass fraction (multi-species mixture with
    Schmidt-number diffusion, e.g. {"H2O": 0.233, "H2": 0.001, "N2": 0.766}).
    `viscous`: navierStokes with a viscousWall on the j-lo face (`wall`: None = adiabatic,
    ("isothermal", T) or ("heatFlux", q)). `turb`: None, "kOmegaWilcox2006" or "sst2003" (RANS,
    implies viscous; farfield turbulence intensity 1 %, eddy viscosity ratio 10)."""
    viscous = viscous or turb is not None
    vel = "[%g, %g, %g]" % IC["velocity"]
    state = "pressure=%g; density=%g; velocity=%s" % (IC["pressure"], IC["density"], vel)
    if species:
        state += "; massFractions=[%s]" % ", ".join("%s=%g" % kv for kv in species.items())
    if turb is not None:
        state += "; turbulenceIntensity=0.01; eddyViscosityRatio=10"
    ic = "icState(tag=-1; %s)" % state if ic_file is None else "icState(tag=-1; file=%s)" % ic_file
    wall_state = "viscousWall(tag=2)"
    if wall is not None and wall[0] == "isothermal":
        wall_state = "viscousWall(tag=2; temperature=%g)" % wall[1]
    elif wall is not None and wall[0] == "heatFlux":
        wall_state = "viscousWall(tag=2; heatFlux=%g)" % wall[1]
    if wall_law:
        wall_state = wall_state[:-1] + "; wallTreatment=wallLaw)"
    nr = "; nonreflecting=true; lengthScale=%g" % nonreflecting if nonreflecting else ""
    lines = [
        "gridName: %s" % name,
        "equationSet: %s" % ("rans" if turb else ("navierStokes" if viscous else "euler")),
        "turbulenceModel: %s" % (turb or "none"),
        "timeIntegration: implicitEuler",
        "cflStart: %g" % cfl, "cflMax: %g" % cfl,
        "faceReconstruction: %s" % recon,
        "limiter: %s" % limiter,
        "inviscidFlux: %s" % flux,
        "inviscidFluxJacobian: %s" % jac,
        "iterations: %d" % iterations,
        "outputFrequency: 1000000",
        "outputVariables: <density, vel_x, vel_y, vel_z, pressure>",
        "referenceTemperature: %g" % REF_T,
        "referenceDensity: %g" % REF_RHO,
        *(["fluids: <%s>" % ", ".join("fluid(name=%s; referenceMassFraction=%g)" % kv
                                      for kv in species.items()),
           "diffusionModel: schmidt"] if species else []),
        "initialConditions: <%s>" % ic,
        "matrixSolver: %s" % solver,
        "matrixSweeps: %d" % sweeps,
        "matrixRelaxation: 1.0",
        "viscousFaceReconstruction: %s" % visc_recon,
        "boundaryStates: <%s>" % ", ".join(
            (["inlet(tag=1; %s; massFractions=[air=1.0]%s)" % (state, nr),
              "pressureOutlet(tag=3; pressure=%g%s)" % (IC["pressure"], nr)] if inlet_outlet
             else ["characteristic(tag=1; %s)" % state]) + ([wall_state] if viscous else []) +
            (["periodic(startTag=4; endTag=5; translation=[%.17g, 0, 0])" % periodic]
             if periodic else [])),
        "boundaryConditions: 1",
        "2 2 2",
        ("periodic %d %d %d %d %d %d 4" if periodic else
         ("inlet %d %d %d %d %d %d 1" if inlet_outlet else "characteristic %d %d %d %d %d %d 1"))
        % (0, 0, 0, nj, 0, nk),
        ("periodic %d %d %d %d %d %d 5" if periodic else
         ("pressureOutlet %d %d %d %d %d %d 3" if inlet_outlet
          else "characteristic %d %d %d %d %d %d 1"))
        % (ni, ni, 0, nj, 0, nk),
        ("viscousWall %d %d %d %d %d %d 2" if viscous else "slipWall %d %d %d %d %d %d 0")
        % (0, ni, 0, 0, 0, nk),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, nj, nj, 0, nk),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, 0, nj, 0, 0),
        "slipWall %d %d %d %d %d %d 0" % (0, ni, 0, nj, nk, nk),
        ""]
    for key, val in (overrides or {}).items():
        hit = [n for n, ln in enumerate(lines) if ln.split(":")[0] == key]
        if hit:
            lines[hit[0]] = "%s: %s" % (key, val)
        else:  # new keys go in front of the boundary-state table
            pos = [n for n, ln in enumerate(lines) if ln.startswith("boundaryStates")][0]
            lines.insert(pos, "%s: %s" % (key, val))
    return "\n".join(lines)


def read_plot3d(path):
    """Blocks of a raw-binary multi-block Plot3D file as (nk+1, nj+1, ni+1, 3) node arrays."""
    with open(path, "rb") as f:
        nb = int(np.fromfile(f, dtype="<i4", count=1)[0])
        dims = np.fromfile(f, dtype="<i4", count=3 * nb).reshape(nb, 3)
        out = []
        for ni1, nj1, nk1 in dims:
            n = int(ni1) * int(nj1) * int(nk1)
            xyz = np.fromfile(f, dtype="<f8", count=3 * n).reshape(3, nk1, nj1, ni1)
            out.append(np.moveaxis(xyz, 0, -1))
    return out


def centroids(x):
    return 0.125 * (x[:-1, :-1, :-1] + x[:-1, :-1, 1:] + x[:-1, 1:, :-1] + x[:-1, 1:, 1:] +
                    x[1:, :-1, :-1] + x[1:, :-1, 1:] + x[1:, 1:, :-1] + x[1:, 1:, 1:])


def _cloud_values(n, seed, amplitude, turb, ic=None):
    """seed-fixed +-amplitude noise on rho, u, v, w, p (+ k, omega for RANS: farfield-like
    k = 1.5 (0.01 |v|)^2, omega = rho k / (10 mu))"""
    rng = np.random.default_rng(seed)
    IC = ic or globals()["IC"]
    base = np.array([IC["density"], *IC["velocity"], IC["pressure"]])
    vals = base[None, :] * (1.0 + amplitude * (2.0 * rng.random((n, 5)) - 1.0))
    kw = np.zeros((n, 2))
    if turb is not None:
        vmag = np.linalg.norm(IC["velocity"])
        k0 = 1.5 * (0.01 * vmag) ** 2
        mu0 = 1.458e-6 * REF_T ** 1.5 / (REF_T + 110.4)
        w0 = IC["density"] * k0 / (10.0 * mu0)
        kw = np.array([k0, w0])[None, :] * (1.0 + amplitude * (2.0 * rng.random((n, 2)) - 1.0))
    return vals, kw


def write_cloud_points(path, cen, seed=0, amplitude=0.01, species="air", turb=None, ic=None):
    """cloud file with seed-fixed noise on the IC state (`ic`: density / velocity / pressure,
    default the synthetic box's) at the given points (see write_cloud)"""
    vals, kw = _cloud_values(cen.shape[0], seed, amplitude, turb, ic)
    with open(path, "w") as f:
        f.write("%d\n%s\n" % (cen.shape[0], species))
        for c, v, t in zip(cen, vals, kw):
            f.write(" ".join("%.17g" % x for x in (*c, *v, *t, 1.0)) + "\n")


def write_cloud(path, nodes, seed=0, amplitude=0.01, species="air", turb=None, mix=None):
    """Initial-condition cloud file (reference src/utility.cpp:513-520: `numberOfPoints`, species
    line, then `x y z rho u v w p tke omega mf...` per point), one point per cell centroid, with
    seed-fixed +-amplitude noise on rho, u, v, w, p. The reference assigns each cell the state of
    its nearest cloud point."""
    x = np.asarray(nodes)
    cen = 0.125 * (x[:-1, :-1, :-1] + x[:-1, :-1, 1:] + x[:-1, 1:, :-1] + x[:-1, 1:, 1:] +
                   x[1:, :-1, :-1] + x[1:, :-1, 1:] + x[1:, 1:, :-1] + x[1:, 1:, 1:]).reshape(-1, 3)
    vals, kw = _cloud_values(cen.shape[0], seed, amplitude, turb)
    mfs = np.ones((cen.shape[0], 1))
    if mix:  # perturbed mass fractions (10x the state noise, so that diffusion is visible)
        species = " ".join(mix)
        rng = np.random.default_rng(seed + 104729)
        mfs = np.array(list(mix.values()))[None, :] * (
            1.0 + 10.0 * amplitude * (2.0 * rng.random((cen.shape[0], len(mix))) - 1.0))
        mfs /= mfs.sum(axis=1, keepdims=True)
    with open(path, "w") as f:
        f.write("%d\n%s\n" % (cen.shape[0], species))
        for c, v, t, m in zip(cen, vals, kw, mfs):
            f.write(" ".join("%.17g" % x for x in (*c, *v, *t, *m)) + "\n")


def write_case(case_dir, name, ni, nj, nk, perturb=None, size=1.0, **kw):
    """Write `<name>.xyz` + `<name>.inp` (+ `ic.dat` when perturb=(seed, amplitude)). `size`: edge
    length of the box in metres (a small box lowers the cell Reynolds number so that the viscous
    fluxes weigh in the residual)."""
    os.makedirs(case_dir, exist_ok=True)
    nodes = box_nodes(ni, nj, nk, lengths=(size, size, size), warp=0.02 * size, period=size)
    write_plot3d(os.path.join(case_dir, name + ".xyz"), [nodes])
    if perturb is not None:
        write_cloud(os.path.join(case_dir, "ic.dat"), nodes, *perturb, turb=kw.get("turb"),
                    mix=kw.get("species"))
        kw["ic_file"] = "ic.dat"
    if kw.get("periodic"):
        kw["periodic"] = size  # the i-faces are one box length apart
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
                recon="thirdOrder", seed=0, amplitude=0.01, viscous=False, visc_recon="central",
                wall=None, size=1.0, turb=None, jac="rusanov"):
    """The synthetic single-block case as a `Problem` (product-side set-up, no reference): Euler,
    or with `viscous` laminar Navier-Stokes with a viscous wall on the j-lo face (`wall`: None =
    adiabatic, ("isothermal", T [K]), ("heatFlux", q)) on a box `size` metres wide."""
    viscous = viscous or turb is not None
    g = {"constant": 1, "weno": 3, "wenoZ": 3}.get(recon, 2)  # input.cpp:1127-1144
    m = block_metrics(box_nodes(ni, nj, nk, lengths=(size,) * 3, warp=0.02 * size, period=size), g)
    fluid = nondim.air(REF_RHO, REF_T)
    free = nondim.nondim_primitive(IC["density"], IC["velocity"], IC["pressure"], REF_RHO, REF_T)
    bc_states = [dict(tag=1, type=abi.BC_CHARACTERISTIC, density=free[0],
                      velocity=list(free[1:4]), pressure=free[4], massFractions=[1.0],
                      turbulenceIntensity=0.01, eddyViscosityRatio=10.0)]
    if viscous:
        ws = dict(tag=2, type=abi.BC_VISCOUS_WALL, velocity=[0.0, 0.0, 0.0], massFractions=[1.0])
        if wall is not None and wall[0] == "isothermal":
            ws.update(isIsothermal=1, temperature=wall[1] / REF_T)
        elif wall is not None and wall[0] == "heatFlux":
            raise ValueError("heat-flux walls need the wall distance: use a reference dump")
        bc_states.append(ws)
    cfg = nondim.euler_cfg(fluid, g=g, solver=solver, sweeps=sweeps, limiter=limiter, flux=flux,
                           recon=recon, bc_states=bc_states, viscous=viscous,
                           visc_recon=visc_recon, turb=turb, jac=jac)
    state = perturbed_state((nk + 2 * g, nj + 2 * g, ni + 2 * g), 5, seed, amplitude)
    wall_dist = None
    if turb is not None:
        # farfield-like turbulence (k = 1.5 (0.01 |v|)^2, omega = rho k / (10 mu)), nondimensional
        # as the reference's cloud reader does (src/utility.cpp:575-578), with the same noise
        rng = np.random.default_rng(seed + 7919)
        vmag = np.linalg.norm(IC["velocity"])
        k0 = 1.5 * (0.01 * vmag) ** 2
        mu0 = nondim.AIR_SUTHERLAND[0] * REF_T ** 1.5 / (REF_T + nondim.AIR_SUTHERLAND[1])
        w0 = IC["density"] * k0 / (10.0 * mu0)
        kw = np.array([k0 / fluid.a_ref ** 2, w0 * mu0 / (REF_RHO * fluid.a_ref ** 2)])
        noise = 1.0 + amplitude * (2.0 * rng.random(state.shape[:3] + (2,)) - 1.0)
        state = np.concatenate([state, kw[None, None, None, :] * noise], axis=-1)
        # wall distance: to the centre of the j-lo wall face of the same (i, k) column -- an
        # approximation of the reference's nearest-wall-face search (src/procBlock.cpp:6030-6107;
        # set-up code outside the hot path) that is good enough for timing runs; parity tests take
        # the wall distance from reference dumps
        cen = m["center"]
        face = 0.5 * (cen[:, g - 1:g, :, :] + cen[:, g:g + 1, :, :])
        wall_dist = np.ascontiguousarray(np.linalg.norm(cen - face, axis=-1)[..., None])
    surfaces = [
        (abi.BC_CHARACTERISTIC, 0, 0, 0, nj, 0, nk, 1),
        (abi.BC_CHARACTERISTIC, ni, ni, 0, nj, 0, nk, 1),
        (abi.BC_VISCOUS_WALL, 0, ni, 0, 0, 0, nk, 2) if viscous
        else (abi.BC_SLIP_WALL, 0, ni, 0, 0, 0, nk, 0),
        (abi.BC_SLIP_WALL, 0, ni, nj, nj, 0, nk, 0),
        (abi.BC_SLIP_WALL, 0, ni, 0, nj, 0, 0, 0),
        (abi.BC_SLIP_WALL, 0, ni, 0, nj, nk, nk, 0),
    ]
    arrays = {k: m[k] for k in ("vol", "fAreaI", "fAreaJ", "fAreaK", "center", "cellWidthI",
                                "cellWidthJ", "cellWidthK")}
    arrays["state"] = state
    arrays["wallDist"] = wall_dist
    return Problem(cfg, [Block(ni, nj, nk, surfaces, arrays)])


# ---- multi-block decomposition of a single-block problem ---------------------------------------
def _cuts(n, parts):
    """cell index of the cuts that split n cells into `parts` nearly equal pieces"""
    return [(n * p) // parts for p in range(parts + 1)]


def split_problem(prob, splits):
    """Cut the single block of `prob` into splits = (pi, pj, pk) sub-blocks joined by `interblock`
    connections, the way the reference's decomposition presents them to the solver (reference
    src/parallel.cpp:95-178 splits blocks; src/boundaryConditions.cpp:2458-2497 names the joins;
    include/boundaryConditions.hpp:324-336 is the connection record).

    Every sub-block's ghost-padded geometry is cut out of the parent's, so geometry ghosts across a
    join equal the neighbour's cells exactly as after the reference's geometry swap
    (src/procBlock.cpp:3149-3330). Block order: i fastest, then j, then k. Connections: all i-joins,
    then j-joins, then k-joins; first side = lower block's upper surface, orientation 1."""
    assert len(prob.blocks) == 1
    parent = prob.blocks[0]
    g = prob.cfg.numGhosts
    pi, pj, pk = splits
    ci, cj, ck = _cuts(parent.ni, pi), _cuts(parent.nj, pj), _cuts(parent.nk, pk)
    n_of = (parent.ni, parent.nj, parent.nk)

    def parent_surface(surf_type, lo, hi):
        """the parent's boundary surface of that side (one surface must cover the sub-face)"""
        for row in parent.surfaces:
            t, imin, imax, jmin, jmax, kmin, kmax, tag = row
            mn, mx = (imin, jmin, kmin), (imax, jmax, kmax)
            d3 = (surf_type - 1) // 2
            if mn[d3] != mx[d3] or mn[d3] != (0 if surf_type % 2 else n_of[d3]):
                continue
            ok = all(mn[d] <= lo[d] and hi[d] <= mx[d] for d in range(3) if d != d3)
            if ok:
                return t, tag
        raise ValueError("no single parent surface covers the sub-block face")

    blocks, index = [], {}
    for c in range(pk):
        for b in range(pj):
            for a in range(pi):
                lo = (ci[a], cj[b], ck[c])
                hi = (ci[a + 1], cj[b + 1], ck[c + 1])
                n = tuple(hi[d] - lo[d] for d in range(3))
                arrays = {}
                for name, arr in parent.arrays.items():
                    if arr is None:
                        arrays[name] = None
                        continue
                    ext = [n[2] + 2 * g, n[1] + 2 * g, n[0] + 2 * g]
                    if name == "fAreaI":
                        ext[2] += 1
                    elif name == "fAreaJ":
                        ext[1] += 1
                    elif name == "fAreaK":
                        ext[0] += 1
                    arrays[name] = np.ascontiguousarray(
                        arr[lo[2]:lo[2] + ext[0], lo[1]:lo[1] + ext[1], lo[0]:lo[0] + ext[2]])
                surfaces = []
                pos = (a, b, c)
                parts = (pi, pj, pk)
                for d3 in range(3):
                    for upper in (0, 1):
                        st = 2 * d3 + 1 + upper
                        rng = [[0, n[0]], [0, n[1]], [0, n[2]]]
                        rng[d3] = [n[d3], n[d3]] if upper else [0, 0]
                        at_edge = pos[d3] == (parts[d3] - 1 if upper else 0)
                        if at_edge:
                            t, tag = parent_surface(st, lo, hi)
                        else:
                            nb = list(pos)
                            nb[d3] += 1 if upper else -1
                            nb_id = nb[0] + pi * (nb[1] + pj * nb[2])
                            partner_surface = st - 1 if upper else st + 1
                            t, tag = abi.BC_INTERBLOCK, partner_surface * 1000 + nb_id
                        surfaces.append((t, rng[0][0], rng[0][1], rng[1][0], rng[1][1], rng[2][0],
                                         rng[2][1], tag))
                bid = len(blocks)
                index[pos] = bid
                blocks.append(Block(n[0], n[1], n[2], surfaces, arrays, parent_block=0,
                                    global_pos=bid))
    conns = lattice_connections([(b.ni, b.nj, b.nk) for b in blocks], splits)
    return Problem(prob.cfg, blocks, conns)


def lattice_connections(dims, splits):
    """`interblock` connections of a pi x pj x pk lattice of blocks (block id = a + pi (b + pj c),
    dims[id] = (ni, nj, nk)): all i-joins, then j-joins, then k-joins; first side = lower block's
    upper surface, second = upper block's lower surface, orientation 1."""
    pi, pj, pk = splits
    conns = []
    for d3 in range(3):
        d1, d2 = (d3 + 1) % 3, (d3 + 2) % 3
        for c in range(pk):
            for b in range(pj):
                for a in range(pi):
                    pos = [a, b, c]
                    if pos[d3] + 1 >= splits[d3]:
                        continue
                    up = list(pos)
                    up[d3] += 1
                    lo_id = pos[0] + pi * (pos[1] + pj * pos[2])
                    up_id = up[0] + pi * (up[1] + pj * up[2])
                    nlo = dims[lo_id]
                    cn = abi.Conn()
                    cn.rank[0] = cn.rank[1] = 0
                    cn.block[0], cn.block[1] = lo_id, up_id
                    cn.localBlock[0], cn.localBlock[1] = lo_id, up_id
                    cn.boundary[0], cn.boundary[1] = 2 * d3 + 2, 2 * d3 + 1
                    for s in range(2):
                        cn.d1Start[s], cn.d1End[s] = 0, nlo[d1]
                        cn.d2Start[s], cn.d2End[s] = 0, nlo[d2]
                    cn.constSurf[0], cn.constSurf[1] = nlo[d3], 0
                    cn.orientation, cn.isInterblock = 1, 1
                    conns.append(cn)
    return conns


def lattice_problem(n, splits, *, only=None, solver="dplur", sweeps=4, limiter="none", flux="roe",
                    recon="thirdOrder", seed=0, amplitude=0.01, viscous=False,
                    visc_recon="central", size=1.0):
    """A pi x pj x pk lattice of blocks of n^3 cells (or n = (ni, nj, nk) cells; each block is one
    period of the warped box, so every block is the benchmark's block) joined by `interblock`
    connections -- the weak-scaling workload.
    `only`: block ids to materialise (default all); the others are dimension-only placeholders so
    that a rank builds just the blocks it owns. Each block's ghost geometry on a joined face is its
    neighbour's real geometry: the block's nodes are generated g cells beyond those faces from the
    same analytic node function, metrics computed, and the extra layers cropped.
    `viscous`: laminar Navier-Stokes with an adiabatic viscous wall on the lattice's j-lo side
    (BASELINE configs[3]'s scheme with recon="weno", visc_recon="centralFourth"); `size`: edge
    length of one block along i in metres (cells are cubes of size / ni)."""
    pi, pj, pk = splits
    nb = pi * pj * pk
    nn = (n, n, n) if np.isscalar(n) else tuple(int(v) for v in n)
    g = {"constant": 1, "weno": 3, "wenoZ": 3}.get(recon, 2)
    fluid = nondim.air(REF_RHO, REF_T)
    free = nondim.nondim_primitive(IC["density"], IC["velocity"], IC["pressure"], REF_RHO, REF_T)
    bc_states = [dict(tag=1, type=abi.BC_CHARACTERISTIC, density=free[0],
                      velocity=list(free[1:4]), pressure=free[4], massFractions=[1.0])]
    if viscous:
        bc_states.append(dict(tag=2, type=abi.BC_VISCOUS_WALL, velocity=[0.0, 0.0, 0.0],
                              massFractions=[1.0]))
    cfg = nondim.euler_cfg(fluid, g=g, solver=solver, sweeps=sweeps, limiter=limiter, flux=flux,
                           recon=recon, bc_states=bc_states, viscous=viscous,
                           visc_recon=visc_recon)
    want = set(range(nb) if only is None else only)
    h = size / nn[0]
    blocks = []
    for bid in range(nb):
        a, b, c = bid % pi, (bid // pi) % pj, bid // (pi * pj)
        pos = (a, b, c)
        surfaces = []
        ext_lo, ext_hi = [0, 0, 0], [0, 0, 0]
        for d3 in range(3):
            for upper in (0, 1):
                st = 2 * d3 + 1 + upper
                rng = [[0, nn[0]], [0, nn[1]], [0, nn[2]]]
                rng[d3] = [nn[d3], nn[d3]] if upper else [0, 0]
                at_edge = pos[d3] == (splits[d3] - 1 if upper else 0)
                if at_edge:
                    t, tag = (abi.BC_CHARACTERISTIC, 1) if d3 == 0 else (abi.BC_SLIP_WALL, 0)
                    if viscous and d3 == 1 and not upper:
                        t, tag = abi.BC_VISCOUS_WALL, 2
                else:
                    nbp = list(pos)
                    nbp[d3] += 1 if upper else -1
                    t = abi.BC_INTERBLOCK
                    tag = (st - 1 if upper else st + 1) * 1000 + nbp[0] + pi * (nbp[1] + pj * nbp[2])
                    (ext_hi if upper else ext_lo)[d3] = g
                surfaces.append((t, rng[0][0], rng[0][1], rng[1][0], rng[1][1], rng[2][0],
                                 rng[2][1], tag))
        arrays = {"state": None}
        if bid in want:
            ne = [nn[d] + ext_lo[d] + ext_hi[d] for d in range(3)]
            nodes = box_nodes(ne[0], ne[1], ne[2],
                              lengths=tuple(ne[d] * h for d in range(3)),
                              origin=tuple(pos[d] * nn[d] * h - ext_lo[d] * h for d in range(3)),
                              warp=0.02 * size, period=size)
            m = block_metrics(nodes, g)
            arrays = {}
            for name in ("vol", "fAreaI", "fAreaJ", "fAreaK", "center", "cellWidthI", "cellWidthJ",
                         "cellWidthK"):
                arr = m[name]
                sl = []
                for ax, d in ((0, 2), (1, 1), (2, 0)):
                    extra = 1 if name == "fArea" + "IJK"[d] else 0
                    sl.append(slice(ext_lo[d], ext_lo[d] + nn[d] + 2 * g + extra))
                arrays[name] = np.ascontiguousarray(arr[tuple(sl)])
            del m
            arrays["state"] = perturbed_state((nn[2] + 2 * g, nn[1] + 2 * g, nn[0] + 2 * g), 5,
                                              seed + bid, amplitude)
            arrays["wallDist"] = None
        blocks.append(Block(nn[0], nn[1], nn[2], surfaces, arrays, parent_block=bid, global_pos=bid))
    conns = lattice_connections([nn] * nb, splits)
    return Problem(cfg, blocks, conns)


def assign_ranks(prob, n_ranks):
    """Place block b on rank b % ... in contiguous groups (the reference's `manual` decomposition:
    one input block per rank, src/parallel.cpp:44-66) and fill the rank / localBlock fields of
    every connection. Returns block ids per rank."""
    nb = len(prob.blocks)
    per = [[] for _ in range(n_ranks)]
    owner, local = {}, {}
    for b in range(nb):
        r = (b * n_ranks) // nb
        owner[b], local[b] = r, len(per[r])
        per[r].append(b)
    for cn in prob.conns:
        for s in range(2):
            gb = cn.block[s]
            cn.rank[s], cn.localBlock[s] = owner[gb], local[gb]
    return per


def reassemble(prob_split, splits, fields):
    """Stitch per-block interior arrays (k, j, i, c) of a split problem back into the parent's
    (nk, nj, ni, c) array."""
    pi, pj, pk = splits
    rows_k = []
    for c in range(pk):
        rows_j = []
        for b in range(pj):
            rows_j.append(np.concatenate([fields[a + pi * (b + pj * c)] for a in range(pi)],
                                         axis=2))
        rows_k.append(np.concatenate(rows_j, axis=1))
    return np.concatenate(rows_k, axis=0)
