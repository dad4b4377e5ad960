"""Shared checker: run a grid level (the CPU oracle or the GPU path -- both expose the same phase
methods) against a committed golden fixture of the UNMODIFIED reference (tests/golden/*.npz, made
by tests/golden/make_golden.py). TEST INFRASTRUCTURE ONLY.
"""
import os

import numpy as np

import refcase
from aither_b200 import ctypes_abi as abi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# normalised L2 residuals the reference's own regression suite pins (1 % tolerance, ignored
# indices set to None): testCases/regressionTests.py:241-242 and :260-261
REGRESSION_GOLDENS = {
    "subsonicCylinder": (100, [1.8751e-01, 2.6727e-01, 3.1217e-01, None, 1.8639e-01]),
    "multiblockCylinder": (100, [2.0529e-01, 3.4540e-01, 5.0153e-01, None, 1.9997e-01]),
    # :356-357
    "viscousFlatPlate": (100, [7.4673e-02, 2.4711e-01, 3.8960e-02, None, 7.7683e-02]),
    # :379-380 (k-omega Wilcox 2006, 20 iterations)
    "turbFlatPlate": (20, [2.2309e-01, 2.9862e-01, None, 3.2376e-01, 2.1910e-01, 2.5208e-07,
                           3.3009e-06]),
    # :444-445 (SST 2003 + BLU-SGS + wall law, 20 iterations)
    "wallLaw": (20, [7.4098e-01, None, 3.1463e-01, 9.2837e-01, 7.2133e-01, 2.6860e-02,
                     2.6250e-07]),
}


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def full_iterations(d):
    return sorted(int(k[2:].split("/")[0]) for k in d if k.startswith("it") and k.endswith("/cfl"))


# Components that are rounding noise in a fixture (named explicitly, as the reference's own
# regression suite does with SetIgnoreIndices, testCases/regressionTests.py): the out-of-plane
# momentum of the 2-D cases -- every term of it is a difference of O(1) fluxes that cancel to the
# last bit, so its value is 1e-16 of the other equations' and carries no information. They are
# measured against the largest component's scale instead of their own. Everything else is held
# to the bar relative to ITS OWN block maximum.
def noise_equations(d, it=None):
    """Equations of a fixture whose residual is rounding noise in the reference's own run: sum(R^2)
    below 1e-20 of the largest equation's (amplitude 1e-10) -- at iteration `it`, or over the whole
    dumped history. Every term of such an equation is a difference of O(1) fluxes that cancel to
    the last bit (the out-of-plane momentum of a 2-D case; every equation but the wall-driven ones
    at the first evaluation from a uniform state), so its value carries no information. For the
    shipped cases the history-wide set is exactly what the reference's regression suite ignores
    (REGRESSION_GOLDENS' None entries; tests/test_oracle_pinned.py::test_noise_equations...)."""
    h = np.asarray(d["hist/residL2"], dtype=np.float64)
    if it is not None:
        h = h[it:it + 1]
    top = h.max()
    return tuple(int(e) for e in range(h.shape[1]) if h[:, e].max() <= 1e-20 * top)


def rel(a, b, noise=(), scale=None, whole_field=False, groups=()):
    """max |a-b| per trailing component relative to that component's OWN max |b| over the block
    (or to `scale`, one value per component). Components listed in `noise` -- and every component
    when `whole_field` (components of one vector / tensor: same units) -- are measured against the
    largest component's scale; the components of each index list in `groups` (the velocity vector
    inside the state) against the largest of the group."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(a.shape)
    ax = tuple(range(a.ndim - 1))
    own = np.abs(b).max(axis=ax)
    scale = own if scale is None else np.maximum(own, np.asarray(scale, dtype=np.float64))
    scale = np.array(scale, dtype=np.float64, copy=True)
    if scale.size:
        if whole_field:
            scale[:] = scale.max()
        for grp in groups:
            grp = [c for c in grp if c < scale.size]
            if grp:
                scale[grp] = scale[grp].max()
        for c in noise:
            if c < scale.size:
                scale[c] = scale.max()
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(a - b).max(axis=ax) / scale).max())


def non_edge_mask(shape, g):
    K, J, I = shape
    kk, jj, ii = np.meshgrid(np.arange(K), np.arange(J), np.arange(I), indexing="ij")
    out = (((kk < g) | (kk >= K - g)).astype(int) + ((jj < g) | (jj >= J - g)).astype(int) +
           ((ii < g) | (ii >= I - g)).astype(int))
    return out <= 1


def non_corner_mask(shape, g):
    """everything but the 8 corner blocks of the ghost shell (never assigned by the solver)"""
    K, J, I = shape
    kk, jj, ii = np.meshgrid(np.arange(K), np.arange(J), np.arange(I), indexing="ij")
    out = (((kk < g) | (kk >= K - g)).astype(int) + ((jj < g) | (jj >= J - g)).astype(int) +
           ((ii < g) | (ii >= I - g)).astype(int))
    return out <= 2


def interior(a, g):
    return a[g:a.shape[0] - g, g:a.shape[1] - g, g:a.shape[2] - g]


def normalised_history(hist_l2, n_first=5):
    """The reference's .resid normalisation (src/output.cpp:1033-1046): sqrt(L2) divided by the
    running maximum of sqrt(L2) over the first 5 iterations."""
    root = np.sqrt(hist_l2)
    first = np.maximum.accumulate(root[:n_first], axis=0)[-1]
    return root / np.where(first > 0, first, 1.0)


def check_phases(make_level, d, it, tol):
    """Restart `make_level(problem)` from the reference's state at the start of iteration `it`
    and compare every phase boundary with the reference's dump. `tol`: dict of per-phase bars."""
    key = "state@it%d.start" % it
    if it == 0 and "b0/" + key not in d:  # dropped from the fixture: identical to the initial state
        key = "state0"
    prob = refcase.problem_from_dump(d, state_key=key)
    lvl = make_level(prob)
    g = prob.cfg.numGhosts
    nb = len(prob.blocks)
    tag = "it%d" % it
    cfl = float(d[tag + "/cfl"][0])
    out = {}

    viscous = bool(prob.cfg.isViscous)
    # equations that are rounding noise at this iteration (named from the reference's own norms)
    noise = noise_equations(d, it)
    per_equation = (abi.FIELD_RESIDUAL, abi.FIELD_UPDATE, abi.FIELD_MATRIX_RESID)
    # components of one vector / tensor share their units: measured against the field's magnitude
    vector_fields = (abi.FIELD_VELOCITY_GRAD, abi.FIELD_TKE_GRAD, abi.FIELD_OMEGA_GRAD)

    # the largest (root-mean-square) magnitude the residual of every equation reaches in the run:
    # where an equation's residual is the rounding remainder of cancelling fluxes at this
    # evaluation (first evaluation from a uniform state: turbFlatPlate's energy equation,
    # sum R^2 = 7e-16 against 2e-7 five iterations later), its own maximum is not a scale
    ncells = sum(int(np.prod(d["b%d/residual@%s" % (bb, tag)].shape[:3])) for bb in range(nb))
    resid_scale = np.sqrt(np.asarray(d["hist/residL2"], dtype=np.float64).max(axis=0) / ncells)

    def cmp(field, key, what, mask_edges=False, inner=False, comps=None, scale_key=None):
        worst = 0.0
        for bb in range(nb):
            a = lvl.field(bb, field)
            b = d["b%d/%s" % (bb, key)]
            if b.shape[:3] != a.shape[:3]:  # the reference keeps this field ghost padded
                b = interior(b, g)
            b = b.reshape(a.shape[:3] + (-1,))
            if comps is not None:
                a, b = a[..., comps], b[..., comps]
            if mask_edges:
                # edge ghost cells are read by the viscous gradient stencils only
                m = (non_corner_mask if viscous else non_edge_mask)(a.shape[:3], g)
                a, b = a[m], b[m]
            if inner:
                a, b = interior(a, g), interior(b, g)
            scale = resid_scale if field == abi.FIELD_RESIDUAL else None
            if scale_key is not None:  # per-equation scale of another dumped field of the block
                sb = d["b%d/%s" % (bb, scale_key)]
                scale = np.abs(sb.reshape(-1, sb.shape[-1])).max(axis=0)
            # the three velocity components of the state are one vector (a 2-D case carries
            # 1e-24 noise in the out-of-plane one)
            ns = int(prob.cfg.numSpecies)
            groups = ([ns, ns + 1, ns + 2],) if field == abi.FIELD_STATE and comps is None else ()
            if field in (abi.FIELD_DIAG, abi.FIELD_DIAG_INV) and a.shape[-1] > 2:
                # block matrices: the flow block and the turbulence block are each one operator,
                # an entry is measured against the largest entry of its block
                nfl = (ns + 4) * (ns + 4)
                groups = (list(range(nfl)), list(range(nfl, a.shape[-1])))
            worst = max(worst, rel(a, b, noise if field in per_equation else (), scale,
                                   whole_field=field in vector_fields, groups=groups))
        out[what] = worst
        assert worst <= tol[what], "%s %s: rel err %.3e > %.1e" % (tag, what, worst, tol[what])

    lvl.store_old_solution(it)
    lvl.get_boundary_conditions()
    cmp(abi.FIELD_STATE, "state@%s.bc" % tag, "ghosts", mask_edges=True)
    lvl.calc_residual()
    if viscous:  # ghost cells as the viscous fluxes saw them (viscous-wall + edge refill)
        cmp(abi.FIELD_STATE, "state@%s.viscbc" % tag, "ghosts", mask_edges=True)
        cmp(abi.FIELD_VISCOSITY, "viscosity@" + tag, "ghosts", mask_edges=True)
    if prob.cfg.numTurb > 0:  # cell averages of the face eddy viscosity, blending, gradients
        for fld, key in ((abi.FIELD_EDDY_VISCOSITY, "eddyViscosity@"), (abi.FIELD_F1, "f1@"),
                         (abi.FIELD_F2, "f2@"), (abi.FIELD_VELOCITY_GRAD, "velocityGrad@")):
            if "b0/" + key + tag in d:
                cmp(fld, key + tag, "turb", inner=True)
        cmp(abi.FIELD_TKE_GRAD, "tkeGrad@" + tag, "turb")
        cmp(abi.FIELD_OMEGA_GRAD, "omegaGrad@" + tag, "turb")
    cmp(abi.FIELD_RESIDUAL, "residual@" + tag, "residual")
    cmp(abi.FIELD_SPEC_RADIUS, "specRadius@" + tag, "specRadius",
        comps=slice(0, 2 if prob.cfg.numTurb > 0 else 1))
    lvl.calc_time_step(cfl)
    cmp(abi.FIELD_DT, "dt@" + tag, "dt")
    lvl.invert_diagonal()
    lvl.initialize_matrix_update()
    cmp(abi.FIELD_DIAG, "diag@" + tag, "diag")
    cmp(abi.FIELD_DIAG_INV, "diagInv@" + tag, "diag")
    if "b0/x0@" + tag in d:
        cmp(abi.FIELD_UPDATE, "x0@" + tag, "x0", inner=True)
    lvl.relax()
    cmp(abi.FIELD_UPDATE, "x@" + tag, "x", inner=True)
    # A x - b is a difference of terms of the size of b = -R / theta + ...: its rounding error
    # scales with the residual of that equation, not with what is left of it after the sweeps
    cmp(abi.FIELD_MATRIX_RESID, "matrixResid@" + tag, "matrixResid", scale_key="residual@" + tag)
    l2, linf = lvl.update_blocks()
    lvl.reset_diagonal()
    cmp(abi.FIELD_STATE, "state@%s.end" % tag, "state", inner=True)
    h = d["hist/residL2"][it]
    # equations whose residual is rounding noise are not compared (as in check_history)
    l2err = float(np.max(np.abs(l2 - h) / np.where(h > 1e-20 * h.max(), h, np.inf)))
    out["l2"] = l2err
    assert l2err <= tol["l2"], (tag, l2err, np.abs(l2 - h) / h, out)
    # L-infinity: the reference takes the largest *signed* residual (src/procBlock.cpp:862-867);
    # when every residual of a uniform flow is <= rounding noise its location is noise too
    loc, lref = d["hist/linfLoc"][it], float(d["hist/linf"][it])
    rscale = float(np.sqrt(h.max()))
    assert abs(linf.linf - lref) <= tol["l2"] * max(abs(lref), rscale)
    if lref > 1e-9 * rscale:
        assert (linf.block, linf.i, linf.j, linf.k, linf.eqn) == tuple(int(v) for v in loc), \
            (tag, (linf.block, linf.i, linf.j, linf.k, linf.eqn), loc)
    lvl.close()
    return out


def check_history(make_level, d, n_iter, tol, name=None):
    """Run n_iter iterations from the reference's initial state; compare the un-normalised
    sum(R^2) history and the matrix residual with the reference's, every iteration."""
    prob = refcase.problem_from_dump(d, state_key="state0")
    lvl = make_level(prob)
    href, mref, cfl = d["hist/residL2"], d["hist/matrixResid"], d["hist/cfl"]
    # one history record per (time step, nonlinear iteration); the old solution is stored once
    # per time step (reference src/main.cpp:231-247)
    nl = int(d["cfg/nonlinearIterations"][0]) if "cfg/nonlinearIterations" in d else 1
    mine = np.zeros((n_iter, prob.neq))
    worst = 0.0
    for it in range(n_iter):
        if it % nl == 0:
            lvl.store_old_solution(it // nl)
        l2, _, mr = lvl.iterate(float(cfl[it]), it % nl)
        mine[it] = l2
        # an equation whose residual is rounding noise (2-D cases: the reference ignores that
        # index too, regressionTests.py SetIgnoreIndices) is not compared
        scale = np.where(href[it] > 1e-20 * href[it].max(), href[it], np.inf)
        err = float(np.max(np.abs(l2 - href[it]) / scale))
        worst = max(worst, err)
        assert err <= tol, "iteration %d: L2 rel err %.3e > %.1e\n%s\n%s" % (it, err, tol, l2,
                                                                            href[it])
        # (a matrix residual at rounding level -- converged dual time steps -- is noise)
        assert abs(mr - mref[it]) <= max(tol, 1e-9) * abs(mref[it]) + 1e-12 * href[it].max(), \
            (it, mr, mref[it])
    lvl.close()
    if name in REGRESSION_GOLDENS and n_iter >= REGRESSION_GOLDENS[name][0]:
        n, gold = REGRESSION_GOLDENS[name]
        norm = normalised_history(mine)[n - 1]
        for e, gv in enumerate(gold):
            if gv is not None:
                assert abs(norm[e] - gv) <= 0.01 * gv, (name, e, norm[e], gv)
    return worst
