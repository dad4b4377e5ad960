"""Device point functions (aither_b200/csrc/physics.cuh, compiled for the host by tests/hostsim)
against the CPU oracle, on seeded random inputs. No GPU needed: this pins the *formulas* that the
kernels inline; the kernels themselves (indexing, tiling, reductions) are covered by -m gpu tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from aither_b200 import ctypes_abi as abi
from aither_b200 import nondim, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
PD = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(PD)


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    out = os.path.join(HERE, "hostsim", "libhostsim.so")
    deps = [src, os.path.join(HERE, "..", "aither_b200", "csrc", "physics.cuh"),
            os.path.join(HERE, "..", "aither_b200", "csrc", "turbulence.cuh"),
            os.path.join(HERE, "..", "include", "aither_gpu.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-fPIC", "-shared",
                               "-o", out, src])
    return C.CDLL(out)


def cfg_for(flux="roe"):
    fl = nondim.air(synthetic.REF_RHO, synthetic.REF_T)
    free = nondim.nondim_primitive(1.2256, (100.0, 20.0, 10.0), 101300.0, synthetic.REF_RHO,
                                   synthetic.REF_T)
    bcs = [dict(tag=1, type=abi.BC_CHARACTERISTIC, density=free[0], velocity=list(free[1:4]),
                pressure=free[4], massFractions=[1.0]),
           dict(tag=2, type=abi.BC_STAGNATION_INLET, stagnationPressure=0.7192,
                stagnationTemperature=1.002, direction=[1.0, 0.0, 0.0], massFractions=[1.0]),
           dict(tag=3, type=abi.BC_PRESSURE_OUTLET, pressure=0.714),
           dict(tag=4, type=abi.BC_SUPERSONIC_INFLOW, density=1.1, velocity=[2.0, 0.1, 0.0],
                pressure=0.8, massFractions=[1.0])]
    return nondim.euler_cfg(fl, flux=flux, bc_states=bcs)


def rand_state(rng, mach=0.3):
    s = np.empty(5)
    s[0] = rng.uniform(0.5, 1.5)
    s[1:4] = rng.uniform(-1, 1, 3) * mach
    s[4] = rng.uniform(0.4, 1.0)
    return s


def unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def close(a, b, tol=1e-13):
    scale = max(np.abs(b).max(), 1e-300)
    assert np.abs(a - b).max() <= tol * scale, (a, b)


@pytest.mark.parametrize("lim", [0, 1, 2])
def test_muscl(hs, lim):
    rng = np.random.default_rng(1 + lim)
    L = oracle.lib()
    for _ in range(200):
        u2, u1, d1 = rand_state(rng), rand_state(rng), rand_state(rng)
        w = rng.uniform(0.5, 2.0, 3)
        a, b = np.empty(5), np.empty(5)
        L.orc_muscl(ptr(u2), ptr(u1), ptr(d1), 5, 1.0 / 3.0, lim, w[0], w[1], w[2], ptr(a))
        hs.hs_muscl(ptr(u2), ptr(u1), ptr(d1), C.c_double(1.0 / 3.0), lim, C.c_double(w[0]),
                    C.c_double(w[1]), C.c_double(w[2]), ptr(b))
        close(b, a)
    # exactly uniform upwind pair: the eps-regularised ratio must give first order, not NaN
    u = rand_state(rng)
    a, b = np.empty(5), np.empty(5)
    d1 = rand_state(rng)
    L.orc_muscl(ptr(u), ptr(u), ptr(d1), 5, 1.0 / 3.0, lim, 1.0, 1.0, 1.0, ptr(a))
    hs.hs_muscl(ptr(u), ptr(u), ptr(d1), C.c_double(1.0 / 3.0), lim, C.c_double(1.0),
                C.c_double(1.0), C.c_double(1.0), ptr(b))
    assert np.isfinite(b).all()
    close(b, a)


@pytest.mark.parametrize("wenoz", [0, 1])
def test_weno(hs, wenoz):
    rng = np.random.default_rng(7 + wenoz)
    L = oracle.lib()
    for _ in range(100):
        ys = np.stack([rand_state(rng) for _ in range(5)])
        w = rng.uniform(0.5, 2.0, 5)
        a, b = np.empty(5), np.empty(5)
        rows = (PD * 5)(*[ptr(np.ascontiguousarray(ys[q])) for q in range(5)])
        keep = [np.ascontiguousarray(ys[q]) for q in range(5)]
        rows = (PD * 5)(*[ptr(k) for k in keep])
        L.orc_weno(rows, ptr(w), 5, wenoz, ptr(a))
        hs.hs_weno(ptr(np.ascontiguousarray(ys)), ptr(w), 5, wenoz, ptr(b))
        close(b, a, 1e-12)


@pytest.mark.parametrize("flux", ["roe", "ausm"])
@pytest.mark.parametrize("mach", [0.3, 1.5])
def test_inviscid_flux(hs, flux, mach):
    rng = np.random.default_rng(11)
    cfg = cfg_for(flux)
    L = oracle.lib()
    for _ in range(300):
        l, r, n = rand_state(rng, mach), rand_state(rng, mach), unit(rng)
        a, b = np.empty(5), np.empty(5)
        L.orc_inviscid_flux(C.byref(cfg), ptr(l), ptr(r), ptr(n), ptr(a))
        hs.hs_inviscid_flux(C.byref(cfg), ptr(l), ptr(r), ptr(n), ptr(b))
        close(b, a)


@pytest.mark.parametrize("bc,tag", [(abi.BC_SLIP_WALL, 0), (abi.BC_CHARACTERISTIC, 1),
                                    (abi.BC_INLET, 1), (abi.BC_STAGNATION_INLET, 2),
                                    (abi.BC_PRESSURE_OUTLET, 3), (abi.BC_SUPERSONIC_INFLOW, 4),
                                    (abi.BC_SUPERSONIC_OUTFLOW, 0)])
def test_ghost_state(hs, bc, tag):
    rng = np.random.default_rng(13)
    cfg = cfg_for()
    L = oracle.lib()
    for trial in range(300):
        mach = 0.3 if trial % 2 == 0 else 1.6  # sub- and supersonic branches
        s, n = rand_state(rng, mach), unit(rng)
        for surf in (1, 2, 5):
            for layer in (1, 2, 3):
                a, b = np.empty(5), np.empty(5)
                L.orc_ghost_state(C.byref(cfg), ptr(s), bc, ptr(n), surf, tag, layer, ptr(a))
                hs.hs_ghost_state(C.byref(cfg), ptr(s), bc, ptr(n), surf, tag, layer, ptr(b))
                if not np.isfinite(a).all():
                    continue  # stagnation inlet formula is undefined for outflow states
                close(b, a, 1e-11 if bc == abi.BC_STAGNATION_INLET else 1e-13)


def test_offdiag(hs):
    rng = np.random.default_rng(17)
    cfg = cfg_for()
    L = oracle.lib()
    for _ in range(300):
        s, n = rand_state(rng), unit(rng)
        du = rng.normal(size=5) * 1e-2
        fa = np.concatenate([n, [rng.uniform(0.1, 2.0)]])
        for pos in (0, 1):
            a, b = np.empty(5), np.empty(5)
            L.orc_offdiag_scalar(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, ptr(a))
            hs.hs_offdiag_scalar(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, ptr(b))
            close(b, a)


@pytest.mark.parametrize("mach", [0.3, 1.5, 0.02])
def test_roe_flux_fast_twin(hs, mach):
    """RoeFluxFast (shared reciprocals; used by ResidualMarchKernel) vs the oracle's Roe flux."""
    rng = np.random.default_rng(23)
    cfg = cfg_for("roe")
    L = oracle.lib()
    for _ in range(500):
        l, r, n = rand_state(rng, mach), rand_state(rng, mach), unit(rng)
        a, b = np.empty(5), np.empty(5)
        L.orc_inviscid_flux(C.byref(cfg), ptr(l), ptr(r), ptr(n), ptr(a))
        hs.hs_roe_flux_fast(C.byref(cfg), ptr(l), ptr(r), ptr(n), ptr(b))
        close(b, a, 2e-14)


def test_offdiag_fast_twin(hs):
    """MakeIngr + OffDiagFromIngr (ImplicitMarchKernel) vs the oracle's scalar off-diagonal."""
    rng = np.random.default_rng(29)
    cfg = cfg_for()
    L = oracle.lib()
    for trial in range(500):
        s, n = rand_state(rng), unit(rng)
        du = rng.normal(size=5) * (1e-2 if trial % 2 else 1e-6)
        fa = np.concatenate([n, [rng.uniform(0.1, 2.0)]])
        for pos in (0, 1):
            a, b = np.empty(5), np.empty(5)
            L.orc_offdiag_scalar(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, ptr(a))
            hs.hs_offdiag_fast(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, ptr(b))
            # absolute bar against the flux magnitude: both compute F(U+dU) - F(U) by cancellation
            scale = fa[3] * (np.abs(s).max() + 1.0)
            assert np.abs(a - b).max() <= 1e-14 * scale + 1e-12 * np.abs(a).max(), (a, b)


# ---- RANS point functions (turbulence.cuh, NT = 2 variants of physics.cuh) ----------------------
def rans_cfg(model):
    """configuration of a committed RANS fixture (transport reference values included)"""
    import goldencheck as gc
    import refcase
    cfg = refcase.cfg_from_dump(gc.load("box_sst"))
    cfg.turbModel = model
    return cfg


def rand_rans_state(rng, mach=0.3):
    s = np.empty(7)
    s[:5] = rand_state(rng, mach)
    s[5] = 10.0 ** rng.uniform(-8, -3)   # k / a_ref^2
    s[6] = 10.0 ** rng.uniform(-4, 0)    # omega mu_ref / (rho_ref a_ref^2)
    return s


def declare_rans(hs, L):
    D = C.c_double
    hs.hs_eddy_visc.argtypes = [C.c_void_p, PD, PD, PD, PD, D, D, PD]
    hs.hs_turb_source.argtypes = [C.c_void_p, PD, PD, PD, PD, D, D, PD]
    hs.hs_offdiag_scalar_rans.argtypes = [C.c_void_p, PD, PD, PD, C.c_int, D, D, D, D, PD]
    L.orc_eddy_visc.argtypes = [C.c_void_p, PD, PD, PD, PD, D, D, PD]
    L.orc_turb_source.argtypes = [C.c_void_p, PD, PD, PD, PD, D, D, PD]
    L.orc_offdiag_scalar_visc.argtypes = [C.c_void_p, PD, PD, PD, C.c_int, D, D, D, D, PD]
    L.orc_ghost_state_visc.argtypes = [C.c_void_p, PD, C.c_int, PD, C.c_int, C.c_int, C.c_int, D, D,
                                       PD]
    for f in (hs.hs_eddy_visc, hs.hs_turb_source, hs.hs_offdiag_scalar_rans, L.orc_eddy_visc,
              L.orc_turb_source, L.orc_offdiag_scalar_visc, L.orc_ghost_state_visc):
        f.restype = None


@pytest.mark.parametrize("model", [abi.TURB_KW_WILCOX, abi.TURB_SST])
def test_eddy_viscosity_and_source(hs, model):
    rng = np.random.default_rng(31 + model)
    cfg = rans_cfg(model)
    L = oracle.lib()
    declare_rans(hs, L)
    for trial in range(400):
        s = rand_rans_state(rng)
        vg = rng.normal(size=9) * 10.0 ** rng.uniform(-2, 3)
        kg = rng.normal(size=3) * s[5] * 1e2
        wg = rng.normal(size=3) * s[6] * 1e2
        mu, wd = rng.uniform(0.5, 1.5), 10.0 ** rng.uniform(-5, 0)
        a, b = np.empty(3), np.empty(3)
        L.orc_eddy_visc(C.byref(cfg), ptr(s), ptr(vg), ptr(kg), ptr(wg), mu, wd, ptr(a))
        hs.hs_eddy_visc(C.byref(cfg), ptr(s), ptr(vg), ptr(kg), ptr(wg), mu, wd, ptr(b))
        assert np.all(np.abs(a - b) <= 1e-12 * np.maximum(np.abs(a), 1e-300) + 1e-15), (a, b)
        sa, sb = np.empty(2), np.empty(2)
        f1 = a[1]
        L.orc_turb_source(C.byref(cfg), ptr(s), ptr(vg), ptr(kg), ptr(wg), a[0], f1, ptr(sa))
        hs.hs_turb_source(C.byref(cfg), ptr(s), ptr(vg), ptr(kg), ptr(wg), a[0], f1, ptr(sb))
        # production and destruction cancel: compare against the larger of the two magnitudes
        scale = abs(sa) + s[0] * s[6] * np.array([s[5], s[6]]) / cfg.nondimScaling
        assert np.all(np.abs(sa - sb) <= 1e-12 * scale), (trial, sa, sb)


@pytest.mark.parametrize("model", [abi.TURB_KW_WILCOX, abi.TURB_SST])
def test_offdiag_rans(hs, model):
    rng = np.random.default_rng(41 + model)
    cfg = rans_cfg(model)
    L = oracle.lib()
    declare_rans(hs, L)
    for _ in range(300):
        s, n = rand_rans_state(rng), unit(rng)
        du = rng.normal(size=7) * 1e-2 * np.array([1, 1, 1, 1, 1, s[5], s[6]])
        fa = np.concatenate([n, [rng.uniform(0.1, 2.0)]])
        mu, mut, f1, dist = rng.uniform(0.5, 1.5), rng.uniform(0, 50), rng.uniform(0, 1), \
            rng.uniform(1e-3, 1.0)
        for pos in (0, 1):
            a, b = np.empty(7), np.empty(7)
            L.orc_offdiag_scalar_visc(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, mu, mut, f1, dist,
                                      ptr(a))
            hs.hs_offdiag_scalar_rans(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, mu, mut, f1, dist,
                                      ptr(b))
            assert np.all(np.abs(a - b) <= 1e-12 * np.abs(a).max() + 1e-13 * np.abs(a)), (a, b)


@pytest.mark.parametrize("bc,tag", [(abi.BC_SLIP_WALL, 0), (abi.BC_CHARACTERISTIC, 1),
                                    (abi.BC_INLET, 1), (abi.BC_SUPERSONIC_INFLOW, 1),
                                    (abi.BC_PRESSURE_OUTLET, 1), (abi.BC_SUPERSONIC_OUTFLOW, 0)])
def test_ghost_state_rans(hs, bc, tag):
    """farfield turbulence (turbulence intensity / eddy viscosity ratio) in the ghost states"""
    rng = np.random.default_rng(43)
    cfg = rans_cfg(abi.TURB_SST)
    L = oracle.lib()
    declare_rans(hs, L)
    for trial in range(200):
        mach = 0.3 if trial % 2 == 0 else 1.6
        s, n = rand_rans_state(rng, mach), unit(rng)
        for surf in (1, 2, 5):
            for layer in (1, 2, 3):
                a, b = np.empty(7), np.empty(7)
                L.orc_ghost_state_visc(C.byref(cfg), ptr(s), bc, ptr(n), surf, tag, layer, 0.0, 0.0,
                                       ptr(a))
                hs.hs_ghost_state_rans(C.byref(cfg), ptr(s), bc, ptr(n), surf, tag, layer, ptr(b))
                if not np.isfinite(a).all():
                    continue
                assert np.all(np.abs(a - b) <= 1e-12 * np.abs(a) + 1e-14 * np.abs(a[:5]).max()), (a, b)


@pytest.mark.parametrize("bc,tag", [(abi.BC_INLET, 2), (abi.BC_PRESSURE_OUTLET, 3)])
def test_ghost_state_nonreflecting(hs, bc, tag):
    """non-reflecting inlet / pressure outlet (LODI relaxation with the state at time n, the time
    step, pressure / velocity gradients and the patch Mach numbers): device point function vs
    the oracle's restatement of src/ghostStates.cpp:435-466, :614-643, with the boundary states
    of the reference's convectingVortex case."""
    import goldencheck as gc
    import refcase
    cfg = refcase.cfg_from_dump(gc.load("convectingVortex"))
    L = oracle.lib()
    for f in (L.orc_ghost_state_nonreflecting, hs.hs_ghost_state_nonreflecting):
        f.argtypes = [C.c_void_p, PD, C.c_int, PD, C.c_int, C.c_int, C.c_int, PD, PD]
        f.restype = None
    rng = np.random.default_rng(61)
    for trial in range(300):
        s, n = rand_state(rng, 0.3), unit(rng)
        extra = np.empty(20)
        extra[0] = 0.0 if trial % 5 == 0 else 10.0 ** rng.uniform(-4, 0)
        extra[1:6] = s * (1.0 + 0.01 * rng.normal(size=5))
        extra[6:9] = rng.normal(size=3)
        extra[9:18] = rng.normal(size=9) * 5.0
        extra[18], extra[19] = rng.uniform(0.0, 0.5), rng.uniform(0.3, 0.9)
        for surf in (1, 4):
            for layer in (1, 2):
                a, b = np.empty(5), np.empty(5)
                L.orc_ghost_state_nonreflecting(C.byref(cfg), ptr(s), bc, ptr(n), surf, tag, layer,
                                                ptr(extra), ptr(a))
                hs.hs_ghost_state_nonreflecting(C.byref(cfg), ptr(s), bc, ptr(n), surf, tag, layer,
                                                ptr(extra), ptr(b))
                if not np.isfinite(a).all():
                    continue
                assert np.all(np.abs(a - b) <= 1e-12 * np.abs(a) + 1e-13 * np.abs(a).max()), (a, b)


@pytest.mark.parametrize("mode,fixture", [(0, "wallLaw"), (1, "box_walllaw_heatflux"),
                                          (2, "box_walllaw_isothermal")])
def test_wall_law(hs, mode, fixture):
    """wall law (walllaw.cuh WallLawEval vs the oracle's restatement of src/wallLaw.cpp): y+ root,
    wall shear stress, heat flux / wall temperature, wall eddy viscosity, k and omega, for random
    wall-adjacent states at wall distances on both sides of the root bracket [10, 1e4] (without a
    root in the bracket the reference keeps the values of its last evaluation, y+ = 1e4)."""
    import goldencheck as gc
    import refcase
    d = gc.load(fixture)
    cfg = refcase.cfg_from_dump(d)
    tag = [cfg.bcStates[q].tag for q in range(cfg.numBCStates) if cfg.bcStates[q].isWallLaw][0]
    L = oracle.lib()
    D = C.c_double
    for f in (L.orc_wall_law, hs.hs_wall_law):
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, PD, D, PD, C.c_int, PD]
        f.restype = None
    rng = np.random.default_rng(59 + mode)
    unbracketed, bracketed = 0, 0
    for trial in range(300):
        s, n = rand_rans_state(rng, 0.25), unit(rng)
        wd = 10.0 ** rng.uniform(-6.5, -2.5)
        a, b = np.empty(11), np.empty(11)
        L.orc_wall_law(C.byref(cfg), mode, tag, ptr(s), wd, ptr(n), trial % 2, ptr(a))
        hs.hs_wall_law(C.byref(cfg), mode, tag, ptr(s), wd, ptr(n), trial % 2, ptr(b))
        if not np.isfinite(a).all():
            continue
        unbracketed += a[0] == 1.0e4
        bracketed += 10.0 <= a[0] < 1.0e4
        scale = np.abs(a)
        scale[1:4] = np.abs(a[1:4]).max()
        assert np.all(np.abs(a - b) <= 1e-11 * scale + 1e-300), (trial, a, b)
    assert unbracketed > 10 and bracketed > 50


@pytest.mark.parametrize("flux", ["roe", "ausm"])
@pytest.mark.parametrize("fast", [0, 1])
def test_inviscid_flux_rans(hs, flux, fast):
    rng = np.random.default_rng(47)
    cfg = rans_cfg(abi.TURB_SST)
    cfg.invFlux = abi.FLUX_ROE if flux == "roe" else abi.FLUX_AUSM
    L = oracle.lib()
    for trial in range(300):
        mach = 0.3 if trial % 2 == 0 else 1.5
        l, r, n = rand_rans_state(rng, mach), rand_rans_state(rng, mach), unit(rng)
        a, b = np.empty(7), np.empty(7)
        L.orc_inviscid_flux(C.byref(cfg), ptr(l), ptr(r), ptr(n), ptr(a))
        hs.hs_inviscid_flux_rans(C.byref(cfg), ptr(l), ptr(r), ptr(n), fast, ptr(b))
        assert np.all(np.abs(a[:5] - b[:5]) <= 3e-14 * np.abs(a[:5]).max()), (a, b)
        tscale = np.maximum(np.abs(a[5:]), np.abs(a[0]) * np.maximum(l[5:], r[5:]))
        assert np.all(np.abs(a[5:] - b[5:]) <= 1e-12 * tscale), (a, b)


# ---- three species ------------------------------------------------------------------------------
def mix3_cfg(flux="ausm"):
    import goldencheck as gc
    import refcase
    cfg = refcase.cfg_from_dump(gc.load("box_mix3_euler"))
    cfg.invFlux = abi.FLUX_ROE if flux == "roe" else abi.FLUX_AUSM
    return cfg


def rand_mix3_state(rng, mach=0.3):
    s = np.empty(7)
    rho = rng.uniform(0.5, 1.5)
    y = rng.dirichlet([2.0, 0.3, 5.0])
    s[:3] = rho * y
    s[3:6] = rng.uniform(-1, 1, 3) * mach
    s[6] = rng.uniform(0.4, 1.0)
    return s


def test_mixture_transport(hs):
    """Wilke's mixing rule for viscosity and conductivity (src/transport.cpp:70-110)."""
    rng = np.random.default_rng(53)
    cfg = mix3_cfg()
    L = oracle.lib()
    for _ in range(500):
        s = rand_mix3_state(rng)
        a, b = np.empty(2), np.empty(2)
        L.orc_mixture_transport(C.byref(cfg), ptr(s), ptr(a))
        hs.hs_mixture_transport3(C.byref(cfg), ptr(s), ptr(b))
        assert np.all(np.abs(a - b) <= 1e-13 * np.abs(a)), (a, b)


@pytest.mark.parametrize("flux", ["roe", "ausm"])
@pytest.mark.parametrize("fast", [0, 1])
def test_inviscid_flux_three_species(hs, flux, fast):
    rng = np.random.default_rng(59)
    cfg = mix3_cfg(flux)
    L = oracle.lib()
    for trial in range(300):
        mach = 0.3 if trial % 2 == 0 else 1.5
        l, r, n = rand_mix3_state(rng, mach), rand_mix3_state(rng, mach), unit(rng)
        a, b = np.empty(7), np.empty(7)
        L.orc_inviscid_flux(C.byref(cfg), ptr(l), ptr(r), ptr(n), ptr(a))
        hs.hs_inviscid_flux3(C.byref(cfg), ptr(l), ptr(r), ptr(n), fast, ptr(b))
        # the energy flux carries the heats of formation (|hf| ~ 100): compare against it
        assert np.all(np.abs(a - b) <= 1e-13 * np.abs(a).max()), (a, b)


def test_offdiag_three_species(hs):
    rng = np.random.default_rng(61)
    cfg = mix3_cfg()
    L = oracle.lib()
    for _ in range(300):
        s, n = rand_mix3_state(rng), unit(rng)
        du = rng.normal(size=7) * 1e-3
        fa = np.concatenate([n, [rng.uniform(0.1, 2.0)]])
        for pos in (0, 1):
            a, b = np.empty(7), np.empty(7)
            L.orc_offdiag_scalar(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, ptr(a))
            hs.hs_offdiag_scalar3(C.byref(cfg), ptr(s), ptr(du), ptr(fa), pos, ptr(b))
            assert np.all(np.abs(a - b) <= 1e-12 * np.abs(a).max() + 1e-13), (a, b)
