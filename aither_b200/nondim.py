"""Nondimensionalisation and `aither_cfg` construction (host set-up, mirrors `.inp` semantics).

Follows reference src/input.cpp:600-621 (reference speed of sound), src/fluid.cpp:89-103
(fluid::Nondimensionalize) and src/inputStates.cpp:464-473,590-599,674-681 (state data):
lengths / lRef, velocities / aRef, rho / rhoRef, p / (rhoRef aRef^2), T / TRef.
"""
from dataclasses import dataclass

import numpy as np

from . import ctypes_abi as abi
from .problem import make_cfg

UNIVERSAL_GAS_CONSTANT = 8.3144598  # J / mol-K, reference include/fluid.hpp:44


@dataclass
class Fluid:
    """One calorically perfect species, nondimensional (reference include/fluid.hpp)."""
    n: float
    gas_constant: float     # nondimensional R = 1/gamma for a single species
    hf: float
    a_ref: float
    rho_ref: float
    t_ref: float
    l_ref: float


def air(rho_ref, t_ref, l_ref=1.0, n=2.5, molar_mass_g=28.97, hf=0.0):
    """fluidDatabase/air.dat constants -> nondimensional fluid (src/fluid.cpp:89-103)."""
    molar_mass = molar_mass_g / 1000.0
    gamma = (n + 1.0) / n
    r_dim = UNIVERSAL_GAS_CONSTANT / molar_mass
    a_ref = np.sqrt(1.0 * gamma * r_dim * t_ref)  # src/input.cpp:616-621
    mm_nd = molar_mass / (rho_ref / l_ref ** 3.0)
    ru_nd = UNIVERSAL_GAS_CONSTANT / (a_ref * a_ref * rho_ref / (t_ref * l_ref ** 3.0))
    hf_nd = hf / (molar_mass * (a_ref * a_ref))
    return Fluid(n=n, gas_constant=ru_nd / mm_nd, hf=hf_nd, a_ref=float(a_ref), rho_ref=rho_ref,
                 t_ref=t_ref, l_ref=l_ref)


def nondim_primitive(density, velocity, pressure, rho_ref, t_ref, fluid=None):
    fl = fluid or air(rho_ref, t_ref)
    return np.array([density / rho_ref, velocity[0] / fl.a_ref, velocity[1] / fl.a_ref,
                     velocity[2] / fl.a_ref, pressure / (rho_ref * fl.a_ref * fl.a_ref)])


_RECON = {"constant": (abi.RECON_CONSTANT, -2.0), "upwind": (abi.RECON_MUSCL, -1.0),
          "fromm": (abi.RECON_MUSCL, 0.0), "quick": (abi.RECON_MUSCL, 0.5),
          "central": (abi.RECON_MUSCL, 1.0), "thirdOrder": (abi.RECON_MUSCL, 1.0 / 3.0),
          # weno / wenoZ leave kappa at its out-of-range default (src/input.cpp:67,288-300)
          "weno": (abi.RECON_WENO, -2.0), "wenoZ": (abi.RECON_WENOZ, -2.0)}
_LIMITER = {"none": abi.LIMITER_NONE, "vanAlbada": abi.LIMITER_VAN_ALBADA,
            "minmod": abi.LIMITER_MINMOD}


# fluidDatabase/air.dat Sutherland coefficients (viscosity C1, S; conductivity C1, S)
AIR_SUTHERLAND = (1.458e-6, 110.4, 2.495e-3, 194.0)


def viscous_terms(fluid, recon_kappa, visc_recon="central", sutherland=AIR_SUTHERLAND):
    """cfg entries of `equationSet: navierStokes`: Sutherland reference values and scaling
    (reference src/transport.cpp:50-68, include/transport.hpp:33-37) and the viscous CFL
    coefficient (src/input.cpp:1110-1118)."""
    c1, s_, k1, ks = sutherland
    mu_ref = c1 * fluid.t_ref ** 1.5 / (fluid.t_ref + s_)
    k_ref = fluid.a_ref * fluid.a_ref * mu_ref / fluid.t_ref
    coeff = 4.0 if recon_kappa == 1.0 else (2.0 if recon_kappa == -2.0 else 1.0)
    return dict(isViscous=1, viscRecon=0 if visc_recon == "central" else 1,
                viscousCFLCoeff=coeff, nondimScaling=mu_ref / (fluid.rho_ref * fluid.a_ref *
                                                               fluid.l_ref),
                suthViscC1=[c1], suthViscS=[s_], suthCondC1=[k1], suthCondS=[ks],
                tRef=fluid.t_ref, muMixRef=mu_ref, kMixRef=k_ref)


def euler_cfg(fluid, *, g=2, solver="dplur", sweeps=4, limiter="none", flux="roe",
              recon="thirdOrder", relaxation=1.0, bc_states=(), viscous=False,
              visc_recon="central", turb=None, jac="rusanov"):
    """`aither_cfg` for `equationSet: euler | navierStokes | rans`, `timeIntegration:
    implicitEuler` (theta=1, zeta=0; src/input.cpp:256-270); `solver`: lusgs / dplur (scalar
    diagonal) or blusgs / bdplur (block matrices); `turb`: None, "kOmegaWilcox2006", "sst2003"."""
    rc, kappa = _RECON[recon]
    viscous = viscous or turb is not None
    turb_id = {None: abi.TURB_NONE, "kOmegaWilcox2006": abi.TURB_KW_WILCOX,
               "sst2003": abi.TURB_SST}[turb]
    is_dplur = solver in ("dplur", "bdplur")
    extra = viscous_terms(fluid, kappa, visc_recon) if viscous else dict(isViscous=0, viscRecon=0,
                                                                         viscousCFLCoeff=1.0)
    return make_cfg(
        numSpecies=1, numTurb=2 if turb else 0, numGhosts=g, isRANS=int(turb is not None),
        isBlockMatrix=int(solver in ("blusgs", "bdplur")), isMultilevelTime=0,
        recon=rc, limiter=_LIMITER[limiter], invFlux=abi.FLUX_ROE if flux == "roe" else abi.FLUX_AUSM,
        invFluxJac=abi.JAC_RUSANOV if jac == "rusanov" else abi.JAC_APPROX_ROE, turbModel=turb_id,
        solver=abi.SOLVER_DPLUR if is_dplur else abi.SOLVER_LUSGS,
        matrixSweeps=sweeps, matrixRequiresInit=int(is_dplur or sweeps > 1),  # input.cpp:1120
        nonlinearIterations=1,
        kappa=kappa, theta=1.0, zeta=0.0, matrixRelaxation=relaxation, dualTimeCFL=-1.0,
        dtNondim=-1.0,
        gasConstant=[fluid.gas_constant], n=[fluid.n], hf=[fluid.hf],
        bcStates=list(bc_states), **extra)
