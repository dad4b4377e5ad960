// physics.cuh -- point-wise numerics of the hot path, as inlineable device functions.
//
// Everything here works on one cell / one face held in registers: thermodynamics, MUSCL / WENO
// reconstruction, Roe and AUSMPW+ fluxes, spectral radii, the Rusanov off-diagonal product and the
// boundary-condition ghost states. Formulas follow mnucci32/aither v0.10.0 (cited per function as
// "ref: file:line"); the structure does not: no heap vectors, no virtual dispatch, no string
// compares -- species / turbulence counts are template parameters so every loop unrolls and every
// state lives in registers.
//
// AITHER_HD lets tests/hostsim compile these same functions with g++ to check them point-wise
// against the oracle without a GPU; the shipped library only ever runs them inside kernels.
#pragma once
#include <math.h>

#include "../../include/aither_gpu.h"

#ifdef __CUDACC__
#define AITHER_HD __host__ __device__ __forceinline__
#else
#define AITHER_HD inline
#endif

namespace aither {

constexpr double kEps = 1.0e-30;        // ref: include/macros.hpp.in:21
constexpr double kEntropyFix = 0.1;     // ref: include/inviscidFlux.hpp:298
constexpr double kTurbMin = 1.0e-20;    // ref: include/turbulence.hpp:72-73

// Gas model constants for <= AITHER_MAX_SPECIES calorically perfect species (nondimensional).
// Sutherland transport of the (single) species; ref: src/transport.cpp:50-68,113-131
struct Transport {
  double tRef, muRef, kRef, scaling;
  // Sutherland coefficients and molar mass per species (src/transport.cpp:32-68)
  double viscC1[AITHER_MAX_SPECIES], viscS[AITHER_MAX_SPECIES];
  double condC1[AITHER_MAX_SPECIES], condS[AITHER_MAX_SPECIES];
  double molarMass[AITHER_MAX_SPECIES];
  double schmidt;  // species diffusion: Schmidt number, <= 0 for `diffusionModel: none`
  int turbModel;   // aither_turb
};
// ref: src/transport.cpp:113-131 (species), :70-110,148-192 (Wilke's mixing rule)
AITHER_HD double SpeciesViscosity(const Transport &tr, double t, int ss) {
  const double temp = t * tr.tRef;
  const double mu = (tr.viscC1[ss] * (temp * sqrt(temp))) / (temp + tr.viscS[ss]);
  return mu / tr.muRef;
}
AITHER_HD double SpeciesConductivity(const Transport &tr, double t, int ss) {
  const double temp = t * tr.tRef;
  const double k = (tr.condC1[ss] * (temp * sqrt(temp))) / (temp + tr.condS[ss]);
  return k / tr.kRef;
}
// mixture viscosity of primitive state s (densities s[0..NS)) at temperature t
template <int NS>
AITHER_HD double MixtureViscosity(const Transport &tr, double t, const double *s) {
  if (NS == 1) return SpeciesViscosity(tr, t, 0);
  double x[NS], sv[NS], sum = 0.0, rho = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) rho += s[q];
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    x[q] = (s[q] / rho) / tr.molarMass[q];
    sum += x[q];
    sv[q] = SpeciesViscosity(tr, t, q);
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) x[q] /= sum;
  double mixtureVisc = 0.0;
#pragma unroll
  for (int ii = 0; ii < NS; ++ii) {
    double denom = 0.0;
#pragma unroll
    for (int jj = 0; jj < NS; ++jj) {
      const double f = 1.0 + sqrt(sv[ii] / sv[jj]) * sqrt(sqrt(tr.molarMass[jj] / tr.molarMass[ii]));
      denom += x[jj] / sqrt(1.0 + tr.molarMass[ii] / tr.molarMass[jj]) * (f * f);
    }
    mixtureVisc += (x[ii] * sv[ii]) / denom;
  }
  return 4.0 / sqrt(2.0) * mixtureVisc;
}
// k_eff = scaling * k_mix, k_mix = (sum x_i k_i + 1 / sum (x_i / k_i)) / 2
template <int NS>
AITHER_HD double MixtureEffConductivity(const Transport &tr, double t, const double *s) {
  if (NS == 1) return SpeciesConductivity(tr, t, 0) * tr.scaling;
  double x[NS], sum = 0.0, rho = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) rho += s[q];
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    x[q] = (s[q] / rho) / tr.molarMass[q];
    sum += x[q];
  }
  double weightedAvg = 0.0, harmonicAvg = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    const double xq = x[q] / sum;
    const double sc = SpeciesConductivity(tr, t, q);
    weightedAvg += xq * sc;
    harmonicAvg += xq / sc;
  }
  harmonicAvg = 1.0 / harmonicAvg;
  return (0.5 * (weightedAvg + harmonicAvg)) * tr.scaling;
}
// single-species forms (the laminar scalar path)
AITHER_HD double SutherlandViscosity(const Transport &tr, double t) {
  return SpeciesViscosity(tr, t, 0);
}
AITHER_HD double EffectiveConductivity(const Transport &tr, double t) {
  return SpeciesConductivity(tr, t, 0) * tr.scaling;
}
// turbulent Prandtl number: 0.9 (include/turbulence.hpp:70), k-omega 2006 8/9 (:398), SST 0.9 (:500)
AITHER_HD double TurbPrandtl(int turbModel) {
  return turbModel == AITHER_TURB_KW_WILCOX ? 8.0 / 9.0 : 0.9;
}

// max(4/(3 rho), gamma/rho) * scaling * mu / Pr: the state-dependent factor shared by
// ViscCellSpectralRadius and ViscFaceSpectralRadius (include/spectralRadius.hpp:94-151);
// Pr = 4 gamma / (9 gamma - 5) (include/thermodynamic.hpp:61-64)
AITHER_HD double ViscSpecFactor(const Transport &tr, double rho, double gamma, double mu,
                                double mut = 0.0) {
  const double maxTerm = fmax(4.0 / (3.0 * rho), gamma / rho);
  const double pr = (4.0 * gamma) / (9.0 * gamma - 5.0);
  const double viscTerm = tr.scaling * (mu / pr + mut / TurbPrandtl(tr.turbModel));
  return maxTerm * viscTerm;
}

struct Gas {
  double R[AITHER_MAX_SPECIES];
  double n[AITHER_MAX_SPECIES];
  double hf[AITHER_MAX_SPECIES];
  // single-species constants used by the marching kernels (march.cuh): 1/R, 1/cv, cp/cv
  double rInv0, cvInv0, gamma0;
};
inline void GasFinalize(Gas *g) {
  g->rInv0 = 1.0 / g->R[0];
  g->cvInv0 = 1.0 / (g->R[0] * g->n[0]);
  g->gamma0 = (g->R[0] * (g->n[0] + 1.0)) / (g->R[0] * g->n[0]);
}

// 1/x for the hot loops. The compiler's IEEE fp64 reciprocal / division is ~12 instructions plus
// a guarded slow-path call (a BSSY/BSYNC region per division: 14 % of the residual kernel's stall
// samples, profiles/r01c_*); this is the hardware seed (MUFU.RCP64H, ~20 bits) and two Newton
// steps: 5 instructions, no branch, <= 1 ulp for normal-range arguments (all we feed it:
// densities, 1 + sqrt(rho_R / rho_L), a^2, eps + slope). Host builds (tests/hostsim) divide.
// Bisect builds (scripts/build_bisect.sh; never shipped): AITHER_BISECT_EXACT_RCP = IEEE division
// here, AITHER_BISECT_REF_ROE = the reference-order Roe flux in the marching residual kernel,
// AITHER_BISECT_MUSCL_DIV = MUSCL through the ratio r as the reference forms it.
AITHER_HD double FastRcp(double x) {
#if defined(__CUDA_ARCH__) && !defined(AITHER_BISECT_EXACT_RCP)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(y, fma(-x, y, 1.0), y);
  y = fma(y, fma(-x, y, 1.0), y);
  return y;
#else
  return 1.0 / x;
#endif
}

// Equation layout for NS species and NT turbulence equations
// (ref: include/varArray.hpp:47-51): [rho_1..rho_NS, u, v, w, p, (k, omega)].
template <int NS, int NT>
struct Eq {
  static constexpr int ns = NS, nt = NT, neq = NS + 4 + NT;
  static constexpr int imx = NS, imy = NS + 1, imz = NS + 2, ie = NS + 3, it = NS + 4;
};

template <int NS>
AITHER_HD double SpeciesSum(const double *s) {
  double r = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) r += s[q];
  return r;
}

// T = p / sum(rho_s R_s); ref: src/eos.cpp:100-109
template <int NS>
AITHER_HD double Temperature(const Gas &g, const double *s) {
  double rhoR = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) rhoR += s[q] * g.R[q];
  return s[NS + 3] / rhoR;
}

// mixture thermodynamic sums; ref: src/thermodynamic.cpp:62-104, include/thermodynamic.hpp:102-119
template <int NS>
struct Mix {
  double cp, cv, hfm, rho;
};
template <int NS>
AITHER_HD Mix<NS> Mixture(const Gas &g, const double *s) {
  Mix<NS> m;
  m.rho = SpeciesSum<NS>(s);
  m.cp = 0.0;
  m.cv = 0.0;
  m.hfm = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    const double mf = s[q] / m.rho;
    m.cp += mf * (g.R[q] * (g.n[q] + 1.0));
    m.cv += mf * (g.R[q] * g.n[q]);
    m.hfm += mf * g.hf[q];
  }
  return m;
}

template <int NS>
AITHER_HD double Gamma(const Gas &g, const double *s) {
  const Mix<NS> m = Mixture<NS>(g, s);
  return m.cp / m.cv;
}

// a = sqrt(gamma p / rho); ref: include/arrayView.hpp:384-391
template <int NS>
AITHER_HD double SoS(const Gas &g, const double *s) {
  const Mix<NS> m = Mixture<NS>(g, s);
  return sqrt(m.cp / m.cv * s[NS + 3] / m.rho);
}

template <int NS>
AITHER_HD double VelMagSq(const double *s) {
  return s[NS] * s[NS] + s[NS + 1] * s[NS + 1] + s[NS + 2] * s[NS + 2];
}

// H = sum Y_s (hf_s + cp_s T) + |v|^2 / 2; ref: include/arrayView.hpp:400-408, src/eos.cpp:83-88.
// The reference forms sum_s Y_s (hf_s + cp_s T); for one species that is hf + cp T exactly, for
// several it differs from hfm + cp T only in rounding.
template <int NS>
AITHER_HD double Enthalpy(const Gas &g, const double *s) {
  const double rho = SpeciesSum<NS>(s);
  const double t = Temperature<NS>(g, s);
  double h = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) h += s[q] / rho * (g.hf[q] + (g.R[q] * (g.n[q] + 1.0)) * t);
  const double vel = sqrt(VelMagSq<NS>(s));
  return h + 0.5 * vel * vel;
}

// E = sum Y_s (hf_s + cv_s T) + |v|^2 / 2; ref: include/arrayView.hpp:432-441
template <int NS>
AITHER_HD double Energy(const Gas &g, const double *s) {
  const double rho = SpeciesSum<NS>(s);
  const double t = Temperature<NS>(g, s);
  double e = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) e += s[q] / rho * (g.hf[q] + (g.R[q] * g.n[q]) * t);
  const double vel = sqrt(VelMagSq<NS>(s));
  return e + 0.5 * vel * vel;
}

// primitive -> conserved; ref: include/primitive.hpp:181-199
template <int NS, int NT>
AITHER_HD void PrimToCons(const Gas &g, const double *s, double *c) {
  using E = Eq<NS, NT>;
  const double rho = SpeciesSum<NS>(s);
#pragma unroll
  for (int q = 0; q < NS; ++q) c[q] = s[q];
  c[E::imx] = rho * s[E::imx];
  c[E::imy] = rho * s[E::imy];
  c[E::imz] = rho * s[E::imz];
  c[E::ie] = rho * Energy<NS>(g, s);
#pragma unroll
  for (int t = 0; t < NT; ++t) c[E::it + t] = rho * s[E::it + t];
}

// conserved -> primitive; ref: include/primitive.hpp:150-177, src/eos.cpp:40-63,
// src/thermodynamic.cpp:107-113, src/primitive.cpp:100-106 (turbulence floor)
template <int NS, int NT>
AITHER_HD void ConsToPrim(const Gas &g, const double *c, double *s) {
  using E = Eq<NS, NT>;
  const double rho = SpeciesSum<NS>(c);
#pragma unroll
  for (int q = 0; q < NS; ++q) s[q] = c[q];
  s[E::imx] = c[E::imx] / rho;
  s[E::imy] = c[E::imy] / rho;
  s[E::imz] = c[E::imz] / rho;
  const double energy = c[E::ie] / rho;
  const double vel = sqrt(VelMagSq<NS>(s));
  const double specEnergy = energy - 0.5 * vel * vel;
  double hfm = 0.0, cv = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    const double mf = s[q] / rho;
    hfm += g.hf[q] * mf;
    cv += mf * (g.R[q] * g.n[q]);
  }
  const double temperature = (specEnergy - hfm) / cv;
  double p = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) p += s[q] * g.R[q] * temperature;
  s[E::ie] = p;
#pragma unroll
  for (int t = 0; t < NT; ++t) s[E::it + t] = fmax(c[E::it + t] / rho, kTurbMin);
}

// state + conserved update -> new primitive state, with the reference's mass-fraction clip and
// renormalisation; ref: include/primitive.hpp:206-231
// second half of UpdatePrimWithCons: `c0` = PrimToCons of the old state (independent of the update,
// so a wavefront sweep can form it before the neighbour's new update is known)
template <int NS, int NT>
AITHER_HD void UpdatePrimFromCons(const Gas &g, const double *c0, const double *du, double *out) {
  using E = Eq<NS, NT>;
  double c[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) c[e] = c0[e] + du[e];
  const double rho = SpeciesSum<NS>(c);
  double mf[NS];
  double total = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    mf[q] = fmax(c[q] / rho, 0.0);
    total += mf[q];
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) c[q] = rho * (mf[q] / total);
  ConsToPrim<NS, NT>(g, c, out);
}

template <int NS, int NT>
AITHER_HD void UpdatePrimWithCons(const Gas &g, const double *s, const double *du, double *out) {
  using E = Eq<NS, NT>;
  double c[E::neq];
  PrimToCons<NS, NT>(g, s, c);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) c[e] += du[e];
  const double rho = SpeciesSum<NS>(c);
  double mf[NS];
  double total = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    mf[q] = fmax(c[q] / rho, 0.0);
    total += mf[q];
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) c[q] = rho * (mf[q] / total);
  ConsToPrim<NS, NT>(g, c, out);
}

// ---------------------------------------------------------------------------------------------
// reconstruction
template <int LIM>
AITHER_HD double Limiter(double r) {
  if (LIM == AITHER_LIMITER_VAN_ALBADA) {  // ref: src/limiter.cpp:37-46
    const double r2 = r * r;
    return fmax(0.0, (r + r2) / (1.0 + r2));
  } else if (LIM == AITHER_LIMITER_MINMOD) {  // ref: src/limiter.cpp:24-34
    return fmax(0.0, fmin(1.0, r));
  }
  return 1.0;
}

// kappa-scheme MUSCL on a non-uniform grid, one component;
// ref: include/reconstruction.hpp:110-154 (FaceReconMUSCL)
template <int LIM>
AITHER_HD double Muscl1(double u2, double u1, double d1, double kappa, double dPlus,
                        double dMinus) {
  const double dm = (u1 - u2) * dMinus;
#if defined(__CUDA_ARCH__) && !defined(AITHER_BISECT_MUSCL_DIV)
  if (LIM == AITHER_LIMITER_NONE) {
    // Without a limiter the ratio r only appears as dm * r = dm (eps + dp) / (eps + dm), which is
    // dp to within eps / |dm| = 1e-30 / |dm| relative (|dm| is 0 or at least an ulp of the
    // variable) -- except for dm == 0 exactly, where the reference's product is 0. No division.
    const double dp = (d1 - u1) * dPlus;
    return u1 + 0.25 * ((1.0 - kappa) * dm + (1.0 + kappa) * (dm != 0.0 ? dp : 0.0));
  }
#endif
  const double r = (kEps + (d1 - u1) * dPlus) * FastRcp(kEps + dm);
  double lim = 1.0, invLim = 1.0;
  if (LIM != AITHER_LIMITER_NONE) {
    lim = Limiter<LIM>(r);
    invLim = Limiter<LIM>(1.0 / r);
  }
  return u1 + 0.25 * dm * ((1.0 - kappa) * lim + (1.0 + kappa) * r * invLim);
}

template <int NEQ, int LIM>
AITHER_HD void Muscl(const double *u2, const double *u1, const double *d1, double kappa,
                     double w2, double w1, double wd, double *face) {
  const double dPlus = (w1 + w1) / (w1 + wd);
  const double dMinus = (w1 + w1) / (w1 + w2);
#pragma unroll
  for (int e = 0; e < NEQ; ++e) face[e] = Muscl1<LIM>(u2[e], u1[e], d1[e], kappa, dPlus, dMinus);
}

// ref: include/utility.hpp:103-114 (StencilWidth)
AITHER_HD double StencilWidth(const double *w, int start, int end) {
  double width = 0.0;
  if (end > start) {
    for (int q = start; q < end; ++q) width += w[q];
  } else if (start > end) {
    for (int q = end; q < start; ++q) width += w[q];
    width = -1.0 * width;
  }
  return width;
}

// Lagrange reconstruction coefficients on a non-uniform stencil;
// ref: src/utility.cpp:449-483 (LagrangeCoeff; Shu ICASE 97-65 eq. 2.20)
template <int DEGREE>
AITHER_HD void LagrangeCoeff(const double *w, int rr, int ii, double *coeffs) {
#pragma unroll
  for (int jj = 0; jj <= DEGREE; ++jj) {
    double cj = 0.0;
#pragma unroll
    for (int mm = jj + 1; mm <= DEGREE + 1; ++mm) {
      double numer = 0.0, denom = 1.0;
#pragma unroll
      for (int ll = 0; ll <= DEGREE + 1; ++ll) {
        if (ll != mm) {
          double numProd = 1.0;
#pragma unroll
          for (int qq = 0; qq <= DEGREE + 1; ++qq) {
            if (qq != mm && qq != ll) numProd *= StencilWidth(w, ii - rr + qq, ii + 1);
          }
          numer += numProd;
          denom *= StencilWidth(w, ii - rr + ll, ii - rr + mm);
        }
      }
      cj += numer * FastRcp(denom);
    }
    coeffs[jj] = cj * w[ii - rr + jj];
  }
}

AITHER_HD double Deriv2nd(double x0, double x1, double x2, double y0, double y1, double y2) {
  // ref: include/utility.hpp:116-122
  const double fwd = (y2 - y1) / (0.5 * (x2 + x1));
  const double bck = (y1 - y0) / (0.5 * (x1 + x0));
  return (fwd - bck) / (0.25 * (x2 + x0) + 0.5 * x1);
}
AITHER_HD double BetaIntegral1(double d1, double d2, double dx, double x) {
  // ref: include/reconstruction.hpp:157-170
  return ((d1 * d1) * x + d1 * d2 * x * x + (d2 * d2) * (x * x * x) / 3.0) * dx +
         (d2 * d2) * x * (dx * dx * dx);
}
AITHER_HD double BetaIntegral(double d1, double d2, double dx, double xl, double xh) {
  return BetaIntegral1(d1, d2, dx, xh) - BetaIntegral1(d1, d2, dx, xl);
}

// WENO5 / WENO-Z quantities that depend only on the five cell widths. Besides the Lagrange
// coefficients and linear weights this holds the reciprocals of every width combination the
// smoothness indicators divide by (the reference re-divides per variable: Deriv2nd,
// include/utility.hpp:116-122, and the d1 terms of include/reconstruction.hpp:185-240 -- 12
// divisions per variable per side) and the two width factors of the closed-form beta integral.
struct WenoGeom {
  double c0[3], c1[3], c2[3];
  double lw0, lw1, lw2;
  double rA, rB, rC, rD;  // 1 / (0.5 (w1 + w0)), ... (w2 + w1), (w3 + w2), (w4 + w3)
  double q0, q1, q2;      // 1 / (0.25 (w2 + w0) + 0.5 w1), ... (w3 + w1) .. w2, (w4 + w2) .. w3
  double hw;              // 0.5 w2
  double bA1, bA2;        // beta = d1^2 bA1 + d2^2 bA2
};
AITHER_HD WenoGeom WenoSetup(const double *w) {
  // ref: include/reconstruction.hpp:256-282
  WenoGeom g;
  double fc[5];
  LagrangeCoeff<2>(w, 2, 2, g.c0);
  LagrangeCoeff<2>(w, 1, 2, g.c1);
  LagrangeCoeff<2>(w, 0, 2, g.c2);
  LagrangeCoeff<4>(w, 2, 2, fc);
  g.lw0 = fc[0] * FastRcp(g.c0[0]);
  g.lw1 = fc[4] * FastRcp(g.c2[2]);
  g.lw2 = 1.0 - g.lw0 - g.lw1;
  g.rA = FastRcp(0.5 * (w[1] + w[0]));
  g.rB = FastRcp(0.5 * (w[2] + w[1]));
  g.rC = FastRcp(0.5 * (w[3] + w[2]));
  g.rD = FastRcp(0.5 * (w[4] + w[3]));
  g.q0 = FastRcp(0.25 * (w[2] + w[0]) + 0.5 * w[1]);
  g.q1 = FastRcp(0.25 * (w[3] + w[1]) + 0.5 * w[2]);
  g.q2 = FastRcp(0.25 * (w[4] + w[2]) + 0.5 * w[3]);
  g.hw = 0.5 * w[2];
  // BetaIntegral(d1, d2, dx = w2, -w2/2, +w2/2) (include/reconstruction.hpp:157-183): the terms odd
  // in x double, the even one cancels: d1^2 (2 h dx) + d2^2 (2 h^3 dx / 3 + 2 h dx^3), h = w2 / 2
  const double dx = w[2], h = g.hw;
  g.bA1 = 2.0 * h * dx;
  g.bA2 = 2.0 * (h * h * h) * dx / 3.0 + 2.0 * h * (dx * dx * dx);
  return g;
}
// one component; y = {upwind3, upwind2, upwind1, downwind1, downwind2};
// ref: include/reconstruction.hpp:185-240 (Beta0/1/2), :284-310
template <bool WENOZ>
AITHER_HD double Weno1(const WenoGeom &g, double y0, double y1, double y2, double y3, double y4) {
  const double st0 = g.c0[0] * y0 + g.c0[1] * y1 + g.c0[2] * y2;
  const double st1 = g.c1[0] * y1 + g.c1[1] * y2 + g.c1[2] * y3;
  const double st2 = g.c2[0] * y2 + g.c2[1] * y3 + g.c2[2] * y4;
  const double e1 = (y1 - y0) * g.rA, e2 = (y2 - y1) * g.rB, e3 = (y3 - y2) * g.rC,
               e4 = (y4 - y3) * g.rD;
  double d2 = (e2 - e1) * g.q0;
  double d1 = e2 + g.hw * d2;
  const double b0 = (d1 * d1) * g.bA1 + (d2 * d2) * g.bA2;
  d2 = (e3 - e2) * g.q1;
  d1 = e3 - g.hw * d2;
  const double b1 = (d1 * d1) * g.bA1 + (d2 * d2) * g.bA2;
  d2 = (e4 - e3) * g.q2;
  d1 = e3 - g.hw * d2;
  const double b2 = (d1 * d1) * g.bA1 + (d2 * d2) * g.bA2;
  double n0, n1, n2;
  if (WENOZ) {
    const double tau5 = fabs(b0 - b2);
    const double eps = 1.0e-40;
    double t = tau5 * FastRcp(eps + b0);
    n0 = g.lw0 * (1.0 + t * t);
    t = tau5 * FastRcp(eps + b1);
    n1 = g.lw1 * (1.0 + t * t);
    t = tau5 * FastRcp(eps + b2);
    n2 = g.lw2 * (1.0 + t * t);
  } else {
    const double eps = 1.0e-6;
    n0 = g.lw0 * FastRcp((eps + b0) * (eps + b0));
    n1 = g.lw1 * FastRcp((eps + b1) * (eps + b1));
    n2 = g.lw2 * FastRcp((eps + b2) * (eps + b2));
  }
  return (n0 * st0 + n1 * st1 + n2 * st2) * FastRcp(n0 + n1 + n2);
}

// ---------------------------------------------------------------------------------------------
// inviscid fluxes (unit normal n)
// ref: include/inviscidFlux.hpp:128-159 (ConstructFromPrim)
template <int NS, int NT>
AITHER_HD void PhysicalFlux(const Gas &g, const double *s, const double *n, double *f) {
  using E = Eq<NS, NT>;
  const double velNorm = s[E::imx] * n[0] + s[E::imy] * n[1] + s[E::imz] * n[2];
#pragma unroll
  for (int q = 0; q < NS; ++q) f[q] = s[q] * velNorm;
  const double rho = SpeciesSum<NS>(s);
  const double p = s[E::ie];
  f[E::imx] = rho * velNorm * s[E::imx] + p * n[0];
  f[E::imy] = rho * velNorm * s[E::imy] + p * n[1];
  f[E::imz] = rho * velNorm * s[E::imz] + p * n[2];
  f[E::ie] = rho * velNorm * Enthalpy<NS>(g, s);
#pragma unroll
  for (int t = 0; t < NT; ++t) f[E::it + t] = rho * velNorm * s[E::it + t];
}

// Roe flux-difference splitting with Harten's entropy fix;
// ref: include/inviscidFlux.hpp:260-382 (RoeFlux), include/primitive.hpp:245-280 (Roe average:
// density-weighted primitive state including *pressure*), src/inviscidFlux.cpp:27-33
template <int NS, int NT>
AITHER_HD void RoeFlux(const Gas &g, const double *l, const double *r, const double *n,
                       double *flux) {
  using E = Eq<NS, NT>;
  double roe[E::neq];
  const double denRatio = sqrt(SpeciesSum<NS>(r) / SpeciesSum<NS>(l));
#pragma unroll
  for (int q = 0; q < NS; ++q) roe[q] = l[q] * denRatio;
#pragma unroll
  for (int e = NS; e < E::neq; ++e) roe[e] = (l[e] + denRatio * r[e]) / (1.0 + denRatio);

  const double hR = Enthalpy<NS>(g, roe);
  const double aR = SoS<NS>(g, roe);
  const double rhoR = SpeciesSum<NS>(roe);
  const double velNormR = roe[E::imx] * n[0] + roe[E::imy] * n[1] + roe[E::imz] * n[2];
  double delta[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) delta[e] = r[e] - l[e];
  const double deltaRho = SpeciesSum<NS>(delta);
  const double normVelDiff = delta[E::imx] * n[0] + delta[E::imy] * n[1] + delta[E::imz] * n[2];
  const double dP = delta[E::ie];

  double diss[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) diss[e] = 0.0;

  // left-moving acoustic wave
  double waveSpeed = fabs(velNormR - aR);
  if (waveSpeed < kEntropyFix) waveSpeed = 0.5 * (waveSpeed * waveSpeed / kEntropyFix + kEntropyFix);
  double waveStrength = (dP - rhoR * aR * normVelDiff) / (2.0 * aR * aR);
  double wss = waveSpeed * waveStrength;
#pragma unroll
  for (int q = 0; q < NS; ++q) diss[q] += wss * (roe[q] / rhoR);
  diss[E::imx] += wss * (roe[E::imx] - aR * n[0]);
  diss[E::imy] += wss * (roe[E::imy] - aR * n[1]);
  diss[E::imz] += wss * (roe[E::imz] - aR * n[2]);
  diss[E::ie] += wss * (hR - aR * velNormR);
#pragma unroll
  for (int t = 0; t < NT; ++t) diss[E::it + t] += wss * roe[E::it + t];

  // entropy wave
  waveSpeed = fabs(velNormR);
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    waveStrength = -dP / (aR * aR);
    wss = waveSpeed * waveStrength;
    diss[q] += wss * (roe[q] / rhoR) + waveSpeed * delta[q];
  }
  waveStrength = deltaRho - dP / (aR * aR);
  wss = waveSpeed * waveStrength;
  diss[E::imx] += wss * roe[E::imx];
  diss[E::imy] += wss * roe[E::imy];
  diss[E::imz] += wss * roe[E::imz];
  diss[E::ie] += wss * 0.5 * VelMagSq<NS>(roe);

  // shear wave
  wss = waveSpeed * rhoR;
  diss[E::imx] += wss * (delta[E::imx] - normVelDiff * n[0]);
  diss[E::imy] += wss * (delta[E::imy] - normVelDiff * n[1]);
  diss[E::imz] += wss * (delta[E::imz] - normVelDiff * n[2]);
  diss[E::ie] += wss * ((roe[E::imx] * delta[E::imx] + roe[E::imy] * delta[E::imy] +
                         roe[E::imz] * delta[E::imz]) -
                        velNormR * normVelDiff);

  // right-moving acoustic wave
  waveSpeed = fabs(velNormR + aR);
  if (waveSpeed < kEntropyFix) waveSpeed = 0.5 * (waveSpeed * waveSpeed / kEntropyFix + kEntropyFix);
  waveStrength = (dP + rhoR * aR * normVelDiff) / (2.0 * aR * aR);
  wss = waveSpeed * waveStrength;
#pragma unroll
  for (int q = 0; q < NS; ++q) diss[q] += wss * (roe[q] / rhoR);
  diss[E::imx] += wss * (roe[E::imx] + aR * n[0]);
  diss[E::imy] += wss * (roe[E::imy] + aR * n[1]);
  diss[E::imz] += wss * (roe[E::imz] + aR * n[2]);
  diss[E::ie] += wss * (hR + aR * velNormR);
#pragma unroll
  for (int t = 0; t < NT; ++t) diss[E::it + t] += wss * roe[E::it + t];

  // turbulence waves
  if (NT > 0) {
    waveSpeed = fabs(velNormR);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      waveStrength = rhoR * delta[E::it + t] + roe[E::it + t] * deltaRho -
                     dP * roe[E::it + t] / (aR * aR);
      diss[E::it + t] += waveSpeed * waveStrength * 1.0;
    }
  }

  double fl[E::neq], fr[E::neq];
  PhysicalFlux<NS, NT>(g, l, n, fl);
  PhysicalFlux<NS, NT>(g, r, n, fr);
#pragma unroll
  for (int e = 0; e < E::neq; ++e) flux[e] = (fl[e] + (fr[e] - diss[e])) * 0.5;
}

AITHER_HD double Sign(double v) { return static_cast<double>((0.0 < v) - (v < 0.0)); }

// AUSMPW+ (Kim, Kim & Rho 1998); ref: include/inviscidFlux.hpp:396-481, :161-208
template <int NS, int NT>
AITHER_HD void AusmFlux(const Gas &g, const double *l, const double *r, const double *n,
                        double *f) {
  using E = Eq<NS, NT>;
  const double velNormL = l[E::imx] * n[0] + l[E::imy] * n[1] + l[E::imz] * n[2];
  const double velNormR = r[E::imx] * n[0] + r[E::imy] * n[1] + r[E::imz] * n[2];
  const double sosL = SoS<NS>(g, l);
  const double sosR = SoS<NS>(g, r);
  const double sosStar = sqrt(sosL * sosR);
  const double vel = 0.5 * (velNormL + velNormR);
  double sos = sosStar;
  if (vel < 0.0) {
    sos = sosStar * sosStar / fmax(velNormR, sosStar);
  } else if (vel > 0.0) {
    sos = sosStar * sosStar / fmax(velNormL, sosStar);
  }
  const double ml = velNormL / sos;
  const double mr = velNormR / sos;
  const double mPlusL = fabs(ml) <= 1.0 ? 0.25 * ((ml + 1.0) * (ml + 1.0)) : 0.5 * (ml + fabs(ml));
  const double mMinusR =
      fabs(mr) <= 1.0 ? -0.25 * ((mr - 1.0) * (mr - 1.0)) : 0.5 * (mr - fabs(mr));
  const double pPlus = fabs(ml) <= 1.0 ? 0.25 * ((ml + 1.0) * (ml + 1.0)) * (2.0 - ml)
                                       : 0.5 * (1.0 + Sign(ml));
  const double pMinus = fabs(mr) <= 1.0 ? 0.25 * ((mr - 1.0) * (mr - 1.0)) * (2.0 + mr)
                                        : 0.5 * (1.0 - Sign(mr));
  const double pl = l[E::ie], pr = r[E::ie];
  const double ps = pPlus * pl + pMinus * pr;
  const double ratio = fmin(pl / pr, pr / pl);
  const double w = 1.0 - ratio * ratio * ratio;
  const double fl = fabs(ml) < 1.0 ? pl / ps - 1.0 : 0.0;
  const double fr = fabs(mr) < 1.0 ? pr / ps - 1.0 : 0.0;
  const double mavg = mPlusL + mMinusR;
  const double mPlusLBar = mavg >= 0.0 ? mPlusL + mMinusR * ((1.0 - w) * (1.0 + fr) - fl)
                                       : mPlusL * w * (1.0 + fl);
  const double mMinusRBar = mavg >= 0.0 ? mMinusR * w * (1.0 + fr)
                                        : mMinusR + mPlusL * ((1.0 - w) * (1.0 + fl) - fr);
  const double vl = mPlusLBar * sos;
  const double vr = mMinusRBar * sos;
  const double rhoL = SpeciesSum<NS>(l);
  const double rhoR = SpeciesSum<NS>(r);
#pragma unroll
  for (int q = 0; q < NS; ++q) f[q] = l[q] * vl + r[q] * vr;
  f[E::imx] = (rhoL * vl * l[E::imx] + pPlus * pl * n[0]) +
              (rhoR * vr * r[E::imx] + pMinus * pr * n[0]);
  f[E::imy] = (rhoL * vl * l[E::imy] + pPlus * pl * n[1]) +
              (rhoR * vr * r[E::imy] + pMinus * pr * n[1]);
  f[E::imz] = (rhoL * vl * l[E::imz] + pPlus * pl * n[2]) +
              (rhoR * vr * r[E::imz] + pMinus * pr * n[2]);
  f[E::ie] = rhoL * vl * Enthalpy<NS>(g, l) + rhoR * vr * Enthalpy<NS>(g, r);
#pragma unroll
  for (int t = 0; t < NT; ++t) f[E::it + t] = rhoL * vl * l[E::it + t] + rhoR * vr * r[E::it + t];
}

template <int NS, int NT, int FLUX>
AITHER_HD void InviscidFlux(const Gas &g, const double *l, const double *r, const double *n,
                            double *f) {
  if (FLUX == AITHER_FLUX_ROE) RoeFlux<NS, NT>(g, l, r, n, f);
  else AusmFlux<NS, NT>(g, l, r, n, f);
}

// ---------------------------------------------------------------------------------------------
// spectral radii; ref: include/spectralRadius.hpp:44-80
template <int NS>
AITHER_HD double InvCellSpectralRadius(const double *s, double sos, const double *fL,
                                       const double *fR) {
  double a0 = 0.5 * (fL[0] + fR[0]), a1 = 0.5 * (fL[1] + fR[1]), a2 = 0.5 * (fL[2] + fR[2]);
  const double rmag = FastRcp(sqrt(a0 * a0 + a1 * a1 + a2 * a2));
  a0 *= rmag;
  a1 *= rmag;
  a2 *= rmag;
  const double fMag = 0.5 * (fL[3] + fR[3]);
  return (fabs(s[NS] * a0 + s[NS + 1] * a1 + s[NS + 2] * a2) + sos) * fMag;
}
// the same with the turbulence-equation convective part |v.n| A beside it
// (turbModel::InviscidCellSpectralRadius, ref: src/turbulence.cpp:152-160)
template <int NS>
AITHER_HD double InvCellSpectralRadii(const double *s, double sos, const double *fL,
                                      const double *fR, double *turb) {
  double a0 = 0.5 * (fL[0] + fR[0]), a1 = 0.5 * (fL[1] + fR[1]), a2 = 0.5 * (fL[2] + fR[2]);
  const double rmag = FastRcp(sqrt(a0 * a0 + a1 * a1 + a2 * a2));
  a0 *= rmag;
  a1 *= rmag;
  a2 *= rmag;
  const double fMag = 0.5 * (fL[3] + fR[3]);
  const double vn = fabs(s[NS] * a0 + s[NS + 1] * a1 + s[NS + 2] * a2);
  *turb = vn * fMag;
  return (vn + sos) * fMag;
}
template <int NS>
AITHER_HD double InvFaceSpectralRadius(const double *s, double sos, const double *fA) {
  return 0.5 * fA[3] * (fabs(s[NS] * fA[0] + s[NS + 1] * fA[1] + s[NS + 2] * fA[2]) + sos);
}

// Scalar (LU-SGS / DPLUR) off-diagonal product for one neighbour, inviscid:
// 0.5 |A| (F(U + dU) - F(U)) +- lambda_face dU, turbulence rows of the flux change zeroed;
// ref: src/fluxJacobian.cpp:122-162 (RusanovScalarOffDiagonal)
template <int NS, int NT>
AITHER_HD void OffDiagScalar(const Gas &g, const double *state, const double *du,
                             const double *fArea, bool positive, double *out,
                             double srExtra = 0.0, double srTurbVisc = 0.0) {
  using E = Eq<NS, NT>;
  double su[E::neq], fo[E::neq], fn[E::neq];
  UpdatePrimWithCons<NS, NT>(g, state, du, su);
  PhysicalFlux<NS, NT>(g, state, fArea, fo);
  PhysicalFlux<NS, NT>(g, su, fArea, fn);
  const double sr = InvFaceSpectralRadius<NS>(state, SoS<NS>(g, state), fArea) + srExtra;
  // turbulence equations: turbModel::FaceSpectralRadius = 0.5 |A| |vn +- |vn|| + viscous part
  // (ref: src/turbulence.cpp:162-171, include/turbulence.hpp:308-329)
  double srT = 0.0;
  if (NT > 0) {
    const double velNorm = state[NS] * fArea[0] + state[NS + 1] * fArea[1] + state[NS + 2] * fArea[2];
    srT = (positive ? 0.5 * fArea[3] * fabs(velNorm + fabs(velNorm))
                    : 0.5 * fArea[3] * fabs(velNorm - fabs(velNorm))) + srTurbVisc;
  }
#pragma unroll
  for (int e = 0; e < E::neq; ++e) {
    const double fc = e < NS + 4 ? 0.5 * fArea[3] * (fn[e] - fo[e]) : 0.0;
    const double srd = (e < NS + 4 ? sr : srT) * du[e];
    out[e] = positive ? fc + srd : fc - srd;
  }
}

// =============================================================================================
// Restructured point functions used by the plane-marching kernels (march.cuh). Same formulas as
// above with shared reciprocals; each differs from its twin by rounding only (tests/hostsim).
// ---------------------------------------------------------------------------------------------
// cheap thermodynamics. cp/cv sums follow src/thermodynamic.cpp:62-104; for one species the mass
// fraction is exactly 1 and the sums collapse to constants.
template <int NS>
struct MixK {
  double rho, rhoInv, cp, cv, cvInv, hf;
  double tFac;    // T = p * tFac          (ref src/eos.cpp:100-109: T = p / sum(rho_s R_s))
  double gamma;   // cp / cv
};
template <int NS>
AITHER_HD MixK<NS> MixOf(const Gas &g, const double *s) {
  MixK<NS> m;
  if (NS == 1) {
    m.rho = s[0];
    m.rhoInv = FastRcp(s[0]);
    m.cp = g.R[0] * (g.n[0] + 1.0);
    m.cv = g.R[0] * g.n[0];
    m.cvInv = g.cvInv0;
    m.hf = g.hf[0];
    m.tFac = m.rhoInv * g.rInv0;
    m.gamma = g.gamma0;
  } else {
    m.rho = SpeciesSum<NS>(s);
    m.rhoInv = 1.0 / m.rho;
    m.cp = 0.0; m.cv = 0.0; m.hf = 0.0;
    double rhoR = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const double mf = s[q] * m.rhoInv;
      m.cp += mf * (g.R[q] * (g.n[q] + 1.0));
      m.cv += mf * (g.R[q] * g.n[q]);
      m.hf += mf * g.hf[q];
      rhoR += s[q] * g.R[q];
    }
    m.cvInv = 1.0 / m.cv;
    m.tFac = 1.0 / rhoR;
    m.gamma = m.cp * m.cvInv;
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// Roe flux, same wave decomposition as RoeFlux (physics.cuh; ref include/inviscidFlux.hpp:260-382)
// with shared reciprocals. Returns flux * 1 (unit normal n), not yet times area.
template <int NS, int NT>
AITHER_HD void RoeFluxFast(const Gas &g, const double *l, const double *r,
                                            const double *n, double *flux) {
  using E = Eq<NS, NT>;
  const MixK<NS> ml = MixOf<NS>(g, l);
  const MixK<NS> mr = MixOf<NS>(g, r);
  const double denRatio = sqrt(mr.rho * ml.rhoInv);
  const double inv1p = FastRcp(1.0 + denRatio);
  double roe[E::neq];
#pragma unroll
  for (int q = 0; q < NS; ++q) roe[q] = l[q] * denRatio;
#pragma unroll
  for (int e = NS; e < E::neq; ++e) roe[e] = (l[e] + denRatio * r[e]) * inv1p;
  const MixK<NS> mm = MixOf<NS>(g, roe);
  const double q2 = VelMagSq<NS>(roe);
  const double tR = roe[E::ie] * mm.tFac;
  const double hR = mm.hf + mm.cp * tR + 0.5 * q2;
  const double a2 = mm.gamma * roe[E::ie] * mm.rhoInv;
  const double aR = sqrt(a2);
  const double a2inv = FastRcp(aR * aR);
  const double rhoR = mm.rho;
  const double velNormR = roe[E::imx] * n[0] + roe[E::imy] * n[1] + roe[E::imz] * n[2];
  double delta[E::neq];
#pragma unroll
  for (int e = 0; e < E::neq; ++e) delta[e] = r[e] - l[e];
  const double deltaRho = SpeciesSum<NS>(delta);
  const double nvd = delta[E::imx] * n[0] + delta[E::imy] * n[1] + delta[E::imz] * n[2];
  const double dP = delta[E::ie];
  double diss[E::neq];
  double mfR[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) mfR[q] = NS == 1 ? 1.0 : roe[q] * mm.rhoInv;

  // left-moving acoustic wave
  double ws = fabs(velNormR - aR);
  if (ws < kEntropyFix) ws = 0.5 * (ws * ws / kEntropyFix + kEntropyFix);
  double wss = ws * ((dP - rhoR * aR * nvd) * (0.5 * a2inv));
#pragma unroll
  for (int q = 0; q < NS; ++q) diss[q] = wss * mfR[q];
  diss[E::imx] = wss * (roe[E::imx] - aR * n[0]);
  diss[E::imy] = wss * (roe[E::imy] - aR * n[1]);
  diss[E::imz] = wss * (roe[E::imz] - aR * n[2]);
  diss[E::ie] = wss * (hR - aR * velNormR);
#pragma unroll
  for (int t = 0; t < NT; ++t) diss[E::it + t] = wss * roe[E::it + t];
  // entropy wave
  ws = fabs(velNormR);
  const double dPa2 = dP * a2inv;
#pragma unroll
  for (int q = 0; q < NS; ++q) diss[q] += (ws * (-dPa2)) * mfR[q] + ws * delta[q];
  wss = ws * (deltaRho - dPa2);
  diss[E::imx] += wss * roe[E::imx];
  diss[E::imy] += wss * roe[E::imy];
  diss[E::imz] += wss * roe[E::imz];
  diss[E::ie] += wss * 0.5 * q2;
  // shear wave
  wss = ws * rhoR;
  diss[E::imx] += wss * (delta[E::imx] - nvd * n[0]);
  diss[E::imy] += wss * (delta[E::imy] - nvd * n[1]);
  diss[E::imz] += wss * (delta[E::imz] - nvd * n[2]);
  diss[E::ie] += wss * ((roe[E::imx] * delta[E::imx] + roe[E::imy] * delta[E::imy] +
                         roe[E::imz] * delta[E::imz]) -
                        velNormR * nvd);
  // right-moving acoustic wave
  ws = fabs(velNormR + aR);
  if (ws < kEntropyFix) ws = 0.5 * (ws * ws / kEntropyFix + kEntropyFix);
  wss = ws * ((dP + rhoR * aR * nvd) * (0.5 * a2inv));
#pragma unroll
  for (int q = 0; q < NS; ++q) diss[q] += wss * mfR[q];
  diss[E::imx] += wss * (roe[E::imx] + aR * n[0]);
  diss[E::imy] += wss * (roe[E::imy] + aR * n[1]);
  diss[E::imz] += wss * (roe[E::imz] + aR * n[2]);
  diss[E::ie] += wss * (hR + aR * velNormR);
#pragma unroll
  for (int t = 0; t < NT; ++t) diss[E::it + t] += wss * roe[E::it + t];
  if (NT > 0) {
    ws = fabs(velNormR);
#pragma unroll
    for (int t = 0; t < NT; ++t)
      diss[E::it + t] += ws * (rhoR * delta[E::it + t] + roe[E::it + t] * deltaRho -
                               dPa2 * roe[E::it + t]);
  }
  // physical fluxes of the two states (ref include/inviscidFlux.hpp:128-159)
  const double vnL = l[E::imx] * n[0] + l[E::imy] * n[1] + l[E::imz] * n[2];
  const double vnR = r[E::imx] * n[0] + r[E::imy] * n[1] + r[E::imz] * n[2];
  const double hL = ml.hf + ml.cp * (l[E::ie] * ml.tFac) + 0.5 * VelMagSq<NS>(l);
  const double hRt = mr.hf + mr.cp * (r[E::ie] * mr.tFac) + 0.5 * VelMagSq<NS>(r);
  const double mL = ml.rho * vnL, mR = mr.rho * vnR;
#pragma unroll
  for (int q = 0; q < NS; ++q) flux[q] = (l[q] * vnL + (r[q] * vnR - diss[q])) * 0.5;
  flux[E::imx] = ((mL * l[E::imx] + l[E::ie] * n[0]) +
                  ((mR * r[E::imx] + r[E::ie] * n[0]) - diss[E::imx])) * 0.5;
  flux[E::imy] = ((mL * l[E::imy] + l[E::ie] * n[1]) +
                  ((mR * r[E::imy] + r[E::ie] * n[1]) - diss[E::imy])) * 0.5;
  flux[E::imz] = ((mL * l[E::imz] + l[E::ie] * n[2]) +
                  ((mR * r[E::imz] + r[E::ie] * n[2]) - diss[E::imz])) * 0.5;
  flux[E::ie] = (mL * hL + (mR * hRt - diss[E::ie])) * 0.5;
#pragma unroll
  for (int t = 0; t < NT; ++t)
    flux[E::it + t] = (mL * l[E::it + t] + (mR * r[E::it + t] - diss[E::it + t])) * 0.5;
}

template <int NS, int NT, int FLUX>
AITHER_HD void InviscidFluxFast(const Gas &g, const double *l, const double *r,
                                                 const double *n, double *f) {
#ifdef AITHER_BISECT_REF_ROE
  if (FLUX == AITHER_FLUX_ROE) RoeFlux<NS, NT>(g, l, r, n, f);
#else
  if (FLUX == AITHER_FLUX_ROE) RoeFluxFast<NS, NT>(g, l, r, n, f);
#endif
  else AusmFlux<NS, NT>(g, l, r, n, f);
}

// ---------------------------------------------------------------------------------------------
// implicit sweep. Ingredients of one cell, as the off-diagonal product of every neighbour needs
// them (ref src/fluxJacobian.cpp:122-162 RusanovScalarOffDiagonal): old primitive state, its
// enthalpy and sound speed, the conserved update, and the updated primitive state + enthalpy.
template <int NS, int NT>
struct Ingr {
  static constexpr int neq = NS + 4 + NT;
  // s[neq] H a | du[neq] | sn[neq] Hn | vt (viscous spectral factor, 0 when inviscid)
  static constexpr int n = 3 * neq + 4;
  static constexpr int ivt = 3 * neq + 3;
};

// state + dU -> Ingr; ref include/primitive.hpp:206-231 (UpdatePrimWithCons), :150-177
template <int NS, int NT>
AITHER_HD void MakeIngr(const Gas &g, const double *s, const double *du,
                                         double *H, double *a, double *sn, double *Hn) {
  using E = Eq<NS, NT>;
  const MixK<NS> m = MixOf<NS>(g, s);
  const double q2 = VelMagSq<NS>(s);
  const double t = s[E::ie] * m.tFac;
  *H = m.hf + m.cp * t + 0.5 * q2;
  *a = sqrt(m.gamma * s[E::ie] * m.rhoInv);
  // conserved + update
  double c[E::neq];
#pragma unroll
  for (int q = 0; q < NS; ++q) c[q] = s[q] + du[q];
  c[E::imx] = m.rho * s[E::imx] + du[E::imx];
  c[E::imy] = m.rho * s[E::imy] + du[E::imy];
  c[E::imz] = m.rho * s[E::imz] + du[E::imz];
  c[E::ie] = m.rho * (m.hf + m.cv * t + 0.5 * q2) + du[E::ie];
#pragma unroll
  for (int tq = 0; tq < NT; ++tq) c[E::it + tq] = m.rho * s[E::it + tq] + du[E::it + tq];
  double rho = SpeciesSum<NS>(c);
  double rhoInv = FastRcp(rho);  // the same reciprocal MixOf(sn) forms below (one species)
  if (NS > 1) {
    double mf[NS], total = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      mf[q] = fmax(c[q] * rhoInv, 0.0);
      total += mf[q];
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) c[q] = rho * (mf[q] / total);
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) sn[q] = c[q];
  sn[E::imx] = c[E::imx] * rhoInv;
  sn[E::imy] = c[E::imy] * rhoInv;
  sn[E::imz] = c[E::imz] * rhoInv;
  const MixK<NS> mn = MixOf<NS>(g, sn);
  const double q2n = VelMagSq<NS>(sn);
  const double tn = ((c[E::ie] * rhoInv - 0.5 * q2n) - mn.hf) * mn.cvInv;
  double rhoRn = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) rhoRn += sn[q] * g.R[q];
  sn[E::ie] = rhoRn * tn;
#pragma unroll
  for (int tq = 0; tq < NT; ++tq) sn[E::it + tq] = fmax(c[E::it + tq] * rhoInv, kTurbMin);
  *Hn = mn.hf + mn.cp * tn + 0.5 * q2n;
}

// off-diagonal product of one neighbour from its ingredients and the shared face's area
// `srExtra`: viscous part of the face spectral radius, |A| / dist * max(4/(3 rho), gamma/rho) *
// scaling * mu / Pr of the neighbour (ref include/spectralRadius.hpp:126-151,180-200)
// `srExtraT`: the same for the turbulence equations (turbModel::ViscousFaceSpectralRadius,
// src/turbulence.cpp:513-527,796-808)
template <int NS, int NT, typename LD>
AITHER_HD void OffDiagFromIngr(LD ld, const double *fA, bool positive,
                                                double *acc, double srExtra = 0.0,
                                                double srExtraT = 0.0) {
  using E = Eq<NS, NT>;
  constexpr int neq = E::neq;
  // layout: [0,neq) s | neq H | neq+1 a | [neq+2, 2neq+2) du | [2neq+2, 3neq+2) sn | 3neq+2 Hn
  const double u = ld(E::imx), v = ld(E::imy), w = ld(E::imz), pr = ld(E::ie);
  const double un = ld(2 * neq + 2 + E::imx), vn_ = ld(2 * neq + 2 + E::imy),
               wn = ld(2 * neq + 2 + E::imz), pn = ld(2 * neq + 2 + E::ie);
  const double vo = u * fA[0] + v * fA[1] + w * fA[2];
  const double vnw = un * fA[0] + vn_ * fA[1] + wn * fA[2];
  double rho = 0.0, rhon = 0.0;
  const double half = 0.5 * fA[3];
  const double sr = half * (fabs(vo) + ld(neq + 1)) + srExtra;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    const double r0 = ld(q), r1 = ld(2 * neq + 2 + q);
    rho += r0;
    rhon += r1;
    const double fc = half * (r1 * vnw - r0 * vo);
    const double srd = sr * ld(neq + 2 + q);
    acc[q] += positive ? fc + srd : fc - srd;
  }
  const double mo = rho * vo, mn = rhon * vnw;
  {
    const double fc = half * ((mn * un + pn * fA[0]) - (mo * u + pr * fA[0]));
    const double srd = sr * ld(neq + 2 + E::imx);
    acc[E::imx] += positive ? fc + srd : fc - srd;
  }
  {
    const double fc = half * ((mn * vn_ + pn * fA[1]) - (mo * v + pr * fA[1]));
    const double srd = sr * ld(neq + 2 + E::imy);
    acc[E::imy] += positive ? fc + srd : fc - srd;
  }
  {
    const double fc = half * ((mn * wn + pn * fA[2]) - (mo * w + pr * fA[2]));
    const double srd = sr * ld(neq + 2 + E::imz);
    acc[E::imz] += positive ? fc + srd : fc - srd;
  }
  {
    const double fc = half * (mn * ld(3 * neq + 2) - mo * ld(neq));
    const double srd = sr * ld(neq + 2 + E::ie);
    acc[E::ie] += positive ? fc + srd : fc - srd;
  }
  // turbulence rows: flux change zeroed (ref src/fluxJacobian.cpp:146-149); their own spectral
  // radius 0.5 |A| |vn +- |vn|| + viscous part (src/turbulence.cpp:162-171,
  // include/turbulence.hpp:308-329), as OffDiagScalar above
  if (NT > 0) {
    const double srT = (positive ? half * fabs(vo + fabs(vo)) : half * fabs(vo - fabs(vo))) + srExtraT;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const double srd = srT * ld(neq + 2 + E::it + t);
      acc[E::it + t] += positive ? srd : 0.0 - srd;
    }
  }
}

// the update-dependent ingredients alone (updated primitive state and its enthalpy)
template <int NS, int NT>
AITHER_HD void MakeIngrDyn(const Gas &g, const double *s, const double *du, double *sn,
                           double *Hn) {
  double H, a;
  MakeIngr<NS, NT>(g, s, du, &H, &a, sn, Hn);
}

// ---------------------------------------------------------------------------------------------
// boundary-condition ghost states; ref: src/ghostStates.cpp:62-708
template <int NS, int NT>
AITHER_HD void ExtrapolateHoldMixture(const double *bnd, double factor, const double *interior,
                                      double *out) {
  // ref: src/ghostStates.cpp:691-708
  using E = Eq<NS, NT>;
  const double bndRho = SpeciesSum<NS>(bnd);
  const double intRho = SpeciesSum<NS>(interior);
  const double ghostRho = factor * bndRho - intRho;
  double tmp[E::neq];
  if (ghostRho <= 0.0) {
#pragma unroll
    for (int e = 0; e < E::neq; ++e) tmp[e] = bnd[e];
  } else {
#pragma unroll
    for (int e = NS; e < E::neq; ++e) tmp[e] = factor * bnd[e] - interior[e];
#pragma unroll
    for (int q = 0; q < NS; ++q) tmp[q] = fmax(ghostRho * (bnd[q] / bndRho), 0.0);
  }
#pragma unroll
  for (int e = 0; e < E::neq; ++e) out[e] = tmp[e];
}

template <int NS, int NT>
AITHER_HD void FreeStateFromBC(const aither_bc_state &bc, double *fs) {
  using E = Eq<NS, NT>;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) fs[e] = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) fs[q] = bc.density * bc.massFractions[q];
  fs[E::imx] = bc.velocity[0];
  fs[E::imy] = bc.velocity[1];
  fs[E::imz] = bc.velocity[2];
  fs[E::ie] = bc.pressure;
}

// Ghost state for one boundary face. `interior` is the reflected cell for walls and the
// boundary-adjacent cell otherwise (ref: src/procBlock.cpp:2512-2514); `areaVec` is the unit
// normal of the boundary face; surf 1..6; layer 1..g.
// primitive::ApplyFarfieldTurbBC; ref: src/primitive.cpp:83-98
template <int NS, int NT>
AITHER_HD void ApplyFarfieldTurb(const Gas &g, const Transport *tr, double *s, double vx, double vy,
                                 double vz, const aither_bc_state &bc) {
  if (NT < 2 || tr == nullptr) return;
  using E = Eq<NS, NT>;
  const double vmag = sqrt(vx * vx + vy * vy + vz * vz);
  const double tv = bc.turbulenceIntensity * vmag;
  s[E::it] = 1.5 * (tv * tv);
  const double mu = MixtureViscosity<NS>(*tr, Temperature<NS>(g, s), s);
  s[E::it + (NT > 1 ? 1 : 0)] = SpeciesSum<NS>(s) * s[E::it] / (bc.eddyViscosityRatio * mu);
#pragma unroll
  for (int t = 0; t < NT; ++t) s[E::it + t] = fmax(s[E::it + t], kTurbMin);
}

// what a non-reflecting inlet / pressure outlet reads besides the interior state (the implicit
// integrators' extra GetGhostState arguments; ref: src/procBlock.cpp:2506-2522, :6235-6285): time
// step, state at time n and the gradients the previous evaluation left in the boundary-adjacent
// cell, average / maximum outward Mach number of the surface patch
struct BcExtra {
  double dt, stateN[AITHER_MAX_SPECIES + 4 + 2], pressGrad[3], velGrad[9], avgMach, maxMach;
};

template <int NS, int NT>
AITHER_HD void GhostState(const Gas &g, const double *interior, int bcType,
                          const double *areaVec, int surf, const aither_bc_state &bc, int layer,
                          double *ghost, const Transport *tr = nullptr,
                          const BcExtra *ex = nullptr) {
  using E = Eq<NS, NT>;
#pragma unroll
  for (int e = 0; e < E::neq; ++e) ghost[e] = interior[e];
  const bool isLower = surf % 2 == 1;
  double nA[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) nA[d] = isLower ? -1.0 * areaVec[d] : areaVec[d];
  const double vn = interior[E::imx] * nA[0] + interior[E::imy] * nA[1] + interior[E::imz] * nA[2];

  if (bcType == AITHER_BC_SLIP_WALL) {  // ref: :118-133
    ghost[E::imx] = interior[E::imx] - 2.0 * nA[0] * vn;
    ghost[E::imy] = interior[E::imy] - 2.0 * nA[1] * vn;
    ghost[E::imz] = interior[E::imz] - 2.0 * nA[2] * vn;
  } else if (bcType == AITHER_BC_CHARACTERISTIC || bcType == AITHER_BC_INLET) {
    // ref: :289-386 (characteristic), :391-484 (inlet, reflecting form)
    double fs[E::neq];
    FreeStateFromBC<NS, NT>(bc, fs);
    const double SoSInt = SoS<NS>(g, interior);
    const double machInt = fabs(vn) / SoSInt;
    const bool isInlet = bcType == AITHER_BC_INLET;
    bool extrapolate = true;
    if (machInt >= 1.0 && (vn < 0.0 || isInlet)) {  // supersonic inflow
#pragma unroll
      for (int e = 0; e < E::neq; ++e) ghost[e] = fs[e];
      ApplyFarfieldTurb<NS, NT>(g, tr, ghost, bc.velocity[0], bc.velocity[1], bc.velocity[2], bc);
      if (isInlet) extrapolate = false;
    } else if (machInt >= 1.0) {  // supersonic outflow: interior
    } else if (vn < 0.0 || isInlet) {  // subsonic inflow
      const double rhoSoSInt = SpeciesSum<NS>(interior) * SoSInt;
      const double vd0 = fs[E::imx] - interior[E::imx], vd1 = fs[E::imy] - interior[E::imy],
                   vd2 = fs[E::imz] - interior[E::imz];
      ghost[E::ie] = 0.5 * (fs[E::ie] + interior[E::ie] -
                            rhoSoSInt * (nA[0] * vd0 + nA[1] * vd1 + nA[2] * vd2));
      const double fsRho = SpeciesSum<NS>(fs);
      if (isInlet && bc.isNonreflecting && ex != nullptr) {
        // LODI relaxation of density and velocity towards the boundary state (ref: :435-466)
        const double *sN = ex->stateN;
        constexpr double sigma = 0.25;
        const double dt = ex->dt;
        const double rhoN = SpeciesSum<NS>(sN);
        const double sosN = SoS<NS>(g, sN);
        const double rhoSoSN = rhoN * sosN;
        const double deltaPressure = ghost[E::ie] - sN[E::ie];
        const double alpha = sigma * sosN / bc.lengthScale;
        const double rhoNp1 =
            (rhoN + dt * alpha * fsRho + deltaPressure / (sosN * sosN)) / (1.0 + dt * alpha);
#pragma unroll
        for (int q = 0; q < NS; ++q) ghost[q] = rhoNp1 * (fs[q] / fsRho);
        const double k = alpha * (1.0 - ex->maxMach * ex->maxMach);
#pragma unroll
        for (int d = 0; d < 3; ++d)
          ghost[E::imx + d] = (sN[E::imx + d] + dt * k * fs[E::imx + d] -
                               nA[d] * deltaPressure / rhoSoSN) /
                              (1.0 + dt * k);
      } else {
        const double deltaPressure = fs[E::ie] - ghost[E::ie];
        const double rho = fsRho - deltaPressure / (SoSInt * SoSInt);
#pragma unroll
        for (int q = 0; q < NS; ++q) ghost[q] = rho * (fs[q] / fsRho);
        ghost[E::imx] = fs[E::imx] - nA[0] * deltaPressure / rhoSoSInt;
        ghost[E::imy] = fs[E::imy] - nA[1] * deltaPressure / rhoSoSInt;
        ghost[E::imz] = fs[E::imz] - nA[2] * deltaPressure / rhoSoSInt;
      }
      ApplyFarfieldTurb<NS, NT>(g, tr, ghost, bc.velocity[0], bc.velocity[1], bc.velocity[2], bc);
    } else {  // subsonic outflow
      const double intRho = SpeciesSum<NS>(interior);
      const double rhoSoSInt = intRho * SoSInt;
      const double deltaPressure = interior[E::ie] - fs[E::ie];
      const double rho = intRho - deltaPressure / (SoSInt * SoSInt);
#pragma unroll
      for (int q = 0; q < NS; ++q) ghost[q] = rho * (interior[q] / intRho);
      ghost[E::imx] = interior[E::imx] + nA[0] * deltaPressure / rhoSoSInt;
      ghost[E::imy] = interior[E::imy] + nA[1] * deltaPressure / rhoSoSInt;
      ghost[E::imz] = interior[E::imz] + nA[2] * deltaPressure / rhoSoSInt;
      ghost[E::ie] = fs[E::ie];
    }
    if (extrapolate) {
      ExtrapolateHoldMixture<NS, NT>(ghost, 2.0, interior, ghost);
      if (layer > 1) {
        ExtrapolateHoldMixture<NS, NT>(ghost, static_cast<double>(layer), interior, ghost);
        // the characteristic BC re-applies the farfield turbulence to the deeper layers, the
        // inlet BC does not (ref: :375-385 vs :484-486)
        if (!isInlet)
          ApplyFarfieldTurb<NS, NT>(g, tr, ghost, bc.velocity[0], bc.velocity[1], bc.velocity[2], bc);
      }
    }
  } else if (bcType == AITHER_BC_SUPERSONIC_INFLOW) {  // ref: :490-515
    double fs[E::neq];
    FreeStateFromBC<NS, NT>(bc, fs);
#pragma unroll
    for (int e = 0; e < NS + 4; ++e) ghost[e] = fs[e];
    ApplyFarfieldTurb<NS, NT>(g, tr, ghost, bc.velocity[0], bc.velocity[1], bc.velocity[2], bc);
  } else if (bcType == AITHER_BC_SUPERSONIC_OUTFLOW) {  // ref: :522-527
    if (layer > 1) {
#pragma unroll
      for (int e = 0; e < E::neq; ++e) ghost[e] = layer * ghost[e] - interior[e];
    }
  } else if (bcType == AITHER_BC_STAGNATION_INLET) {  // ref: :533-598 (Blazek)
    const double gam = Gamma<NS>(g, interior);
    const double gm1 = gam - 1.0;
    const double sosI = SoS<NS>(g, interior);
    const double rNeg = vn - 2.0 * sosI / gm1;
    const double magSq = VelMagSq<NS>(interior);
    const double cosTheta = -1.0 * vn / sqrt(magSq);
    const double stagSoSsq = sosI * sosI + 0.5 * gm1 * magSq;
    const double sosB = -1.0 * rNeg * gm1 / (gm1 * cosTheta * cosTheta + 2.0) *
                        (1.0 + cosTheta * sqrt((gm1 * cosTheta * cosTheta + 2.0) * stagSoSsq /
                                                   (gm1 * rNeg * rNeg) -
                                               0.5 * gm1));
    const double tb = bc.stagnationTemperature * (sosB * sosB / stagSoSsq);
    const double pb = bc.stagnationPressure * pow(sosB * sosB / stagSoSsq, gam / gm1);
    const double vbMag = sqrt(2.0 / gm1 * (bc.stagnationTemperature - tb));
    double Rmix = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) Rmix += bc.massFractions[q] * g.R[q];
    const double rhoGhost = pb / (Rmix * tb);
#pragma unroll
    for (int q = 0; q < NS; ++q) ghost[q] = rhoGhost * bc.massFractions[q];
    ghost[E::imx] = vbMag * bc.direction[0];
    ghost[E::imy] = vbMag * bc.direction[1];
    ghost[E::imz] = vbMag * bc.direction[2];
    ghost[E::ie] = pb;
    ApplyFarfieldTurb<NS, NT>(g, tr, ghost, ghost[E::imx], ghost[E::imy], ghost[E::imz], bc);
    ExtrapolateHoldMixture<NS, NT>(ghost, 2.0, interior, ghost);
    if (layer > 1) {
      ExtrapolateHoldMixture<NS, NT>(ghost, static_cast<double>(layer), interior, ghost);
      ApplyFarfieldTurb<NS, NT>(g, tr, ghost, ghost[E::imx], ghost[E::imy], ghost[E::imz], bc);
    }
  } else if (bcType == AITHER_BC_PRESSURE_OUTLET) {  // ref: :604-664
    const double SoSInt = SoS<NS>(g, interior);
    const double intRho = SpeciesSum<NS>(interior);
    const double rhoSoSInt = intRho * SoSInt;
    ghost[E::ie] = bc.pressure;
    if (bc.isNonreflecting && ex != nullptr) {
      // LODI relaxation of the pressure with transverse terms (ref: :614-643)
      const double *sN = ex->stateN;
      const double dt = ex->dt;
      const double deltaVel = (interior[E::imx] - sN[E::imx]) * nA[0] +
                              (interior[E::imy] - sN[E::imy]) * nA[1] +
                              (interior[E::imz] - sN[E::imz]) * nA[2];
      constexpr double sigma = 0.25;
      const double rhoN = SpeciesSum<NS>(sN);
      const double sosN = SoS<NS>(g, sN);
      const double rhoSoSN = rhoN * sosN;
      const double k = sigma * sosN * (1.0 - ex->maxMach * ex->maxMach) / bc.lengthScale;
      const double beta = ex->avgMach;
      const double pgn = ex->pressGrad[0] * nA[0] + ex->pressGrad[1] * nA[1] + ex->pressGrad[2] * nA[2];
      const double vnN = sN[E::imx] * nA[0] + sN[E::imy] * nA[1] + sN[E::imz] * nA[2];
      double pGradT[3], velT[3], vgT[9], dVelN[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        pGradT[d] = ex->pressGrad[d] - pgn * nA[d];
        velT[d] = sN[E::imx + d] - vnN * nA[d];
      }
      // tensor::RemoveComponent (include/tensor.hpp:371-379): every row loses its normal part
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double *row = ex->velGrad + 3 * r;
        const double rn = row[0] * nA[0] + row[1] * nA[1] + row[2] * nA[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) vgT[3 * r + c] = row[c] - rn * nA[c];
      }
      // tensor::LinearCombination (:384-389): rows scaled by the normal and summed
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double comb = vgT[c] * nA[0];
        comb += vgT[3 + c] * nA[1];
        comb += vgT[6 + c] * nA[2];
        dVelN[c] = comb;
      }
      double vgSum = 0.0;
#pragma unroll
      for (int q = 0; q < 9; ++q) vgSum += vgT[q];
      double dnSum = 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) dnSum += dVelN[c];
      const double dVelT_dTrans = vgSum - dnSum;
      const double gam = Gamma<NS>(g, sN);
      const double dotT = velT[0] * (pGradT[0] - rhoSoSN * dVelN[0]) +
                          velT[1] * (pGradT[1] - rhoSoSN * dVelN[1]) +
                          velT[2] * (pGradT[2] - rhoSoSN * dVelN[2]);
      const double trans = -0.5 * (dotT + gam * sN[E::ie] * dVelT_dTrans);
      ghost[E::ie] = (sN[E::ie] + rhoSoSN * deltaVel + dt * k * bc.pressure - dt * beta * trans) /
                     (1.0 + dt * k);
    }
    const double deltaPressure = interior[E::ie] - ghost[E::ie];
    const double rho = intRho - deltaPressure / (SoSInt * SoSInt);
#pragma unroll
    for (int q = 0; q < NS; ++q) ghost[q] = rho * (interior[q] / intRho);
    ghost[E::imx] = interior[E::imx] + nA[0] * deltaPressure / rhoSoSInt;
    ghost[E::imy] = interior[E::imy] + nA[1] * deltaPressure / rhoSoSInt;
    ghost[E::imz] = interior[E::imz] + nA[2] * deltaPressure / rhoSoSInt;
    const double gvn = ghost[E::imx] * nA[0] + ghost[E::imy] * nA[1] + ghost[E::imz] * nA[2];
    if (gvn / SoS<NS>(g, ghost) >= 1.0) {
#pragma unroll
      for (int e = 0; e < E::neq; ++e) ghost[e] = interior[e];
    }
#pragma unroll
    for (int e = 0; e < E::neq; ++e) ghost[e] = 2.0 * ghost[e] - interior[e];
    if (layer > 1) {
#pragma unroll
      for (int e = 0; e < E::neq; ++e) ghost[e] = layer * ghost[e] - interior[e];
    }
  }
  // interblock / periodic: filled by the halo exchange
}

}  // namespace aither
