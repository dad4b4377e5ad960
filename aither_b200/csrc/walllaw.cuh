// walllaw.cuh -- wall law of the viscous-wall boundary condition (wallTreatment=wallLaw) as an
// inlineable host/device point function: White & Christoph's compressible law of the wall in the
// form of Nichols & Nelson (2004), y+ of the wall-adjacent cell found with Ridder's method, then
// the wall shear stress, wall heat flux / temperature, wall eddy viscosity and the k and omega the
// law implies at the wall.
//
// Reference: mnucci32/aither v0.10.0 src/wallLaw.cpp:31-289 (AdiabaticBCs :31-88, HeatFluxBCs
// :90-145, IsothermalBCs :147-200 and their helpers), include/wallLaw.hpp:36-95 (constructor:
// yplus0 = exp(-kappa B)), include/utility.hpp:130-184 (FindRoot), include/wallData.hpp:40-57
// (wallVars, SwitchToLowRe = y+ < 10). The reference's wallLaw object keeps the values of its LAST
// function evaluation, and the wall variables are built from those: WallLawCtx does the same.
#pragma once
#include "turbulence.cuh"

namespace aither {

// one record per boundary face of a block (kWallVarsStride doubles), written by the viscous-wall
// ghost-cell kernel for the first ghost layer and read by the viscous flux of that face
// (ref: src/procBlock.cpp:6287-6290, :1286-1300)
constexpr int kWallVarsStride = 16;
enum WallVarSlot { kWvYplus = 0, kWvTau = 1, kWvHeatFlux = 4, kWvMu = 5, kWvMut = 6, kWvRho = 7,
                   kWvT = 8, kWvTke = 9, kWvSdr = 10, kWvVelWall = 11, kWvUtau = 14 };

struct WallVars {
  double yplus, tau[3], heatFlux, mu, mut, rho, t, tke, sdr, utau;
  AITHER_HD bool SwitchToLowRe() const { return yplus < 10.0; }  // include/wallData.hpp:57
};

enum WallLawMode { kWallAdiabatic = 0, kWallHeatFlux = 1, kWallIsothermal = 2 };

template <int NS>
struct WallLawCtx {
  const Gas *g;
  const Transport *tr;
  const double *state;
  double wallDist, vonKarmen, yplus0, velTanMag, tInt, cp, R;
  double beta, gamma, q, phi, yplusWhite, uStar, uplus, tW, rhoW, muW, kW, recovery;
  double heatFlux, yplus, temperature;
  int mode;

  AITHER_HD void SetWallVars(double t) {  // ref: src/wallLaw.cpp:229-237
    tW = t;
    rhoW = state[NS + 3] / (R * t);
    muW = MixtureViscosity<NS>(*tr, t, state) * tr->scaling;
    kW = MixtureEffConductivity<NS>(*tr, t, state);
  }
  AITHER_HD double Func(double yp) {
    // CalcVelocities :264-268
    uplus = (wallDist * rhoW * velTanMag) / (muW * yp);
    uStar = velTanMag / uplus;
    if (mode == kWallHeatFlux) {  // CalcWallTemperature :220-227, SetWallVars
      temperature = tInt + recovery * uStar * uStar * uplus * uplus /
                               (2.0 * cp + heatFlux * muW / (rhoW * kW * uStar));
      SetWallVars(temperature);
    }
    gamma = recovery * uStar * uStar / (2.0 * cp * tW);  // UpdateGamma :186-191
    if (mode == kWallIsothermal) {  // CalcHeatFlux :211-218
      const double tmp = (tInt / tW - 1.0 + gamma * uplus * uplus) / uplus;
      heatFlux = tmp * (rhoW * tW * kW * uStar) / muW;
    }
    // UpdateConstants :193-198
    beta = heatFlux * muW / (rhoW * tW * kW * uStar);
    q = sqrt(beta * beta + 4.0 * gamma);
    phi = asin(-beta / q);
    // CalcYplusWhite :200-205
    yplusWhite = exp((vonKarmen / sqrt(gamma)) * (asin((2.0 * gamma * uplus - beta) / q) - phi)) * yplus0;
    yplus = yp;
    // CalcYplusRoot :239-244
    const double ku = vonKarmen * uplus;
    return yp - (uplus + yplusWhite - yplus0 * (1.0 + ku + 0.5 * ku * ku + (1.0 / 6.0) * (ku * ku * ku)));
  }
  AITHER_HD static double Sgn(double v) { return static_cast<double>((0.0 < v) - (v < 0.0)); }
  // Ridder's method; ref: include/utility.hpp:130-184
  AITHER_HD void FindRoot(double x1, double x2, double tol) {
    double f1 = Func(x1);
    double f2 = Func(x2);
    if (Sgn(f1) == Sgn(f2) && Sgn(f1) != 0.0) return;
    for (int ii = 0; ii < 100; ++ii) {
      const double x3 = 0.5 * (x1 + x2);
      const double f3 = Func(x3);
      if (f3 == 0.0) return;
      const double denom = sqrt(fabs(f3 * f3 - f1 * f2));
      if (denom == 0.0) return;
      const double x4 = x3 + (x3 - x1) * (Sgn(f1 - f2) * f3) / denom;
      const double f4 = Func(x4);
      if (f4 == 0.0) return;
      if (Sgn(f4) != Sgn(f3)) {
        x1 = x3;
        f1 = f3;
        x2 = x4;
        f2 = f4;
      } else if (Sgn(f4) != Sgn(f1)) {
        x2 = x4;
        f2 = f4;
      } else {
        x1 = x4;
        f1 = f4;
      }
      if (fabs(x2 - x1) <= tol) return;
    }
  }
};

// wallLaw::AdiabaticBCs / HeatFluxBCs / IsothermalBCs. `area`: unit normal pointing out of the
// domain; `interior`: the state the ghost cell mirrors (ref: src/ghostStates.cpp:149-256)
template <int NS, int NT>
AITHER_HD void WallLawEval(const Gas &g, const Transport &tr, const aither_bc_state &bc, int mode,
                           const double *interior, double wallDist, const double *area,
                           bool isLower, WallVars &wv) {
  WallLawCtx<NS> c;
  c.g = &g;
  c.tr = &tr;
  c.state = interior;
  c.mode = mode;
  c.wallDist = wallDist;
  c.vonKarmen = bc.vonKarmen;
  c.yplus0 = exp(-bc.vonKarmen * bc.wallConstant);
  c.beta = c.gamma = c.q = c.phi = c.yplusWhite = c.uStar = c.uplus = 0.0;
  c.yplus = 0.0;
  double vel[3], velTan[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) vel[d] = interior[NS + d] - bc.velocity[d];
  const double vn = vel[0] * area[0] + vel[1] * area[1] + vel[2] * area[2];
#pragma unroll
  for (int d = 0; d < 3; ++d) velTan[d] = vel[d] - vn * area[d];
  c.velTanMag = sqrt(velTan[0] * velTan[0] + velTan[1] * velTan[1] + velTan[2] * velTan[2]);
  c.tInt = Temperature<NS>(g, interior);
  const auto mix = Mixture<NS>(g, interior);
  c.cp = mix.cp;
  c.R = 0.0;  // eos DensityTP: rho = p / (sum Y_s R_s T), src/eos.cpp:111-115
#pragma unroll
  for (int q = 0; q < NS; ++q) c.R += interior[q] / mix.rho * g.R[q];
  const double gam = mix.cp / mix.cv;
  c.recovery = pow((4.0 * gam) / (9.0 * gam - 5.0), 1.0 / 3.0);  // CalcRecoveryFactor :286-289
  if (mode == kWallAdiabatic) {  // Crocco-Busemann :46-52
    c.heatFlux = 0.0;
    c.temperature = 0.0;
    c.SetWallVars(c.tInt + 0.5 * c.recovery * c.velTanMag * c.velTanMag / c.cp);
  } else if (mode == kWallHeatFlux) {  // :96-105
    c.heatFlux = bc.heatFlux;
    c.temperature = c.tInt;
    c.SetWallVars(c.tInt);
  } else {  // :148-160
    c.heatFlux = 0.0;
    c.temperature = bc.temperature;
    c.SetWallVars(bc.temperature);
  }
  c.FindRoot(1.0e1, 1.0e4, 1.0e-8);
  double mutW = 0.0;
  wv.tke = 0.0;
  wv.sdr = 0.0;
  if (NT > 0) {  // CalcTurbVars :270-284, EddyVisc :246-262
    const double a = 2.0 * c.gamma * c.uplus - c.beta;
    const double dYplusWhite = 2.0 * c.yplusWhite * c.vonKarmen * sqrt(c.gamma) / c.q *
                               sqrt(fmax(1.0 - (a * a) / (c.q * c.q), 0.0));
    const double ku = c.vonKarmen * c.uplus;
    mutW = c.muW * (1.0 + dYplusWhite - c.vonKarmen * c.yplus0 * (1.0 + ku + 0.5 * ku * ku)) -
           MixtureViscosity<NS>(tr, c.tInt, interior) * tr.scaling;
    mutW = fmax(mutW, 0.0);
    double wi = 6.0 * c.muW / (TurbWallBeta(tr.turbModel) * c.rhoW * wallDist * wallDist);
    wi *= tr.scaling;
    double wo = c.uStar / (sqrt(kw::betaStar) * c.vonKarmen * wallDist);
    wo *= tr.scaling;
    wv.sdr = sqrt(wi * wi + wo * wo);
    wv.tke = wv.sdr * mutW / SpeciesSum<NS>(interior) * (1.0 / tr.scaling);
  }
  wv.heatFlux = c.heatFlux;
  wv.yplus = c.yplus;
  wv.rho = c.rhoW;
  wv.t = mode == kWallAdiabatic ? c.tW : c.temperature;
  wv.mu = c.muW;
  wv.mut = mutW;
  wv.utau = c.uStar;  // wallVars::frictionVelocity_ (src/wallLaw.cpp:81,139,194)
  const double tauMag = c.uStar * c.uStar * c.rhoW;  // ShearStressMag
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    wv.tau[d] = tauMag * velTan[d] / c.velTanMag;
    if (!isLower) wv.tau[d] *= -1.0;
  }
}

}  // namespace aither
